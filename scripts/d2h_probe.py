"""Probe host<->device transfer strategies on the GPU box (used to choose fit_em's output path)."""
import time
import numpy as np
import torch

n = 800_000_000  # 3.2 GB fp32
x = torch.rand(n, device="cuda")
torch.cuda.synchronize()
def t(f, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print("%-40s %.3f s  (%.1f GB/s)" % (name, time.perf_counter() - t0, n * 4 / 1e9 / (time.perf_counter() - t0)), flush=True)
    return r
a = t(lambda: x.cpu().numpy(), "x.cpu().numpy() (pageable)")
a = t(lambda: x.cpu().numpy(), "x.cpu().numpy() again")
pin = t(lambda: torch.empty(n, dtype=torch.float32, pin_memory=True), "alloc pinned 3.2GB")
t(lambda: pin.copy_(x, non_blocking=True), "copy into pinned")
t(lambda: pin.copy_(x, non_blocking=True), "copy into pinned again")
b = t(lambda: np.empty(n, np.float32), "np.empty")
t(lambda: b.fill(0), "first touch np (page faults)")
host = torch.from_numpy(b)
t(lambda: host.copy_(x), "copy into touched pageable")
y = np.random.poisson(0.5, size=(1_000_000, 500)).astype(np.float32)
t(lambda: torch.from_numpy(y).cuda(), "H2D pageable 2GB")
t(lambda: torch.from_numpy(y).cuda(), "H2D pageable 2GB again")
