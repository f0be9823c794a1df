"""Host-side profile (cProfile) of the time-sharded EM iteration at the strong-scaling shape: where does the launching
thread spend its time when the per-rank GPU work is ~1.4 ms?  Run under torchrun; rank 0 prints."""
import os, sys, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200.core import EMLoop
from poor_man_gplvm_b200.shard import TimeShard
from poor_man_gplvm_b200.synthetic import make_dataset_torch
N, K, T = 500, 400, int(os.environ.get("T", 1000000))
Tr = int(os.environ.get("T_RANK", T // world))           # bins per rank (default: strong split; override to emulate 8 ranks)
model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, device=dev)
model.params = np.random.default_rng(1).standard_normal((model.n_basis, N)).astype(np.float32)
P, logP, M, logM, op = model._transition_pack({})
y = make_dataset_torch(Tr, N, K, dev, seed=1234 + rank, tuning_seed=1234)["y"].to(torch.float32).contiguous()
g = torch.Generator(device=dev); g.manual_seed(99 + rank)
p0 = torch.rand((Tr, K), generator=g, device=dev); lp0 = torch.log(p0 / p0.sum(dim=1, keepdim=True))
ma_n, ma_l = model._masks(None, None, Tr)
loop = EMLoop(model, y, op, ma_n, ma_l, 1.0, model.tuning_basis, lp0, model.param_prior_std, 0.01, 1000, 1e-6, shard=TimeShard())
for _ in range(12):
    loop.iteration(speculate=True)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    loop.iteration(speculate=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    loop.iteration(speculate=True)
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    print("world %d, %d bins per rank: %.3f ms per EM iteration (wall, unprofiled)" % (world, Tr, dt * 1e3))
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue())
dist.barrier()
dist.destroy_process_group()
