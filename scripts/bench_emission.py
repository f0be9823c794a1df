"""Times the emission GEMM alone at the headline shape (CUDA events, repeated launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poor_man_gplvm_b200 import ops
from poor_man_gplvm_b200.synthetic import make_dataset_torch
T, N, K = 1000000, 500, 400
dev = torch.device("cuda")
d = make_dataset_torch(T, N, K, dev, seed=1)
y = d["y"].to(torch.float32).contiguous()
em = ops.EmissionOperands(y, None, ones_col=True)
tun = torch.rand((K, N), device=dev) + 0.05
ll = torch.empty((T, K), device=dev)
extra = [torch.empty(int(x), device=dev) for x in (os.environ.get("PAD_MB", "0").split(","))] if os.environ.get("PAD_MB") else []
for _ in range(3):
    em.loglik(tun, None, 1.0, out=ll)
torch.cuda.synchronize()
ts = []
for _ in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); em.loglik(tun, None, 1.0, out=ll); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.environ.get("TAG", ""), "emission ms: min %.3f med %.3f max %.3f" % (min(ts), float(np.median(ts)), max(ts)),
      "ptrs", hex(em.A16.data.data_ptr()), hex(ll.data_ptr()), flush=True)
