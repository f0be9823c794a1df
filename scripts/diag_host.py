"""Host-side time of each operator call inside the EM loop at the headline shape (where does the launch thread wait?)."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200 import ops
from poor_man_gplvm_b200.core import EMLoop
from poor_man_gplvm_b200.synthetic import make_dataset_torch
T, N, K = int(os.environ.get("T", 1000000)), 500, 400
dev = torch.device("cuda")
data = make_dataset_torch(T, N, K, dev, seed=1234)
y = data["y"].to(torch.float32).contiguous()
model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, movement_variance=1.0, device=dev)
model.params = np.random.default_rng(1).standard_normal((model.n_basis, N)).astype(np.float32)
P, logP, M, logM, op = model._transition_pack({})
ma_n, ma_l = model._masks(None, None, T)
post0 = torch.rand((T, K), device=dev)
lp0 = torch.log(post0 / post0.sum(dim=1, keepdim=True)); del post0
loop = EMLoop(model, y, op, ma_n, ma_l, 1.0, model.tuning_basis, lp0, model.param_prior_std, 0.01, 1000, 1e-6)
del lp0
acc = collections.defaultdict(list)
def wrap(name):
    f = getattr(ops, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[name].append((time.perf_counter() - t0) * 1e3); return r
    setattr(ops, name, g)
for n in ["emission_prepare_f16", "emission_poisson_f16", "atb_f16", "mstep_adam", "forward_compact",
          "backward_compact", "seam_check"]:
    wrap(n)
orig_read = loop.es._read_err
def read_err():
    t0 = time.perf_counter(); r = orig_read(); acc["_read_err(sync)"].append((time.perf_counter() - t0) * 1e3); return r
loop.es._read_err = read_err
for i in range(12):
    if i == 4:
        acc.clear(); torch.cuda.synchronize(); t_begin = time.perf_counter()
    loop.iteration()
torch.cuda.synchronize()
print("ms/iter wall %.3f" % ((time.perf_counter() - t_begin) * 1e3 / 8))
for k, v in acc.items():
    print("%-24s calls/iter %.1f  mean %.3f ms  max %.3f ms  sum/iter %.3f" % (k, len(v) / 8, np.mean(v), np.max(v), np.sum(v) / 8))
