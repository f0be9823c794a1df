"""Phase breakdown of decode_latent at the headline shape (device-resident inputs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200 import ops, core
from poor_man_gplvm_b200.estep import EStep
from poor_man_gplvm_b200.synthetic import make_dataset_torch
T, N, K = int(os.environ.get("T", 1000000)), 500, 400
dev = torch.device("cuda")
d = make_dataset_torch(T, N, K, dev, seed=1234)
y = d["y"].to(torch.float32).contiguous()
model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, movement_variance=1.0, device=dev)
em = model.fit_em(y, n_iter=6, return_device=True)
tun = em["tuning"] if isinstance(em["tuning"], torch.Tensor) else torch.as_tensor(em["tuning"], device=dev)
del em
marks = []
def hook(name):
    torch.cuda.synchronize(); marks.append((name, time.perf_counter()))
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = sync()
    P, logP, M, logM, op = model._transition_pack({})
    ma_n, ma_l = model._masks(None, None, T)
    t1 = sync()
    es = EStep(y, op, ma_n, ma_l, 1.0)
    t2 = sync()
    ops.PHASE_HOOK = hook; marks.clear(); marks.append(("start", time.perf_counter()))
    res = es.run(tun, want_gamma=True, want_gamma_lat=True, want_dyn=True, want_r=True)
    t3 = sync(); ops.PHASE_HOOK = None; marks.append(("end", t3))
    print("   run phases:", " ".join("%s %.2f" % (n1, (t1 - t0) * 1e3) for (n0, t0), (n1, t1) in zip(marks[:-1], marks[1:])),
          "relays", res.n_relay_fwd, res.n_relay_bwd)
    lg = torch.log(res.gamma)
    t4 = sync()
    log_acc = model._transition_counts(es, res, logP, logM)
    t5 = sync()
    tp = core.compute_transition_posterior_prob(log_acc)
    t6 = sync()
    print("rep %d: pack %.1f  EStep() %.1f  run %.1f  log %.1f  xi %.1f  post %.1f   total %.1f ms" %
          (rep, *(1e3 * (b - a) for a, b in [(t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6)]), 1e3 * (t6 - t0)), flush=True)
    del es, res, lg, log_acc, tp
