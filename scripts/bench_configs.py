"""Timing of the other BASELINE.json configs on one GPU (parity-test shapes, not bench lines)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200.synthetic import make_dataset_torch
dev = torch.device("cuda")
def sync(): torch.cuda.synchronize(); return time.perf_counter()
out = {}
# configs[1]: single-session scale, fit_em n_iter=50 through the public API (device-resident y, host results)
N, K, T = 200, 100, 100000
y = make_dataset_torch(T, N, K, dev, seed=1)["y"].to(torch.float32)
m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, device=dev)
m.fit_em(y, n_iter=3)
t0 = sync(); r = m.fit_em(y, n_iter=50); t1 = sync()
out["session_fit_em_50"] = {"N": N, "K": K, "T": T, "wall_s": t1 - t0, "bins_iters_per_s": T * 50 / (t1 - t0),
                            "lml_first_last": [float(r["log_marginal_l"][0]), float(r["log_marginal_l"][-1])]}
del y, m, r
# configs[2]: naive Bayes decode, per-GPU share of T=1e7 over 8 GPUs
N, K, T = 1000, 200, 1250000
y = make_dataset_torch(T, N, K, dev, seed=2)["y"].to(torch.float32)
m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, device=dev)
tun = torch.rand((K, N), device=dev) + 0.05
m.decode_latent_naive_bayes(y, tuning=tun, return_device=True)
t0 = sync()
for _ in range(3): nb = m.decode_latent_naive_bayes(y, tuning=tun, return_device=True)
t1 = sync()
out["naive_bayes_share_of_1e7_over_8"] = {"N": N, "K": K, "T": T, "ms": (t1 - t0) / 3 * 1e3, "bins_per_s": T * 3 / (t1 - t0)}
del y, m, nb
# configs[4]: large-state stress, per-GPU share of T=1e6 over 8 GPUs
N, K, T = 300, 2000, 125000
y = make_dataset_torch(T, N, K, dev, seed=3)["y"].to(torch.float32)
m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, device=dev)
m.fit_em(y, n_iter=2, return_device=True)
t0 = sync(); r = m.fit_em(y, n_iter=6, return_device=True); t1 = sync()
info = m._last_estep_info
out["stress_share_of_1e6_over_8_fit_em_6"] = {"N": N, "K": K, "T": T, "wall_s": t1 - t0, "bins_iters_per_s": T * 6 / (t1 - t0),
                                              "n_chain": info["n_chain"], "relays": [p[:2] for p in info["per_iter"]]}
print(json.dumps(out, indent=1))
