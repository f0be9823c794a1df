"""Host-side profile of decode_latent_naive_bayes at the config-C shape (where does the wall time between kernels go)."""
import sys, os, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200.synthetic import make_dataset_torch
N, K, T = 1000, 200, int(os.environ.get("T", 1250000))
dev = torch.device("cuda")
d = make_dataset_torch(T, N, K, dev, seed=4321)
y = d["y"].to(torch.float32).contiguous()
tun = (d["tuning_true"] * 1.05).contiguous()
m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, device=dev)
for _ in range(3):
    out = m.decode_latent_naive_bayes(y, tuning=tun, return_device=True)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    out = m.decode_latent_naive_bayes(y, tuning=tun, return_device=True)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue())
