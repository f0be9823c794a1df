"""decode_latent + decode_latent_naive_bayes at the headline shape, twice (second call = warm): the profiling target
for the general scan kernels and the transition-count GEMM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200.synthetic import make_dataset_torch
T, N, K = int(os.environ.get("T", 1000000)), 500, 400
dev = torch.device("cuda")
d = make_dataset_torch(T, N, K, dev, seed=1234)
y = d["y"].to(torch.float32).contiguous()
m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, device=dev)
tun = (d["tuning_true"] * 1.03).contiguous()
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); r = m.decode_latent(y, tuning=tun, return_device=True); e1.record()
    nb = m.decode_latent_naive_bayes(y, tuning=tun, return_device=True); e2.record()
    torch.cuda.synchronize()
    print("rep %d: decode_latent %.2f ms, naive bayes %.2f ms, lml %.6e" % (rep, e0.elapsed_time(e1), e1.elapsed_time(e2), r["log_marginal_final"]), flush=True)
    del r, nb
