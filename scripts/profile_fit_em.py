"""cProfile of fit_em(y_host) at the headline shape: where does the end-to-end wall time go on the host?"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200.synthetic import make_dataset_torch
T, N, K = int(os.environ.get("T", 1000000)), 500, 400
dev = torch.device("cuda")
y_host = make_dataset_torch(T, N, K, dev, seed=1234)["y"].to(torch.float32).cpu().numpy()
model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, movement_variance=1.0, device=dev)
os.environ.setdefault("PMG_TIMING", "1")
for rep in range(int(os.environ.get("REPS", 3))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if rep == -1:
        pr = cProfile.Profile(); pr.enable()
    em = model.fit_em(y_host, key=5, n_iter=(int(os.environ.get("WARM_ITERS", 20)) if rep == 0 else 20))
    torch.cuda.synchronize()
    if rep == -1:
        pr.disable()
    print("rep", rep, "wall %.3f s" % (time.perf_counter() - t0), flush=True)
    del em
pass
