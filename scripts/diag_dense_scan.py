"""Precision of the lockstep tensor-core scan vs the CUDA-core general path and the fp64 oracle on random inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poor_man_gplvm_b200 import ops, gp_kernel as gpk
from poor_man_gplvm_b200.estep import EStep
dev = torch.device("cuda")


def run(K, T, chunk, halo, sigma, tc):
    x = np.arange(K, dtype=np.float64)
    ck = (np.exp(-np.abs(x[:, None] - x[None, :]) / 150.0) + 0.02).astype(np.float32)
    P, logP, M, logM = gpk.create_transition_prob_1d(np.arange(K), np.arange(2), 1.0, 0.02, 0.05, custom_kernel=ck)
    hostop = gpk.move_operator_host(K, 1.0, ck, p_move_to_jump=0.02)
    rng = np.random.default_rng(K + T)
    ll = (rng.standard_normal((T, K)) * sigma - 40.0).astype(np.float32)
    op = ops.MoveOperator(hostop, M, dev, P0=P[0], dense_tc=tc)
    es = EStep(torch.zeros((T, 1), device=dev), op, None, None, 1.0, halo=halo, chunk_len=chunk)
    es.ll.copy_(torch.from_numpy(ll).to(dev))
    es.emission = lambda tuning, es=es: es.ll
    es.device_repair = False
    es.seam_tol = 1.0            # never repair: look at the raw seam errors
    res = es.run(None, want_gamma=True, want_gamma_lat=True)
    S = es.S
    return (res.alpha.cpu().numpy().astype(np.float64), res.gamma.cpu().numpy().astype(np.float64),
            es.err_host[:S].numpy().copy(), es.err_host[S:2 * S].numpy().copy(), ll, P, M)


for K, T, chunk, halo, sigma in ((1024, 400, 400, 0, 3.0), (1088, 400, 400, 0, 3.0), (2000, 400, 400, 0, 3.0),
                                 (2000, 400, 400, 0, 0.5), (2000, 1200, 100, 256, 3.0), (512, 1200, 100, 256, 3.0)):
    a1, g1, ef1, eb1, ll, P, M = run(K, T, chunk, halo, sigma, True)
    a0, g0, ef0, eb0, _, _, _ = run(K, T, chunk, halo, sigma, False)
    Kf = K
    # oracle in fp64 from the same ll
    Pd, Md = P.astype(np.float64), M.astype(np.float64)
    al = np.zeros((T, 2, K)); be = np.zeros((T, 2, K))
    prev = np.full((2, K), 0.5 / K)
    for t in range(T):
        L = np.exp(ll[t].astype(np.float64) - ll[t].max())
        pr0 = (Md[0, 0] * prev[0] + Md[1, 0] * prev[1]) @ Pd[0]
        pr1 = (Md[0, 1] * prev[0].sum() + Md[1, 1] * prev[1].sum()) / K
        v = np.stack([pr0 * L, pr1 * L]); prev = v / v.sum(); al[t] = prev
    print("K=%d T=%d chunk=%d halo=%d sigma=%.1f" % (K, T, chunk, halo, sigma))
    for name, a in (("tensor-core", a1), ("cuda-core", a0)):
        rel = np.abs(a - al) / np.maximum(al, 1e-30)
        big = al > 1e-6
        print("   %-11s alpha: max abs err %.3g, max rel err over entries > 1e-6: %.3g, median rel %.3g"
              % (name, np.abs(a - al).max(), rel[big].max(), np.median(rel[big])))
    print("   gamma: max |tc - cuda| %.3g" % np.abs(g1 - g0).max())
    if halo:
        print("   seam errors tc   fwd max %.3g median %.3g | bwd max %.3g median %.3g" % (ef1.max(), np.median(ef1), eb1.max(), np.median(eb1)))
        print("   seam errors cuda fwd max %.3g median %.3g | bwd max %.3g median %.3g" % (ef0.max(), np.median(ef0), eb0.max(), np.median(eb0)))
