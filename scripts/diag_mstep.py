import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from poor_man_gplvm_b200 import ops, gp_kernel as gpk
dev = lambda a: torch.as_tensor(a, device="cuda")
for (K, N, ls, maxiter, tol) in [(100, 30, 10.0, 1000, 1e-6), (400, 500, 10.0, 1000, 1e-6)]:
    rng = np.random.default_rng(K + N)
    basis = gpk.generate_basis(ls, K); B = basis.shape[1]
    tw = (rng.random(K) * 50 + 1).astype(np.float32)
    yw = (rng.random((K, N)) * tw[:, None] * 1.5).astype(np.float32)
    W0 = rng.standard_normal((B, N)).astype(np.float32)
    outs = {}
    for lag in ("0", "1"):
        os.environ["PMG_MSTEP_LAG"] = lag
        W = dev(W0.copy()); st = ops.AdamState(W)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lh, eh, n_it, fin, tuning = ops.mstep_adam(dev(basis), dev(yw), dev(tw), W, st, 1.0, 0.01, maxiter, tol)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        outs[lag] = [x.cpu().numpy().copy() for x in (lh, eh, n_it, fin, tuning, W, st.mu, st.nu, st.count)]
        print("K %d N %d lag %s: n_iter %d, %.2f ms, %.2f us/step" % (K, N, lag, int(n_it), dt * 1e3, dt * 1e6 / int(n_it)))
    a, b = outs["0"], outs["1"]
    names = ["loss_hist", "err_hist", "n_iter", "final", "tuning", "W", "mu", "nu", "count"]
    for nm, x, y in zip(names, a, b):
        if not np.array_equal(x, y):
            d = np.nonzero(np.ravel(x != y))[0]
            print("   differs:", nm, "first idx", d[0], "n", d.size, "vals", np.ravel(x)[d[0]], np.ravel(y)[d[0]],
                  "max abs", np.max(np.abs(x.astype(np.float64) - y.astype(np.float64))))
