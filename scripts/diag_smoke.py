import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poor_man_gplvm_b200 as pmg
from oracle import ref_numpy as ref
from poor_man_gplvm_b200.synthetic import make_dataset
N, K, T = 20, 64, 600
d = make_dataset(T, N, K, seed=1)
model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0)
oracle = ref.OraclePoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0, dtype=np.float64, tuning_basis=model.tuning_basis, params=model.params)
lp0, _ = model.init_latent_posterior(T, key=0)
kw = dict(n_iter=2, log_posterior_init=lp0, m_step_maxiter=20, m_step_tol=-1)
want = oracle.fit_em(d["y"], **kw)
for env in [{}, {"PMG_MSTEP_LAG": "0"}, {"PMG_NO_SPECULATE": "1"}, {"PMG_MSTEP_LAG": "0", "PMG_NO_SPECULATE": "1"}, {"PMG_SEAM_FLOOR": "1e-20"}]:
    os.environ.update(env)
    m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0)
    got = m.fit_em(d["y"], **kw)
    lg, lw = np.array(got["log_marginal_l"]), np.array(want["log_marginal_l"])
    print(env, "lml rel", np.max(np.abs(lg - lw) / np.abs(lw)), "post err", np.max(np.abs(got["posterior_latent_marg"] - want["posterior_latent_marg"])),
          "tuning rel", np.max(np.abs(got["tuning"] - want["tuning"]) / want["tuning"]), m._last_estep_info["per_iter"], m._last_estep_info["n_chain"])
    for k in env: os.environ.pop(k)
