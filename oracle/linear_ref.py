"""Independent linear-space derivation of the E-step (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED — see ``oracle/__init__.py``.  This is SURVEY.md Appendix A: the
max-shifted, per-step-normalised alpha/beta recursion that the CUDA kernels
implement, written in NumPy (fp64 by default) so that it can arbitrate between
the log-space restatement (``ref_numpy.py`` following decoder.py:151-332) and the
GPU results.  It is *not* a restatement of reference code; it is the algebraic
identity the reference's log-space scan satisfies.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln


def emission_gemm_form(y, tuning, ma_neuron, ma_latent, dtype=np.float64):
    """ll = (Y*m) log(lam)^T - m.lam^T - sum_n m lgamma(Y+1)   (decoder.py:30-48 in GEMM form)."""
    y = np.asarray(y, dtype=dtype)
    lam = np.asarray(tuning, dtype=dtype) + dtype(1e-20)
    m = np.asarray(ma_neuron, dtype=dtype)
    if m.ndim == 1:
        ll = y @ (np.log(lam) * m[None, :]).T - (lam * m[None, :]).sum(axis=1)[None, :]
        ll -= (gammaln(y + 1.0) * m[None, :]).sum(axis=1)[:, None]
    else:
        ll = (y * m) @ np.log(lam).T - m @ lam.T - (gammaln(y + 1.0) * m).sum(axis=1)[:, None]
    return np.where(np.asarray(ma_latent).astype(bool)[None, :], ll, dtype(-1e20))


def forward(ll, P0, M, likelihood_scale=1.0, carry=None, dtype=np.float64):
    """alpha-hat recursion.  Returns alpha[T,2,K], lmr[T] (= log c_t + s*m_t), prior[T,2,K]."""
    ll = np.asarray(ll, dtype=dtype)
    T, K = ll.shape
    P0 = np.asarray(P0, dtype=dtype); M = np.asarray(M, dtype=dtype)
    a_prev = np.full((2, K), 1.0 / (2 * K), dtype=dtype) if carry is None else np.asarray(carry, dtype=dtype)
    alpha = np.empty((T, 2, K), dtype=dtype)
    prior_all = np.empty((T, 2, K), dtype=dtype)
    lmr = np.empty(T, dtype=dtype)
    s = dtype(likelihood_scale)
    for t in range(T):
        a = M.T @ a_prev                               # [d', x]
        prior = np.stack([a[0] @ P0, np.full(K, a[1].sum() / K, dtype=dtype)])
        m = ll[t].max()
        L = np.exp(s * (ll[t] - m))
        u = prior * L[None, :]
        c = u.sum()
        a_prev = u / c
        alpha[t], prior_all[t] = a_prev, prior
        lmr[t] = np.log(c) + s * m
    return alpha, lmr, prior_all


def backward(ll, alpha, P0, M, likelihood_scale=1.0, beta_init=None, dtype=np.float64, want_r=False):
    """beta recursion scaled so that sum(alpha_t * beta_t) = 1.  Returns gamma[T,2,K]
    (and r[T,2,K] with r[t] = L_t * beta_t / c_t scaled consistently, r[0] unused)."""
    ll = np.asarray(ll, dtype=dtype)
    T, K = ll.shape
    P0 = np.asarray(P0, dtype=dtype); M = np.asarray(M, dtype=dtype)
    s = dtype(likelihood_scale)
    beta = np.ones((2, K), dtype=dtype) if beta_init is None else np.asarray(beta_init, dtype=dtype)
    gamma = np.empty((T, 2, K), dtype=dtype)
    r_all = np.zeros((T, 2, K), dtype=dtype)
    g = alpha[T - 1] * beta
    z = g.sum()
    gamma[T - 1] = g / z
    beta = beta / z
    for t in range(T - 2, -1, -1):
        m = ll[t + 1].max()
        L = np.exp(s * (ll[t + 1] - m))
        r = L[None, :] * beta                          # [d', x'] (unnormalised)
        w = np.stack([P0 @ r[0], np.full(K, r[1].sum() / K, dtype=dtype)])
        b = M @ w                                      # [d, x]
        z = (alpha[t] * b).sum()
        beta = b / z
        gamma[t] = alpha[t] * beta
        r_all[t + 1] = r / z
    return (gamma, r_all) if want_r else gamma


def xi_from_alpha_r(alpha, r, P, M):
    """sum_t xi_t[d,d',x,x'] = M[d,d'] P_{d'}[x,x'] sum_t alpha_t[d,x] r_{t+1}[d',x']  (SURVEY S4)."""
    T, _, K = alpha.shape
    A = alpha[:-1].reshape(T - 1, 2 * K)
    R = r[1:].reshape(T - 1, 2 * K)
    G = (A.T @ R).reshape(2, K, 2, K).transpose(0, 2, 1, 3)      # [d, d', x, x']
    return G * M[:, :, None, None] * np.asarray(P)[None, :, :, :]


def e_step(y, tuning, P, M, ma_neuron, ma_latent, likelihood_scale=1.0, dtype=np.float64, want_xi=False):
    ll = emission_gemm_form(y, tuning, ma_neuron, ma_latent, dtype)
    alpha, lmr, prior = forward(ll, P[0], M, likelihood_scale, dtype=dtype)
    out = backward(ll, alpha, P[0], M, likelihood_scale, dtype=dtype, want_r=want_xi)
    res = {"ll": ll, "alpha": alpha, "lmr": lmr, "log_marginal": lmr.sum(), "prior": prior}
    if want_xi:
        res["gamma"], r = out
        res["xi"] = xi_from_alpha_r(alpha, r, np.asarray(P, dtype=dtype), np.asarray(M, dtype=dtype))
    else:
        res["gamma"] = out
    return res


def fit_em_linear(model, y, hyperparam={}, n_iter=20, log_posterior_init=None, ma_neuron=None, ma_latent=None,
                  likelihood_scale=1.0, m_step_step_size=0.01, m_step_maxiter=1000, m_step_tol=1e-6,
                  m_step_schedule=None):
    """The reference EM driver (core.py:592-713, :802-849) with the restated M-step of ``ref_numpy`` and the
    E-step in linear space (``e_step`` above) instead of the per-step log-space joint.  The log-space restatement
    costs 4K^2 exponentials per bin; this one two K x K mat-vecs, which is what makes parity runs at the real
    shapes (K=400 / K=2000, thousands of bins) affordable.  ``model``: an ``OraclePoissonGPLVMJump1D`` in
    fp64.  PINNED through ``tests/test_oracle_golden.py::test_linear_em_driver_matches_reference_source``: it
    reproduces the reference source's README run (golden fixture) to 1e-9.
    ``m_step_schedule``: optional list of Adam step counts, one per EM iteration, that replaces the stopping rule
    (``n_iter`` of that M-step is pinned: maxiter = the count, tol = -1) -- the stopping step of the default rule is
    decided by a relative loss change of 1e-6, i.e. by rounding, so implementations are compared step for step."""
    from . import ref_numpy as ref
    fd = np.float64
    y_ = np.asarray(y, dtype=fd)
    hp = dict(hyperparam)
    prior_std = hp.get("param_prior_std", model.param_prior_std)
    P, _, M, _ = model._transitions(hp)
    P = np.asarray(P, dtype=fd); M = np.asarray(M, dtype=fd)
    ma_neuron = model.ma_neuron_default if ma_neuron is None else np.asarray(ma_neuron, dtype=fd)
    ma_latent = model.ma_latent_default if ma_latent is None else np.asarray(ma_latent, dtype=fd)
    basis = np.asarray(model.tuning_basis, dtype=fd)
    params = np.asarray(model.params, dtype=fd)
    lp_curr = np.asarray(log_posterior_init, dtype=fd)
    opt_state = ref.adam_init(params)
    lml_l, n_it, losses = [], [], []
    for it in range(n_iter):
        yw, tw = ref.get_statistics(lp_curr, y_)
        mi, tl = (m_step_maxiter, m_step_tol) if m_step_schedule is None else (int(m_step_schedule[it]), -1.0)
        m_res = ref.adam_run(params, opt_state, prior_std, basis, yw, tw, step_size=m_step_step_size,
                             maxiter=mi, tol=tl)
        params, opt_state = m_res["params"], m_res["opt_state"]
        n_it.append(int(m_res["n_iter"])); losses.append(float(m_res["final_loss"]))
        tuning = ref.get_tuning_softplus(params, basis)
        es = e_step(y_, tuning, P, M, ma_neuron, ma_latent, likelihood_scale, dtype=fd)
        gamma = es["gamma"]
        with np.errstate(divide="ignore"):
            lp_curr = np.log(gamma.sum(axis=1))
        lml_l.append(float(es["log_marginal"]))
    return {"params": params, "tuning": tuning, "log_marginal_l": lml_l, "posterior": gamma,
            "posterior_latent_marg": gamma.sum(axis=1), "posterior_dynamics_marg": gamma.sum(axis=2),
            "m_step_n_iter": n_it, "m_step_final_loss": losses, "alpha": es["alpha"], "ll": es["ll"]}
