"""Oracle self-consistency (CPU).  The reference pins nothing, so the
restatement is validated by invariants and by agreement between the log-space
restatement (oracle/ref_numpy.py) and the independent linear-space derivation
(oracle/linear_ref.py)."""
import numpy as np
import pytest

from oracle import linear_ref as lin
from oracle import ref_numpy as ref
from poor_man_gplvm_b200.synthetic import make_dataset


@pytest.fixture(scope="module")
def small():
    d = make_dataset(T=240, n_neuron=12, n_latent_bin=24, seed=3)
    P, logP, M, logM = ref.create_transition_prob_1d(24, 1.5, 0.02, 0.03, dtype=np.float64)
    return d, P, logP, M, logM


def test_transition_rows_normalised():
    P, logP, M, logM = ref.create_transition_prob_1d(50, 1.0, 0.01, 0.01, dtype=np.float64)
    assert np.allclose(P.sum(axis=2), 1.0)
    assert np.allclose(M.sum(axis=1), 1.0)
    # analytic log kernel stays finite where the linear one underflows
    assert np.isfinite(logP).all()
    assert np.allclose(np.exp(logP[0, 10, 8:13]), P[0, 10, 8:13])


def test_basis_shape_and_bias():
    B = ref.generate_basis(10.0, 100, dtype=np.float64)
    assert B.shape[0] == 100 and np.all(B[:, 0] == 1.0)
    assert 10 <= B.shape[1] <= 30          # B = 18 at K=100, ls=10
    B1 = ref.generate_basis(1.0, 100, dtype=np.float64)
    assert B1.shape[1] > 90


def test_emission_gemm_form_matches_elementwise(small):
    d, *_ = small
    tun = d["tuning_true"].astype(np.float64)
    ma_n = np.ones(12); ma_n[3] = 0
    ma_l = np.ones(24); ma_l[5] = 0
    a = ref.get_loglikelihood_ma_all(d["y"].astype(np.float64), tun, ma_n, ma_l)
    b = lin.emission_gemm_form(d["y"], tun, ma_n, ma_l)
    assert np.allclose(a, b, rtol=0, atol=1e-10)
    assert np.all(a[:, 5] == -1e20)
    # spatio-temporal mask
    rng = np.random.default_rng(0)
    ma_tn = (rng.random(d["y"].shape) > 0.2).astype(np.float64)
    a = ref.get_loglikelihood_ma_all(d["y"].astype(np.float64), tun, ma_tn, np.ones(24))
    b = lin.emission_gemm_form(d["y"], tun, ma_tn, np.ones(24))
    assert np.allclose(a, b, rtol=0, atol=1e-10)


def test_logspace_vs_linear_estep(small):
    d, P, logP, M, logM = small
    tun = d["tuning_true"].astype(np.float64)
    y = d["y"].astype(np.float64)
    ma_n, ma_l = np.ones(12), np.ones(24)
    out = ref.smooth_all_step_combined_ma_chunk(y, tun, logP, logM, ma_n, ma_l,
                                                likelihood_scale=1.3, n_time_per_chunk=100)
    lp_all, lml, lf_all, lmr, acc, ll = out
    res = lin.e_step(y, tun, P, M, ma_n, ma_l, likelihood_scale=1.3, want_xi=True)
    assert np.allclose(np.exp(lf_all), res["alpha"], atol=1e-12)
    assert np.allclose(np.exp(lp_all), res["gamma"], atol=1e-12)
    assert np.allclose(lmr, res["lmr"], atol=1e-9)
    assert np.isclose(lml, res["log_marginal"], rtol=1e-12)
    assert np.allclose(np.exp(acc), res["xi"], atol=1e-10)
    # invariants
    assert np.allclose(np.exp(lp_all).sum(axis=(1, 2)), 1.0)
    assert np.isclose(lmr.sum(), lml)
    assert np.isclose(np.exp(acc).sum(), y.shape[0] - 1)


def test_chunk_size_is_a_noop(small):
    d, P, logP, M, logM = small
    tun = d["tuning_true"].astype(np.float64)
    y = d["y"].astype(np.float64)
    a = ref.smooth_all_step_combined_ma_chunk(y, tun, logP, logM, np.ones(12), None, n_time_per_chunk=37)
    b = ref.smooth_all_step_combined_ma_chunk(y, tun, logP, logM, np.ones(12), None, n_time_per_chunk=10000)
    for u, v in zip(a, b):
        assert np.allclose(u, v, atol=1e-9)


def test_transition_posterior_postproc(small):
    d, P, logP, M, logM = small
    tun = d["tuning_true"].astype(np.float64)
    y = d["y"].astype(np.float64)
    out = ref.smooth_all_step_combined_ma_chunk(y, tun, logP, logM, np.ones(12), None)
    res = ref.compute_transition_posterior_prob(out[4])
    assert np.isclose(res["p_joint_full"].sum(), 1.0)
    assert np.allclose(res["p_transition_latent"].sum(axis=1), 1.0)
    assert np.allclose(res["p_transition_dynamics"].sum(axis=1), 1.0)
    assert np.allclose(res["p_transition_full"].sum(axis=(1, 3)), 1.0)
    gam = np.exp(out[0])
    row = gam[:-1].sum(axis=0) / (y.shape[0] - 1)                 # [d, x]
    assert np.allclose(res["p_joint_full"].sum(axis=(1, 3)), row, atol=1e-10)


def test_naive_bayes_equals_smoother_with_uniform_transitions():
    d = make_dataset(T=60, n_neuron=8, n_latent_bin=16, seed=1)
    K = 16
    tun = d["tuning_true"].astype(np.float64)
    y = d["y"].astype(np.float64)
    unif = np.full((K, K), 1.0 / K)
    P, logP, M, logM = ref.create_transition_prob_1d(K, 1.0, 0.5, 0.5, custom_kernel=unif, dtype=np.float64)
    out = ref.smooth_all_step_combined_ma_chunk(y, tun, logP, logM, np.ones(8), None)
    nb = ref.get_naive_bayes_ma_chunk(y, tun, np.ones(8), np.ones(K))
    assert np.allclose(ref.lse(out[0], axis=1), nb[0], atol=1e-9)


def test_adam_gradient_matches_finite_difference():
    rng = np.random.default_rng(0)
    K, B, N = 20, 6, 5
    basis = ref.generate_basis(4.0, K, dtype=np.float64)[:, :B]
    W = rng.standard_normal((B, N))
    yw = rng.random((K, N)) * 3
    tw = rng.random(K) + 0.5
    loss, g = ref.poisson_m_step_value_and_grad(W, 1.3, basis, yw, tw)
    num = np.zeros_like(W)
    for i in range(B):
        for j in range(N):
            Wp, Wm = W.copy(), W.copy()
            Wp[i, j] += 1e-6; Wm[i, j] -= 1e-6
            num[i, j] = (ref.poisson_m_step_objective(Wp, 1.3, basis, yw, tw)
                         - ref.poisson_m_step_objective(Wm, 1.3, basis, yw, tw)) / 2e-6
    assert np.allclose(g, num, rtol=1e-5, atol=1e-6)


def test_adam_loop_semantics():
    rng = np.random.default_rng(1)
    K, N = 30, 4
    basis = ref.generate_basis(5.0, K, dtype=np.float64)
    B = basis.shape[1]
    W = rng.standard_normal((B, N))
    yw = rng.random((K, N)) * 3
    tw = rng.random(K) + 0.5
    # tol=-1 pins the step count: exactly maxiter-1 updates, n_iter == maxiter
    res = ref.adam_run(W, ref.adam_init(W), 1.0, basis, yw, tw, maxiter=25, tol=-1)
    assert res["n_iter"] == 25 and res["opt_state"]["count"] == 24
    # history[0] and history[1] are both the loss at the initial params (:148,:168-175)
    assert res["loss_history"][0] == res["loss_history"][1]
    assert res["loss_history"][-1] < res["loss_history"][0]
    # final_loss lags params by one update
    assert res["final_loss"] == res["loss_history"][24]
    # default rule stops early on a converged problem and never before 6 evaluations
    res2 = ref.adam_run(W, ref.adam_init(W), 1.0, basis, yw, tw, maxiter=4000, tol=1e-3)
    assert 6 <= res2["n_iter"] < 4000


def test_em_increases_marginal_likelihood_and_recovers_structure():
    d = make_dataset(T=400, n_neuron=20, n_latent_bin=30, seed=5)
    m = ref.OraclePoissonGPLVMJump1D(20, 30, tuning_lengthscale=5.0, movement_variance=1.0, dtype=np.float64)
    res = m.fit_em(d["y"], n_iter=6, m_step_maxiter=200, m_step_tol=-1, seed=0)
    lml = np.array(res["log_marginal_l"])
    assert np.all(np.diff(lml)[1:] > -1e-3 * abs(lml[-1]))       # approximately monotone (GEM)
    assert lml[-1] > lml[0]
    assert res["posterior"].shape == (400, 2, 30)
    assert np.allclose(res["posterior_latent_marg"].sum(axis=1), 1.0)
    assert set(res["m_step_res_l"]) == {"n_iter", "final_loss", "final_error", "loss_history", "error_history"}
