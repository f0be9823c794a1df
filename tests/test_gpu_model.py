"""End-to-end parity of the drop-in model class against the oracle's EM driver.

Both sides get identical spikes, ``tuning_basis``, ``params`` and
``log_posterior_init`` and a pinned Adam step count
(``m_step_tol=-1``).  Tolerances are BASELINE.json's: log_marginal_l
1e-4 relative per iteration, posterior marginals 1e-5 absolute (vs the fp64
oracle; the fp32 log-space restatement is itself only ~1e-5 from fp64, SURVEY
H5), tuning 1e-3 relative.
"""
import pickle

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ref_numpy as ref
from poor_man_gplvm_b200.synthetic import make_dataset

pytestmark = pytest.mark.gpu

EM_KEYS = {'log_posterior_all_saved', 'log_posterior_init', 'params_saved', 'tuning_saved', 'iter_saved', 'params',
           'tuning', 'log_posterior_final', 'log_marginal', 'log_marginal_l', 'log_marginal_saved', 'posterior',
           'posterior_latent_marg', 'posterior_dynamics_marg', 'm_step_res_l'}
DECODE_KEYS = {'log_posterior_all', 'log_marginal_final', 'posterior_all', 'posterior_latent_marg',
               'posterior_dynamics_marg', 'log_one_step_predictive_marginals_all', 'log_likelihood_all',
               'p_joint_full', 'p_joint_latent', 'p_joint_dynamics', 'p_transition_full', 'p_transition_latent',
               'p_transition_dynamics', 'log_joint_full', 'log_joint_latent', 'log_joint_dynamics',
               'log_transition_full', 'log_transition_latent', 'log_transition_dynamics'}
NB_KEYS = {'log_posterior_latent', 'log_marginal_l', 'log_marginal_total', 'posterior_latent', 'll_per_pos_l'}


def _pair(N, K, T, ls, seed):
    import poor_man_gplvm_b200 as pmg
    d = make_dataset(T, N, K, seed=seed)
    model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=ls, movement_variance=1.0)
    rng = np.random.default_rng(seed + 1)
    params = rng.standard_normal((model.n_basis, N)).astype(np.float32)
    model.params = params.copy()
    model.tuning = np.logaddexp(model.tuning_basis @ params, np.float32(0)).astype(np.float32)
    oracle = ref.OraclePoissonGPLVMJump1D(N, K, tuning_lengthscale=ls, movement_variance=1.0, dtype=np.float64,
                                          tuning_basis=model.tuning_basis, params=params)
    lp0, _ = model.init_latent_posterior(T, key=7)
    return d, model, oracle, lp0


def test_fit_em_readme_config_matches_oracle():
    """configs[0]: N=30, K=100, T=1000, n_iter=20 (README example), Adam pinned at 50 steps."""
    d, model, oracle, lp0 = _pair(30, 100, 1000, 10.0, seed=0)
    kw = dict(n_iter=20, log_posterior_init=lp0, m_step_maxiter=50, m_step_tol=-1)
    want = oracle.fit_em(d["y"], **kw)
    got = model.fit_em(d["y"], **kw)
    assert set(got) == EM_KEYS
    lw, lg = np.array(want["log_marginal_l"]), np.array(got["log_marginal_l"])
    assert np.max(np.abs(lg - lw) / np.abs(lw)) < 1e-4
    assert np.max(np.abs(got["tuning"] - want["tuning"]) / want["tuning"]) < 1e-3
    assert np.max(np.abs(got["posterior_latent_marg"] - want["posterior_latent_marg"])) < 5e-5
    assert np.max(np.abs(got["posterior_dynamics_marg"] - want["posterior_dynamics_marg"])) < 5e-5
    assert got["posterior"].shape == (1000, 2, 100) and got["log_posterior_final"].shape == (1000, 2, 100)
    assert np.allclose(got["posterior"].sum(axis=(1, 2)), 1.0, atol=1e-5)
    assert got["m_step_res_l"]["n_iter"] == [50] * 20
    assert len(got["m_step_res_l"]["loss_history"][0]) == 50
    assert got["iter_saved"] == [0] and len(got["log_posterior_all_saved"]) == 1
    # attributes the reference mutates (core.py:679-686)
    assert model.tuning.shape == (100, 30) and model.params.shape == (model.n_basis, 30)
    assert model.log_latent_transition_kernel_l.shape == (2, 100, 100)


def test_one_em_iteration_teacher_forced_posterior_tolerance():
    """One M+E step from identical state meets the 1e-5 absolute posterior tolerance."""
    d, model, oracle, lp0 = _pair(40, 100, 800, 10.0, seed=3)
    kw = dict(n_iter=1, log_posterior_init=lp0, m_step_maxiter=30, m_step_tol=-1)
    want = oracle.fit_em(d["y"], **kw)
    got = model.fit_em(d["y"], **kw)
    assert np.max(np.abs(got["posterior_latent_marg"] - want["posterior_latent_marg"])) < 1e-5
    assert np.max(np.abs(got["posterior_dynamics_marg"] - want["posterior_dynamics_marg"])) < 1e-5
    assert abs(got["log_marginal"] - want["log_marginal"]) < 1e-5 * abs(want["log_marginal"])


def test_decode_latent_keys_and_values():
    d, model, oracle, _ = _pair(25, 64, 500, 8.0, seed=5)
    ma_l = np.ones(64, np.float32); ma_l[10:14] = 0
    ma_n = np.ones(25, np.float32); ma_n[3] = 0
    want = oracle.decode_latent(d["y"], ma_neuron=ma_n, ma_latent=ma_l, likelihood_scale=1.5)
    got = model.decode_latent(d["y"], ma_neuron=ma_n, ma_latent=ma_l, likelihood_scale=1.5)
    assert set(got) == DECODE_KEYS
    assert abs(got["log_marginal_final"] - want["log_marginal_final"]) < 1e-4 * abs(want["log_marginal_final"])
    assert np.max(np.abs(got["posterior_all"] - want["posterior_all"])) < 1e-5
    assert np.max(np.abs(got["posterior_latent_marg"] - want["posterior_latent_marg"])) < 1e-5
    assert np.max(np.abs(got["posterior_dynamics_marg"] - want["posterior_dynamics_marg"])) < 1e-5
    assert np.max(np.abs(got["log_one_step_predictive_marginals_all"]
                         - want["log_one_step_predictive_marginals_all"])) < 1e-3
    live = ma_l.astype(bool)
    assert np.all(got["log_likelihood_all"][:, ~live] == np.float32(-1e20))
    assert np.max(np.abs(got["log_likelihood_all"][:, live] - want["log_likelihood_all"][:, live])) < 2e-3
    for k in ("p_joint_full", "p_joint_latent", "p_joint_dynamics", "p_transition_dynamics"):
        assert np.max(np.abs(got[k] - want[k])) < 2e-5, k
    # rows of masked latent bins carry no mass; the reference normalises rounding noise there
    assert np.max(np.abs(got["p_transition_latent"][live] - want["p_transition_latent"][live])) < 2e-5
    assert np.all(np.isfinite(got["p_transition_latent"])) and np.all(np.isfinite(got["log_transition_full"]))
    # log outputs agree wherever the posterior is not in the deep tail
    big = want["posterior_all"] > 1e-6
    assert np.max(np.abs(got["log_posterior_all"][big] - want["log_posterior_all"][big])) < 1e-3


def test_decode_naive_bayes_keys_values_and_argmax():
    d, model, oracle, _ = _pair(60, 120, 4000, 8.0, seed=6)
    want = oracle.decode_latent_naive_bayes(d["y"])
    got = model.decode_latent_naive_bayes(d["y"])
    assert set(got) == NB_KEYS
    assert np.max(np.abs(got["ll_per_pos_l"] - want["ll_per_pos_l"]) / np.maximum(1, np.abs(want["ll_per_pos_l"]))) < 2e-6
    assert abs(got["log_marginal_total"] - want["log_marginal_total"]) < 1e-5 * abs(want["log_marginal_total"])
    assert np.max(np.abs(got["posterior_latent"] - want["posterior_latent"])) < 1e-4
    am_g, am_w = got["log_posterior_latent"].argmax(axis=1), want["log_posterior_latent"].argmax(axis=1)
    mism = np.nonzero(am_g != am_w)[0]
    # bit-exact argmax except at fp32-unresolvable ties: any mismatch must have a top-2 gap below 4 ulp(|ll|)
    for t in mism:
        row = want["ll_per_pos_l"][t]
        gap = abs(row[am_g[t]] - row[am_w[t]])
        assert gap < 4 * np.spacing(np.float32(abs(row).max())), (t, gap)
    assert mism.size <= 2


def test_naive_bayes_equals_smoother_under_uniform_transitions():
    import poor_man_gplvm_b200 as pmg
    N, K, T = 12, 32, 300
    d = make_dataset(T, N, K, seed=8)
    unif = np.full((K, K), 1.0 / K, dtype=np.float32)
    model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=6.0, custom_transition_kernel=unif,
                                   p_move_to_jump=0.5, p_jump_to_move=0.5)
    model.tuning = d["tuning_true"]
    nb = model.decode_latent_naive_bayes(d["y"])
    sm = model.decode_latent(d["y"])
    assert np.max(np.abs(sm["posterior_latent_marg"] - nb["posterior_latent"])) < 1e-5


def test_hyperparam_override_save_every_and_pickle():
    d, model, oracle, lp0 = _pair(20, 50, 300, 6.0, seed=9)
    hp = {"movement_variance": 2.5, "p_move_to_jump": 0.05, "param_prior_std": 2.0}
    kw = dict(hyperparam=hp, n_iter=4, log_posterior_init=lp0, m_step_maxiter=20, m_step_tol=-1, save_every=2)
    want = oracle.fit_em(d["y"], **kw)
    got = model.fit_em(d["y"], **kw)
    assert got["iter_saved"] == [0, 2] and len(got["tuning_saved"]) == 2
    assert model.movement_variance == 2.5 and model.p_move_to_jump == 0.05
    lw, lg = np.array(want["log_marginal_l"]), np.array(got["log_marginal_l"])
    assert np.max(np.abs(lg - lw) / np.abs(lw)) < 1e-4
    assert np.max(np.abs(np.exp(got["log_posterior_all_saved"][1]) - np.exp(want["log_posterior_all_saved"][1]))) < 5e-5
    # every saved snapshot is self-consistent and matches the oracle's (the M-step of the NEXT iteration is
    # enqueued ahead of time: the saved weights must be the ones that produced the saved tuning)
    for j in range(2):
        z = model.tuning_basis.astype(np.float64) @ got["params_saved"][j].astype(np.float64)
        sp = np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))
        assert np.max(np.abs(sp - got["tuning_saved"][j]) / sp) < 1e-5
        assert np.max(np.abs(got["params_saved"][j] - want["params_saved"][j])) < 3e-3
        assert np.max(np.abs(got["tuning_saved"][j] - want["tuning_saved"][j]) / want["tuning_saved"][j]) < 3e-3
    assert np.max(np.abs(got["params"] - want["params"])) < 3e-3
    m2 = pickle.loads(pickle.dumps(model))
    assert np.array_equal(m2.tuning, model.tuning) and m2.adam_runner is None


def test_device_posterior_init_matches_host_threefry():
    """pmg_threefry_posterior_init == jaxprng (jax.random's bit stream) incl. odd sizes, row offsets, pieces."""
    from poor_man_gplvm_b200 import jaxprng as jr, ops
    dev = torch.device("cuda")
    for T, K, key in [(7, 5, 0), (33, 100, 123), (64, 401, 9)]:
        u = jr.uniform(jr.PRNGKey(key), (T, K)) * np.float32(0.1)
        want = u / u.sum(axis=1, keepdims=True)
        post, logp, tw = ops.threefry_posterior_init(T, K, key, 0.1, dev, want_post=True, want_log=True, want_tw=True)
        assert np.allclose(post.cpu().numpy(), want, rtol=3e-7, atol=0)
        assert np.allclose(np.exp(logp.cpu().numpy()), want, rtol=2e-6)
        assert np.allclose(tw.cpu().numpy(), want.sum(axis=0), rtol=1e-6)
        # rows 10..19 of the same global draw, written as fp16 pieces into rows 3.. of a larger buffer
        if T > 20:
            g16 = ops.new_gamma16(16, K, dev)
            ops.threefry_posterior_init(10, K, key, 0.1, dev, t_offset=10, T_total=T, g16=g16, g16_row0=3)
            rec = (g16[0].float() + g16[1].float()).cpu().numpy()[:, :K]
            assert np.allclose(rec[3:13], want[10:20], rtol=0, atol=1e-7)
            assert np.all(rec[:3] == 0) and np.all(rec[13:] == 0)


def test_fit_em_default_key_draws_reference_posterior_on_device():
    """fit_em(y) without log_posterior_init: the posterior is drawn on the device from `key` with jax's stream;
    the result equals a run that is handed the same array explicitly, and em_res exposes it lazily."""
    import poor_man_gplvm_b200 as pmg
    N, K, T = 20, 64, 700
    d = make_dataset(T, N, K, seed=4)
    m1 = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0)
    m2 = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0)
    assert np.array_equal(m1.params, m2.params)                 # initialize_params is deterministic in rng_init_int
    r1 = m1.fit_em(d["y"], key=11, n_iter=3, m_step_maxiter=30, m_step_tol=-1)
    lp0, _ = m2.init_latent_posterior(T, 11)
    r2 = m2.fit_em(d["y"], n_iter=3, log_posterior_init=lp0, m_step_maxiter=30, m_step_tol=-1)
    assert np.allclose(np.asarray(r1["log_posterior_init"]), lp0, rtol=0, atol=2e-6)
    assert np.allclose(r1["log_marginal_l"], r2["log_marginal_l"], rtol=2e-6)
    assert np.max(np.abs(r1["posterior_latent_marg"] - r2["posterior_latent_marg"])) < 2e-5


def test_results_pickle_and_reject_bad_dt():
    """em_res / decoding_res round-trip through pickle as host arrays; per-bin dt must be positive."""
    d, model, oracle, lp0 = _pair(12, 32, 200, 6.0, seed=21)
    em = model.fit_em(d["y"], n_iter=2, m_step_maxiter=10, m_step_tol=-1)      # default key: lazy initial posterior
    dec = model.decode_latent(d["y"])
    em2, dec2 = pickle.loads(pickle.dumps(em)), pickle.loads(pickle.dumps(dec))
    assert set(em2) == EM_KEYS and set(dec2) == DECODE_KEYS
    assert np.array_equal(np.asarray(em["log_posterior_final"]), em2["log_posterior_final"])
    assert np.array_equal(np.asarray(em["log_posterior_init"]), em2["log_posterior_init"])
    assert np.array_equal(np.asarray(dec["p_joint_full"]), dec2["p_joint_full"])
    dt = np.ones(200, np.float32); dt[7] = 0.0
    with pytest.raises(ValueError):
        model.decode_latent_naive_bayes(d["y"], dt_l=dt)


def test_model_on_a_device_that_is_not_current():
    """device='cuda:1' while cuda:0 is current: every launch must land on the model's device (needs 2 GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import poor_man_gplvm_b200 as pmg
    N, K, T = 20, 64, 500
    d = make_dataset(T, N, K, seed=22)
    torch.cuda.set_device(0)
    m0 = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0, device="cuda:0")
    m1 = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0, device="cuda:1")
    kw = dict(n_iter=3, key=3, m_step_maxiter=15, m_step_tol=-1)
    r0, r1 = m0.fit_em(d["y"], **kw), m1.fit_em(d["y"], **kw)
    assert torch.cuda.current_device() == 0
    assert np.array_equal(r0["tuning"], r1["tuning"]) and np.array_equal(r0["posterior"], r1["posterior"])
    assert np.array_equal(np.asarray(m0.decode_latent(d["y"])["posterior_all"]),
                          np.asarray(m1.decode_latent(d["y"])["posterior_all"]))
