"""Pins the NumPy oracle (oracle/ref_numpy.py) against golden vectors produced by the REFERENCE's own source
files (tests/golden/make_golden.py: /root/reference/poor_man_gplvm/*.py executed unmodified, with
oracle/jaxshim supplying the jax / optax names on torch CPU).  fp64 runs must agree to rounding; fp32 runs
to fp32 rounding accumulated over the EM iterations.  CPU only."""
import ast
import glob
import os

import numpy as np
import pytest

from oracle import ref_numpy as ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = ["masked_chunked", "odd_wide", "mask_tn_dt", "t1", "t2"]


def load(name, prec):
    g = np.load(os.path.join(GOLD, "%s_%s.npz" % (name, prec)))
    return g, ast.literal_eval(str(g["meta_case"]))


def make_oracle(g, c, dtype):
    mk = {}
    if "in_custom_transition_kernel" in g:
        mk["custom_transition_kernel"] = g["in_custom_transition_kernel"]
    return ref.OraclePoissonGPLVMJump1D(c["N"], c["K"], tuning_lengthscale=c["ls"],
                                        movement_variance=c.get("mv", 1.0), p_move_to_jump=c.get("pmj", 0.01),
                                        p_jump_to_move=c.get("pjm", 0.01), dtype=dtype,
                                        tuning_basis=g["tuning_basis"], params=g["in_params"], **mk)


def em_kwargs(g, c):
    kw = dict(n_iter=c["n_iter"], log_posterior_init=g["in_log_posterior_init"], ma_neuron=g["in_ma_neuron"],
              ma_latent=g["in_ma_latent"], n_time_per_chunk=c.get("n_time_per_chunk", 10000),
              likelihood_scale=c.get("likelihood_scale", 1.0))
    for k in ("m_step_maxiter", "m_step_tol"):
        if k in c:
            kw[k] = c[k]
    return kw


def test_fixtures_present():
    names = {os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*.npz"))}
    for n in SMALL + ["readme_pinned", "readme_default"]:
        assert "%s_f64.npz" % n in names and "%s_f32.npz" % n in names


@pytest.mark.parametrize("name", SMALL)
def test_oracle_fp64_matches_reference_source(name):
    g, c = load(name, "f64")
    o = make_oracle(g, c, np.float64)
    # transition matrices (gp_kernel.py:42-89)
    P, logP, M, logM = o._transitions({})
    assert np.allclose(P, g["tr_P"], rtol=1e-12, atol=1e-300)
    assert np.allclose(logP, g["tr_logP"], rtol=1e-12, atol=1e-12)
    assert np.allclose(M, g["tr_M"], rtol=1e-12) and np.allclose(logM, g["tr_logM"], rtol=1e-12)
    # tuning from the injected params (fit_tuning_helper.py:11-25)
    assert np.allclose(o.tuning, g["tuning_init"], rtol=1e-12)
    em = o.fit_em(g["in_y"], **em_kwargs(g, c))
    assert np.allclose(np.array(em["log_marginal_l"]), g["em_log_marginal_l"], rtol=1e-10)
    assert np.allclose(em["params"], g["em_params"], rtol=1e-8, atol=1e-10)
    assert np.allclose(em["tuning"], g["em_tuning"], rtol=1e-8)
    assert np.allclose(em["posterior"], g["em_posterior"], rtol=0, atol=1e-10)
    assert np.allclose(em["posterior_dynamics_marg"], g["em_posterior_dynamics_marg"], rtol=0, atol=1e-10)
    assert [int(n) for n in em["m_step_res_l"]["n_iter"]] == list(g["em_m_n_iter"])
    assert np.allclose(np.array(em["m_step_res_l"]["final_loss"], dtype=float), g["em_m_final_loss"], rtol=1e-10)
    assert np.allclose(np.array(em["m_step_res_l"]["final_error"], dtype=float), g["em_m_final_error"], rtol=1e-7)
    assert np.allclose(em["m_step_res_l"]["loss_history"][0], g["em_m_loss_history0"][:len(em["m_step_res_l"]["loss_history"][0])],
                       rtol=1e-10)
    # decode_latent with the fitted tuning (core.py:454-497) and the 12 transition summaries (decoder.py:334-375)
    kw = em_kwargs(g, c)
    dec = o.decode_latent(g["in_y"], ma_neuron=kw["ma_neuron"], ma_latent=kw["ma_latent"],
                          likelihood_scale=kw["likelihood_scale"], n_time_per_chunk=kw["n_time_per_chunk"])
    assert np.isclose(dec["log_marginal_final"], float(g["dec_log_marginal_final"]), rtol=1e-10)
    for k in ("posterior_all", "posterior_latent_marg", "posterior_dynamics_marg"):
        assert np.allclose(dec[k], g["dec_" + k], rtol=0, atol=1e-10), k
    assert np.allclose(dec["log_one_step_predictive_marginals_all"], g["dec_log_one_step_predictive_marginals_all"],
                       rtol=1e-9, atol=1e-9)
    assert np.allclose(dec["log_likelihood_all"], g["dec_log_likelihood_all"], rtol=1e-10)
    assert np.allclose(dec["log_causal_posterior_all"], g["dec_log_causal_posterior_all"], rtol=1e-8, atol=1e-8)
    if c["T"] > 1:
        for k in ("p_joint_full", "p_joint_latent", "p_joint_dynamics", "p_transition_dynamics"):
            assert np.allclose(dec[k], g["dec_" + k], rtol=0, atol=1e-10), k
        live = g["in_ma_latent"].astype(bool)
        assert np.allclose(dec["p_transition_latent"][live], g["dec_p_transition_latent"][live], rtol=0, atol=1e-9)
        fin = np.isfinite(g["dec_log_accumulated_joint_total"])
        assert np.allclose(dec["log_accumulated_joint_total"][fin], g["dec_log_accumulated_joint_total"][fin],
                           rtol=1e-8, atol=1e-8)
    # naive Bayes (core.py:499-524), per-bin dt where the case has it (decoder.py:73-85)
    nb_kw = dict(ma_neuron=kw["ma_neuron"], ma_latent=kw["ma_latent"], n_time_per_chunk=kw["n_time_per_chunk"])
    if "in_dt_l" in g.files:
        nb_kw["dt_l"] = g["in_dt_l"]
    o.tuning = np.asarray(g["em_tuning"], dtype=np.float64)
    nb = o.decode_latent_naive_bayes(g["in_y"], **nb_kw)
    assert np.allclose(nb["ll_per_pos_l"], g["nb_ll_per_pos_l"], rtol=1e-9)
    assert np.allclose(nb["log_marginal_l"], g["nb_log_marginal_l"], rtol=1e-9)
    assert np.isclose(nb["log_marginal_total"], float(g["nb_log_marginal_total"]), rtol=1e-10)
    assert np.array_equal(np.argmax(nb["log_posterior_latent"], axis=1), g["nb_argmax"])


@pytest.mark.parametrize("name", ["masked_chunked", "odd_wide", "t2"])
def test_oracle_fp32_matches_reference_source_fp32(name):
    """Same comparison with both sides in fp32 (the reference's dtype): agreement to accumulated rounding."""
    g, c = load(name, "f32")
    o = make_oracle(g, c, np.float32)
    em = o.fit_em(g["in_y"], **em_kwargs(g, c))
    assert np.allclose(np.array(em["log_marginal_l"], dtype=np.float64), g["em_log_marginal_l"], rtol=2e-5)
    assert np.max(np.abs(em["tuning"] - g["em_tuning"]) / g["em_tuning"]) < 2e-3
    assert np.max(np.abs(em["posterior"] - g["em_posterior"])) < 2e-3


def test_oracle_readme_config_matches_reference_source():
    """BASELINE.json configs[0] (N=30, K=100, T=1000, n_iter=20), Adam pinned at 50 steps: fp64 oracle vs the
    reference source in fp64 (T-sized fixtures are stored in float32)."""
    g, c = load("readme_pinned", "f64")
    o = make_oracle(g, c, np.float64)
    kw = em_kwargs(g, c)
    kw["n_iter"] = 3            # the CPU suite stays short; the GPU suite checks all 20 iterations
    em = o.fit_em(g["in_y"].astype(np.float64), **kw)
    assert np.allclose(np.array(em["log_marginal_l"]), g["em_log_marginal_l"][:3], rtol=1e-10)
    assert [int(n) for n in em["m_step_res_l"]["n_iter"]] == [50, 50, 50]


def test_linear_em_driver_matches_reference_source():
    """oracle/linear_ref.fit_em_linear (restated M-step + linear-space E-step; the oracle of the real-shape GPU
    parity tests, tests/test_gpu_shapes.py) against the reference source's README run in fp64."""
    from oracle import linear_ref as lin
    g, c = load("readme_pinned", "f64")
    o = make_oracle(g, c, np.float64)
    kw = em_kwargs(g, c)
    kw.pop("n_time_per_chunk")
    em = lin.fit_em_linear(o, g["in_y"].astype(np.float64), **kw)
    assert np.allclose(np.array(em["log_marginal_l"]), g["em_log_marginal_l"], rtol=1e-9)
    assert np.allclose(em["tuning"], g["em_tuning"], rtol=1e-6)
    assert np.allclose(em["posterior"], g["em_posterior"], rtol=0, atol=1e-7)      # fixture stored in float32
    assert em["m_step_n_iter"] == [int(v) for v in g["em_m_n_iter"]]


def test_linear_em_driver_step_schedule_reproduces_the_free_run():
    """``m_step_schedule`` pins each M-step's Adam step count (what the real-shape GPU tests use to compare fits
    whose default stopping rule would otherwise be decided by rounding): fed the counts of the reference source's
    default-Adam README run, the driver must land on that run's results."""
    from oracle import linear_ref as lin
    g, c = load("readme_default", "f64")
    o = make_oracle(g, c, np.float64)
    kw = em_kwargs(g, c)
    kw.pop("n_time_per_chunk")
    for k in ("m_step_maxiter", "m_step_tol"):
        kw.pop(k, None)
    n = 4
    kw["n_iter"] = n
    em = lin.fit_em_linear(o, g["in_y"].astype(np.float64), m_step_schedule=[int(v) for v in g["em_m_n_iter"][:n]], **kw)
    assert em["m_step_n_iter"] == [int(v) for v in g["em_m_n_iter"][:n]]
    # (hundreds of Adam steps per M-step amplify the last-digit differences between the linear-space E-step and the
    # reference's log-space one: 3e-9 here, 1e-9 on the pinned 50-step run above)
    assert np.allclose(np.array(em["log_marginal_l"]), g["em_log_marginal_l"][:n], rtol=1e-7)


def test_linear_em_driver_with_a_dense_custom_kernel_matches_reference_source():
    """The oracle of the lockstep-scan GPU tests (oracle/linear_ref with a dense custom transition kernel, reference
    gp_kernel.py:61-66) against the reference source's own run of that model (fixture `dense_custom`, K = 256)."""
    from oracle import linear_ref as lin
    g, c = load("dense_custom", "f64")
    o = make_oracle(g, c, np.float64)
    kw = em_kwargs(g, c)
    kw.pop("n_time_per_chunk")
    em = lin.fit_em_linear(o, g["in_y"].astype(np.float64), **kw)
    assert np.allclose(np.array(em["log_marginal_l"]), g["em_log_marginal_l"], rtol=1e-9)
    assert np.allclose(em["tuning"], g["em_tuning"], rtol=1e-6)
    assert np.allclose(em["posterior"], g["em_posterior"], rtol=0, atol=1e-7)          # fixture stored in float32
    assert em["m_step_n_iter"] == [int(v) for v in g["em_m_n_iter"]]


def test_linear_em_driver_at_the_headline_shape_matches_reference_source():
    """oracle/linear_ref at N=500, K=400 (fixture `headline_shape`: the reference source's own fp64 run, T=640, two EM
    iterations): the oracle the real-shape GPU tests trust is pinned at that shape too."""
    from oracle import linear_ref as lin
    g, c = load("headline_shape", "f64")
    o = make_oracle(g, c, np.float64)
    kw = em_kwargs(g, c)
    kw.pop("n_time_per_chunk")
    em = lin.fit_em_linear(o, g["in_y"].astype(np.float64), **kw)
    assert np.allclose(np.array(em["log_marginal_l"]), g["em_log_marginal_l"], rtol=1e-9)
    assert np.allclose(em["tuning"], g["em_tuning"], rtol=2e-6)                        # fixture stored in float32
    assert np.allclose(em["posterior"], g["em_posterior"], rtol=0, atol=1e-6)
    assert em["m_step_n_iter"] == [20, 20]
