"""The reference's other model families against golden vectors produced by the reference's own
source (tests/golden/make_golden_families.py: core.py:852-1093 + decoder_latentonly.py on oracle/jaxshim, fp64)."""
import ast
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    g = np.load(os.path.join(GOLD, "%s_f64.npz" % name))
    return g, ast.literal_eval(str(g["meta_case"]))


def make_model(g, c):
    from poor_man_gplvm_b200 import families
    kw = dict(n_latent_bin=c["K"], tuning_lengthscale=c["ls"], movement_variance=c["mv"])
    if "noise_std" in c:
        kw["noise_std"] = c["noise_std"]
    if "pmj" in c:
        kw.update(p_move_to_jump=c["pmj"], p_jump_to_move=c["pjm"])
    m = getattr(families, c["cls"])(c["N"], **kw)
    m.tuning_basis = g["tuning_basis"].astype(np.float32)
    m.n_basis = m.tuning_basis.shape[1]
    m.params = g["in_params"].astype(np.float32)
    m.tuning = np.asarray(m.get_tuning(m.params, {}, m.tuning_basis))
    return m


@pytest.mark.parametrize("name", ["fam_poisson1d", "fam_gauss_jump", "fam_gauss1d"])
def test_family_fit_decode_match_reference_source(name):
    g, c = load(name)
    m = make_model(g, c)
    kw = dict(n_iter=c["n_iter"], log_posterior_init=g["in_log_posterior_init"], ma_neuron=g["in_ma_neuron"],
              ma_latent=g["in_ma_latent"], likelihood_scale=c.get("likelihood_scale", 1.0))
    for k in ("m_step_maxiter", "m_step_tol"):
        if k in c:
            kw[k] = c[k]
    em = m.fit_em(g["in_y"], **kw)
    assert sorted(em.keys()) == list(g["em_keys"])
    lw, lg = g["em_log_marginal_l"], np.array(em["log_marginal_l"], dtype=np.float64)
    assert np.max(np.abs(lg - lw) / np.abs(lw)) < 1e-4, (lg, lw)
    # Gaussian tuning is linear and crosses zero: compare on the scale of the observations
    scale = np.maximum(np.abs(g["em_tuning"]), 0.05)
    assert np.max(np.abs(em["tuning"] - g["em_tuning"]) / scale) < 2e-3
    assert np.max(np.abs(np.asarray(em["posterior"]) - g["em_posterior"])) < 1e-4
    if "em_posterior_dynamics_marg" in g.files:
        assert np.max(np.abs(em["posterior_dynamics_marg"] - g["em_posterior_dynamics_marg"])) < 1e-4
    # decode with the REFERENCE's fitted tuning: one E-step from identical state, 1e-5 absolute
    dec = m.decode_latent(g["in_y"], tuning=g["em_tuning"].astype(np.float32), ma_neuron=g["in_ma_neuron"],
                          ma_latent=g["in_ma_latent"], likelihood_scale=c.get("likelihood_scale", 1.0))
    assert sorted(dec.keys()) == list(g["dec_keys"])
    ref_lml = float(g["dec_log_marginal_final"])
    assert abs(dec["log_marginal_final"] - ref_lml) < 1e-4 * abs(ref_lml)
    assert np.max(np.abs(np.asarray(dec["posterior_all"]) - g["dec_posterior_all"])) < 1e-5
    live = g["in_ma_latent"].astype(bool)
    ll_ref = g["dec_log_likelihood_all"]
    assert np.max(np.abs(dec["log_likelihood_all"][:, live] - ll_ref[:, live]) / np.maximum(1, np.abs(ll_ref[:, live]))) < 5e-6
    assert np.max(np.abs(np.asarray(dec["log_one_step_predictive_marginals_all"])
                         - g["dec_log_one_step_predictive_marginals_all"])) < 2e-3
    assert np.max(np.abs(np.asarray(dec["p_joint_latent"]) - g["dec_p_joint_latent"])) < 2e-5
    ptl = np.asarray(dec["p_transition_latent"])
    rows = live & (g["dec_p_joint_latent"].sum(axis=1) > 1e-4)        # rows the recording actually visits
    assert np.max(np.abs(ptl[rows] - g["dec_p_transition_latent"][rows])) < 1e-3
    if "dec_posterior_dynamics_marg" in g.files:
        assert np.max(np.abs(dec["posterior_dynamics_marg"] - g["dec_posterior_dynamics_marg"])) < 1e-5
        assert np.max(np.abs(np.asarray(dec["p_joint_full"]) - g["dec_p_joint_full"])) < 2e-5
    # naive Bayes with the reference's tuning
    nb = m.decode_latent_naive_bayes(g["in_y"], tuning=g["em_tuning"].astype(np.float32), ma_neuron=g["in_ma_neuron"],
                                     ma_latent=g["in_ma_latent"])
    assert np.max(np.abs(nb["ll_per_pos_l"][:, live] - g["nb_ll_per_pos_l"][:, live])
                  / np.maximum(1, np.abs(g["nb_ll_per_pos_l"][:, live]))) < 5e-6
    assert abs(nb["log_marginal_total"] - float(g["nb_log_marginal_total"])) < 1e-5 * abs(float(g["nb_log_marginal_total"]))
    am = nb["log_posterior_latent"].argmax(axis=1)
    mism = np.nonzero(am != g["nb_argmax"])[0]
    for t in mism:
        row = g["nb_ll_per_pos_l"][t]
        assert abs(row[am[t]] - row[g["nb_argmax"][t]]) < 4 * np.spacing(np.float32(np.abs(row[live]).max())), t


def test_latent_only_raises_when_the_smooth_prior_cannot_explain_the_data():
    """A latent that teleports: the reference's log-space filter stays finite, the fp32 linear-space one cannot;
    the latent-only classes must say so instead of returning NaNs."""
    from poor_man_gplvm_b200 import families
    from poor_man_gplvm_b200.synthetic import bump_tuning
    rng = np.random.default_rng(0)
    K, N, T = 64, 40, 120
    tuning = bump_tuning(K, N, rng, peak=(3.0, 6.0), width_frac=0.05)
    lat = np.where(np.arange(T) < T // 2, 5, 58)
    y = rng.poisson(tuning[lat]).astype(np.float32)
    m = families.PoissonGPLVM1D(N, n_latent_bin=K, movement_variance=0.5)
    with pytest.raises(RuntimeError):
        m.decode_latent(y, tuning=tuning)
    # the jump model explains the same recording
    import poor_man_gplvm_b200 as pmg
    mj = pmg.PoissonGPLVMJump1D(N, K, movement_variance=0.5)
    dec = mj.decode_latent(y, tuning=tuning)
    assert np.isfinite(dec["log_marginal_final"])
    assert dec["posterior_latent_marg"][T // 2 + 3].argmax() in range(55, 62)
