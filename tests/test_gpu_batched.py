"""Batched callers: model_selection_helper.get_downsampled_lml / get_lml_test_history and the
shuffle tests of test.py against plain loops over the single-call API and against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ref_numpy as ref
from poor_man_gplvm_b200.synthetic import make_dataset

pytestmark = pytest.mark.gpu


def _fitted(N=24, K=64, T=900, seed=31):
    import poor_man_gplvm_b200 as pmg
    d = make_dataset(T, N, K, seed=seed)
    m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0)
    em = m.fit_em(d["y"], key=2, n_iter=4, m_step_maxiter=25, m_step_tol=-1, save_every=2)
    return d, m, em


def test_downsampled_lml_matches_single_calls_and_oracle():
    from poor_man_gplvm_b200 import model_selection_helper as msh
    d, m, em = _fitted()
    masks = msh.draw_latent_masks(m.n_latent_bin, 0.25, 5, key=4)
    assert masks.shape == (5, 64) and np.all(masks.sum(axis=1) == 16)
    got = msh.get_downsampled_lml(m, d["y"], downsample_frac=0.25, n_repeat=5, key=4)
    want = [m.decode_latent(d["y"], ma_latent=mask)["log_marginal_final"] for mask in masks]
    assert np.allclose(got["lml_l"], want, rtol=2e-6)
    assert np.isclose(got["value"], np.mean(want), rtol=2e-6) and np.isclose(got["std"], np.std(want), rtol=1e-3)
    o = ref.OraclePoissonGPLVMJump1D(24, 64, tuning_lengthscale=8.0, dtype=np.float64, tuning_basis=m.tuning_basis,
                                     params=m.params)
    o_lml = o.decode_latent(d["y"], tuning=m.tuning.astype(np.float64), ma_latent=masks[0])["log_marginal_final"]
    assert abs(got["lml_l"][0] - o_lml) < 1e-4 * abs(o_lml)
    # explicit masks + a neuron mask + likelihood_scale go through the same session
    ma_n = np.ones(24, np.float32); ma_n[5] = 0
    got2 = msh.get_downsampled_lml(m, d["y"], latent_masks=masks[:2], ma_neuron=ma_n, likelihood_scale=0.7)
    want2 = [m.decode_latent(d["y"], ma_latent=mk, ma_neuron=ma_n, likelihood_scale=0.7)["log_marginal_final"]
             for mk in masks[:2]]
    assert np.allclose(got2["lml_l"], want2, rtol=2e-6)


def test_lml_test_history_matches_single_calls():
    from poor_man_gplvm_b200 import model_selection_helper as msh
    d, m, em = _fitted()
    y_test = make_dataset(500, 24, 64, seed=32)["y"]
    tunings = em["tuning_saved"]
    assert len(tunings) == 2
    nb = msh.get_lml_test_history(y_test, m, tunings, do_nb=True)
    want_nb = [m.decode_latent_naive_bayes(y_test, tuning=t)["log_marginal_total"] for t in tunings]
    assert np.allclose(nb, want_nb, rtol=2e-6)
    ma_t = (np.arange(500) % 3 != 0).astype(np.float32)
    dyn = msh.get_lml_test_history(y_test, m, tunings, do_nb=False, ma_temporal=ma_t)
    ma_tn = np.ones((1, 24), np.float32) * ma_t[:, None]
    want_dyn = [m.decode_latent(y_test, tuning=t, ma_neuron=ma_tn)["log_marginal_final"] for t in tunings]
    assert np.allclose(dyn, want_dyn, rtol=2e-6)


def test_shuffle_and_decode_matches_manual_shuffles():
    from poor_man_gplvm_b200 import test as shuf
    d, m, em = _fitted(T=400)
    y = d["y"]
    shifts = shuf.draw_shifts(400, 24, 3, seed=7)
    res = shuf.shuffle_and_decode(m, y, n_shuffle=3, decoder_type='naive_bayes', shifts=shifts)
    assert res["log_marginal_l"].shape == (3, 400) and res["posterior_latent"].shape == (3, 400, 64)
    for i in range(3):
        ys = np.stack([np.roll(y[:, j], int(shifts[i, j])) for j in range(24)], axis=1)
        one = m.decode_latent_naive_bayes(ys)
        assert np.array_equal(res["log_marginal_l"][i], one["log_marginal_l"])
        assert np.isclose(res["log_marginal_total"][i], one["log_marginal_total"])
    dyn = shuf.shuffle_and_decode(m, y, n_shuffle=2, decoder_type='dynamics', shifts=shifts,
                                  keys=("log_one_step_predictive_marginals_all", "log_marginal_final"))
    assert set(dyn) == {"log_one_step_predictive_marginals_all", "log_marginal_final"}
    ys = np.stack([np.roll(y[:, j], int(shifts[1, j])) for j in range(24)], axis=1)
    assert np.isclose(dyn["log_marginal_final"][1], m.decode_latent(ys)["log_marginal_final"], rtol=2e-6)
    with pytest.raises(ValueError):
        shuf.shuffle_and_decode(m, y, n_shuffle=1, decoder_type='other')
    out = shuf.test_one_model(y, m, n_shuffle=8, decoder_type='naive_bayes', seed=3)
    assert out["is_sig_tsd"].shape == (400,) and out["log_marg_thresh"].shape == (400,)
    assert out["is_sig_tsd"].dtype == bool and out["is_sig_tsd"].any()
    ent = shuf.compute_entropy(np.asarray(out["decode_res_true"]["log_posterior_latent"]), axis=-1)
    assert ent.shape == (400,) and np.all(ent >= 0) and np.all(ent <= np.log(64) + 1e-5)
