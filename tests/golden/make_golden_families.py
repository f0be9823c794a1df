#!/usr/bin/env python
"""Golden vectors for the other model families: GaussianGPLVMJump1D, PoissonGPLVM1D,
GaussianGPLVM1D of the REFERENCE's own source (/root/reference/poor_man_gplvm/core.py:852-1093 with
decoder_latentonly.py), executed unmodified on `oracle/jaxshim` in fp64.  Same conventions as make_golden.py:
every random input is generated here with NumPy and stored in the fixture.

    python tests/golden/make_golden_families.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    # class, N, K, T, ls, mv, noise_std, n_iter, Adam pinned
    "fam_poisson1d": dict(cls="PoissonGPLVM1D", N=14, K=40, T=260, ls=6.0, mv=1.5, n_iter=3, m_step_maxiter=20,
                          m_step_tol=-1.0, seed=21, mask_latent=(3,), mask_neuron=(5,)),
    "fam_gauss_jump": dict(cls="GaussianGPLVMJump1D", N=12, K=32, T=220, ls=5.0, mv=1.0, noise_std=0.6, n_iter=3,
                           seed=22, pmj=0.03, pjm=0.1, likelihood_scale=0.8),
    "fam_gauss1d": dict(cls="GaussianGPLVM1D", N=12, K=32, T=220, ls=5.0, mv=2.0, noise_std=0.6, n_iter=3, seed=23),
}


def make_inputs(c):
    from poor_man_gplvm_b200.synthetic import bump_tuning, walk_latent
    rng = np.random.default_rng(2000 + c["seed"])
    tuning = bump_tuning(c["K"], c["N"], rng)
    jump = c["cls"].endswith("Jump1D")
    lat, _ = walk_latent(c["T"], c["K"], rng, p_jump=0.02 if jump else 0.0)
    if c["cls"].startswith("Gaussian"):
        y = (tuning[lat] + c["noise_std"] * rng.standard_normal((c["T"], c["N"]))).astype(np.float32)
    else:
        y = rng.poisson(tuning[lat]).astype(np.float32)
    inp = {"y": y, "ma_neuron": np.ones(c["N"], np.float32), "ma_latent": np.ones(c["K"], np.float32)}
    for i in c.get("mask_neuron", ()):
        inp["ma_neuron"][i] = 0
    for i in c.get("mask_latent", ()):
        inp["ma_latent"][i] = 0
    post = (1.0 / c["K"] + rng.random((c["T"], c["K"])) * 0.1).astype(np.float32)
    inp["log_posterior_init"] = np.log(post / post.sum(axis=1, keepdims=True))
    return inp, rng


def run_case(name, c, ref):
    import jax.numpy as jnp
    inp, rng = make_inputs(c)
    kw = dict(n_neuron=c["N"], n_latent_bin=c["K"], tuning_lengthscale=c["ls"], movement_variance=c["mv"])
    if "noise_std" in c:
        kw["noise_std"] = c["noise_std"]
    if "pmj" in c:
        kw.update(p_move_to_jump=c["pmj"], p_jump_to_move=c["pjm"])
    m = getattr(ref.core, c["cls"])(**kw)
    basis = np.asarray(m.tuning_basis)
    params0 = (0.3 * rng.standard_normal((basis.shape[1], c["N"]))).astype(np.float32)
    m.params = jnp.array(params0)
    m.tuning = m.get_tuning(m.params, {}, m.tuning_basis)
    out = {"in_" + k: v for k, v in inp.items()}
    out["in_params"] = params0
    out["tuning_basis"] = basis.astype(np.float64)
    fit_kw = dict(n_iter=c["n_iter"], log_posterior_init=jnp.array(inp["log_posterior_init"]), verboase=False,
                  ma_neuron=jnp.array(inp["ma_neuron"]), ma_latent=jnp.array(inp["ma_latent"]),
                  likelihood_scale=c.get("likelihood_scale", 1.0))
    for k in ("m_step_maxiter", "m_step_tol"):
        if k in c:
            fit_kw[k] = c[k]
    em = m.fit_em(inp["y"].astype(np.float64), **fit_kw)
    A = lambda x: np.asarray(x).astype(np.float64)
    out["em_log_marginal_l"] = np.array([float(v) for v in em["log_marginal_l"]])
    out["em_params"], out["em_tuning"], out["em_posterior"] = A(em["params"]), A(em["tuning"]), A(em["posterior"])
    out["em_keys"] = np.array(sorted(em.keys()))
    if "posterior_dynamics_marg" in em:
        out["em_posterior_dynamics_marg"] = A(em["posterior_dynamics_marg"])
    dec = m.decode_latent(inp["y"].astype(np.float64), ma_neuron=fit_kw["ma_neuron"], ma_latent=fit_kw["ma_latent"],
                          likelihood_scale=fit_kw["likelihood_scale"])
    out["dec_keys"] = np.array(sorted(dec.keys()))
    for k, v in dec.items():
        if k == "log_marginal_final":
            out["dec_log_marginal_final"] = np.array(float(v))
        else:
            out["dec_" + k] = A(v)
    nb = m.decode_latent_naive_bayes(inp["y"].astype(np.float64), ma_neuron=fit_kw["ma_neuron"],
                                     ma_latent=fit_kw["ma_latent"])
    out["nb_ll_per_pos_l"] = A(nb["ll_per_pos_l"])
    out["nb_log_marginal_l"] = A(nb["log_marginal_l"])
    out["nb_log_marginal_total"] = np.array(float(nb["log_marginal_total"]))
    out["nb_argmax"] = np.argmax(np.asarray(nb["log_posterior_latent"]), axis=1).astype(np.int32)
    out["meta_case"] = np.array(repr(c))
    print("  %s: lml %s" % (name, out["em_log_marginal_l"]), flush=True)
    return out


def main():
    os.environ["JAXSHIM_X64"] = "1"
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, "%s_f64.npz" % name), **run_case(name, c, ref))


if __name__ == "__main__":
    main()
