#!/usr/bin/env python
"""Generates tests/golden/*.npz by executing the REFERENCE's own source files
(/root/reference/poor_man_gplvm/{core,decoder,fit_tuning_helper,gp_kernel}.py, unmodified, read-only) with
`oracle/jaxshim` standing in for the jax / optax import names (JAX is not installable in this image; see
oracle/jaxshim/README.md for what that does and does not pin).

    python tests/golden/make_golden.py            # fp64 ("jax_enable_x64") and fp32 runs of every case

Run in the build container only (needs /root/reference); the .npz fixtures it writes are committed and are
what the tests on the GPU box read.  Every random input (spikes, params, log_posterior_init) is generated
here with NumPy and stored in the fixture, so nothing depends on a PRNG implementation.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name: dict(N, K, T, ls, mv, pmj, pjm, n_iter, m_step kwargs, masks / options)
CASES = {
    # BASELINE.json configs[0]: README example, Adam pinned at 50 steps per EM iteration
    "readme_pinned": dict(N=30, K=100, T=1000, ls=10.0, n_iter=20, m_step_maxiter=50, m_step_tol=-1.0, seed=0),
    # same model, the reference's default optimiser settings (maxiter=1000, tol=1e-6): data-dependent stop
    "readme_default": dict(N=30, K=100, T=1000, ls=10.0, n_iter=5, seed=0, em_only=True),
    # masks, likelihood scale, several chunks with a ragged last one, non-default dynamics
    "masked_chunked": dict(N=17, K=48, T=203, ls=6.0, mv=2.0, pmj=0.03, pjm=0.2, n_iter=3, m_step_maxiter=25,
                           m_step_tol=-1.0, n_time_per_chunk=37, likelihood_scale=0.7, mask_neuron=(2, 9),
                           mask_latent=(0, 11, 47), seed=3),
    # K not a multiple of 8, wide movement kernel, class-default tuning_lengthscale=1 (B ~ K)
    "odd_wide": dict(N=9, K=37, T=150, ls=1.0, mv=6.0, n_iter=2, m_step_maxiter=20, m_step_tol=-1.0, seed=5),
    # spatio-temporal neuron mask [T,N] (decoder.py:291-294) and per-bin dt for naive Bayes (decoder.py:73-85)
    "mask_tn_dt": dict(N=11, K=32, T=120, ls=5.0, n_iter=2, m_step_maxiter=20, m_step_tol=-1.0, mask_tn=True,
                       dt_l=True, seed=7),
    # dense custom transition kernel (gp_kernel.py:61-66): K = 256 puts the CUDA path on the lockstep tensor-core scan
    "dense_custom": dict(N=16, K=256, T=400, ls=12.0, pmj=0.02, pjm=0.05, n_iter=2, m_step_maxiter=25,
                         m_step_tol=-1.0, custom_kernel="dense", seed=12),
    # BASELINE.json configs[3] shape (N=500, K=400) on a recording just long enough for the production kernels
    # (cta_group::2 emission units of 512 bins, tensor-core statistics, 125-CTA M-step, compact scans): EM outputs only
    "headline_shape": dict(N=500, K=400, T=640, ls=10.0, n_iter=2, m_step_maxiter=20, m_step_tol=-1.0, seed=21,
                           em_only=True, x64_only=True),
    # shortest recordings
    "t1": dict(N=5, K=16, T=1, ls=4.0, n_iter=1, m_step_maxiter=5, m_step_tol=-1.0, seed=9),
    "t2": dict(N=5, K=16, T=2, ls=4.0, n_iter=1, m_step_maxiter=5, m_step_tol=-1.0, seed=10),
}


def make_inputs(c):
    from poor_man_gplvm_b200.synthetic import make_dataset
    rng = np.random.default_rng(1000 + c["seed"])
    d = make_dataset(c["T"], c["N"], c["K"], seed=c["seed"])
    inp = {"y": d["y"].astype(np.float32)}
    inp["ma_neuron"] = np.ones(c["N"], np.float32)
    for i in c.get("mask_neuron", ()):
        inp["ma_neuron"][i] = 0
    inp["ma_latent"] = np.ones(c["K"], np.float32)
    for i in c.get("mask_latent", ()):
        inp["ma_latent"][i] = 0
    if c.get("mask_tn"):
        inp["ma_neuron"] = (rng.random((c["T"], c["N"])) < 0.8).astype(np.float32)
    if c.get("dt_l"):
        inp["dt_l"] = rng.uniform(0.5, 1.5, size=c["T"]).astype(np.float32)
    if c.get("custom_kernel") == "dense":
        x = np.arange(c["K"])
        dist = np.abs(x[:, None] - x[None, :])
        inp["custom_transition_kernel"] = (np.exp(-dist / 40.0) * (1 + 0.3 * rng.random((c["K"], c["K"]))) + 0.01
                                           ).astype(np.float32)
    post = rng.random((c["T"], c["K"])).astype(np.float32) * np.float32(0.1)
    post = post / post.sum(axis=1, keepdims=True)
    inp["log_posterior_init"] = np.log(post)
    return inp, rng


def run_case(name, c, ref, fdt):
    import jax.numpy as jnp
    inp, rng = make_inputs(c)
    mk = {}
    if "custom_transition_kernel" in inp:
        mk["custom_transition_kernel"] = jnp.array(inp["custom_transition_kernel"])
    m = ref.core.PoissonGPLVMJump1D(n_neuron=c["N"], n_latent_bin=c["K"], tuning_lengthscale=c["ls"],
                                    movement_variance=c.get("mv", 1.0), p_move_to_jump=c.get("pmj", 0.01),
                                    p_jump_to_move=c.get("pjm", 0.01), **mk)
    basis = np.asarray(m.tuning_basis)
    params0 = rng.standard_normal((basis.shape[1], c["N"])).astype(np.float32)
    m.params = jnp.array(params0)
    m.tuning = m.get_tuning(m.params, {}, m.tuning_basis)
    out = {"in_" + k: v for k, v in inp.items()}
    out["in_params"] = params0
    out["tuning_basis"] = basis.astype(fdt)
    out["tuning_init"] = np.asarray(m.tuning).astype(fdt)
    kw = dict(n_iter=c["n_iter"], log_posterior_init=jnp.array(inp["log_posterior_init"]), verboase=False,
              ma_neuron=jnp.array(inp["ma_neuron"]), ma_latent=jnp.array(inp["ma_latent"]),
              n_time_per_chunk=c.get("n_time_per_chunk", 10000), likelihood_scale=c.get("likelihood_scale", 1.0))
    for k in ("m_step_maxiter", "m_step_tol"):
        if k in c:
            kw[k] = c[k]
    t0 = time.time()
    em = m.fit_em(inp["y"], **kw)
    A = lambda x: np.asarray(x).astype(fdt)
    out["em_log_marginal_l"] = np.array([float(v) for v in em["log_marginal_l"]], dtype=fdt)
    out["em_params"] = A(em["params"])
    out["em_tuning"] = A(em["tuning"])
    out["em_posterior"] = A(em["posterior"])
    out["em_posterior_latent_marg"] = A(em["posterior_latent_marg"])
    out["em_posterior_dynamics_marg"] = A(em["posterior_dynamics_marg"])
    out["em_m_n_iter"] = np.array([int(v) for v in em["m_step_res_l"]["n_iter"]])
    out["em_m_final_loss"] = np.array([float(v) for v in em["m_step_res_l"]["final_loss"]], dtype=fdt)
    out["em_m_final_error"] = np.array([float(v) for v in em["m_step_res_l"]["final_error"]], dtype=fdt)
    out["em_m_loss_history0"] = A(em["m_step_res_l"]["loss_history"][0])
    if c.get("em_only"):
        for k in ("em_posterior",) + (("em_tuning", "tuning_init", "tuning_basis", "em_posterior_latent_marg",
                                       "em_m_loss_history0") if c.get("x64_only") else ()):
            out[k] = out[k].astype(np.float32)
        out["in_y"] = out["in_y"].astype(np.uint8)
        out["meta_case"] = np.array(repr(c))
        return out
    # decode with the fitted tuning (same masks / scale / chunking)
    dec = m.decode_latent(inp["y"], ma_neuron=kw["ma_neuron"], ma_latent=kw["ma_latent"],
                          likelihood_scale=kw["likelihood_scale"], n_time_per_chunk=kw["n_time_per_chunk"])
    for k in ("log_posterior_all", "posterior_all", "posterior_latent_marg", "posterior_dynamics_marg",
              "log_one_step_predictive_marginals_all", "log_likelihood_all", "p_joint_full", "p_joint_latent",
              "p_joint_dynamics", "p_transition_full", "p_transition_latent", "p_transition_dynamics",
              "log_joint_full", "log_transition_latent", "log_transition_dynamics"):
        out["dec_" + k] = A(dec[k])
    out["dec_log_marginal_final"] = np.array(float(dec["log_marginal_final"]), dtype=fdt)
    # the 6-tuple of _decode_latent: the causal (filtered) posterior is only visible there
    tup = m._decode_latent(jnp.array(inp["y"]), m.tuning, {}, m.log_latent_transition_kernel_l,
                           m.log_dynamics_transition_kernel, kw["ma_neuron"], kw["ma_latent"],
                           likelihood_scale=kw["likelihood_scale"], n_time_per_chunk=kw["n_time_per_chunk"])
    out["dec_log_causal_posterior_all"] = A(tup[2])
    out["dec_log_accumulated_joint_total"] = A(tup[4])
    nb_kw = dict(ma_neuron=kw["ma_neuron"], ma_latent=kw["ma_latent"], n_time_per_chunk=kw["n_time_per_chunk"])
    if "dt_l" in inp:
        nb_kw["dt_l"] = jnp.array(inp["dt_l"])
    nb = m.decode_latent_naive_bayes(inp["y"], **nb_kw)
    for k in ("log_posterior_latent", "log_marginal_l", "posterior_latent", "ll_per_pos_l"):
        out["nb_" + k] = A(nb[k])
    out["nb_log_marginal_total"] = np.array(float(nb["log_marginal_total"]), dtype=fdt)
    out["nb_argmax"] = np.argmax(np.asarray(nb["log_posterior_latent"]), axis=1).astype(np.int32)
    # transition matrices as the reference builds them
    P, logP, M, logM = ref.gpk.create_transition_prob_1d(m.possible_latent_bin, m.possible_dynamics,
                                                         c.get("mv", 1.0), c.get("pmj", 0.01), c.get("pjm", 0.01),
                                                         **({"custom_kernel": mk["custom_transition_kernel"]} if mk else {}))
    out["tr_P"], out["tr_logP"], out["tr_M"], out["tr_logM"] = A(P), A(logP), A(M), A(logM)
    out["meta_case"] = np.array(repr(c))
    # keep the fixtures small: for the README-sized cases the T-sized arrays are stored as float32 (6e-8
    # relative, far below every tolerance they are compared at) and redundant ones are dropped; the fp32
    # run of those cases keeps only the small arrays
    if c["T"] * c["K"] > 50000:
        for k in ("dec_log_posterior_all", "nb_ll_per_pos_l", "nb_posterior_latent", "dec_posterior_latent_marg",
                  "em_posterior_latent_marg"):
            out.pop(k, None)
        for k in list(out):
            a = out[k]
            if a.ndim >= 2 and a.shape[0] == c["T"]:
                if fdt == np.float32 and not k.startswith("in_"):
                    out.pop(k)
                elif a.dtype == np.float64:
                    out[k] = a.astype(np.float32)
        out["in_y"] = out["in_y"].astype(np.uint8)
    if c.get("custom_kernel"):
        # K x K x 2 x 2 outputs of a K = 256 model: keep what the tests read, in float32 where a tolerance allows
        for k in ("dec_p_transition_full", "dec_log_joint_full", "dec_p_joint_full", "tr_logP", "tr_P",
                  "dec_log_transition_latent", "dec_p_transition_latent", "dec_log_causal_posterior_all",
                  "dec_log_likelihood_all", "nb_log_posterior_latent", "dec_log_transition_dynamics"):
            out.pop(k, None)
        for k in ("dec_log_accumulated_joint_total", "dec_p_joint_latent"):
            out[k] = out[k].astype(np.float32)
    print("  %s [%s]: %.1f s, lml %s" % (name, np.dtype(fdt).name, time.time() - t0, out["em_log_marginal_l"][-1]),
          flush=True)
    return out


def main():
    if "--child" not in sys.argv:
        for x64 in ("1", "0"):
            env = dict(os.environ, JAXSHIM_X64=x64)
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child"] + sys.argv[1:], env=env)
        return
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    x64 = os.environ.get("JAXSHIM_X64") == "1"
    fdt = np.float64 if x64 else np.float32
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        if (c.get("custom_kernel") or c.get("x64_only")) and not x64:
            continue                                # (fp64 fixture only: a few MB)
        out = run_case(name, c, ref, fdt)
        np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (name, "f64" if x64 else "f32")), **out)


if __name__ == "__main__":
    main()
