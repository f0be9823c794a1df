"""GPU parity tests of the individual operators (through the C ABI) against the oracle.

Tolerances: the oracle runs in fp64; the kernels compute in fp32 (linear space),
so posteriors are compared at 1e-5 absolute (BASELINE.json north_star), log
marginals at 1e-4 relative, emission log-likelihoods at a few fp32 ulps of |ll|.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

torch = pytest.importorskip("torch")

from oracle import linear_ref as lin
from oracle import ref_numpy as ref
from poor_man_gplvm_b200.synthetic import make_dataset

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t.to(dtype or torch.float32).contiguous()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from poor_man_gplvm_b200 import ops as _ops
    return _ops


# ----------------------------------------------------------------------------- emission / NB
@pytest.mark.parametrize("T,N,K", [(1, 5, 3), (257, 37, 53), (1000, 30, 100), (3000, 200, 100), (513, 130, 400),
                                   (40000, 500, 400), (300, 70, 600)])
@pytest.mark.parametrize("impl", [0, 1])
def test_emission_matches_oracle(ops, T, N, K, impl):
    d = make_dataset(T, N, K, seed=T + N)
    rng = np.random.default_rng(0)
    ma_n = (rng.random(N) > 0.15).astype(np.float32)
    ma_l = (rng.random(K) > 0.1).astype(np.float32)
    want = ref.get_loglikelihood_ma_all(d["y"].astype(np.float64), d["tuning_true"].astype(np.float64), ma_n, ma_l)
    y = dev(d["y"])
    lgam = ops.lgamma_rowsum(y, dev(ma_n))
    y16 = ops.CountsF16(y)
    assert y16.exact
    got = host(ops.emission(y, dev(d["tuning_true"]), lgam, dev(ma_n), dev(ma_l), y16=y16, impl=impl))
    live = ma_l.astype(bool)
    assert np.all(got[:, ~live] == np.float32(-1e20))
    scale = np.maximum(1.0, np.abs(want[:, live]))
    assert np.max(np.abs(got[:, live] - want[:, live]) / scale) < 2e-6


def test_emission_noninteger_and_large_counts(ops):
    rng = np.random.default_rng(3)
    T, N, K = 300, 40, 64
    y = (rng.gamma(2.0, 3.0, size=(T, N))).astype(np.float32)      # non-integer "counts" (decoder.py:37-38)
    y[:5] = np.floor(y[:5]) + 300                                    # counts > 256
    tun = (rng.random((K, N)) * 4 + 0.01).astype(np.float32)
    tun[3, 4] = 0.0                                                   # zero rate -> log(1e-20)
    ones_n, ones_k = np.ones(N, np.float32), np.ones(K, np.float32)
    want = ref.get_loglikelihood_ma_all(y.astype(np.float64), tun.astype(np.float64), ones_n, ones_k)
    lgam = ops.lgamma_rowsum(dev(y), None)
    y16 = ops.CountsF16(dev(y))
    assert not y16.exact                                              # routed to the fp32 tiles
    got = host(ops.emission(dev(y), dev(tun), lgam, y16=y16))
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 3e-6


@pytest.mark.parametrize("T,N,ones", [(1, 5, True), (257, 500, True), (1000, 37, False), (513, 128, True)])
def test_counts_prepare_matches_separate_passes(ops, T, N, ones):
    """Fused pass over the counts (fp16 copy + ones column + exactness + lgamma row term, decoder.py:40) against
    the two-kernel path and NumPy/SciPy."""
    from scipy.special import gammaln
    rng = np.random.default_rng(T + N)
    y = rng.poisson(1.3, size=(T, N)).astype(np.float32)
    y[rng.random((T, N)) < 0.01] = 40.0            # a few large counts (beyond the small-count table)
    ma = (rng.random(N) > 0.2).astype(np.float32)
    for mask in (None, ma):
        c = ops.CountsF16(dev(y), ones_col=ones, ma_vec=None if mask is None else dev(mask), want_lgam=True,
                          want_ysum=True)
        assert c.exact
        got = host(c.data.float())
        assert np.array_equal(got[:, :N], y)
        if ones:
            assert np.all(got[:, N] == 1.0)
        assert np.all(got[:, N + (1 if ones else 0):] == 0.0)
        w = np.ones(N) if mask is None else mask.astype(np.float64)
        want = (gammaln(y.astype(np.float64) + 1.0) * w).sum(axis=1)
        assert np.max(np.abs(host(c.lgam) - want) / np.maximum(1.0, np.abs(want))) < 2e-6
        assert np.allclose(host(c.ysum), (y * w).sum(axis=1), rtol=1e-6)
        ref = host(ops.lgamma_rowsum(dev(y), None if mask is None else dev(mask)))
        assert np.max(np.abs(host(c.lgam) - ref) / np.maximum(1.0, np.abs(ref))) < 2e-6
    y2 = y.copy(); y2[T // 2, N // 2] = 0.3          # not representable exactly as a count
    assert not ops.CountsF16(dev(y2), ones_col=ones, want_lgam=True).exact


def test_naive_bayes_normalize(ops):
    d = make_dataset(700, 25, 90, seed=2)
    ones_n, ones_k = np.ones(25, np.float32), np.ones(90, np.float32)
    lp, lml_l, lml_tot, ll = ref.get_naive_bayes_ma_chunk(d["y"].astype(np.float64), d["tuning_true"].astype(np.float64),
                                                          ones_n, ones_k, n_time_per_chunk=300)
    got_lp, got_lml = ops.naive_bayes_normalize(dev(ll))
    assert np.max(np.abs(host(got_lp) - lp)) < 1e-4
    assert np.max(np.abs(host(got_lml) - lml_l) / np.abs(lml_l)) < 1e-6
    assert np.array_equal(host(got_lp).argmax(axis=1), lp.argmax(axis=1))


# ----------------------------------------------------------------------------- forward / backward
def _transition(K, mv, custom=None, pmj=0.02, pjm=0.05):
    from poor_man_gplvm_b200 import gp_kernel as gpk
    P, logP, M, logM = gpk.create_transition_prob_1d(np.arange(K), np.arange(2), mv, pmj, pjm, custom_kernel=custom)
    # band truncation against the jump floor (W=5 at movement_variance=1); K=200 keeps the full fp32 support (W=9)
    hostop = gpk.move_operator_host(K, mv, custom, p_move_to_jump=None if K == 200 else pmj)
    return P, logP, M, logM, hostop


def _run_scan(ops, ll, hostop, M, scale, chunk_len, halo, want_r=True):
    from poor_man_gplvm_b200.estep import EStep
    T, K = ll.shape
    op = ops.MoveOperator(hostop, M, torch.device("cuda"))
    y_dummy = torch.zeros((T, 1), device="cuda")
    es = EStep(y_dummy, op, None, None, scale, halo=halo, chunk_len=chunk_len)
    es.ll.copy_(dev(ll))
    es.emission = lambda tuning: es.ll          # scan-only test: keep the injected ll
    return es, es.run(None, want_gamma=True, want_gamma_lat=True, want_dyn=True, want_r=want_r)


SCAN_CASES = [
    # K, mv, custom, T, chunk_len, halo  -> exercises (Q, WPC, path)
    (24, 1.0, None, 200, 200, 0),        # Q=4 reg path, single exact chain
    (100, 1.0, None, 900, 150, 64),      # Q=4 reg, chains + halo
    (200, 1.0, None, 600, 128, 64),      # Q=8 reg
    (400, 1.0, None, 500, 125, 96),      # Q=13 reg
    (500, 0.7, None, 260, 90, 64),       # Q=16 reg
    (96, 1.0, None, 400, 100, 64),       # Q=4, bulk-copy kernels (K % 8 == 0)
    (512, 1.0, None, 300, 100, 64),      # Q=16, bulk-copy kernels
    (400, 1.0, None, 3000, 256, 128),    # Q=13 bulk, many ring wraps
    (100, 3.0, None, 700, 175, 128),     # Toeplitz generic path (W=27)
    (600, 1.0, None, 150, 50, 40),       # WPC=2
    (1100, 1.0, None, 100, 100, 0),      # WPC=4
    (64, 1.0, "band", 500, 125, 96),     # general banded custom kernel
    (48, 1.0, "dense", 400, 100, 80),    # dense custom kernel
]


@pytest.mark.parametrize("K,mv,custom,T,chunk_len,halo", SCAN_CASES)
def test_forward_backward_matches_linear_oracle(ops, K, mv, custom, T, chunk_len, halo):
    rng = np.random.default_rng(K + T)
    N = 20
    d = make_dataset(T, N, K, seed=K)
    ck = None
    if custom == "band":
        x = np.arange(K)
        ck = np.exp(-np.abs(x[:, None] - x[None, :]) / 1.5) * (np.abs(x[:, None] - x[None, :]) <= 6)
        ck = ck * (1 + 0.3 * rng.random((K, K)))
    elif custom == "dense":
        ck = rng.random((K, K)) + 0.05
    P, logP, M, logM, hostop = _transition(K, mv, ck)
    ma_l = np.ones(K); ma_l[K // 3] = 0
    ll = lin.emission_gemm_form(d["y"], d["tuning_true"], np.ones(N), ma_l).astype(np.float32)
    scale = 0.9
    want = lin.e_step(d["y"], d["tuning_true"], P.astype(np.float64), M.astype(np.float64), np.ones(N), ma_l,
                      likelihood_scale=scale, want_xi=True)
    es, res = _run_scan(ops, ll, hostop, M, scale, chunk_len, halo)
    assert np.max(np.abs(host(res.alpha) - want["alpha"])) < 1e-5
    assert np.max(np.abs(host(res.gamma) - want["gamma"])) < 1e-5
    assert np.max(np.abs(host(res.gamma_lat) - want["gamma"].sum(axis=1))) < 1e-5
    assert np.max(np.abs(host(res.dyn_marg) - want["gamma"].sum(axis=2))) < 1e-5
    assert abs(float(res.log_marginal) - want["log_marginal"]) < 1e-4 * abs(want["log_marginal"])
    assert np.max(np.abs(host(res.lmr) - want["lmr"])) < 1e-3
    assert np.max(np.abs(host(res.tw) - want["gamma"].sum(axis=(0, 1)))) < 1e-3
    # transition counts through the time-reduction GEMM + finalize
    G = ops.atb(res.alpha.view(T, 2 * K)[:T - 1], res.r.view(T, 2 * K)[1:])
    log_acc = host(ops.xi_finalize(G, dev(logP), logM))
    xi = np.exp(log_acc.astype(np.float64))
    assert abs(xi.sum() - (T - 1)) < 1e-3 * (T - 1)
    assert np.max(np.abs(xi - want["xi"])) < 2e-4 * max(1.0, want["xi"].max())
    # the same counts from bf16 hi/lo pieces written by the backward kernel itself (pmg_backward_xi16 +
    # pmg_atb_bf16x2_pieces: decode_latent's path on long recordings) -- where the bulk kernel runs
    if ops.xi16_supported(es.op, scale):
        res2 = es.run(None, want_gamma=True, want_gamma_lat=True, want_dyn=True, want_r=True, xi16_ok=True)
        assert res2.xi16 is not None and res2.r_ext is None
        x16 = host(res2.xi16.float()).astype(np.float64)
        a_rec, r_rec = x16[0] + x16[1], x16[2] + x16[3]
        a_ref, r_ref = host(res.alpha).reshape(T, 2 * K), host(res.r).reshape(T, 2 * K)
        assert np.max(np.abs(a_rec - a_ref)) <= 2.0 ** -16 * np.abs(a_ref).max() + 2e-6     # (second pass: warm starts)
        assert np.max(np.abs(r_rec[1:] - r_ref[1:])) <= (2.0 ** -16 + 1e-5) * np.abs(r_ref).max()
        G2 = host(ops.atb_bf16x2_pieces(res2.xi16, 0, T - 1))
        Gh = host(G)
        assert np.max(np.abs(G2 - Gh)) < 1e-4 * np.abs(Gh).max()
        assert np.max(np.abs(host(res2.gamma) - host(res.gamma))) < 2e-6
    else:
        assert custom is not None or K % 8 or K > 512 or mv > 2.0


DENSE_TC_CASES = [
    # K, kernel, T, chunk_len, halo      -> (column tiles, K blocks) of the lockstep GEMM
    (256, "dense", 1500, 125, 96),       # 1 column tile, 4 K blocks, 12 chains
    (400, "wide", 900, 100, 80),         # 2 column tiles of 208, Kk = 448; Toeplitz kernel with a wide band
    (520, "dense", 700, 64, 64),         # 3 column tiles
    (320, "band", 800, 200, 128),        # block-banded custom kernel: K-block ranges skip the zero blocks
    (256, "dense", 300, 300, 0),         # one exact chain
]


@pytest.mark.parametrize("K,kernel,T,chunk_len,halo", DENSE_TC_CASES)
def test_lockstep_tensor_core_scan_matches_linear_oracle(ops, K, kernel, T, chunk_len, halo):
    """pmg_forward_dense / pmg_backward_dense (dense and wide-band move kernels as one tcgen05 GEMM per time step
    over all chains, reference decoder.py:151-256 with gp_kernel.py:61-66 kernels) against the fp64 oracle, and
    against the CUDA-core general path on the same inputs."""
    from poor_man_gplvm_b200.estep import EStep
    rng = np.random.default_rng(K + T)
    N = 20
    d = make_dataset(T, N, K, seed=K + 1)
    x = np.arange(K)
    dist = np.abs(x[:, None] - x[None, :])
    ck, mv = None, 1.0
    if kernel == "dense":
        ck = np.exp(-dist / 40.0) * (1 + 0.3 * rng.random((K, K))) + 0.01
    elif kernel == "band":
        ck = np.exp(-dist / 30.0) * (dist <= 70) * (1 + 0.3 * rng.random((K, K)))
    else:
        mv = 25.0
    P, logP, M, logM, hostop = _transition(K, mv, ck)
    ma_l = np.ones(K); ma_l[K // 3] = 0
    ll = lin.emission_gemm_form(d["y"], d["tuning_true"], np.ones(N), ma_l).astype(np.float32)
    scale = 0.9
    want = lin.e_step(d["y"], d["tuning_true"], P.astype(np.float64), M.astype(np.float64), np.ones(N), ma_l,
                      likelihood_scale=scale, want_xi=True)
    outs = {}
    for tc in (True, False):
        op = ops.MoveOperator(hostop, M, torch.device("cuda"), P0=P[0], dense_tc=tc)
        assert (op.dense is not None) == tc
        es = EStep(torch.zeros((T, 1), device="cuda"), op, None, None, scale, halo=halo, chunk_len=chunk_len)
        es.ll.copy_(dev(ll))
        es.emission = lambda tuning, es=es: es.ll
        g16 = ops.new_gamma16(T, K, torch.device("cuda"))
        res = es.run(None, want_gamma=True, want_gamma_lat=True, want_dyn=True, want_r=True, gamma16=g16, want_tw=True)
        outs[tc] = res
        assert res.n_relay_fwd == 0 and res.n_relay_bwd == 0
        assert np.max(np.abs(host(res.alpha) - want["alpha"])) < 1e-5
        assert np.max(np.abs(host(res.gamma) - want["gamma"])) < 1e-5
        assert np.max(np.abs(host(res.gamma_lat) - want["gamma"].sum(axis=1))) < 1e-5
        assert np.max(np.abs(host(res.dyn_marg) - want["gamma"].sum(axis=2))) < 1e-5
        assert abs(float(res.log_marginal) - want["log_marginal"]) < 1e-4 * abs(want["log_marginal"])
        assert np.max(np.abs(host(res.lmr) - want["lmr"])) < 1e-3
        assert np.max(np.abs(host(res.tw) - want["gamma"].sum(axis=(0, 1)))) < 1e-3
        pieces = host(g16.float())
        assert np.max(np.abs(pieces[0, :, :K] + pieces[1, :, :K] - want["gamma"].sum(axis=1))) < 1e-5
        G = ops.atb(res.alpha.view(T, 2 * K)[:T - 1], res.r.view(T, 2 * K)[1:])
        xi = np.exp(host(ops.xi_finalize(G, dev(logP), logM)).astype(np.float64))
        assert abs(xi.sum() - (T - 1)) < 1e-3 * (T - 1)
        assert np.max(np.abs(xi - want["xi"])) < 2e-4 * max(1.0, want["xi"].max())
    # the two paths agree far inside the tolerance (22-bit operand pieces vs fp32)
    assert np.max(np.abs(host(outs[True].gamma) - host(outs[False].gamma))) < 5e-6
    assert abs(float(outs[True].log_marginal) - float(outs[False].log_marginal)) < 2e-6 * abs(want["log_marginal"])


def test_relay_repairs_failed_seams(ops):
    """With a useless halo every seam fails; the relay must reproduce the sequential answer."""
    K, T, N = 100, 600, 15
    d = make_dataset(T, N, K, seed=11)
    P, logP, M, logM, hostop = _transition(K, 1.0)
    ll = lin.emission_gemm_form(d["y"], d["tuning_true"], np.ones(N), np.ones(K)).astype(np.float32)
    want = lin.e_step(d["y"], d["tuning_true"], P.astype(np.float64), M.astype(np.float64), np.ones(N), np.ones(K))
    es, res = _run_scan(ops, ll, hostop, M, 1.0, chunk_len=100, halo=2, want_r=False)
    assert res.n_relay_fwd >= 1 and res.n_relay_bwd >= 1
    assert np.max(np.abs(host(res.alpha) - want["alpha"])) < 1e-5
    assert np.max(np.abs(host(res.gamma) - want["gamma"])) < 1e-5
    assert abs(float(res.log_marginal) - want["log_marginal"]) < 1e-4 * abs(want["log_marginal"])


def test_chunking_is_a_numerical_noop(ops):
    """reference invariant: results do not depend on the chunking (decoder.py:299,322)."""
    K, T, N = 200, 2000, 30
    d = make_dataset(T, N, K, seed=5)
    P, logP, M, logM, hostop = _transition(K, 1.0)
    ll = lin.emission_gemm_form(d["y"], d["tuning_true"], np.ones(N), np.ones(K)).astype(np.float32)
    _, a = _run_scan(ops, ll, hostop, M, 1.0, chunk_len=T, halo=0, want_r=False)
    ga, lma = host(a.gamma).copy(), float(a.log_marginal)
    _, b = _run_scan(ops, ll, hostop, M, 1.0, chunk_len=250, halo=200, want_r=False)
    assert np.max(np.abs(host(b.gamma) - ga)) < 2e-6
    assert abs(float(b.log_marginal) - lma) < 1e-6 * abs(lma)


# ----------------------------------------------------------------------------- compact (EM fast path) kernels
def _run_scan_compact(ops, ll, hostop, M, scale, chunk_len, halo):
    from poor_man_gplvm_b200.estep import EStep
    T, K = ll.shape
    op = ops.MoveOperator(hostop, M, torch.device("cuda"))
    es = EStep(torch.zeros((T, 1), device="cuda"), op, None, None, scale, halo=halo, chunk_len=chunk_len)
    assert es.compact_ok
    es.ll.copy_(dev(ll))
    es.emission = lambda tuning: es.ll
    g16 = ops.new_gamma16(T, K, torch.device("cuda"))
    res = es.run(None, want_gamma_lat=False, gamma16=g16)
    return es, res, (g16[0].float() + g16[1].float())[:, :K]


COMPACT_CASES = [
    # K, mv, T, chunk_len, halo -> (pairs per lane, band capacity)
    (24, 1.0, 200, 200, 0),         # QP=4, single exact chain
    (96, 1.0, 400, 100, 64),        # QP=4, chains + halo
    (200, 1.0, 603, 128, 64),       # QP=4, ragged last chain, odd chunk count
    (248, 1.0, 300, 75, 50),        # QP=7 (K + 4 > 31*8)
    (400, 1.0, 3001, 256, 128),     # QP=7, many ring wraps, headline K
    (432, 1.0, 260, 90, 64),        # QP=8
    (488, 0.7, 150, 50, 40),        # QP=8, narrow kernel
    (400, 1.6, 500, 125, 96),       # W in (5, 10]: wide-band instantiation
    (104, 1.8, 300, 100, 64),       # wide band, small K -> QP=7
]


@pytest.mark.parametrize("K,mv,T,chunk_len,halo", COMPACT_CASES)
def test_compact_scan_matches_linear_oracle(ops, K, mv, T, chunk_len, halo):
    N = 20
    d = make_dataset(T, N, K, seed=K + 1)
    P, logP, M, logM, hostop = _transition(K, mv)
    assert hostop["W"] <= 10
    ma_l = np.ones(K); ma_l[K // 3] = 0
    ll = lin.emission_gemm_form(d["y"], d["tuning_true"], np.ones(N), ma_l).astype(np.float32)
    scale = 0.9
    want = lin.e_step(d["y"], d["tuning_true"], P.astype(np.float64), M.astype(np.float64), np.ones(N), ma_l,
                      likelihood_scale=scale)
    es, res, glat = _run_scan_compact(ops, ll, hostop, M, scale, chunk_len, halo)
    assert res.alpha is None                                  # the compact path was taken
    assert np.max(np.abs(host(glat) - want["gamma"].sum(axis=1))) < 1e-5
    assert abs(float(res.log_marginal) - want["log_marginal"]) < 1e-4 * abs(want["log_marginal"])
    assert np.max(np.abs(host(res.lmr) - want["lmr"])) < 1e-3
    assert res.tw is None                                     # left to the statistics GEMM (ones column)
    # compact filtered posterior: alpha[0,:] stored, alpha[1,:] = a1s * E
    ax = host(es.ax)
    assert np.max(np.abs(ax[:, :K] - want["alpha"][:, 0])) < 1e-5
    E = np.exp2((ll - ll.max(axis=1, keepdims=True)) * np.float32(scale * 1.4426950408889634))
    assert np.max(np.abs(ax[:, K:K + 1] * E - want["alpha"][:, 1])) < 1e-5


def test_compact_scan_equals_general_path_and_relays(ops):
    """Same inputs through both kernel families; a useless halo forces relays on the compact path too."""
    K, T, N = 200, 1500, 25
    d = make_dataset(T, N, K, seed=21)
    P, logP, M, logM, hostop = _transition(K, 1.0)
    ll = lin.emission_gemm_form(d["y"], d["tuning_true"], np.ones(N), np.ones(K)).astype(np.float32)
    _, full = _run_scan(ops, ll, hostop, M, 1.0, chunk_len=250, halo=128, want_r=False)
    gl_full, lm_full = host(full.gamma_lat).copy(), float(full.log_marginal)
    _, res, glat = _run_scan_compact(ops, ll, hostop, M, 1.0, chunk_len=250, halo=128)
    assert np.max(np.abs(host(glat) - gl_full)) < 3e-6
    assert abs(float(res.log_marginal) - lm_full) < 1e-6 * abs(lm_full)
    _, res2, glat2 = _run_scan_compact(ops, ll, hostop, M, 1.0, chunk_len=100, halo=2)
    assert res2.n_relay_fwd >= 1 and res2.n_relay_bwd >= 1
    assert np.max(np.abs(host(glat2) - gl_full)) < 1e-5
    assert abs(float(res2.log_marginal) - lm_full) < 1e-5 * abs(lm_full)


# ----------------------------------------------------------------------------- time-reduction GEMM
@pytest.mark.parametrize("T,M,N", [(1, 3, 5), (1000, 100, 30), (4097, 130, 70), (20000, 400, 500)])
@pytest.mark.parametrize("impl", [0, 1])
def test_atb_matches_numpy(ops, T, M, N, impl):
    rng = np.random.default_rng(T)
    A = rng.random((T, M)).astype(np.float32) ** 4
    B = rng.poisson(0.7, size=(T, N)).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    got = host(ops.atb(dev(A), dev(B), impl=impl))
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 5e-6


@pytest.mark.parametrize("T,M,N", [(1, 3, 5), (1000, 96, 40), (4099, 130, 70), (6001, 800, 800), (3000, 264, 520)])
def test_atb_bf16x2_tensor_core_matches_numpy(ops, T, M, N):
    """Transition-count GEMM on tcgen05: both operands as bf16 hi/lo pieces, three products.  Entries
    span many orders of magnitude (filtered posteriors); every term keeps 16 significant bits at any scale."""
    rng = np.random.default_rng(T + M)
    A = (rng.random((T, M)) ** 8 * 10.0 ** rng.integers(-12, 1, size=(T, M))).astype(np.float32)
    B = (rng.random((T, N)) ** 8 * 10.0 ** rng.integers(-12, 3, size=(T, N))).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    # row-offset views, as core._transition_counts passes them
    Ad, Bd = dev(np.concatenate([A[:1], A])), dev(np.concatenate([B, B[:1]]))
    got = host(ops.atb_bf16x2(Ad[1:], Bd[:-1]))
    assert got.shape == (M, N)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-30)) < 6e-5
    # tiny inputs are not flushed: scale everything down by 1e-20
    got_s = host(ops.atb_bf16x2(dev(A * np.float32(1e-10)), dev(B * np.float32(1e-10))))
    assert np.max(np.abs(got_s - want * 1e-20) / np.maximum(np.abs(want * 1e-20), 1e-36)) < 6e-5


def test_seam_check_is_scale_invariant(ops):
    """Messages are defined up to a positive factor (both passes renormalise every step): the seam check compares
    them after normalisation, so a rescaled copy passes and a perturbed entry is reported with its relative size."""
    rng = np.random.default_rng(3)
    a = (rng.random((5, 64)) ** 4).astype(np.float32)
    b = a * np.array([1.0, 3.7, 1e-3, 2.0, 1.0], np.float32)[:, None]
    b[3, 10] *= 1.01
    b[4, :] = 0.0
    ad, bd = dev(a), dev(b)
    err = torch.zeros(5, device="cuda")
    ops.seam_check(5, 64, ad.data_ptr(), 64, bd.data_ptr(), 64, err)
    e = host(err)
    assert np.all(e[:3] < 1e-6)
    assert 5e-3 < e[3] < 2e-2
    assert not np.isfinite(e[4])


@pytest.mark.parametrize("T,K,N", [(1, 3, 5), (1000, 100, 30), (4099, 130, 70), (20000, 400, 500), (3000, 600, 200)])
def test_atb_f16_tensor_core_matches_numpy(ops, T, K, N):
    """Statistics GEMM on tcgen05: posterior as fp16 hi/lo pieces (22 bits), counts exact in fp16."""
    rng = np.random.default_rng(T + K)
    G = rng.random((T, K)).astype(np.float32) ** 6
    G /= G.sum(axis=1, keepdims=True)
    Y = rng.poisson(0.7, size=(T, N)).astype(np.float32)
    want = G.astype(np.float64).T @ Y.astype(np.float64)
    y16 = ops.CountsF16(dev(Y))
    g16 = ops.split_f16(dev(G))
    got = host(ops.atb_f16(g16, y16, K))
    # fp16 pieces: per-element error <= max(2^-22 rel, 3e-8 abs); sums of T terms
    assert np.max(np.abs(got - want)) < 2e-6 * np.max(np.abs(want)) + 1e-5
    ref32 = host(ops.atb(dev(G), dev(Y), impl=1))
    assert np.max(np.abs(got - ref32)) < 4e-6 * np.max(np.abs(want)) + 1e-5


# ----------------------------------------------------------------------------- boundary exchange kernels
@pytest.mark.parametrize("K", [24, 400, 2000])
def test_boundary_pack_unpack_match_the_host_restatement(ops, K):
    """pmg_boundary_pack_fwd / pmg_boundary_unpack_fwd (one kernel on each side of the time-sharded boundary
    exchange) against the NumPy restatement the CPU orchestration tests run on (tests/cpu_scan_emulation.py)."""
    import cpu_scan_emulation as emu
    rng = np.random.default_rng(K)
    scale = 0.8
    first, last, warm = (rng.random(2 * K).astype(np.float32) for _ in range(3))
    ax_row = np.concatenate([rng.random(K), [0.37, -3.0, 0.0, 0.0]]).astype(np.float32)
    ll_row = (rng.standard_normal(K) * 4 - 30).astype(np.float32)
    for kw in ({}, {"warm_src": warm}, {"ax_row": ax_row, "ll_row": ll_row}):
        want = np.full(8 * K, 7.0, np.float32)
        emu.boundary_pack_fwd(K, want, first, last, scale=scale, **kw)
        got = torch.full((8 * K,), 7.0, device="cuda")
        ops.boundary_pack_fwd(K, got, dev(first), dev(last), scale=scale, **{k: dev(v) for k, v in kw.items()})
        assert np.allclose(host(got), want, rtol=2e-5, atol=1e-30), kw.keys()     # (fp32 exp2 of arguments down to -100)
    from_left, from_right = rng.random(4 * K).astype(np.float32), rng.random(4 * K).astype(np.float32)
    for compact in (False, True):
        w_end, w_warm = np.zeros((2, K), np.float32), np.zeros((2, K), np.float32)
        g_end, g_warm = torch.zeros((2, K), device="cuda"), torch.zeros((2, K), device="cuda")
        if compact:
            w_ax, g_ax = np.zeros(K + 4, np.float32), torch.zeros(K + 4, device="cuda")
            emu.boundary_unpack_fwd(K, from_left, from_right, w_end, w_warm, ax_stop=w_ax, ll_stop=ll_row, scale=scale)
            ops.boundary_unpack_fwd(K, dev(from_left), dev(from_right), g_end, g_warm, ax_stop=g_ax, ll_stop=dev(ll_row),
                                    scale=scale)
            assert np.allclose(host(g_ax), w_ax, rtol=2e-5)
        else:
            w_al, g_al = np.zeros(2 * K, np.float32), torch.zeros(2 * K, device="cuda")
            emu.boundary_unpack_fwd(K, from_left, from_right, w_end, w_warm, alpha_stop=w_al)
            ops.boundary_unpack_fwd(K, dev(from_left), dev(from_right), g_end, g_warm, alpha_stop=g_al)
            assert np.array_equal(host(g_al), w_al)
        assert np.array_equal(host(g_end), w_end) and np.array_equal(host(g_warm), w_warm)
    # a rank at the edge of the recording: nothing arrives from that side, nothing is touched
    g_end = torch.full((2, K), 5.0, device="cuda")
    ops.boundary_unpack_fwd(K, None, None, g_end, None)
    assert float(g_end.min()) == 5.0


# ----------------------------------------------------------------------------- M-step
@pytest.mark.parametrize("K,N,ls", [(30, 7, 5.0), (100, 30, 10.0), (100, 33, 1.0), (400, 50, 10.0)])
def test_mstep_adam_matches_oracle(ops, K, N, ls):
    rng = np.random.default_rng(K + N)
    basis = ref.generate_basis(ls, K, dtype=np.float32)
    B = basis.shape[1]
    W0 = rng.standard_normal((B, N)).astype(np.float32)
    yw = (rng.random((K, N)) * 5).astype(np.float32)
    yw[0, 0] = 0.0
    tw = (rng.random(K) + 0.5).astype(np.float32)
    maxiter = 60
    want = ref.adam_run(W0.astype(np.float64), ref.adam_init(W0.astype(np.float64)), 1.2, basis.astype(np.float64),
                        yw.astype(np.float64), tw.astype(np.float64), step_size=0.01, maxiter=maxiter, tol=-1)
    W = dev(W0).clone()
    st = ops.AdamState(W)
    lh, eh, n_it, fin, tuning = ops.mstep_adam(dev(basis), dev(yw), dev(tw), W, st, 1.2, 0.01, maxiter, -1.0)
    assert int(n_it.item()) == want["n_iter"] == maxiter
    assert int(st.count.item()) == maxiter - 1
    assert np.max(np.abs(host(W) - want["params"])) < 2e-4
    assert np.max(np.abs(host(lh) - want["loss_history"]) / np.abs(want["loss_history"])) < 1e-5
    assert np.max(np.abs(host(eh) - want["error_history"]) / np.abs(want["error_history"])) < 1e-3
    assert abs(float(fin[0]) - want["final_loss"]) < 1e-5 * abs(want["final_loss"])
    tun_want = ref.get_tuning_softplus(want["params"], basis.astype(np.float64))
    assert np.max(np.abs(host(tuning) - tun_want) / tun_want) < 1e-3
    # second call continues the optimiser state (count, mu, nu) like core.py:662
    want2 = ref.adam_run(want["params"], want["opt_state"], 1.2, basis.astype(np.float64), yw.astype(np.float64),
                         tw.astype(np.float64), step_size=0.01, maxiter=20, tol=-1)
    ops.mstep_adam(dev(basis), dev(yw), dev(tw), W, st, 1.2, 0.01, 20, -1.0)
    assert int(st.count.item()) == maxiter - 1 + 19
    assert np.max(np.abs(host(W) - want2["params"])) < 4e-4


@pytest.mark.parametrize("K,N,ls,maxiter,tol", [(100, 30, 10.0, 1000, 1e-6), (400, 500, 10.0, 300, 1e-6),
                                                  (64, 7, 4.0, 40, -1.0), (48, 20, 1.0, 12, 1e-3)])
def test_mstep_lagged_kernel_matches_barrier_kernel(ops, K, N, ls, maxiter, tol):
    """The M-step kernel that keeps stepping while the stop rule of step i-4 is being reduced (and rolls back to
    the step the rule selects) against the barrier-per-step kernel: same rule, same step count, same optimum up
    to the rounding of two separately compiled fp32 loops (which Adam amplifies to ~1e-3 in W after hundreds of
    steps)."""
    import os
    from poor_man_gplvm_b200 import gp_kernel as gpk
    rng = np.random.default_rng(K + N)
    basis = gpk.generate_basis(ls, K)
    B = basis.shape[1]
    tw = (rng.random(K) * 50 + 1).astype(np.float32)
    yw = (rng.random((K, N)) * tw[:, None] * 1.5).astype(np.float32)
    W0 = rng.standard_normal((B, N)).astype(np.float32)
    outs = []
    for lag in ("0", "1"):
        os.environ["PMG_MSTEP_LAG"] = lag
        try:
            W = dev(W0.copy())
            st = ops.AdamState(W)
            res = []
            for rep in range(2):                      # second call: Adam state carried over (count > 0)
                lh, eh, n_it, fin, tuning = ops.mstep_adam(dev(basis), dev(yw), dev(tw), W, st, 1.0, 0.01, maxiter, tol)
                res.append([host(x).copy() for x in (lh, eh, n_it, fin, tuning, W, st.mu, st.nu, st.count)])
            outs.append(res)
        finally:
            os.environ.pop("PMG_MSTEP_LAG", None)
    for a, b in zip(outs[0], outs[1]):
        na, nb = int(a[2][0]), int(b[2][0])
        assert 1 <= nb <= maxiter and abs(na - nb) <= max(1, na // 50)
        assert int(a[8][0]) - na == int(b[8][0]) - nb                       # Adam count advanced by n_iter - 1
        n = min(na, nb)
        assert np.max(np.abs(a[0][:n] - b[0][:n]) / np.abs(a[0][:n])) < 1e-5  # loss history
        assert np.all(b[0][nb:] == 0) and np.all(b[1][nb:] == 0)             # nothing written past the stop
        assert np.max(np.abs(a[4] - b[4]) / np.maximum(a[4], 1e-3)) < 2e-3     # tuning
        assert np.max(np.abs(a[5] - b[5])) < 1e-2                              # W


def test_mstep_default_stop_rule_close_to_oracle(ops):
    rng = np.random.default_rng(9)
    K, N = 60, 12
    basis = ref.generate_basis(6.0, K, dtype=np.float32)
    B = basis.shape[1]
    W0 = rng.standard_normal((B, N)).astype(np.float32)
    yw = (rng.random((K, N)) * 5).astype(np.float32)
    tw = (rng.random(K) + 0.5).astype(np.float32)
    want = ref.adam_run(W0.copy(), ref.adam_init(W0), 1.0, basis, yw, tw, maxiter=1000, tol=1e-6)
    W = dev(W0).clone()
    st = ops.AdamState(W)
    lh, eh, n_it, fin, _ = ops.mstep_adam(dev(basis), dev(yw), dev(tw), W, st, 1.0, 0.01, 1000, 1e-6)
    n = int(n_it.item())
    # the stop test compares fp32 losses at ~8 ulp: step counts agree only roughly
    assert 6 <= n <= 1000 and abs(n - want["n_iter"]) <= 0.25 * want["n_iter"] + 10
    assert np.all(host(lh)[n:] == 0)
    assert abs(float(fin[0]) - want["final_loss"]) < 1e-4 * abs(want["final_loss"])
