"""Host-side orchestration of the E-step on CPU: chain plan, warm starts carried from pass to pass, seam
verification, Jacobi repair sweeps and the rank-boundary exchanges of `estep.EStep`, run under gloo with 1, 3
and 8 ranks against the sequential (single-chain) answer of the linear-space oracle.

The CUDA scan operators are replaced by NumPy stand-ins with the same interface conventions
(tests/cpu_scan_emulation.py); everything else -- `EStep.run`, `TimeShard`, the plan -- is the product code.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _patch():
    """Route the operators EStep uses to the CPU stand-ins; no CUDA call is left on its path."""
    os.environ["PMG_SCAN_COMPACT"] = "0"
    import cpu_scan_emulation as emu
    from poor_man_gplvm_b200 import ops

    class _Props:
        multi_processor_count = 2

    class _Stream:
        def synchronize(self):
            pass

    torch.cuda.get_device_properties = lambda dev=None: _Props()
    torch.cuda.current_stream = lambda dev=None: _Stream()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    ops.forward, ops.backward, ops.seam_check = emu.forward, emu.backward, emu.seam_check
    ops.seam_check_fix = emu.seam_check_fix
    ops.boundary_pack_fwd, ops.boundary_unpack_fwd = emu.boundary_pack_fwd, emu.boundary_unpack_fwd
    ops.strided_sum_workspace = lambda device: None
    ops.strided_sum = lambda src, dst, ws: dst.copy_(src.sum(dim=0, keepdim=True, dtype=torch.float64))
    ops.EmissionOperands = emu.FakeEmission
    ops.scan_compact_supported = lambda op, scale: False
    return emu


def _problem(T_total, N, K, seed):
    from poor_man_gplvm_b200.synthetic import make_dataset
    from poor_man_gplvm_b200 import gp_kernel as gpk
    d = make_dataset(T_total, N, K, seed=seed)
    P, logP, M, logM = gpk.create_transition_prob_1d(np.arange(K), np.arange(2), 1.0, 0.02, 0.05)
    host = gpk.move_operator_host(K, 1.0, None, p_move_to_jump=0.02)
    return d, P, M, host


def _tunings(d, n_pass):
    """A slowly changing tuning, as successive EM iterations produce (exercises the carried warm starts)."""
    rng = np.random.default_rng(3)
    base = d["tuning_true"].astype(np.float64)
    out = []
    freeze = int(os.environ.get("PMG_TEST_FREEZE_AFTER", n_pass))      # passes >= freeze see the converged tuning
    for i in range(n_pass):
        k = 0 if i >= freeze else (n_pass - 1 - i)
        out.append((base * (1.0 + 0.05 * k * rng.standard_normal(base.shape) * 0.2 + 0.0)).clip(1e-3))
    return out


def _worker(rank, world, port, cfg, q, device_repair=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PMG_DEVICE_REPAIR"] = str(device_repair)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _patch()
        from poor_man_gplvm_b200 import ops
        from poor_man_gplvm_b200.estep import EStep
        from poor_man_gplvm_b200.shard import TimeShard
        T_total, N, K, halo, chunk, n_pass, seed = cfg
        d, P, M, host = _problem(T_total, N, K, seed)
        per = T_total // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else T_total
        y = torch.from_numpy(d["y"][lo:hi].copy())
        op = ops.MoveOperator(host, M, torch.device("cpu"), P0=P[0])
        es = EStep(y, op, None, None, 1.0, halo=halo, chunk_len=chunk, shard=TimeShard() if world > 1 else None,
                   adaptive=bool(os.environ.get("PMG_TEST_ADAPTIVE")))
        out = []
        for tun in _tunings(d, n_pass):
            res = es.run(torch.from_numpy(tun.astype(np.float32)), want_gamma=True, want_gamma_lat=True,
                         want_dyn=True, want_r=False)
            out.append({"gamma": res.gamma.numpy().copy(), "lm": float(res.log_marginal),      # global already
                        "relay": (res.n_relay_fwd + res.n_fix_fwd, res.n_relay_bwd + res.n_fix_bwd),
                        "host_relay": (res.n_relay_fwd, res.n_relay_bwd),
                        "repaired": res.repaired, "err": (res.seam_err_fwd, res.seam_err_bwd),
                        "n_chain": res.plan.n_chain, "halo": res.halo,
                        "boosted": (int((es.boost_f > 0).sum() + (es.boost_b > 0).sum()) if es.adaptive else 0)})
        q.put((rank, out))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if world > 1:
            dist.destroy_process_group()


def _run(world, cfg, device_repair=1):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q, device_repair)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert isinstance(res[r], list), res[r]
    return res


def _oracle(cfg):
    from oracle import linear_ref as lin
    T_total, N, K, halo, chunk, n_pass, seed = cfg
    d, P, M, host = _problem(T_total, N, K, seed)
    out = []
    for tun in _tunings(d, n_pass):
        r = lin.e_step(d["y"], tun, P.astype(np.float64), M.astype(np.float64), np.ones(N), np.ones(K))
        # FakeEmission drops the lgamma row term (constant in k): same posteriors, shifted log marginal
        from scipy.special import gammaln
        out.append({"gamma": r["gamma"], "lm": r["log_marginal"] + gammaln(d["y"].astype(np.float64) + 1).sum()})
    return out


@pytest.mark.parametrize("world,device_repair", [(1, 1), (3, 1), (8, 1), (1, 0), (3, 0)])
def test_estep_orchestration_matches_sequential_answer(world, device_repair):
    # 8 ranks x 96 bins, chains of 24 bins, warm-up of 16: several chains per rank, rank boundaries with both
    # neighbours, a warm-up too short for the first (cold) pass -> repairs (tier 1: conditional relaunch selected
    # "on the device"; tier 2 / device_repair=0: host sweeps); later passes start warm
    cfg = (768, 12, 24, 16, 24, 3, 5)
    want = _oracle(cfg)
    got = _run(world, cfg, device_repair)
    n_pass = cfg[5]
    for i in range(n_pass):
        gamma = np.concatenate([got[r][i]["gamma"] for r in range(world)])
        assert gamma.shape == want[i]["gamma"].shape
        assert np.max(np.abs(gamma - want[i]["gamma"])) < 2e-5
        for r in range(world):
            assert abs(got[r][i]["lm"] - want[i]["lm"]) < 1e-5 * abs(want[i]["lm"])
            assert max(got[r][i]["err"]) <= 1e-5
            # the verdict is global: every rank reports the same
            assert got[r][i]["repaired"] == got[0][i]["repaired"]
    assert got[0][0]["n_chain"] >= 2
    # the cold first pass needs repairs somewhere (that path is the point of the test) ...
    assert any(sum(got[r][0]["relay"]) > 0 for r in range(world))
    # ... and warm starts carried across passes (and across rank boundaries) reduce them
    first = sum(sum(got[r][0]["relay"]) for r in range(world))
    last = sum(sum(got[r][n_pass - 1]["relay"]) for r in range(world))
    assert last <= first


def test_adaptive_warm_up_shrinks_and_boosts(monkeypatch):
    """Adaptive warm-up (EM mode): the common base halves while no seam needs a repair, chains whose seam fails get
    their own longer warm-up, and the posteriors stay equal to the sequential answer throughout (3 ranks)."""
    monkeypatch.setenv("PMG_TEST_ADAPTIVE", "1")
    monkeypatch.setenv("PMG_HALO_MIN", "4")
    monkeypatch.setenv("PMG_TEST_FREEZE_AFTER", "3")
    cfg = (960, 12, 24, 32, 40, 12, 7)      # halo 32, chains of 40 bins, 12 passes; the tuning is converged from pass 3
    want = _oracle(cfg)
    got = _run(3, cfg)
    for i in range(cfg[5]):
        gamma = np.concatenate([got[r][i]["gamma"] for r in range(3)])
        assert np.max(np.abs(gamma - want[i]["gamma"])) < 2e-5, i
        for r in range(3):
            assert abs(got[r][i]["lm"] - want[i]["lm"]) < 1e-5 * abs(want[i]["lm"])
            assert got[r][i]["halo"] == got[0][i]["halo"]            # the base is global
    halos = [got[0][i]["halo"] for i in range(cfg[5])]
    assert halos[0] == 32 and halos[-1] < 32, halos                 # it shrank ...
    assert all(a >= b or a * 2 >= b for a, b in zip(halos, halos[1:]))
    assert sum(got[r][-1]["boosted"] for r in range(3)) > 0, halos  # ... and some chains needed their own boost


# ---------------------------------------------------------------------------------------------------------
# EM loop (core.EMLoop): the M-step of the next iteration is enqueued ahead of the seam verdict and rolled back
# when chains are re-run; time-sharded ranks broadcast the M-step result.  With deterministic stand-ins the
# speculative loop must reproduce the plain loop bit for bit.
# ---------------------------------------------------------------------------------------------------------
def _em_worker(rank, world, port, cfg, speculate, q, device_repair=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PMG_DEVICE_REPAIR"] = str(device_repair)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emu = _patch()
        import poor_man_gplvm_b200 as pmg
        from poor_man_gplvm_b200 import ops
        from poor_man_gplvm_b200.core import EMLoop
        from poor_man_gplvm_b200.shard import TimeShard
        ops.EmissionOperands = emu.FakeEmissionTC
        ops.backward, ops.atb_f16, ops.split_f16, ops.mstep_adam = (emu.backward_with_pieces, emu.atb_f16,
                                                                    emu.split_f16, emu.mstep_adam)
        T_total, N, K, halo, chunk, n_iter, seed = cfg
        d, P, M, host = _problem(T_total, N, K, seed)
        per = T_total // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else T_total
        cpu = torch.device("cpu")
        model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=6.0, device=cpu)
        rng = np.random.default_rng(11)
        model.params = rng.standard_normal((model.n_basis, N)).astype(np.float32)
        post0 = rng.random((T_total, K)) + 0.05
        lp0 = np.log(post0 / post0.sum(axis=1, keepdims=True)).astype(np.float32)
        op = ops.MoveOperator(host, M, cpu, P0=P[0])
        ma_n, ma_l = model._masks(None, None, hi - lo)
        loop = EMLoop(model, torch.from_numpy(d["y"][lo:hi].copy()), op, ma_n, ma_l, 1.0, model.tuning_basis,
                      lp0[lo:hi], 1.0, 0.01, 15, -1.0, halo=halo, chunk_len=chunk,
                      shard=TimeShard() if world > 1 else None)
        assert loop.use_tc
        out = []
        for i in range(n_iter):
            res, m_res = loop.iteration(speculate=(speculate and i < n_iter - 1))
            out.append({"tuning": m_res[4].numpy().copy(), "W_iter": loop.W_iter.numpy().copy(),
                        "n_adam": int(m_res[2].item()), "lm": float(res.log_marginal),
                        "repaired": bool(res.repaired), "relay": (res.n_relay_fwd, res.n_relay_bwd)})
        out.append({"W": loop.W.numpy().copy(), "count": int(loop.state.count.item()),
                    "Phi": loop.Phi.numpy().copy()})
        q.put((rank, out))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if world > 1:
            dist.destroy_process_group()


def _run_em(world, cfg, speculate, device_repair=0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_em_worker, args=(r, world, port, cfg, speculate, q, device_repair))
             for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert isinstance(res[r], list), res[r]
    return res


@pytest.mark.parametrize("world,device_repair", [(1, 0), (3, 0), (3, 1)])
def test_em_loop_speculative_mstep_equals_plain_loop(world, device_repair):
    """device_repair=0: every repair is a host sweep, so the scenario contains rolled-back speculative M-steps."""
    cfg = (288, 12, 24, 16, 24, 5, 5)
    plain = _run_em(world, cfg, speculate=False, device_repair=device_repair)
    spec = _run_em(world, cfg, speculate=True, device_repair=device_repair)
    n_iter = cfg[5]
    for r in range(world):
        for i in range(n_iter):
            a, b = plain[r][i], spec[r][i]
            assert np.array_equal(a["tuning"], b["tuning"]), (r, i)
            assert np.array_equal(a["W_iter"], b["W_iter"]), (r, i)
            assert a["n_adam"] == b["n_adam"] == 15 and a["lm"] == b["lm"]
            assert a["repaired"] == b["repaired"] and a["relay"] == b["relay"]
            # the reported weights are the ones that produced the reported tuning
            z = plain[r][-1]["Phi"].astype(np.float64) @ b["W_iter"].astype(np.float64)
            sp = np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))
            assert np.max(np.abs(sp - b["tuning"]) / sp) < 1e-5
        assert np.array_equal(plain[r][-1]["W"], spec[r][-1]["W"])
        assert plain[r][-1]["count"] == spec[r][-1]["count"] == 14 * n_iter
        # replicated M-step: identical on every rank
        assert np.array_equal(spec[r][n_iter - 1]["tuning"], spec[0][n_iter - 1]["tuning"])
    # the scenario must contain at least one repaired iteration (rollback path) and one clean one
    rep = [spec[0][i]["repaired"] for i in range(n_iter)]
    assert any(rep) or device_repair


def _fit_worker(q):
    try:
        os.environ["PMG_HALO"] = "16"           # read when the package is imported
        emu = _patch()
        import poor_man_gplvm_b200 as pmg
        from poor_man_gplvm_b200 import ops
        from oracle import ref_numpy as ref
        from poor_man_gplvm_b200.synthetic import make_dataset
        ops.EmissionOperands = emu.FakeEmissionTC
        ops.backward, ops.atb_f16, ops.split_f16, ops.mstep_adam = (emu.backward_with_pieces, emu.atb_f16,
                                                                    emu.split_f16, emu.mstep_adam)
        N, K, T = 10, 24, 160
        d = make_dataset(T, N, K, seed=9)
        cpu = torch.device("cpu")
        model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=6.0, device=cpu)
        rng = np.random.default_rng(2)
        params = rng.standard_normal((model.n_basis, N)).astype(np.float32)
        model.params = params.copy()
        oracle = ref.OraclePoissonGPLVMJump1D(N, K, tuning_lengthscale=6.0, dtype=np.float64,
                                              tuning_basis=model.tuning_basis, params=params)
        post0 = rng.random((T, K)) + 0.05
        lp0 = np.log(post0 / post0.sum(axis=1, keepdims=True)).astype(np.float32)
        kw = dict(n_iter=4, log_posterior_init=lp0, m_step_maxiter=12, m_step_tol=-1, save_every=2)
        want = oracle.fit_em(d["y"], **kw)
        got = model.fit_em(d["y"], **kw)
        out = {"iter_saved": got["iter_saved"],
               "params_saved": [np.asarray(a) for a in got["params_saved"]],
               "tuning_saved": [np.asarray(a) for a in got["tuning_saved"]],
               "params": np.asarray(got["params"]), "tuning": np.asarray(got["tuning"]),
               "post": np.asarray(got["posterior_latent_marg"]), "dyn": np.asarray(got["posterior_dynamics_marg"]),
               "lml": np.array(got["log_marginal_l"], dtype=np.float64),
               "n_adam": list(got["m_step_res_l"]["n_iter"]),
               "w_params_saved": want["params_saved"], "w_tuning_saved": want["tuning_saved"],
               "w_params": want["params"], "w_post": want["posterior_latent_marg"],
               "w_dyn": want["posterior_dynamics_marg"], "w_lml": np.array(want["log_marginal_l"]),
               "y": d["y"], "basis": model.tuning_basis}
        q.put(out)
    except Exception:  # pragma: no cover
        import traceback
        q.put(traceback.format_exc())


def test_fit_em_control_flow_on_cpu_matches_oracle():
    """fit_em end to end (snapshots, speculative M-steps, last iteration with full outputs, result dictionary)
    with the CPU stand-ins, against the oracle's fit_em."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_fit_worker, args=(q,))
    p.start()
    out = q.get(timeout=900)
    p.join(timeout=60)
    assert isinstance(out, dict), out
    assert out["iter_saved"] == [0, 2] and out["n_adam"] == [12] * 4
    for j in range(2):
        assert np.max(np.abs(out["params_saved"][j] - out["w_params_saved"][j])) < 1e-4
        assert np.max(np.abs(out["tuning_saved"][j] - out["w_tuning_saved"][j]) / out["w_tuning_saved"][j]) < 1e-4
        z = out["basis"].astype(np.float64) @ out["params_saved"][j].astype(np.float64)
        sp = np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))
        assert np.max(np.abs(sp - out["tuning_saved"][j]) / sp) < 1e-5
    assert np.max(np.abs(out["params"] - out["w_params"])) < 1e-4
    assert np.max(np.abs(out["post"] - out["w_post"])) < 2e-5
    assert np.max(np.abs(out["dyn"] - out["w_dyn"])) < 2e-5
    # FakeEmission drops the lgamma row term: constant shift of every log marginal
    from scipy.special import gammaln
    shift = gammaln(out["y"].astype(np.float64) + 1).sum()
    assert np.max(np.abs(out["lml"] - shift - out["w_lml"]) / np.abs(out["w_lml"])) < 1e-5


def _fit_sharded_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        os.environ["PMG_HALO"] = "16"           # read when the package is imported
        emu = _patch()
        import poor_man_gplvm_b200 as pmg
        from poor_man_gplvm_b200 import ops
        from poor_man_gplvm_b200.synthetic import make_dataset
        ops.EmissionOperands = emu.FakeEmissionTC
        ops.backward, ops.atb_f16, ops.split_f16, ops.mstep_adam = (emu.backward_with_pieces, emu.atb_f16,
                                                                    emu.split_f16, emu.mstep_adam)
        N, K, T = 10, 24, 192
        d = make_dataset(T, N, K, seed=4)
        model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=6.0, device=torch.device("cpu"))
        rng = np.random.default_rng(6)
        model.params = rng.standard_normal((model.n_basis, N)).astype(np.float32)
        post0 = rng.random((T, K)) + 0.05
        lp0 = np.log(post0 / post0.sum(axis=1, keepdims=True)).astype(np.float32)
        per = T // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else T
        got = model.fit_em(d["y"][lo:hi], n_iter=3, log_posterior_init=lp0[lo:hi], m_step_maxiter=10, m_step_tol=-1,
                           time_sharded=world > 1)
        q.put((rank, {"tuning": np.asarray(got["tuning"]), "params": np.asarray(got["params"]),
                      "post": np.asarray(got["posterior_latent_marg"]), "dyn": np.asarray(got["posterior_dynamics_marg"]),
                      "lml": np.array(got["log_marginal_l"], dtype=np.float64)}))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if world > 1:
            dist.destroy_process_group()


def _run_fit(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert isinstance(res[r], dict), res[r]
    return res


def test_time_sharded_fit_em_on_cpu_matches_single_process():
    """fit_em(time_sharded=True) under gloo with 3 ranks (blocks of one recording) against the single-process fit:
    same tuning and log marginals on every rank, posteriors of the blocks concatenate to the single-process ones."""
    one = _run_fit(1)[0]
    three = _run_fit(3)
    for r in range(3):
        assert np.max(np.abs(three[r]["tuning"] - one["tuning"]) / one["tuning"]) < 1e-4
        assert np.max(np.abs(three[r]["lml"] - one["lml"]) / np.abs(one["lml"])) < 1e-6
        assert np.array_equal(three[r]["tuning"], three[0]["tuning"])
        assert np.array_equal(three[r]["params"], three[0]["params"])
    post = np.concatenate([three[r]["post"] for r in range(3)])
    dyn = np.concatenate([three[r]["dyn"] for r in range(3)])
    assert np.max(np.abs(post - one["post"])) < 2e-5
    assert np.max(np.abs(dyn - one["dyn"])) < 2e-5


def _dense_plan_worker(q):
    """EStep with a lockstep-scan operand attached: chain plan sized to one wave of accumulator tiles, and the longest
    warm-up of each pass handed to the (stand-in) scan operators as ``halo_max``."""
    try:
        emu = _patch()
        from poor_man_gplvm_b200 import ops
        from poor_man_gplvm_b200.estep import EStep
        seen = []
        fwd0, bwd0 = emu.forward, emu.backward

        def fwd(*a, **k):
            seen.append(("f", k.get("mode", 0), k.get("halo_max")))
            return fwd0(*a, **k)

        def bwd(*a, **k):
            seen.append(("b", k.get("mode", 0), k.get("halo_max")))
            return bwd0(*a, **k)

        ops.forward, ops.backward = fwd, bwd
        T, N, K = 4000, 8, 320
        d, P, M, host = _problem(T, N, K, 3)
        op = ops.MoveOperator(host, M, torch.device("cpu"), P0=P[0], dense_tc=True)
        assert op.dense is not None and op.dense.n_ntiles == 2 and op.dense.chains(148) == 128 * 74
        y = torch.from_numpy(d["y"].copy())
        es = EStep(y, op, None, None, 1.0, halo=32, adaptive=True)
        # 2 "SMs" (patched): one wave = 128 chains; 4000 bins / 128 chains < the 64-bin minimum chunk
        plan = (es.S, es.chunk_len)
        tun = torch.from_numpy(d["tuning_true"].astype(np.float32))
        out = []
        for _ in range(3):
            res = es.run(tun, want_gamma=True, want_gamma_lat=True)
            out.append((res.gamma.numpy().copy(), res.halo))
        q.put(("ok", plan, seen, [o[1] for o in out], out[-1][0]))
    except Exception:  # pragma: no cover
        import traceback
        q.put(("err", traceback.format_exc()))


def test_lockstep_scan_plan_and_halo_max_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_dense_plan_worker, args=(q,))
    p.start()
    res = q.get(timeout=600)
    p.join(timeout=60)
    assert res[0] == "ok", res[1]
    _, (S, chunk), seen, halos, gamma = res
    assert chunk == 64 and S == (4000 + 63) // 64
    # every mode-0 pass carries the longest per-chain warm-up of that pass (adaptive: base <= 32, boosts <= 32)
    assert all(hm is not None and 0 < hm <= 32 for kind, mode, hm in seen if mode == 0), seen
    assert {k for k, m, _ in seen if m == 0} == {"f", "b"}
    from oracle import linear_ref as lin
    d, P, M, host = _problem(4000, 8, 320, 3)
    want = lin.e_step(d["y"], d["tuning_true"].astype(np.float64), P.astype(np.float64), M.astype(np.float64),
                      np.ones(8), np.ones(320))
    assert np.max(np.abs(gamma - want["gamma"])) < 2e-5
