"""Seam diagnostics of the time-parallel E-step (round 2).

  python scripts/diag_seams.py halo   [T]          # seam-error distribution vs halo on a fitted model (1 process)
  python scripts/diag_seams.py pooled [T] [NBLK]   # one process fits NBLK bench blocks concatenated: relays?
  torchrun --nproc-per-node R scripts/diag_seams.py ranks [T]   # R ranks (BACKEND=gloo: all on cuda:0)

Prints, per EM iteration, which seams were over the tolerance at the FIRST verdict (before any repair).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200 import ops
from poor_man_gplvm_b200.core import EMLoop
from poor_man_gplvm_b200.estep import EStep
from poor_man_gplvm_b200.shard import TimeShard
from poor_man_gplvm_b200.synthetic import make_dataset_torch

N, K = int(os.environ.get("N", 500)), int(os.environ.get("K", 400))
ITERS = int(os.environ.get("ITERS", 16))


def block(T, dev, r):
    y = make_dataset_torch(T, N, K, dev, seed=1234 + r, tuning_seed=1234)["y"].to(torch.float32).contiguous()
    g = torch.Generator(device=dev); g.manual_seed(99 + 7919 * r)
    post0 = torch.rand((T, K), generator=g, device=dev)
    lp0 = torch.log(post0 / post0.sum(dim=1, keepdim=True))
    return y, lp0


def make_loop(y, lp0, dev, shard=None, halo=None, chunk_len=None):
    model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, movement_variance=1.0, device=dev)
    model.params = np.random.default_rng(1).standard_normal((model.n_basis, N)).astype(np.float32)
    P, logP, M, logM, op = model._transition_pack({})
    ma_n, ma_l = model._masks(None, None, y.shape[0])
    loop = EMLoop(model, y, op, ma_n, ma_l, 1.0, model.tuning_basis, lp0, model.param_prior_std, 0.01, 1000, 1e-6,
                  shard=shard, halo=halo, chunk_len=chunk_len)
    return model, op, loop


def dissect(tag, est, truth):
    """where does the relative error of a failing seam live?"""
    a = est.reshape(-1).double(); b = truth.reshape(-1).double()
    a = a / a.sum(); b = b / b.sum()
    mx = float(b.max())
    rel = (a - b).abs() / torch.minimum(a, b).clamp_min(1e-300)
    rel = torch.where(torch.maximum(a, b) > 1e-12, rel, torch.zeros_like(rel))
    j = int(rel.argmax())
    out = ["%s max rel %.2e at j=%d (est %.2e truth %.2e, max entry %.2e) | L1 abs %.2e max abs %.2e" %
           (tag, float(rel.max()), j, float(a[j]), float(b[j]), mx, float((a - b).abs().sum()), float((a - b).abs().max()))]
    for fl in (1e-10, 1e-8, 1e-6, 2.5e-5, 1e-4, 1e-3):
        big = torch.maximum(a, b) > fl
        r1 = float(rel[big].max()) if big.any() else 0.0
        r2 = float(((a - b).abs() / torch.maximum(b, torch.tensor(fl, dtype=b.dtype, device=b.device))).max())
        out.append("floor %.1e: rel(entries>floor) %.2e | abs/max(truth,floor) %.2e | n>floor %d" % (fl, r1, r2, int(big.sum())))
    K2 = a.numel() // 2
    out.append("dyn mass est (%.6f, %.6f) truth (%.6f, %.6f)" % (float(a[:K2].sum()), float(a[K2:].sum()),
                                                                  float(b[:K2].sum()), float(b[K2:].sum())))
    print("\n    ".join(out), flush=True)


DISSECT_FROM = int(os.environ.get("DISSECT_FROM", "-1"))


def tap_first_verdict(es, store):
    """record the seam errors of the first verdict of every E-step"""
    orig = es._verdict
    state = {"first": True, "n": 0}

    def verdict(before_sync=None):
        out = orig(before_sync)
        if state["first"]:
            err = out[0]
            ef, eb = err[es.f_lo:es.S].clone(), err[es.S:es.S + es.b_hi].clone()
            store.append((ef, eb))
            state["first"] = False
            state["n"] += 1
            if 0 <= DISSECT_FROM < state["n"]:
                rk = es.shard.rank
                for i in torch.nonzero(ef > es.seam_tol).flatten()[:3]:
                    c = int(i) + es.f_lo
                    dissect("[r%d it%d FWD chain %d]" % (rk, state["n"] - 1, c), es.halo_state[c], es.fwd_end_ext[c])
                for i in torch.nonzero(eb > es.seam_tol).flatten()[:3]:
                    c = int(i)
                    dissect("[r%d it%d BWD chain %d]" % (rk, state["n"] - 1, c), es.beta_halo[c], es.beta_end[c + 1])
        return out
    es._verdict = verdict
    return state


def describe(e, tol):
    if e.numel() == 0:
        return "-"
    q = torch.quantile(e.double(), torch.tensor([0.5, 0.9, 0.99, 1.0], dtype=torch.float64))
    bad = torch.nonzero(e > tol).flatten()
    return "med %.1e p90 %.1e p99 %.1e max %.1e | >tol %d %s" % (q[0], q[1], q[2], q[3], bad.numel(),
                                                                [(int(i), "%.1e" % float(e[i])) for i in bad[:5]])


def run_loop(loop, rank, label, iters=ITERS):
    es = loop.es
    store = []
    st = tap_first_verdict(es, store)
    for i in range(iters):
        st["first"] = True
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res, m = loop.iteration(speculate=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        ef, eb = store[-1]
        print("[%s r%d] it %2d %.1f ms adam %d devfix f %d b %d hostrelay f %d b %d S=%d chunk=%d halo=%d | F %s | B %s" %
              (label, rank, i, dt, int(m[2].item()), res.n_fix_fwd, res.n_fix_bwd, res.n_relay_fwd, res.n_relay_bwd,
               es.S, es.chunk_len, res.halo,
               describe(ef, es.seam_tol), describe(eb, es.seam_tol)), flush=True)
    return store


def mode_halo(T):
    dev = torch.device("cuda", 0)
    y, lp0 = block(T, dev, 0)
    model, op, loop = make_loop(y, lp0, dev)
    del lp0
    for i in range(ITERS):
        res, m = loop.iteration(speculate=True)
    tuning = m[4].clone()
    print("fitted %d iterations at T=%d: default plan S=%d chunk=%d halo=%d" % (ITERS, T, loop.es.S, loop.es.chunk_len,
                                                                              loop.es.halo), flush=True)
    del loop
    torch.cuda.empty_cache()
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    for halo in (16, 32, 64, 128, 256):
        for chains_per_sm in (12,):
            chunk = max(32, -(-T // (sm * chains_per_sm)))
            es = EStep(y, op, None, None, 1.0, halo=halo, chunk_len=chunk, em_mode=True)
            g16 = ops.new_gamma16(es.T, K, dev)
            store = []
            st = tap_first_verdict(es, store)
            times = []
            for rep in range(3):            # pass 0 starts from the stationary message, passes 1-2 from the carried ones
                st["first"] = True
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = es.run(tuning, want_gamma_lat=False, gamma16=g16)
                torch.cuda.synchronize(); times.append((time.perf_counter() - t0) * 1e3)
                ef, eb = store[-1]
                print("halo %3d chunk %4d S %5d pass %d: %.2f ms devfix f %d b %d hostrelay f %d b %d | F %s | B %s" %
                      (halo, chunk, es.S, rep, times[-1], r.n_fix_fwd, r.n_fix_bwd, r.n_relay_fwd, r.n_relay_bwd,
                       describe(ef, es.seam_tol),
                       describe(eb, es.seam_tol)), flush=True)
            del es, g16
            torch.cuda.empty_cache()


def mode_pooled(T, nblk):
    dev = torch.device("cuda", 0)
    ys, lps = zip(*[block(T, dev, r) for r in range(nblk)])
    y = torch.cat(ys); lp0 = torch.cat(lps)
    del ys, lps
    model, op, loop = make_loop(y, lp0, dev)
    del lp0
    run_loop(loop, 0, "pooled%d" % nblk)


def mode_ranks(T):
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("BACKEND", "nccl")
    dev = torch.device("cuda", 0 if backend == "gloo" else lr)
    torch.cuda.set_device(dev)
    if backend == "gloo":
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    y, lp0 = block(T, dev, rank)
    model, op, loop = make_loop(y, lp0, dev, shard=TimeShard())
    del lp0
    run_loop(loop, rank, "%s%d" % (backend, world))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    mode = sys.argv[1]
    T = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1000000
    if mode == "halo":
        mode_halo(T)
    elif mode == "pooled":
        mode_pooled(T, int(sys.argv[3]) if len(sys.argv) > 3 else 8)
    else:
        mode_ranks(T)
