"""Per-iteration seam statistics of the time-sharded EM loop (torchrun --nproc-per-node 2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import poor_man_gplvm_b200 as pmg
from poor_man_gplvm_b200.core import EMLoop
from poor_man_gplvm_b200.shard import TimeShard
from poor_man_gplvm_b200.synthetic import make_dataset_torch
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
T, N, K = int(os.environ.get("T", 1000000)), 500, 400
y = make_dataset_torch(T, N, K, dev, seed=int(os.environ.get("SEED", 1234)) + rank, tuning_seed=1234)["y"].to(torch.float32).contiguous()
model = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=10.0, movement_variance=1.0, device=dev)
model.params = np.random.default_rng(1).standard_normal((model.n_basis, N)).astype(np.float32)
P, logP, M, logM, op = model._transition_pack({})
ma_n, ma_l = model._masks(None, None, T)
g = torch.Generator(device=dev); g.manual_seed(99 + 7919 * rank)
post0 = torch.rand((T, K), generator=g, device=dev)
lp0 = torch.log(post0 / post0.sum(dim=1, keepdim=True)); del post0
loop = EMLoop(model, y, op, ma_n, ma_l, 1.0, model.tuning_basis, lp0, model.param_prior_std, 0.01, 1000, 1e-6, shard=TimeShard())
es = loop.es
_orig_check = es._check_fwd
state = {"n": 0}
def check_fwd(compact=False):
    _orig_check(compact)
    if rank == 1 and state["n"] < 40 and os.environ.get("DBG"):
        state["n"] += 1
        torch.cuda.synchronize()
        a = es.halo_state[0].reshape(-1).double(); b = es.truth[0].reshape(-1).double()
        a = a / a.sum(); b = b / b.sum()
        rel = (a - b).abs() / torch.maximum(torch.minimum(a, b), torch.tensor(1e-37, device=a.device, dtype=a.dtype))
        rel = torch.where(torch.maximum(a, b) > 1e-20, rel, torch.zeros_like(rel))
        j = int(rel.argmax())
        big = (torch.maximum(a, b) > 1e-6)
        print("  [rank1 chain0 check %d] err %.3e at j=%d (a=%.3e b=%.3e) | max rel err over entries>1e-6: %.3e | err[0]=%.3e sum_a1=%.4f sum_b1=%.4f" %
              (state["n"], float(rel.max()), j, float(a[j]), float(b[j]), float(rel[big].max()) if big.any() else 0.0, float(es.err[0]),
               float(a[es.K:].sum()), float(b[es.K:].sum())), flush=True)
es._check_fwd = check_fwd
from poor_man_gplvm_b200 import ops as _ops
if os.environ.get("DUMP_SEAMS"):
    # first verdict of every E-step (before any repair): which seams are over the tolerance, and by how much
    _orig_verdict = es._read_err_global
    _calls = {"n": 0}
    def _verdict():
        out = _orig_verdict()
        err, any_f, any_b = out
        ef = err[es.f_lo:es.S].clone(); eb = err[es.S:es.S + es.b_hi].clone()
        bf = torch.nonzero(ef > es.seam_tol).flatten(); bb = torch.nonzero(eb > es.seam_tol).flatten()
        _calls["n"] += 1
        if bf.numel() or bb.numel():
            print("  [rank %d verdict %d] fwd over tol: %s | bwd over tol: %s | S=%d" %
                  (rank, _calls["n"], [(int(i) + es.f_lo, "%.2e" % float(ef[i])) for i in bf[:6]],
                   [(int(i), "%.2e" % float(eb[i])) for i in bb[:6]], es.S), flush=True)
        return out
    es._read_err_global = _verdict
marks = []
if os.environ.get("HOSTMARKS"):
    _ops.PHASE_HOOK = lambda name: marks.append((name, time.perf_counter()))
    _o = es._read_err_global
    def _r():
        marks.append(("enqueued", time.perf_counter())); out = _o(); marks.append(("verdict", time.perf_counter())); return out
    es._read_err_global = _r
    for k in range(12):
        loop.iteration(speculate=True)
    torch.cuda.synchronize(); dist.barrier(); marks.clear()
    t_begin = time.perf_counter()
    for k in range(8):
        marks.append(("begin", time.perf_counter())); loop.iteration(speculate=True)
    torch.cuda.synchronize(); t_end = time.perf_counter()
    if rank == 0:
        print("free-running: %.3f ms/iter" % ((t_end - t_begin) * 1e3 / 8))
        acc = {}
        for (n0, t0), (n1, t1) in zip(marks[:-1], marks[1:]):
            acc[n1] = acc.get(n1, 0.0) + (t1 - t0) * 1e3 / 8
        print("host ms per iteration by segment (time up to the mark):", {k: round(v, 3) for k, v in acc.items()})
    dist.destroy_process_group(); sys.exit(0)
for i in range(int(os.environ.get("ITERS", 10))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, m = loop.iteration()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    err = es.err_host.clone()
    ef, eb = err[es.f_lo:es.S], err[es.S:es.S + es.b_hi]
    print("rank %d it %d: %.1f ms relay f %d b %d | final seam err f max %.2e (arg %d) b max %.2e (arg %d) | S=%d f_lo=%d b_hi=%d" %
          (rank, i, dt, res.n_relay_fwd, res.n_relay_bwd, float(ef.max()), int(ef.argmax()), float(eb.max()), int(eb.argmax()), es.S, es.f_lo, es.b_hi), flush=True)
dist.destroy_process_group()
