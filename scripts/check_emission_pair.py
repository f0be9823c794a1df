"""Pair (cta_group::2) emission kernel against the single-CTA kernel: bit-identical ll, and timing.
usage: check_emission_pair.py run TAG   (PMG_EM_PAIR selects the kernel; writes gpurun_out/ll_TAG_*.npy)
       check_emission_pair.py cmp A B"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
SHAPES = [(1000000, 500, 400), (100003, 501, 400), (70001, 200, 104), (5000, 30, 96), (513, 64, 208), (250000, 1000, 200)]
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def run(tag):
    import torch
    from poor_man_gplvm_b200 import ops
    dev = torch.device("cuda")
    for i, (T, N, K) in enumerate(SHAPES):
        g = torch.Generator(device=dev); g.manual_seed(i)
        y = torch.poisson(torch.rand((T, N), generator=g, device=dev) * 1.5, generator=g).contiguous()
        tun = torch.rand((K, N), generator=g, device=dev) + 0.05
        ma_l = torch.ones(K, device=dev); ma_l[3] = 0
        em = ops.EmissionOperands(y, None, ones_col=True)
        ll = torch.full((T, K), 7.0, device=dev)
        em.loglik(tun, ma_l, 1.0, out=ll)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); em.loglik(tun, ma_l, 1.0, out=ll); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        step = max(1, T // 4000)
        rows = torch.cat([ll[::step], ll[-600:]]).cpu().numpy()
        np.save(os.path.join(OUT, "ll_%s_%d.npy" % (tag, i)), rows)
        # fp64 reference on a few rows
        idx = torch.tensor([0, 1, T // 2, T - 1], device=dev)
        ref = (y[idx].double() @ torch.log(tun.double() + 1e-20).T - (tun.double() + 1e-20).sum(1)[None]
               - torch.lgamma(y[idx].double() + 1).sum(1, keepdim=True))
        err = ((ll[idx].double() - ref).abs() / ref.abs().clamp_min(1))[:, ma_l.bool()].max().item()
        print("%s shape %s: %.3f ms (min of 5; incl. prepare), rel err vs fp64 %.2e, masked col ok %s, sum %.6e" %
              (tag, (T, N, K), min(ts), err, bool((ll[:, 3] == -1e20).all()), float(ll[:, ma_l.bool()].double().sum())), flush=True)
        del y, ll, em


def cmp(a, b):
    ok = True
    for i in range(len(SHAPES)):
        x = np.load(os.path.join(OUT, "ll_%s_%d.npy" % (a, i))); y = np.load(os.path.join(OUT, "ll_%s_%d.npy" % (b, i)))
        same = np.array_equal(x, y)
        ok &= same
        print("shape %s: identical %s (max abs diff %.3e)" % (SHAPES[i], same, float(np.max(np.abs(x - y)))))
    print("ALL IDENTICAL" if ok else "MISMATCH")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        cmp(sys.argv[2], sys.argv[3])
