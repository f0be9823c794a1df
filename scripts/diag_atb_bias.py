"""Does the tensor-core accumulator's truncation bias show in the statistics GEMM (T-long reductions)?
atb_f16 (tcgen05, fp16 pieces, split over time) vs an fp64 GEMM at the headline shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poor_man_gplvm_b200 import ops
dev = torch.device("cuda")
for T in (20000, 200000, 1000000):
    K, N = 400, 500
    g = torch.Generator(device=dev); g.manual_seed(T)
    G = torch.rand((T, K), generator=g, device=dev) ** 6
    G = (G / G.sum(dim=1, keepdim=True)).contiguous()
    Y = torch.poisson(torch.full((T, N), 0.7, device=dev), generator=g).contiguous()
    y16 = ops.CountsF16(Y, ones_col=True)
    g16 = ops.split_f16(G)
    got = ops.atb_f16(g16, y16, K).double()
    want = torch.zeros((K, N + 1), dtype=torch.float64, device=dev)
    for s in range(0, T, 100000):
        Ys = torch.cat([Y[s:s + 100000], torch.ones((min(100000, T - s), 1), device=dev)], dim=1).double()
        want += G[s:s + 100000].double().T @ Ys
    rel = (got[:, :N + 1] - want) / want
    print("T=%d: statistics GEMM relative error: mean %.3g, min %.3g, max %.3g; ones column (sum_t gamma): mean %.3g"
          % (T, rel.mean().item(), rel.min().item(), rel.max().item(), rel[:, N].mean().item()), flush=True)
    del G, Y, y16, g16, got, want
