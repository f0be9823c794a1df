"""E-step scan with a dense K x K move kernel: lockstep tensor-core path vs the CUDA-core general path (same inputs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poor_man_gplvm_b200 import ops, gp_kernel as gpk
from poor_man_gplvm_b200.estep import EStep
K, T = int(os.environ.get("K", 2000)), int(os.environ.get("T", 20000))
dev = torch.device("cuda")
x = np.arange(K, dtype=np.float64)
ck = (np.exp(-np.abs(x[:, None] - x[None, :]) / 150.0) + 0.02).astype(np.float32)
P, logP, M, logM = gpk.create_transition_prob_1d(np.arange(K), np.arange(2), 1.0, 0.02, 0.05, custom_kernel=ck)
hostop = gpk.move_operator_host(K, 1.0, ck, p_move_to_jump=0.02)
g = torch.Generator(device=dev); g.manual_seed(0)
ll = (torch.randn((T, K), generator=g, device=dev) * 3.0 - 40.0).contiguous()
out = {}
for tc in (True, False):
    op = ops.MoveOperator(hostop, M, dev, P0=P[0], dense_tc=tc)
    es = EStep(torch.zeros((T, 1), device=dev), op, None, None, 1.0)
    es.ll.copy_(ll)
    es.emission = lambda tuning, es=es: es.ll
    g16 = ops.new_gamma16(T, K, dev)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = es.run(None, want_gamma=False, want_gamma_lat=True, gamma16=g16)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("tensor_cores=%s rep %d: %.1f ms for T=%d K=%d (%d chains x %d bins, halo %d) lml %.6e relays %d/%d"
              % (tc, rep, dt * 1e3, T, K, es.plan.n_chain, es.chunk_len, es.halo, float(res.log_marginal),
                 res.n_relay_fwd, res.n_relay_bwd), flush=True)
    out[tc] = res.gamma_lat.clone()
    del es, op
print("max |gamma_lat difference| between the paths: %.3g" % float((out[True] - out[False]).abs().max()))
