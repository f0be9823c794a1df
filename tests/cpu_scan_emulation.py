"""CPU stand-ins for the scan operators, used ONLY to exercise the host-side orchestration of the E-step
(`poor_man_gplvm_b200.estep.EStep`: chain plan, warm starts, seam checks, Jacobi repair sweeps, rank-boundary
exchanges) without a GPU, under gloo with several ranks.

They follow the interface conventions of the CUDA kernels they replace (`csrc/pmg_scan.cu`: `fwd_kernel`,
`bwd_kernel`, `seam_check_kernel`) -- where a chain's warm-up starts, which message goes to `halo_state`,
`beta_halo`, `beta_end`, `warm_out`, what mode 1 restarts from -- in plain float64 NumPy.  Test infrastructure:
the product path never imports this module.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch


def _np(t):
    return t.numpy() if isinstance(t, torch.Tensor) else t


def _dense_move(op):
    """P0[x,x'] = taps[|x-x'|] * inv_z[x] (Toeplitz numerator, row normaliser)."""
    K = op.K
    taps = _np(op.taps).astype(np.float64)
    inv_z = _np(op.inv_z).astype(np.float64)
    d = np.abs(np.arange(K)[:, None] - np.arange(K)[None, :])
    G = np.where(d < taps.shape[0], taps[np.minimum(d, taps.shape[0] - 1)], 0.0)
    return G * inv_z[:, None]


def _chains(plan, mode, chain_ids, sel_err=None, sel_tol=0.0):
    if mode == 1:
        ids = [int(i) for i in _np(chain_ids).reshape(-1)]
    elif mode == 2:           # selection "on the device": chain s iff !(sel_err[s] <= sel_tol)
        e = _np(sel_err)
        ids = [s for s in range(plan.n_chain) if not (e[s] <= sel_tol)]
    else:
        ids = range(plan.n_chain)
    for s in ids:
        t_begin = plan.core_begin + s * plan.chunk_len
        t_end = min(t_begin + plan.chunk_len, plan.core_end)
        if 0 <= s < plan.n_chain and t_begin < t_end:
            yield s, t_begin, t_end


def _slot(buf, s):
    """warm_in may be one [2,K] vector shared by all chains or one slot per chain."""
    b = _np(buf)
    return b if b.ndim == 2 else b[s]


def _halo_next(plan, j=None):
    """warm-up length chain j uses in the next pass (per-chain array of the plan, else the scalar)"""
    if j is not None and plan.halo_next_arr and 0 <= j < plan.n_chain:
        return int(ctypes.c_int32.from_address(int(plan.halo_next_arr) + 4 * j).value)
    return int(plan.halo_next) if int(plan.halo_next) > 0 else int(plan.halo)


def _halo_own(plan, s):
    if plan.halo_arr:
        return int(ctypes.c_int32.from_address(int(plan.halo_arr) + 4 * s).value)
    return int(plan.halo)


def forward(plan, op, ll, alpha, lmr, halo_state=None, carry_in=None, mode=0, chain_ids=None, warm_in=None,
            warm_out=None, sel_err=None, sel_tol=0.0, halo_max=0):
    K = op.K
    P0 = _dense_move(op)
    M = op.M.reshape(2, 2).astype(np.float64)
    scale = float(plan.likelihood_scale)
    ll_, al_, lmr_ = _np(ll), _np(alpha), _np(lmr)
    for s, t_begin, t_end in _chains(plan, mode, chain_ids, sel_err, sel_tol):
        msg = None
        if mode != 0:
            t0 = t_begin
            if warm_in is not None:
                msg = _slot(warm_in, s)
            elif t0 > 0:
                msg = al_[t0 - 1]
            elif carry_in is not None:
                msg = _np(carry_in)
        else:
            t0 = t_begin - _halo_own(plan, s)
            if t0 <= 0 and plan.left_exact:
                t0 = 0
                if carry_in is not None:
                    msg = _np(carry_in)
            else:
                t0 = max(t0, 0)
                if warm_in is not None:
                    msg = _slot(warm_in, s)
        m = np.full((2, K), 0.5 / K) if msg is None else np.array(msg, dtype=np.float64).reshape(2, K)
        tot = m.sum()
        m = m / tot if (tot > 0 and np.isfinite(tot)) else np.full((2, K), 0.5 / K)
        for t in range(t0, t_end):
            a0 = M[0, 0] * m[0] + M[1, 0] * m[1]
            a1 = M[0, 1] * m[0] + M[1, 1] * m[1]
            row = ll_[t].astype(np.float64)
            mx = row.max()
            L = np.exp(scale * (row - mx))
            u = np.stack([(a0 @ P0) * L, (a1.sum() / K) * L])
            c = u.sum()
            m = u / c
            if t >= t_begin:
                al_[t] = m.astype(np.float32)
                lmr_[t] = np.float32(np.log(c) + scale * mx)
            elif t == t_begin - 1 and halo_state is not None:
                _np(halo_state)[s] = m.astype(np.float32)
            if (warm_out is not None and t == t_end - _halo_next(plan, s + 1) - 1
                    and (s + 1 < plan.n_chain or not plan.right_exact)):
                _np(warm_out)[s + 1] = m.astype(np.float32)


def backward(plan, op, ll, alpha, gamma=None, gamma_lat=None, dyn_marg=None, r_out=None, tw_partial=None,
             beta_halo=None, beta_end=None, beta_in=None, mode=0, chain_ids=None, gamma16=None, warm_in=None,
             warm_out=None, sel_err=None, sel_tol=0.0, halo_max=0, xi16=None):
    K, T = op.K, int(plan.T)
    P0 = _dense_move(op)
    M = op.M.reshape(2, 2).astype(np.float64)
    scale = float(plan.likelihood_scale)
    ll_, al_ = _np(ll), _np(alpha)
    for s, t_begin, t_end in _chains(plan, mode, chain_ids, sel_err, sel_tol):
        init = None
        if mode != 0:
            if t_end < T:
                t_hi = t_end
                init = _slot(warm_in, s) if warm_in is not None else _np(beta_end)[s + 1]
            else:
                t_hi, init = T - 1, (None if beta_in is None else _np(beta_in))
        else:
            t_hi = t_end - 1 + _halo_own(plan, s)
            if t_hi >= T - 1 and plan.right_exact:
                t_hi, init = T - 1, (None if beta_in is None else _np(beta_in))
            else:
                t_hi = min(t_hi, T - 1)
                if warm_in is not None:
                    init = _slot(warm_in, s)
        be = None
        Lb = None
        RL = 0.0
        tw = np.zeros(K)
        for t in range(t_hi, t_begin - 1, -1):
            row = ll_[t].astype(np.float64)
            L = np.exp(scale * (row - row.max()))
            if t == t_hi:
                b = np.ones((2, K)) if init is None else np.array(init, dtype=np.float64).reshape(2, K)
                r = np.zeros((2, K))
            else:
                r = np.stack([Lb * be[0], Lb * be[1]])
                w0 = P0 @ r[0]
                w1 = RL / K
                b = np.stack([M[0, 0] * w0 + M[0, 1] * w1, M[1, 0] * w0 + M[1, 1] * w1])
            if t <= t_end and t < T:
                a = al_[t].astype(np.float64)
                s0, s1 = (a[0] * b[0]).sum(), (a[1] * b[1]).sum()
            else:
                a = None
                s0, s1 = b[0].sum(), b[1].sum()
            inv = 1.0 / (s0 + s1)
            be, Lb, RL = b * inv, L, (L * b[1]).sum() * inv
            if t < t_end:
                g = a * be
                if gamma is not None:
                    _np(gamma)[t] = g.astype(np.float32)
                if gamma_lat is not None:
                    _np(gamma_lat)[t] = (g[0] + g[1]).astype(np.float32)
                tw += g[0] + g[1]
                if dyn_marg is not None:
                    _np(dyn_marg)[t] = np.array([s0 * inv, s1 * inv], dtype=np.float32)
                if r_out is not None and t != t_hi:
                    _np(r_out)[t + 1] = (r * inv).astype(np.float32)
            elif t == t_end and beta_halo is not None:
                _np(beta_halo)[s] = be.astype(np.float32)
            if t == t_begin and beta_end is not None:
                _np(beta_end)[s] = be.astype(np.float32)
            if warm_out is not None and t == t_begin + _halo_next(plan, s - 1) - 1 and (s >= 1 or not plan.left_exact):
                # slot s-1 of the view; s == 0 writes the slot in front of the view (the caller passes buf[1:])
                wo = warm_out
                if s >= 1:
                    _np(wo)[s - 1] = be.astype(np.float32)
                else:
                    base = wo.data_ptr() - 2 * K * 4
                    dst = np.ctypeslib.as_array((ctypes.c_float * (2 * K)).from_address(base))
                    dst[:] = be.astype(np.float32).reshape(-1)
        if tw_partial is not None:
            _np(tw_partial)[s] = tw.astype(np.float32)


def boundary_pack_fwd(K, out, first, last, warm_src=None, ax_row=None, ll_row=None, scale=1.0):
    """pmg_boundary_pack_fwd on host tensors: out [8K] = [first | 0 | last | warm]"""
    o = _np(out)
    K2 = 2 * K
    o[:] = 0.0
    if first is not None:
        o[:K2] = _np(first).reshape(-1)
    if last is not None:
        o[2 * K2:3 * K2] = _np(last).reshape(-1)
    if ax_row is not None:
        a, l = _np(ax_row), _np(ll_row).astype(np.float64)
        E = np.exp2((l - l.max()) * (scale * 1.4426950408889634))
        o[3 * K2:3 * K2 + K] = a[:K]
        o[3 * K2 + K:] = (a[K] * E).astype(np.float32)
    elif warm_src is not None:
        o[3 * K2:] = _np(warm_src).reshape(-1)


def boundary_unpack_fwd(K, from_left, from_right, fwd_end0, fwarm0, ax_stop=None, ll_stop=None, scale=1.0,
                        alpha_stop=None):
    K2 = 2 * K
    if from_left is not None:
        fl = _np(from_left)
        _np(fwd_end0).reshape(-1)[:] = fl[:K2]
        if fwarm0 is not None:
            _np(fwarm0).reshape(-1)[:] = fl[K2:2 * K2]
    if from_right is not None:
        fr = _np(from_right)
        if ax_stop is not None:
            l = _np(ll_stop).astype(np.float64)
            E = np.exp2((l - l.max()) * (scale * 1.4426950408889634))
            a = _np(ax_stop)
            a[:K] = fr[:K]
            a[K] = fr[K:K2].sum() / E.sum()
        else:
            _np(alpha_stop).reshape(-1)[:] = fr[:K2]


def seam_check(n, length, est_ptr, ld_est, truth_ptr, ld_truth, err, floor_val=1e-12):
    """Same contract as ops.seam_check: raw pointers + row strides (here: host memory)."""
    def rows(ptr, ld):
        flat = np.ctypeslib.as_array((ctypes.c_float * (int(n) * int(ld))).from_address(int(ptr)))
        return flat.reshape(int(n), int(ld))[:, :int(length)].astype(np.float64)
    a, b = rows(est_ptr, ld_est), rows(truth_ptr, ld_truth)
    out = _np(err)
    for i in range(int(n)):
        sa, sb = a[i].sum(), b[i].sum()
        if not (sa > 0 and sb > 0 and np.isfinite(sa) and np.isfinite(sb)):
            out[i] = np.inf
            continue
        u, v = a[i] / sa, b[i] / sb
        hi, lo = np.maximum(u, v), np.minimum(u, v)
        sel = hi > floor_val
        out[i] = np.float32(((hi - lo)[sel] / np.maximum(lo[sel], 1e-37)).max()) if sel.any() else 0.0


def seam_check_fix(n, length, est_ptr, ld_est, truth_ptr, ld_truth, err, tol, fix=False, counter=None,
                   floor_val=1e-12):
    """Same contract as ops.seam_check_fix (host memory): count seams over the tolerance, optionally replace their
    estimate by the truth."""
    seam_check(n, length, est_ptr, ld_est, truth_ptr, ld_truth, err, floor_val)
    out = _np(err)
    for i in range(int(n)):
        if not (out[i] <= tol):
            if counter is not None:
                _np(counter)[0] += 1.0
            if fix:
                dst = np.ctypeslib.as_array((ctypes.c_float * int(length)).from_address(int(est_ptr) + 4 * i * int(ld_est)))
                src = np.ctypeslib.as_array((ctypes.c_float * int(length)).from_address(int(truth_ptr) + 4 * i * int(ld_truth)))
                dst[:] = src


class FakeEmission:
    """Stands in for ops.EmissionOperands: ll = y log(lam)^T - sum(lam) (the lgamma row term is constant in k)."""
    mode = 0
    A16 = None

    def __init__(self, y, ma_neuron=None, impl=0, ones_col=False, dt_l=None):
        self.y = y

    def loglik(self, tuning, ma_latent=None, dt=1.0, out=None):
        lam = tuning.double() + 1e-20
        ll = self.y.double() @ torch.log(lam).T - lam.sum(dim=1)[None, :]
        out.copy_(ll.float())
        return out


# ---------------------------------------------------------------------------------------------------------
# stand-ins for the operators of the EM loop (core.EMLoop): fp16 counts, fp16 posterior pieces, statistics
# GEMM, Adam M-step (the oracle's restatement of the reference optimiser)
# ---------------------------------------------------------------------------------------------------------
class FakeCounts:
    """Stands in for ops.CountsF16 (counts are small integers: exact)."""
    exact = True

    def __init__(self, y, ones_col=False, **kw):
        self.T, self.N = y.shape
        self.ones_col = bool(ones_col)
        self.y = y


class FakeEmissionTC(FakeEmission):
    """Emission stand-in that also advertises the fp16 counts, so that EMLoop takes its tensor-core code path
    (fp16 posterior pieces written by the backward pass, statistics from the pieces, speculation allowed)."""

    def __init__(self, y, ma_neuron=None, impl=0, ones_col=False, dt_l=None):
        super().__init__(y)
        self.A16 = FakeCounts(y, ones_col=ones_col)


def backward_with_pieces(plan, op, ll, alpha, gamma=None, gamma_lat=None, gamma16=None, **kw):
    """backward() that also fills the fp16 hi/lo pieces of gamma_lat, as the CUDA kernels do."""
    T, K = int(plan.T), op.K
    gl = gamma_lat if gamma_lat is not None else torch.zeros((T, K), dtype=torch.float32)
    if gamma16 is not None and kw.get("mode", 0) != 0:
        gl.copy_(gamma16[0, :, :K].float() + gamma16[1, :, :K].float())     # rows of chains that are not re-run
    backward(plan, op, ll, alpha, gamma=gamma, gamma_lat=gl, gamma16=None, **kw)
    if gamma16 is not None:
        for s, t_begin, t_end in _chains(plan, kw.get("mode", 0), kw.get("chain_ids"), kw.get("sel_err"),
                                         kw.get("sel_tol", 0.0)):
            g = gl[t_begin:t_end]
            hi = g.half()
            gamma16[0, t_begin:t_end, :K] = hi
            gamma16[1, t_begin:t_end, :K] = (g - hi.float()).half()


def split_f16(src, out=None):
    T, K = src.shape
    if out is None:
        out = torch.zeros((2, T, (K + 7) // 8 * 8), dtype=torch.float16)
    hi = src.half()
    out[0, :, :K] = hi
    out[1, :, :K] = (src - hi.float()).half()
    return out


def atb_f16(g16, y16, K, out=None):
    """[K, N (+1)] = sum_t gamma_lat[t,:K]^T [y | 1]."""
    g = (g16[0, :, :K].double() + g16[1, :, :K].double())
    Y = y16.y.double()
    if y16.ones_col:
        Y = torch.cat([Y, torch.ones((Y.shape[0], 1), dtype=torch.float64)], dim=1)
    if g.shape[0] != Y.shape[0]:          # pieces cover the extended block; halo rows are zero
        raise ValueError("row mismatch")
    res = (g.T @ Y).float()
    if out is not None:
        out.copy_(res)
        return out
    return res


def mstep_adam(Phi, yw, tw, W, state, prior_std, step_size=0.01, maxiter=1000, tol=1e-6, min_iters=5,
               b1=0.9, b2=0.999, eps=1e-8, out=None):
    from oracle import ref_numpy as ref
    opt = {"count": int(state.count.item()), "mu": state.mu.numpy().astype(np.float64),
           "nu": state.nu.numpy().astype(np.float64)}
    r = ref.adam_run(W.numpy().astype(np.float64), opt, float(prior_std), Phi.numpy().astype(np.float64),
                     yw.numpy().astype(np.float64), tw.numpy().astype(np.float64), step_size, int(maxiter), tol,
                     b1, b2, eps)
    W.copy_(torch.from_numpy(r["params"].astype(np.float32)))
    state.mu.copy_(torch.from_numpy(r["opt_state"]["mu"].astype(np.float32)))
    state.nu.copy_(torch.from_numpy(r["opt_state"]["nu"].astype(np.float32)))
    state.count.fill_(int(r["opt_state"]["count"]))
    tuning_v = ref.get_tuning_softplus(r["params"], Phi.numpy().astype(np.float64)).astype(np.float32)
    if out is None:
        out = (torch.zeros(int(maxiter)), torch.zeros(int(maxiter)), torch.zeros(1, dtype=torch.int32),
               torch.zeros(2), torch.zeros(tuning_v.shape))
    lh, eh, ni, fin, tuning = out
    lh.copy_(torch.from_numpy(r["loss_history"].astype(np.float32)))
    eh.copy_(torch.from_numpy(r["error_history"].astype(np.float32)))
    ni.fill_(int(r["n_iter"]))
    fin.copy_(torch.tensor([float(r["final_loss"]), float(r["final_error"])]))
    tuning.copy_(torch.from_numpy(tuning_v))
    return lh, eh, ni, fin, tuning
