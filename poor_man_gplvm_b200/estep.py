"""E-step engine: emission -> time-parallel forward/backward -> seam verification.

Host-side orchestration of reference ``decoder.smooth_all_step_combined_ma_chunk``
(poor_man_gplvm/decoder.py:258-332).  The reference walks 10 000-bin chunks
sequentially; here the time axis is cut into ``n_chain`` chunks that run
concurrently.  A chain warms up over ``halo`` bins from the uniform message;
``pmg_seam_check`` then compares its warmed-up message with the neighbouring
chain's true one, and any seam above tolerance is repaired by relaying the
exact carry (worst case this degenerates to the reference's sequential walk,
so the result never depends on the chunking beyond ``seam_tol``).
"""
from __future__ import annotations

import os

import torch

from . import ops

DEFAULT_HALO = int(os.environ.get("PMG_HALO", "256"))
DEFAULT_SEAM_TOL = float(os.environ.get("PMG_SEAM_TOL", "1e-5"))
MIN_CHUNK_OVER_HALO = int(os.environ.get("PMG_MIN_CHUNK_OVER_HALO", "4"))


def plan_chunks(n_core, halo, sm_count, chains_per_sm=8):
    """chunk length: at least MIN_CHUNK_OVER_HALO x halo (bounded warm-up overhead),
    at most what fills `sm_count * chains_per_sm` chains."""
    if halo <= 0:
        return n_core
    target = max(1, sm_count * chains_per_sm)
    chunk = max(MIN_CHUNK_OVER_HALO * halo, (n_core + target - 1) // target)
    return min(chunk, n_core)


class EStepResult:
    __slots__ = ("ll", "alpha", "lmr", "gamma", "gamma_lat", "dyn_marg", "r", "tw", "log_marginal",
                 "n_relay_fwd", "n_relay_bwd", "seam_err_fwd", "seam_err_bwd", "plan")


class EStep:
    """Buffers and launch plan for repeated E-steps over the same spike matrix."""

    def __init__(self, y, op, ma_neuron=None, ma_latent=None, likelihood_scale=1.0, halo=None, seam_tol=None,
                 chunk_len=None, emission_impl=0):
        self.y = y
        self.op = op
        self.T, self.N = y.shape
        self.K = op.K
        self.dev = y.device
        self.ma_neuron = ma_neuron
        self.ma_latent = ma_latent
        self.scale = float(likelihood_scale)
        self.halo = DEFAULT_HALO if halo is None else int(halo)
        self.seam_tol = DEFAULT_SEAM_TOL if seam_tol is None else float(seam_tol)
        self.emission_impl = emission_impl
        self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count
        if chunk_len is None:
            chunk_len = plan_chunks(self.T, self.halo, self.sm_count)
        self.chunk_len = int(min(max(1, chunk_len), self.T))
        self.plan = ops.make_plan(self.T, 0, self.T, self.chunk_len, self.halo, True, True, self.scale)
        S = self.plan.n_chain
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.lgam = ops.lgamma_rowsum(y, ma_neuron)
        self.y16 = ops.CountsF16(y) if emission_impl == 0 else None
        self.ll = torch.empty((self.T, self.K), **f32)
        self.alpha = torch.empty((self.T, 2, self.K), **f32)
        self.lmr = torch.empty(self.T, **f32)
        self.halo_state = torch.zeros((S, 2, self.K), **f32)
        self.beta_halo = torch.zeros((S, 2, self.K), **f32)
        self.beta_end = torch.zeros((S, 2, self.K), **f32)
        self.tw_partial = torch.zeros((S, self.K), **f32)
        # warm-up starts: ping-pong buffers holding, for every chain, the message of the previous pass
        # at the bin where its warm-up starts (forward and backward); before the first pass the forward
        # warm-up starts from the stationary distribution of the prior chain (exact for flat likelihoods)
        self.fwarm = [torch.zeros((S, 2, self.K), **f32), torch.zeros((S, 2, self.K), **f32)]
        self.bwarm = [torch.zeros((S, 2, self.K), **f32), torch.zeros((S, 2, self.K), **f32)]
        self.warm_cur = 0
        self.warm_valid = False
        self.err = torch.zeros(2 * max(S, 1), **f32)
        self.err_host = torch.zeros(2 * max(S, 1), dtype=torch.float32).pin_memory()

    # -- pieces ---------------------------------------------------------------------------
    def emission(self, tuning):
        ops.emission(self.y, tuning, self.lgam, self.ma_neuron, self.ma_latent, 1.0, out=self.ll, y16=self.y16,
                     impl=self.emission_impl)
        return self.ll

    def _check_fwd(self, n, first_chain=1):
        # est = halo_state[s], truth = alpha[t_begin(s)-1], s = first_chain..first_chain+n-1
        K2 = 2 * self.K
        est = self.halo_state.data_ptr() + first_chain * K2 * 4
        truth = self.alpha.data_ptr() + (first_chain * self.chunk_len - 1) * K2 * 4
        ops.seam_check(n, K2, est, K2, truth, self.chunk_len * K2, self.err[first_chain:first_chain + n])

    def _check_bwd(self, n, first_chain=0):
        # est = beta_halo[s], truth = beta_end[s+1], s = first_chain..first_chain+n-1
        K2 = 2 * self.K
        S = self.plan.n_chain
        est = self.beta_halo.data_ptr() + first_chain * K2 * 4
        truth = self.beta_end.data_ptr() + (first_chain + 1) * K2 * 4
        ops.seam_check(n, K2, est, K2, truth, K2, self.err[S + first_chain:S + first_chain + n])

    def _read_err(self):
        self.err_host.copy_(self.err, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.err_host

    def run(self, tuning, want_gamma=False, want_gamma_lat=True, want_dyn=False, want_r=False, gamma16=None):
        """One E-step.  Returns an EStepResult whose tensors alias this object's buffers.
        gamma16: optional [2,T,ldg] fp16 buffer that receives the hi/lo pieces of the latent posterior."""
        S = self.plan.n_chain
        K = self.K
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.emission(tuning)
        ops.phase("emission")
        gamma = torch.empty((self.T, 2, K), **f32) if want_gamma else None
        gamma_lat = torch.empty((self.T, K), **f32) if want_gamma_lat else None
        dyn = torch.empty((self.T, 2), **f32) if want_dyn else None
        r = torch.zeros((self.T, 2, K), **f32) if want_r else None

        cur, nxt = self.warm_cur, 1 - self.warm_cur
        f_in = self.fwarm[cur] if self.warm_valid else getattr(self.op, "stationary", None)
        b_in = self.bwarm[cur] if self.warm_valid else None

        def bwd(mode=0, ids=None, carry=None):
            ops.backward(self.plan, self.op, self.ll, self.alpha, gamma=gamma, gamma_lat=gamma_lat, dyn_marg=dyn,
                         r_out=r, tw_partial=self.tw_partial, beta_halo=self.beta_halo, beta_end=self.beta_end,
                         mode=mode, chain_ids=ids, gamma16=gamma16, warm_in=(b_in if mode == 0 else carry),
                         warm_out=self.bwarm[nxt])

        ops.forward(self.plan, self.op, self.ll, self.alpha, self.lmr, halo_state=self.halo_state, warm_in=f_in,
                    warm_out=self.fwarm[nxt])
        n_relay_f = n_relay_b = 0
        if S > 1:
            self._check_fwd(S - 1)
        ops.phase("forward")
        bwd()
        ops.phase("backward")
        if S > 1:
            self._check_bwd(S - 1)
            err = self._read_err()
            ef = err[1:S].clone()            # ef[c-1]: seam in front of chain c
            eb = err[S:2 * S - 1].clone()    # eb[c]:   seam behind chain c
            # Seam repair = parallel (Jacobi) sweeps: every chain whose incoming message was off restarts,
            # all at once, from a snapshot of its neighbour's current boundary message; the seams are then
            # re-verified against the messages those restarts produced.  Each sweep extends the effective
            # warm-up by one chunk, so the number of sweeps is ~ mixing length / chunk length (in the worst,
            # non-mixing case it degenerates to the reference's sequential walk).
            rows = torch.arange(1, S, device=self.dev) * self.chunk_len - 1      # bin t_begin(c)-1, c = 1..S-1
            redo_bwd = False
            for _ in range(S):
                bad = torch.nonzero(ef > self.seam_tol).flatten() + 1
                if not bad.numel():
                    break
                redo_bwd = True
                n_relay_f += int(bad.numel())
                ids = bad.to(device=self.dev, dtype=torch.int32)
                self.halo_state[ids.long()] = self.alpha[rows[ids.long() - 1]]   # carry snapshot = new "estimate"
                ops.forward(self.plan, self.op, self.ll, self.alpha, self.lmr, halo_state=None, mode=1,
                            chain_ids=ids, warm_in=self.halo_state, warm_out=self.fwarm[nxt])
                self._check_fwd(S - 1)
                ef = self._read_err()[1:S].clone()
            if redo_bwd:
                bwd()
                self._check_bwd(S - 1)
                eb = self._read_err()[S:2 * S - 1].clone()
            for _ in range(S):
                bad = torch.nonzero(eb > self.seam_tol).flatten()
                if not bad.numel():
                    break
                n_relay_b += int(bad.numel())
                ids = bad.to(device=self.dev, dtype=torch.int32)
                self.beta_halo[ids.long()] = self.beta_end[ids.long() + 1]
                bwd(mode=1, ids=ids, carry=self.beta_halo)
                self._check_bwd(S - 1)
                eb = self._read_err()[S:2 * S - 1].clone()
        else:
            ef = eb = torch.zeros(0)

        if S > 1:
            self.warm_cur, self.warm_valid = nxt, True
        res = EStepResult()
        res.ll, res.alpha, res.lmr = self.ll, self.alpha, self.lmr
        res.gamma, res.gamma_lat, res.dyn_marg, res.r = gamma, gamma_lat, dyn, r
        res.tw = self.tw_partial.sum(dim=0, dtype=torch.float64).to(torch.float32)
        res.log_marginal = self.lmr.sum(dtype=torch.float64)
        res.n_relay_fwd, res.n_relay_bwd = n_relay_f, n_relay_b
        res.seam_err_fwd = float(ef.max()) if ef.numel() else 0.0
        res.seam_err_bwd = float(eb.max()) if eb.numel() else 0.0
        res.plan = self.plan
        return res
