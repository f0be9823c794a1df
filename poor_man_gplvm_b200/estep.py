"""E-step engine: emission -> time-parallel forward/backward -> seam verification.

Host-side orchestration of reference ``decoder.smooth_all_step_combined_ma_chunk``
(poor_man_gplvm/decoder.py:258-332).  The reference walks 10 000-bin chunks
sequentially; here the time axis is cut into ``n_chain`` chunks ("chains") that run
concurrently, on one GPU or on a contiguous time block per rank.  A chain warms
up over ``halo`` bins starting from the previous pass's message at that bin (or
the stationary distribution of the prior chain on the first pass);
``pmg_seam_check`` then compares the warmed-up message with the neighbouring
chain's true one, and failing chains are restarted in parallel sweeps from a
snapshot of the neighbour's boundary message until every seam agrees to
``seam_tol`` (worst case this degenerates to the reference's sequential walk,
so the result never depends on the chunking beyond ``seam_tol``).
"""
from __future__ import annotations

import os

import torch

from . import ops
from .shard import TimeShard

DEFAULT_HALO = int(os.environ.get("PMG_HALO", "256"))
DEFAULT_SEAM_TOL = float(os.environ.get("PMG_SEAM_TOL", "1e-5"))
MIN_CHUNK_OVER_HALO = int(os.environ.get("PMG_MIN_CHUNK_OVER_HALO", "2"))
# chains per SM of an EM-mode plan on the compact kernels (12 warps per CTA, 3 per scheduler); 8 elsewhere
EM_CHAINS_PER_SM = int(os.environ.get("PMG_EM_CHAINS_PER_SM", "12"))


def plan_chunks(n_core, halo, sm_count, chains_per_sm=8):
    """chunk length: at least MIN_CHUNK_OVER_HALO x halo (bounded warm-up overhead),
    at most what fills `sm_count * chains_per_sm` chains."""
    if halo <= 0:
        return n_core
    target = max(1, sm_count * chains_per_sm)
    chunk = max(MIN_CHUNK_OVER_HALO * halo, (n_core + target - 1) // target)
    return min(chunk, n_core)


class EStepResult:
    """Tensors cover this rank's core bins only (views into the E-step buffers)."""
    __slots__ = ("ll", "alpha", "lmr", "gamma", "gamma_lat", "dyn_marg", "r", "tw", "log_marginal",
                 "n_relay_fwd", "n_relay_bwd", "seam_err_fwd", "seam_err_bwd", "plan", "alpha_ext", "r_ext", "core",
                 "repaired")


class EStep:
    """Buffers and launch plan for repeated E-steps over the same spike matrix (this rank's block)."""

    def __init__(self, y, op, ma_neuron=None, ma_latent=None, likelihood_scale=1.0, halo=None, seam_tol=None,
                 chunk_len=None, emission_impl=0, shard=None, em_mode=False):
        """em_mode: the plan is sized for the compact EM kernels (more, shorter chains) when they apply."""
        self.op = op
        self.K = op.K
        self.dev = y.device
        self.ma_neuron = ma_neuron
        self.ma_latent = ma_latent
        self.scale = float(likelihood_scale)
        self.halo = DEFAULT_HALO if halo is None else int(halo)
        self.seam_tol = DEFAULT_SEAM_TOL if seam_tol is None else float(seam_tol)
        self.emission_impl = emission_impl
        self.shard = shard if shard is not None else TimeShard(None, single=True)
        self.T_core, self.N = y.shape
        # neighbours' bins for the warm-ups that cross the block boundary (fetched once; y is constant)
        y_ext, self.h_left, self.h_right = self.shard.halo_exchange(y, self.halo)
        self.y = y_ext
        self.T = y_ext.shape[0]
        self.core = slice(self.h_left, self.h_left + self.T_core)
        self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.compact_ok = (os.environ.get("PMG_SCAN_COMPACT", "1") != "0"
                           and ops.scan_compact_supported(op, self.scale))
        if chunk_len is None:
            wide = em_mode and self.compact_ok and self.K > 31 * 8       # the 12-chain variants exist for K > 248
            chunk_len = plan_chunks(self.T_core, self.halo, self.sm_count, EM_CHAINS_PER_SM if wide else 8)
        self.chunk_len = int(min(max(1, chunk_len), self.T_core))
        self.plan = ops.make_plan(self.T, self.core.start, self.core.stop, self.chunk_len, self.halo,
                                  self.shard.is_first, self.shard.is_last, self.scale)
        S = self.plan.n_chain
        self.S = S
        f32 = dict(dtype=torch.float32, device=self.dev)
        # emission operands (constant across EM iterations); a [T,N] neuron mask arrives for the core bins
        # only, so the neighbours' halo rows of the mask are exchanged like the halo rows of y
        if ma_neuron is not None and ma_neuron.dim() == 2:
            ma_neuron, _, _ = self.shard.halo_exchange(ma_neuron.contiguous(), self.halo)
            self.ma_neuron = ma_neuron
        self.em = ops.EmissionOperands(self.y, ma_neuron, impl=emission_impl, ones_col=True)
        # fp16 counts for the statistics GEMM (the M-step uses the unmasked counts, reference core.py:807)
        if self.em.mode == 0:
            self.y16 = self.em.A16
        else:
            self.y16 = ops.CountsF16(self.y, ones_col=True) if emission_impl == 0 else None
        self.ll = torch.empty((self.T, self.K), **f32)
        self._alpha = None                 # [T,2,K] filtered posterior of the general path (allocated on first use)
        self._ax = None                    # [T,K+4] compact filtered posterior of the EM fast path
        self.lmr = torch.zeros(self.T, **f32)
        self.fwd_end = torch.zeros((S, 2, self.K), **f32)       # true message at the last bin of each chain
        self.first_out = torch.zeros((2, self.K), **f32)        # true message at the first core bin
        self.halo_state = torch.zeros((S, 2, self.K), **f32)
        self.beta_halo = torch.zeros((S, 2, self.K), **f32)
        self.beta_end = torch.zeros((S + 1, 2, self.K), **f32)     # [S] = the right neighbour's first chain
        self.truth = torch.zeros((S, 2, self.K), **f32)
        self.truth_left = None
        self.tw_partial = torch.zeros((S, self.K), **f32)
        # warm-up starts: ping-pong buffers holding, for every chain, the message of the previous pass
        # at the bin where its warm-up starts (forward and backward); before the first pass the forward
        # warm-up starts from the stationary distribution of the prior chain (exact for flat likelihoods)
        # One extra slot each for the neighbour rank's boundary chain: forward slot c belongs to chain c and slot
        # S to the right neighbour's first chain; backward slot c+1 belongs to chain c and slot 0 to the left
        # neighbour's last chain (the kernels see the backward buffers through a view that starts at slot 1).
        self.fwarm = [torch.zeros((S + 1, 2, self.K), **f32), torch.zeros((S + 1, 2, self.K), **f32)]
        self.bwarm = [torch.zeros((S + 1, 2, self.K), **f32), torch.zeros((S + 1, 2, self.K), **f32)]
        # entries no local chain ever writes (the warm-up of the first chain starts in the left
        # neighbour's bins, the last chain's backward warm-up in the right neighbour's): keep them at the
        # first-pass defaults — stationary prior for the forward message, all-ones for the backward one
        stat = getattr(op, "stationary", None)
        for buf in self.fwarm:
            buf[0] = stat if stat is not None else 0.5 / self.K
        for buf in self.bwarm:
            buf[S] = 1.0
        self.warm_cur = 0
        self.warm_valid = False
        self.err = torch.zeros(2 * S, **f32)
        self.err_host = torch.zeros(2 * S, dtype=torch.float32).pin_memory()
        self._gmax_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        # rows of alpha holding the true message in front of chain c (c >= 1): bin t_begin(c) - 1
        self.rows_f = self.core.start + torch.arange(1, S, device=self.dev) * self.chunk_len - 1
        # which seams exist: forward seam c sits in front of chain c; backward seam c behind chain c
        self.f_lo = 0 if not self.shard.is_first else 1
        self.b_hi = S if not self.shard.is_last else S - 1

    @property
    def alpha(self):
        if self._alpha is None:
            self._alpha = torch.zeros((self.T, 2, self.K), dtype=torch.float32, device=self.dev)
        return self._alpha

    @property
    def ax(self):
        if self._ax is None:
            self._ax = torch.zeros((self.T, self.K + 4), dtype=torch.float32, device=self.dev)
        return self._ax

    # -- pieces ---------------------------------------------------------------------------
    def emission(self, tuning):
        self.em.loglik(tuning, self.ma_latent, 1.0, out=self.ll)
        return self.ll

    def _exchange_fwd(self, compact=False, nxt=None):
        """After a forward pass: the last true alpha goes right (seam truth of the neighbour's first
        chain) together with this pass's message at the bin where that chain's next warm-up starts; the first
        true alpha goes left (normaliser of the neighbour's last backward seam)."""
        if not self.shard.active:
            return
        K2 = 2 * self.K
        if compact:
            first, last = self.first_out.reshape(-1), self.fwd_end[self.S - 1].reshape(-1)
        else:
            first = self.alpha[self.core.start].reshape(-1)
            last = self.alpha[self.core.stop - 1].reshape(-1)
        # message at the bin in front of the right neighbour's first warm-up bin.  The kernels' slot S holds it
        # only when that bin lies in the last chain's own range; with a ragged (short) last chunk it lies in an
        # earlier chain, so it is read back from the stored filtered posterior instead.
        t_star = self.core.stop - self.halo - 1
        if nxt is None:
            warm = torch.zeros_like(last)
        elif t_star < self.core.start:
            warm = self.fwarm[nxt][self.S].reshape(-1)
        elif compact:
            row = self.ll[t_star]
            E = torch.exp2((row - row.max()) * (self.scale * 1.4426950408889634))
            warm = torch.cat([self.ax[t_star, :self.K], self.ax[t_star, self.K] * E])
        else:
            warm = self.alpha[t_star].reshape(-1)
        from_left, from_right = self.shard.boundary(torch.cat([first, torch.zeros_like(first)]),
                                                    torch.cat([last, warm]))
        if from_left is not None:
            self.truth_left = from_left[:K2].view(2, self.K)
            w = from_left[K2:]
            if nxt is not None:
                self.fwarm[nxt][0].copy_(w.view(2, self.K))
        if from_right is not None:
            msg = from_right[:K2].view(2, self.K)
            if compact:
                # row of the compact buffer for the neighbour's first bin: alpha[0,:] and the scalar a1s with
                # alpha[1,x] = a1s * exp2(s*log2e*(ll[x] - max ll))  (the factor the kernels recompute)
                row = self.ll[self.core.stop]
                E = torch.exp2((row - row.max()) * (self.scale * 1.4426950408889634))
                self.ax[self.core.stop, :self.K] = msg[0]
                self.ax[self.core.stop, self.K] = msg[1].sum() / E.sum()
            else:
                self.alpha[self.core.stop].copy_(msg)

    def _exchange_bwd(self, nxt=None):
        """After a backward pass: beta at the first core bin goes left (seam truth of the neighbour's last
        chain) together with this pass's message at the bin where that chain's next warm-up starts."""
        if not self.shard.active:
            return
        K2 = 2 * self.K
        first = self.beta_end[0].reshape(-1)
        warm = self.bwarm[nxt][0].reshape(-1) if nxt is not None else torch.zeros_like(first)
        _, from_right = self.shard.boundary(torch.cat([first, warm]), None)
        if from_right is not None:
            self.beta_end[self.S].copy_(from_right[:K2].view(2, self.K))
            if nxt is not None:
                self.bwarm[nxt][self.S].copy_(from_right[K2:].view(2, self.K))

    def _check_fwd(self, compact=False):
        S, K2 = self.S, 2 * self.K
        if S > 1:
            self.truth[1:] = self.fwd_end[:-1] if compact else self.alpha[self.rows_f]
        if self.f_lo == 0:
            self.truth[0] = self.truth_left
        n = S - self.f_lo
        if n > 0:
            ops.seam_check(n, K2, self.halo_state[self.f_lo:].data_ptr(), K2, self.truth[self.f_lo:].data_ptr(), K2,
                           self.err[self.f_lo:S])

    def _check_bwd(self):
        S, K2 = self.S, 2 * self.K
        n = self.b_hi
        if n > 0:
            ops.seam_check(n, K2, self.beta_halo.data_ptr(), K2, self.beta_end[1:].data_ptr(), K2,
                           self.err[S:S + n])

    def _read_err(self):
        self.err_host.copy_(self.err, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.err_host

    def _read_err_global(self):
        """Seam errors of this rank plus, for time-sharded runs, whether ANY rank has a failing forward /
        backward seam -- one collective and one synchronisation for the common case of no repair at all."""
        if not self.shard.active:
            err = self._read_err()
            ef, eb = err[self.f_lo:self.S], err[self.S:self.S + self.b_hi]
            return err, bool(ef.numel() and float(ef.max()) > self.seam_tol), \
                bool(eb.numel() and float(eb.max()) > self.seam_tol)
        S = self.S
        zero = self.err.new_zeros(())
        m = torch.stack([self.err[self.f_lo:S].max() if S > self.f_lo else zero,
                         self.err[S:S + self.b_hi].max() if self.b_hi > 0 else zero])
        m = torch.nan_to_num(m, nan=float("inf"))
        self.shard.allreduce_max_(m)
        self._gmax_host.copy_(m, non_blocking=True)
        err = self._read_err()
        return err, bool(self._gmax_host[0] > self.seam_tol), bool(self._gmax_host[1] > self.seam_tol)

    def run(self, tuning, want_gamma=False, want_gamma_lat=True, want_dyn=False, want_r=False, gamma16=None,
            before_sync=None):
        """One E-step.  gamma16: optional [2,T,ldg] fp16 buffer (T = local bins incl. halos) that receives
        the hi/lo pieces of the latent posterior.  before_sync: optional callable invoked once both passes and
        the seam checks are enqueued, before the launching thread waits for the verdict (work enqueued there
        keeps the GPU busy during the synchronisation; ``res.repaired`` tells whether chains were re-run after
        it, i.e. whether what it read from this E-step's outputs was final)."""
        S, K = self.S, self.K
        f32 = dict(dtype=torch.float32, device=self.dev)
        # EM fast path: only the fp16 posterior pieces and sum_t gamma are wanted -> compact kernels
        compact = (self.compact_ok and gamma16 is not None
                   and not (want_gamma or want_gamma_lat or want_dyn or want_r))
        self.emission(tuning)
        ops.phase("emission")
        gamma = torch.empty((self.T, 2, K), **f32) if want_gamma else None
        gamma_lat = torch.empty((self.T, K), **f32) if want_gamma_lat else None
        dyn = torch.empty((self.T, 2), **f32) if want_dyn else None
        r = torch.zeros((self.T, 2, K), **f32) if want_r else None

        cur, nxt = self.warm_cur, 1 - self.warm_cur
        f_in = self.fwarm[cur] if self.warm_valid else getattr(self.op, "stationary", None)
        b_in = self.bwarm[cur][1:] if self.warm_valid else None
        b_out = self.bwarm[nxt][1:]

        def fwd(mode=0, ids=None):
            if compact:
                ops.forward_compact(self.plan, self.op, self.ll, self.ax,
                                    halo_state=(self.halo_state if mode == 0 else None), fwd_end=self.fwd_end,
                                    first_out=self.first_out, mode=mode, chain_ids=ids,
                                    warm_in=(f_in if mode == 0 else self.halo_state), warm_out=self.fwarm[nxt])
                return
            ops.forward(self.plan, self.op, self.ll, self.alpha, self.lmr,
                        halo_state=(self.halo_state if mode == 0 else None), mode=mode, chain_ids=ids,
                        warm_in=(f_in if mode == 0 else self.halo_state), warm_out=self.fwarm[nxt])

        def bwd(mode=0, ids=None):
            if compact:
                ops.backward_compact(self.plan, self.op, self.ll, self.ax, gamma16, beta_halo=self.beta_halo, beta_end=self.beta_end, mode=mode, chain_ids=ids,
                                     warm_in=(b_in if mode == 0 else self.beta_halo), warm_out=b_out)
                return
            ops.backward(self.plan, self.op, self.ll, self.alpha, gamma=gamma, gamma_lat=gamma_lat, dyn_marg=dyn,
                         r_out=r, tw_partial=self.tw_partial, beta_halo=self.beta_halo, beta_end=self.beta_end,
                         mode=mode, chain_ids=ids, gamma16=gamma16,
                         warm_in=(b_in if mode == 0 else self.beta_halo), warm_out=b_out)

        fwd()
        self._exchange_fwd(compact, nxt)
        self._check_fwd(compact)
        ops.phase("forward")
        bwd()
        self._exchange_bwd(nxt)
        self._check_bwd()
        ops.phase("backward")

        n_relay_f = n_relay_b = 0
        ef = eb = torch.zeros(0)
        any_f = any_b = False
        if before_sync is not None:
            before_sync()
        if S > 1 or self.shard.active:
            err, any_f, any_b = self._read_err_global()
            ef = err[self.f_lo:S].clone()         # ef[i]: seam in front of chain f_lo + i
            eb = err[S:S + self.b_hi].clone()     # eb[c]: seam behind chain c
            # Seam repair = parallel (Jacobi) sweeps: every chain whose incoming message was off restarts,
            # all at once (on all ranks), from a snapshot of its neighbour's current boundary message; the
            # seams are then re-verified against the messages those restarts produced.  Each sweep extends
            # the effective warm-up by one chunk, so the number of sweeps is ~ mixing length / chunk length.
            # One sweep = forward restarts, then the backward pass of (a) the chains whose filtered posterior
            # just changed -- from their own, already verified, incoming beta -- and (b) the chains whose
            # incoming beta was off -- from the neighbour's; other chains' backward results do not depend on
            # the re-run chains (beta does not involve alpha; normalisers are scale only).  One verdict
            # (synchronisation + collective) per sweep.
            repaired = bool(any_f or any_b)
            for _ in range(S * self.shard.world + 2):
                if not (any_f or any_b):
                    break
                bad_f = torch.nonzero(ef > self.seam_tol).flatten() + self.f_lo
                bad_b = torch.nonzero(eb > self.seam_tol).flatten()
                n_relay_f += int(bad_f.numel())
                n_relay_b += int(bad_b.numel())
                if bad_f.numel():
                    ids = bad_f.to(device=self.dev, dtype=torch.int32)
                    self.halo_state[ids.long()] = self.truth[ids.long()]      # carry snapshot = new "estimate"
                    fwd(mode=1, ids=ids)
                self._exchange_fwd(compact, nxt)
                if bad_b.numel():
                    idb = bad_b.to(device=self.dev)
                    self.beta_halo[idb] = self.beta_end[idb + 1]
                both = torch.unique(torch.cat([bad_f, bad_b]))
                if both.numel():
                    bwd(mode=1, ids=both.to(device=self.dev, dtype=torch.int32))
                self._exchange_bwd(nxt)
                self._check_fwd(compact)
                self._check_bwd()
                err, any_f, any_b = self._read_err_global()
                ef = err[self.f_lo:S].clone()
                eb = err[S:S + self.b_hi].clone()
            any_f = any_b = repaired
            self.warm_cur, self.warm_valid = nxt, True

        c = self.core
        res = EStepResult()
        res.core = c
        lmr = self.ax[:, K + 1] if compact else self.lmr
        res.ll, res.lmr = self.ll[c], lmr[c]
        res.alpha = None if compact else self.alpha[c]
        res.alpha_ext, res.r_ext = (None if compact else self.alpha), r
        res.gamma = gamma[c] if gamma is not None else None
        res.gamma_lat = gamma_lat[c] if gamma_lat is not None else None
        res.dyn_marg = dyn[c] if dyn is not None else None
        res.r = r[c] if r is not None else None
        # local sums; the caller all-reduces them together with the spike-weighted statistics
        # (the compact path leaves sum_t gamma to the statistics GEMM: ones column of the fp16 counts)
        res.tw = None if compact else self.tw_partial.sum(dim=0, dtype=torch.float64).to(torch.float32)
        res.log_marginal = lmr[c].sum(dtype=torch.float64)
        res.n_relay_fwd, res.n_relay_bwd = n_relay_f, n_relay_b
        res.repaired = bool(any_f or any_b)                   # same verdict on every rank
        res.seam_err_fwd = float(ef.max()) if ef.numel() else 0.0
        res.seam_err_bwd = float(eb.max()) if eb.numel() else 0.0
        res.plan = self.plan
        return res
