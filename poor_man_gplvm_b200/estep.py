"""E-step engine: emission -> time-parallel forward/backward -> seam verification.

Host-side orchestration of reference ``decoder.smooth_all_step_combined_ma_chunk``
(poor_man_gplvm/decoder.py:258-332).  The reference walks 10 000-bin chunks
sequentially; here the time axis is cut into ``n_chain`` chunks ("chains") that run
concurrently, on one GPU or on a contiguous time block per rank.  A chain warms
up over ``halo`` bins starting from the previous pass's message at that bin (or
the stationary distribution of the prior chain on the first pass);
``pmg_seam_check_fix`` then compares the warmed-up message with the neighbouring
chain's true one.

Seams that miss ``seam_tol`` are repaired in two tiers:

* on the device, with no host round trip: the check overwrites the failed estimate
  by the neighbour's true boundary message and a *conditional* relaunch of the pass
  (scan mode 2) re-runs exactly those chains from that snapshot; the other chains'
  warps return at once.  The forward repair runs before the backward pass, so the
  smoother already sees the repaired filtered posterior.  This is the common case
  (a few slow-mixing seams per thousand on a fitted model, most seams of a flat one);
* on the host (Jacobi sweeps, as in round 1) if a seam still fails the re-check:
  failing chains restart all at once from their neighbours' current boundary
  messages until every seam passes; worst case this degenerates to the reference's
  sequential walk, so the result never depends on the chunking beyond ``seam_tol``.

One host synchronisation per E-step: a 5-float record (log marginal, seams repaired on
the device and seams still failing, per pass) plus the seam errors.  On time-sharded
runs that record travels in the caller's packed statistics all-reduce
(``before_sync``), so an EM iteration has ONE collective besides the two boundary
exchanges.  The warm-up length adapts between ``halo_min`` and the initial ``halo``
(EM mode): it halves while no seam needs a repair and grows back when repairs would
cost more than the longer warm-up.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .shard import TimeShard

DEFAULT_HALO = int(os.environ.get("PMG_HALO", "256"))
DEFAULT_HALO_MIN = int(os.environ.get("PMG_HALO_MIN", "32"))
DEFAULT_SEAM_TOL = float(os.environ.get("PMG_SEAM_TOL", "1e-5"))
MIN_CHUNK = int(os.environ.get("PMG_MIN_CHUNK", "64"))
# chains per SM of an EM-mode plan on the compact kernels (12 warps per CTA, 3 per scheduler); 8 elsewhere
EM_CHAINS_PER_SM = 12

# record shared with the caller's collective: log marginal, then (repaired on device, still failing) per pass
_NO_CHAINS = __import__("numpy").zeros(0, dtype="int64")
TAIL = 5
T_LML, T_FIX_F, T_FAIL_F, T_FIX_B, T_FAIL_B = range(TAIL)


def plan_chunks(n_core, halo, sm_count, chains_per_sm=8, min_chunk=None):
    """chunk length: what fills `sm_count * chains_per_sm` chains, but at least `min_chunk` bins (bounded
    warm-up overhead: the warm-up adapts down to DEFAULT_HALO_MIN in EM mode; one-shot passes keep
    chunks of at least 2 x halo)."""
    if halo <= 0:
        return n_core
    if min_chunk is None:
        min_chunk = 2 * halo
    target = max(1, sm_count * chains_per_sm)
    chunk = max(int(min_chunk), (n_core + target - 1) // target)
    return min(chunk, n_core)


class EStepResult:
    """Tensors cover this rank's core bins only (views into the E-step buffers)."""
    __slots__ = ("ll", "alpha", "lmr", "gamma", "gamma_lat", "dyn_marg", "r", "tw", "log_marginal",
                 "n_relay_fwd", "n_relay_bwd", "n_fix_fwd", "n_fix_bwd", "seam_err_fwd", "seam_err_bwd", "plan",
                 "alpha_ext", "r_ext", "core", "repaired", "halo", "xi16")


class EStep:
    """Buffers and launch plan for repeated E-steps over the same spike matrix (this rank's block)."""

    def __init__(self, y, op, ma_neuron=None, ma_latent=None, likelihood_scale=1.0, halo=None, seam_tol=None,
                 chunk_len=None, emission_impl=0, shard=None, em_mode=False, tail=None, adaptive=None,
                 emission_factory=None, carry_in=None):
        """em_mode: the plan is sized for the compact EM kernels (more, shorter chains) when they apply and the
        warm-up length adapts from pass to pass.  tail: optional 5-float device view the E-step writes its record
        into (the caller all-reduces it together with its own data inside ``before_sync``).
        emission_factory: callable(y_ext, ma_neuron) -> object with ``loglik`` (default: Poisson EmissionOperands);
        carry_in: [2,K] message in front of bin 0 of the recording (default: uniform, reference decoder.py:181-183)."""
        self.op = op
        self.carry_in = carry_in
        self.K = op.K
        self.dev = y.device
        self.ma_neuron = ma_neuron
        self.ma_latent = ma_latent
        self.scale = float(likelihood_scale)
        self.halo = DEFAULT_HALO if halo is None else int(halo)       # widest warm-up = rows fetched from neighbours
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
        if adaptive is None:
            adaptive = em_mode and os.environ.get("PMG_HALO_ADAPT", "1") != "0"
        self.adaptive = bool(adaptive) and self.halo > 0
        # tier 1 of the seam repair (conditional relaunch on the device); off = every repair is a host sweep
        self.device_repair = os.environ.get("PMG_DEVICE_REPAIR", "1") != "0"
        self.halo_min = min(self.halo, DEFAULT_HALO_MIN)     # raised below once the chunk length is known
        self.dense = getattr(op, "dense", None)      # lockstep tensor-core scan (dense / wide-band move kernels)
        if chunk_len is None and self.dense is not None and self.halo > 0:
            # one wave of 128-chain accumulator tiles; a pass costs (warm-up + chunk) GEMM steps
            mc = min(MIN_CHUNK, 2 * self.halo) if self.adaptive else 2 * self.halo
            target = self.dense.chains(self.sm_count)
            chunk_len = min(self.T_core, max(mc, (self.T_core + target - 1) // target))
        if chunk_len is None:
            wide = em_mode and self.compact_ok and self.K > 31 * 8       # the 12-chain variants exist for K > 248
            chunk_len = plan_chunks(self.T_core, self.halo, self.sm_count, EM_CHAINS_PER_SM if wide else 8,
                                    min_chunk=(min(MIN_CHUNK, 2 * self.halo) if self.adaptive else None))
        self.chunk_len = int(min(max(1, chunk_len), self.T_core))
        # A seam that fails is repaired by re-running its chain: about one chunk of scan time at low occupancy,
        # whatever the number of failing seams.  A shorter warm-up saves every chain `halo` steps at full occupancy.
        # Halving the warm-up below ~chunk/5 cannot pay for even occasional repairs, so that is the floor.
        floor = self.halo
        while floor // 2 >= max(self.halo_min, self.chunk_len // 5) and floor // 2 >= 1:
            floor //= 2
        self.halo_min = min(self.halo, max(self.halo_min, floor))
        # common warm-up of this pass and of the next one (the pass writes the next pass's warm-start messages)
        self.halos = [self.halo, self.halo]
        self._calm, self._streak, self._hold, self._bounces = 0, 0, 0, 0
        self.plan = ops.make_plan(self.T, self.core.start, self.core.stop, self.chunk_len, self.halo,
                                  self.shard.is_first, self.shard.is_last, self.scale, halo_next=self.halo)
        S = self.plan.n_chain
        self.S = S
        f32 = dict(dtype=torch.float32, device=self.dev)
        # emission operands (constant across EM iterations); a [T,N] neuron mask arrives for the core bins
        # only, so the neighbours' halo rows of the mask are exchanged like the halo rows of y
        if ma_neuron is not None and ma_neuron.dim() == 2:
            ma_neuron, _, _ = self.shard.halo_exchange(ma_neuron.contiguous(), self.halo)
            self.ma_neuron = ma_neuron
        if emission_factory is not None:
            self.em = emission_factory(self.y, ma_neuron)
        else:
            self.em = ops.EmissionOperands(self.y, ma_neuron, impl=emission_impl, ones_col=True)
        # fp16 counts for the statistics GEMM (the M-step uses the unmasked counts, reference core.py:807)
        if self.em.mode == 0:
            self.y16 = self.em.A16
        else:
            self.y16 = ops.CountsF16(self.y, ones_col=True) if (emission_impl == 0 and emission_factory is None) else None
        self.ll = torch.empty((self.T, self.K), **f32)
        self._alpha = None                 # [T,2,K] filtered posterior of the general path (allocated on first use)
        self._ax = None                    # [T,K+4] compact filtered posterior of the EM fast path
        self.lmr = torch.zeros(self.T, **f32)
        # true message at the last bin of each chain, one slot ahead: slot 0 holds the left neighbour rank's last
        # message, chain c writes slot c+1, so seam c (in front of chain c) always finds its truth in slot c
        self.fwd_end_ext = torch.zeros((S + 1, 2, self.K), **f32)
        self.fwd_end = self.fwd_end_ext[1:]
        self.first_out = torch.zeros((2, self.K), **f32)        # true message at the first core bin
        self.halo_state = torch.zeros((S, 2, self.K), **f32)
        self.beta_halo = torch.zeros((S, 2, self.K), **f32)
        self.beta_end = torch.zeros((S + 1, 2, self.K), **f32)     # [S] = the right neighbour's first chain
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
        # seam errors: [final check of every seam | first check of the interior seams (before repairs)], one buffer so
        # that one copy brings both to the host
        self._errs = torch.zeros(4 * S, **f32)
        self.err, self.err1 = self._errs[:2 * S], self._errs[2 * S:]
        self._errs_host = torch.zeros(4 * S, dtype=torch.float32).pin_memory()
        self.err_host, self.err1_host = self._errs_host[:2 * S], self._errs_host[2 * S:]
        # per-chain warm-up lengths (adaptive mode): [this pass, next pass] for each direction, device + host copies
        self.hf = self.hb = None
        if self.adaptive:
            import numpy as np
            i32 = dict(dtype=torch.int32, device=self.dev)
            self.hf = [torch.full((S,), self.halo, **i32), torch.full((S,), self.halo, **i32)]
            self.hb = [torch.full((S,), self.halo, **i32), torch.full((S,), self.halo, **i32)]
            self.hf_host = [np.full(S, self.halo, np.int32), np.full(S, self.halo, np.int32)]
            self.hb_host = [np.full(S, self.halo, np.int32), np.full(S, self.halo, np.int32)]
            self.boost_f, self.boost_b = np.zeros(S, np.int32), np.zeros(S, np.int32)
            self.h_cur = 0
            # A boosted chain runs longer than its neighbours and the pass ends with the longest chain, so a boost
            # never exceeds the initial warm-up (the length every chain had before the base came down); a seam that
            # still fails then is repaired on the device every pass.
            self.boost_cap = self.halo
        self.tail = tail if tail is not None else torch.zeros(TAIL, **f32)
        if self.tail.numel() != TAIL or self.tail.dtype != torch.float32 or not self.tail.is_contiguous():
            raise ValueError("tail must be a contiguous float32 view of %d entries" % TAIL)
        self.tail_host = torch.zeros(TAIL, dtype=torch.float32).pin_memory()
        self._xbuf_f = self._xbuf_b = None          # message buffers of the boundary exchanges (time-sharded runs)
        self._sum_ws = None
        # which seams exist: forward seam c sits in front of chain c; backward seam c behind chain c
        self.f_lo = 0 if not self.shard.is_first else 1
        self.b_hi = S if not self.shard.is_last else S - 1
        # CUDA graphs of the steady-state E-step (+ whatever the caller enqueues in before_sync), see run()
        # (time-sharded ranks capture their NCCL exchanges along; PMG_EM_GRAPH_DIST=0 keeps those runs eager)
        # Opt-in (PMG_EM_GRAPH=1): measured on B200 the EM iteration is GPU-bound even at 125 000 bins per rank
        # (1.32 ms eager vs 1.37 ms replayed), and every new warm-up plan costs a capture.
        self.use_graphs = (em_mode and os.environ.get("PMG_EM_GRAPH", "0") != "0" and self.dev.type == "cuda"
                           and not self.shard.staged
                           and (not self.shard.active or os.environ.get("PMG_EM_GRAPH_DIST", "0") != "0"))
        self._graphs, self._graph_seen = {}, set()
        # host repair sweeps: same bound on every rank (block lengths, hence S, may differ between ranks)
        self.max_sweeps = self.shard.max_int(S, self.dev) * self.shard.world + 2

    @property
    def alpha(self):
        if self._alpha is None:
            # every row a pass reads is written first (core rows by the forward pass, the row behind the block by the
            # boundary exchange): no 3.2 GB zero fill
            self._alpha = torch.empty((self.T, 2, self.K), dtype=torch.float32, device=self.dev)
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
        true alpha goes left (normaliser of the neighbour's last backward seam).  One pack kernel, one all-gather of
        the ranks' fixed-size buffers, one unpack kernel."""
        sh = self.shard
        if not sh.active:
            return
        K = self.K
        if self._xbuf_f is None:
            self._xbuf_f = torch.zeros(8 * K, dtype=torch.float32, device=self.dev)
            self._xbuf_b = torch.zeros(4 * K, dtype=torch.float32, device=self.dev)
        if compact:
            first, last = self.first_out, self.fwd_end[self.S - 1]
        else:
            first, last = self.alpha[self.core.start], self.alpha[self.core.stop - 1]
        # message at the bin in front of the right neighbour's first warm-up bin (of the NEXT pass).  The kernels'
        # slot S holds it only when that bin lies in the last chain's own range; with a ragged (short) last chunk
        # it lies in an earlier chain, so it is read back from the stored filtered posterior instead.
        t_star = self.core.stop - self.halos[1] - 1
        kw = {}
        if nxt is None:
            pass
        elif t_star < self.core.start:
            kw = dict(warm_src=self.fwarm[nxt][self.S])
        elif compact:
            kw = dict(ax_row=self.ax[t_star], ll_row=self.ll[t_star])
        else:
            kw = dict(warm_src=self.alpha[t_star])
        ops.boundary_pack_fwd(K, self._xbuf_f, first, last, scale=self.scale, **kw)
        g = sh.neighbour_gather(self._xbuf_f)
        from_left = g[sh.rank - 1, 4 * K:] if not sh.is_first else None
        from_right = g[sh.rank + 1, :4 * K] if not sh.is_last else None
        # what arrives: seam truth + warm start of chain 0; the neighbour's first message = the row behind the block
        # (compact: alpha[0,:] and the scalar a1s with alpha[1,x] = a1s * exp2(s*log2e*(ll[x] - max ll)))
        stop = {}
        if from_right is not None:
            stop = (dict(ax_stop=self.ax[self.core.stop], ll_stop=self.ll[self.core.stop]) if compact
                    else dict(alpha_stop=self.alpha[self.core.stop]))
        ops.boundary_unpack_fwd(K, from_left, from_right, self.fwd_end_ext[0],
                                self.fwarm[nxt][0] if nxt is not None else None, scale=self.scale, **stop)

    def _exchange_bwd(self, nxt=None):
        """After a backward pass: beta at the first core bin goes left (seam truth of the neighbour's last
        chain) together with this pass's message at the bin where that chain's next warm-up starts."""
        sh = self.shard
        if not sh.active:
            return
        K2 = 2 * self.K
        buf = self._xbuf_b
        buf[:K2].copy_(self.beta_end[0].reshape(-1))
        if nxt is not None:
            buf[K2:].copy_(self.bwarm[nxt][0].reshape(-1))
        g = sh.neighbour_gather(buf)
        if not sh.is_last:
            from_right = g[sh.rank + 1]
            self.beta_end[self.S].copy_(from_right[:K2].view(2, self.K))
            if nxt is not None:
                self.bwarm[nxt][self.S].copy_(from_right[K2:].view(2, self.K))

    def _check_fwd(self, compact, lo, fix, counter, err=None):
        """Forward seams lo..S-1 (seam c sits in front of chain c): warmed-up estimate vs the true message.
        fix: failed estimates are replaced by the truth (the snapshot a mode-2 restart starts from)."""
        S, K2 = self.S, 2 * self.K
        if lo >= S:
            return
        err = self.err if err is None else err
        est = self.halo_state
        if compact:
            ops.seam_check_fix(S - lo, K2, est[lo].data_ptr(), K2, self.fwd_end_ext[lo].data_ptr(), K2,
                               err[lo:S], self.seam_tol, fix, counter)
            return
        if lo == 0:         # the left neighbour rank's message
            ops.seam_check_fix(1, K2, est[0].data_ptr(), K2, self.fwd_end_ext[0].data_ptr(), K2, err[0:1],
                               self.seam_tol, fix, counter)
            lo = 1
        if lo < S:          # true message in front of chain c >= 1: row t_begin(c) - 1 of the filtered posterior
            row = self.alpha[self.core.start + lo * self.chunk_len - 1]
            ops.seam_check_fix(S - lo, K2, est[lo].data_ptr(), K2, row.data_ptr(), self.chunk_len * K2,
                               err[lo:S], self.seam_tol, fix, counter)

    def _check_bwd(self, hi, fix, counter, err=None):
        """Backward seams 0..hi-1 (seam c sits behind chain c; its truth is the next chain's first message)."""
        S, K2 = self.S, 2 * self.K
        err = self.err if err is None else err
        if hi > 0:
            ops.seam_check_fix(hi, K2, self.beta_halo.data_ptr(), K2, self.beta_end[1].data_ptr(), K2,
                               err[S:S + hi], self.seam_tol, fix, counter)

    def _truth_fwd(self, ids, compact):
        """true forward messages in front of the chains `ids` (host repair sweeps)"""
        if compact:
            return self.fwd_end_ext[ids]
        rows = self.core.start + ids * self.chunk_len - 1
        out = self.alpha[rows.clamp_min(0)]
        if self.f_lo == 0:
            out = torch.where((ids == 0).view(-1, 1, 1), self.fwd_end_ext[0].unsqueeze(0), out)
        return out

    def _lml_to_tail(self, lmr):
        if self._sum_ws is None:
            self._sum_ws = ops.strided_sum_workspace(self.dev)
        ops.strided_sum(lmr[self.core], self.tail[T_LML:T_LML + 1], self._sum_ws)

    def _verdict_enqueue(self, before_sync=None):
        """Enqueues what brings the record (and the seam errors) to the host.  On time-sharded runs the record is
        summed over ranks -- inside the caller's collective when ``before_sync`` performs one (it must all-reduce
        ``self.tail``), else here."""
        if before_sync is not None:
            before_sync()
        elif self.shard.active:
            self.shard.allreduce_flat_sum_(self.tail)
        self.tail_host.copy_(self.tail, non_blocking=True)
        self._errs_host.copy_(self._errs, non_blocking=True)

    def _verdict_wait(self):
        """The ONE synchronisation of an E-step."""
        torch.cuda.current_stream().synchronize()
        t = self.tail_host
        return self.err_host, bool(not (t[T_FAIL_F] == 0)), bool(not (t[T_FAIL_B] == 0))

    def _verdict(self, before_sync=None):
        self._verdict_enqueue(before_sync)
        return self._verdict_wait()

    def _replay_or_capture(self, key, enqueue):
        """Steady-state E-steps are a fixed launch sequence on fixed buffers: the second time a configuration
        (buffer parities, warm-up plan) comes up it is captured into a CUDA graph, from then on replayed -- one
        launch instead of ~30 (the host, not the GPU, bounds short recordings and time-sharded ranks otherwise)."""
        g = self._graphs.get(key)
        if g is None:
            if key not in self._graph_seen:
                self._graph_seen.add(key)
                enqueue()                                   # first sight: run it the plain way
                return
            if len(self._graphs) >= 16:                     # a drifting warm-up plan: stop hoarding graph memory
                self._graphs.clear()
            g = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(g, capture_error_mode="thread_local"):    # (host I/O threads may be busy)
                enqueue()
            g.n_kernels = ops.LAUNCHES - n0                 # this library's kernels among the graph's nodes
            self._graphs[key] = g
        else:
            ops._count(g.n_kernels)                         # replayed launches are launches
        g.replay()

    def _adapt(self, n_fail, failed_f, failed_b):
        """Plans the warm-ups of the pass after next from this pass's record.  n_fail: seams repaired anywhere
        (global: the common base must be the same on every rank, neighbours exchange messages at positions derived
        from it); failed_f / failed_b: this rank's chains whose forward / backward seam failed."""
        cur, nxt = self.halos
        new = nxt
        if not self.adaptive:
            self.halos = [nxt, new]
            return
        import numpy as np
        S = self.S
        S_tot = S * self.shard.world
        mass = n_fail > 0.25 * S_tot                          # e.g. a nearly flat model: nothing forgets quickly
        # A repair pass re-runs the failing chains alone: ~one chunk of scan steps.  A warm-up step is paid by every
        # chain.  While chunks are short compared with the warm-up, a few repairs per pass are the cheaper side.
        few = n_fail <= max(1, S_tot // 100)
        cheap = self.chunk_len <= 2 * nxt
        if n_fail == 0 or (few and cheap):
            self._calm += 1
        else:
            self._calm = 0
        self._streak = self._streak + 1 if (n_fail > 0 and not cheap) else 0
        if self._hold:
            self._hold -= 1
        if (mass or self._streak >= 4) and nxt < self.halo:
            # expensive repairs in four passes running (the per-chain boosts did not absorb them): this base is too
            # short for the model as it is now -- go back up, wait before trying again (early EM iterations change the
            # model quickly), and after three such bounces stop coming down this far
            new = min(self.halo, 2 * nxt)
            self._bounces += 1
            self._hold = 6 * self._bounces
            if self._bounces >= 3:
                self.halo_min = max(self.halo_min, new)
            self._streak = 0
        elif (self._calm >= 2 and self._hold == 0 and nxt > self.halo_min
              and (n_fail == 0 or self.chunk_len <= nxt)):
            new = max(self.halo_min, nxt // 2)
            self._calm = 0
        c, n = self.h_cur, 1 - self.h_cur
        for failed, boost, host, dev in ((failed_f, self.boost_f, self.hf_host, self.hf),
                                         (failed_b, self.boost_b, self.hb_host, self.hb)):
            bumped = False
            if failed.size and not mass:
                # the chains that failed double the warm-up they had in this pass -- for the next pass too (its
                # warm-start message was taken for the shorter warm-up: it starts from a message of a nearby bin)
                want = np.minimum(self.boost_cap, 2 * host[c][failed]).astype(np.int32)
                boost[failed] = np.maximum(boost[failed], want)
                bumped = bool((host[n][failed] < boost[failed]).any())
                host[n][failed] = np.maximum(host[n][failed], boost[failed])
            if bumped:
                dev[n].copy_(torch.from_numpy(host[n]))
            plan2 = np.maximum(np.int32(new), boost).astype(np.int32)       # the pass after next
            if not np.array_equal(plan2, host[c]):
                host[c][:] = plan2
                dev[c].copy_(torch.from_numpy(host[c]))
        self.h_cur = n
        self.halos = [nxt, new]

    def run(self, tuning, want_gamma=False, want_gamma_lat=True, want_dyn=False, want_r=False, gamma16=None,
            before_sync=None, forward_only=False, graph_ok=False, want_tw=None, xi16_ok=False):
        """One E-step.  gamma16: optional [2,T,ldg] fp16 buffer (T = local bins incl. halos) that receives
        the hi/lo pieces of the latent posterior.  before_sync: optional callable invoked once both passes, the
        device repairs and the seam checks are enqueued, before the launching thread waits for the verdict (work
        enqueued there keeps the GPU busy during the synchronisation; on time-sharded runs it must all-reduce
        ``self.tail`` with its own data; ``res.repaired`` tells whether chains were re-run by the HOST after it,
        i.e. whether what it read from this E-step's outputs was final).
        forward_only: emission + filter only (log marginal, one-step predictive marginals): what the batched callers
        of the reference (model selection, shuffle tests) read from decode_latent.
        want_tw: per-chain sums of the latent posterior (``res.tw``; general kernels only).  Default: only when no
        ``gamma16`` is requested -- with the fp16 pieces the statistics GEMM returns sum_t gamma through the column
        of ones of the counts.
        xi16_ok: with ``want_r``, the caller accepts the transition-count operands as bf16 hi/lo pieces
        (``res.xi16`` [4, T, 2K]: alpha hi/lo, r hi/lo; ``res.r_ext`` is then None) where the backward kernel can
        write them itself -- two passes over [T,2K] arrays less than splitting the fp32 arrays afterwards.
        graph_ok: the caller vouches that ``tuning``, ``gamma16`` and everything ``before_sync`` touches are fixed
        buffers and that ``before_sync`` only enqueues GPU work (no Python state): the launch sequence may then be
        captured into a CUDA graph and replayed."""
        S, K = self.S, self.K
        if forward_only:
            want_gamma = want_gamma_lat = want_dyn = want_r = False
        if want_tw is None:
            want_tw = gamma16 is None
        f32 = dict(dtype=torch.float32, device=self.dev)
        # EM fast path: only the fp16 posterior pieces and sum_t gamma are wanted -> compact kernels
        compact = (self.compact_ok and (gamma16 is not None or forward_only)
                   and not (want_gamma or want_gamma_lat or want_dyn or want_r))
        self.plan.halo, self.plan.halo_next = int(self.halos[0]), int(self.halos[1])
        gamma = torch.empty((self.T, 2, K), **f32) if want_gamma else None
        gamma_lat = torch.empty((self.T, K), **f32) if want_gamma_lat else None
        dyn = torch.empty((self.T, 2), **f32) if want_dyn else None
        r = xi16 = None
        if want_r and xi16_ok and not compact and ops.xi16_supported(self.op, self.scale):
            # every row the count GEMM reads is written by the pass (alpha: core rows; r: rows core.start+1 ..)
            xi16 = torch.empty((4, self.T, 2 * K), dtype=torch.bfloat16, device=self.dev)
        elif want_r:      # rows core.start+1 .. core.stop are written by the backward pass; row core.start is unused
            r = torch.empty((self.T, 2, K), **f32)
            r[:self.core.start + 1].zero_()
            r[self.core.stop:].zero_()

        cur, nxt = self.warm_cur, 1 - self.warm_cur
        f_in = self.fwarm[cur] if self.warm_valid else getattr(self.op, "stationary", None)
        b_in = self.bwarm[cur][1:] if self.warm_valid else None
        b_out = self.bwarm[nxt][1:]
        err_f, err_b = self.err1[0:S], self.err1[S:2 * S]           # first check: selects the device repairs
        tol = self.seam_tol
        hf = (self.hf[self.h_cur], self.hf[1 - self.h_cur]) if self.adaptive else (None, None)
        hb = (self.hb[self.h_cur], self.hb[1 - self.h_cur]) if self.adaptive else (None, None)
        # longest warm-up of this pass (the lockstep scan runs that many + chunk_len steps)
        hmax_f = int(self.hf_host[self.h_cur].max()) if (self.adaptive and self.dense is not None) else 0
        hmax_b = int(self.hb_host[self.h_cur].max()) if (self.adaptive and self.dense is not None) else 0

        def fwd(mode=0, ids=None):
            sel = dict(sel_err=err_f, sel_tol=tol) if mode == 2 else {}
            ops.set_chain_halos(self.plan, *hf)
            if compact:
                ops.forward_compact(self.plan, self.op, self.ll, self.ax, carry_in=self.carry_in,
                                    halo_state=(self.halo_state if mode == 0 else None), fwd_end=self.fwd_end,
                                    first_out=self.first_out, mode=mode, chain_ids=ids,
                                    warm_in=(f_in if mode == 0 else self.halo_state), warm_out=self.fwarm[nxt], **sel)
                return
            ops.forward(self.plan, self.op, self.ll, self.alpha, self.lmr, carry_in=self.carry_in,
                        halo_state=(self.halo_state if mode == 0 else None), mode=mode, chain_ids=ids,
                        warm_in=(f_in if mode == 0 else self.halo_state), warm_out=self.fwarm[nxt],
                        halo_max=hmax_f, **sel)

        def bwd(mode=0, ids=None):
            sel = dict(sel_err=err_b, sel_tol=tol) if mode == 2 else {}
            ops.set_chain_halos(self.plan, *hb)
            if compact:
                ops.backward_compact(self.plan, self.op, self.ll, self.ax, gamma16, beta_halo=self.beta_halo,
                                     beta_end=self.beta_end, mode=mode, chain_ids=ids,
                                     warm_in=(b_in if mode == 0 else self.beta_halo), warm_out=b_out, **sel)
                return
            ops.backward(self.plan, self.op, self.ll, self.alpha, gamma=gamma, gamma_lat=gamma_lat, dyn_marg=dyn,
                         r_out=r, tw_partial=(self.tw_partial if want_tw else None), beta_halo=self.beta_halo,
                         beta_end=self.beta_end,
                         mode=mode, chain_ids=ids, gamma16=gamma16, xi16=xi16,
                         warm_in=(b_in if mode == 0 else self.beta_halo), warm_out=b_out, halo_max=hmax_b, **sel)

        seams = S > 1 or self.shard.active
        lmr = self.ax[:, K + 1] if compact else self.lmr

        def enqueue():
            self.emission(tuning)
            ops.phase("emission")
            self.tail.zero_()
            # ---- forward: all chains; seams between this rank's own chains are verified and repaired on the
            # device (conditional relaunch) before anything consumes the filtered posterior; the boundary seam to
            # the left neighbour rank is verified after the exchange (its repair, rare, is the host's)
            fwd()
            if S > 1 and self.device_repair:
                self._check_fwd(compact, 1, True, self.tail[T_FIX_F:T_FIX_F + 1], err=self.err1)
                fwd(mode=2)
            self._exchange_fwd(compact, nxt)
            if seams:
                self._check_fwd(compact, self.f_lo, False, self.tail[T_FAIL_F:T_FAIL_F + 1])
            self._lml_to_tail(lmr)
            ops.phase("forward")
            # ---- backward, same structure
            if not forward_only:
                bwd()
                if S > 1 and self.device_repair:
                    self._check_bwd(S - 1, True, self.tail[T_FIX_B:T_FIX_B + 1], err=self.err1)
                    bwd(mode=2)
                self._exchange_bwd(nxt)
                if seams:
                    self._check_bwd(self.b_hi, False, self.tail[T_FAIL_B:T_FAIL_B + 1])
                ops.phase("backward")
            self._verdict_enqueue(before_sync)

        if (graph_ok and self.use_graphs and compact and self.warm_valid and not forward_only and seams
                and ops.PHASE_HOOK is None):
            key = (cur, self.h_cur if self.adaptive else 0, int(self.halos[0]), int(self.halos[1]),
                   tuning.data_ptr(), gamma16.data_ptr())
            self._replay_or_capture(key, enqueue)
        else:
            enqueue()

        n_relay_f = n_relay_b = 0
        host_bad_f, host_bad_b = [], []
        err, any_f, any_b = self._verdict_wait()
        n_fix_f, n_fix_b = int(self.tail_host[T_FIX_F]), int(self.tail_host[T_FIX_B])
        # seams still failing after the device repairs, over ALL ranks (what the warm-up planning may depend on:
        # the common base must come out the same everywhere)
        n_left = int(self.tail_host[T_FAIL_F]) + int(self.tail_host[T_FAIL_B])
        repaired = bool(any_f or any_b)
        # Host repair = parallel (Jacobi) sweeps: every chain whose incoming message is still off restarts,
        # all at once (on all ranks), from a snapshot of its neighbour's current boundary message; the
        # seams are then re-verified against the messages those restarts produced.  Each sweep extends
        # the effective warm-up by one chunk, so the number of sweeps is ~ mixing length / chunk length.
        # One sweep = forward restarts, then the backward pass of (a) the chains whose filtered posterior
        # just changed -- from their own, already verified, incoming beta -- and (b) the chains whose
        # incoming beta was off -- from the neighbour's; other chains' backward results do not depend on
        # the re-run chains (beta does not involve alpha; normalisers are scale only).  One verdict
        # (synchronisation + collective) per sweep.
        for _ in range(self.max_sweeps):
            if not (any_f or any_b):
                break
            ef, eb = err[self.f_lo:S], err[S:S + self.b_hi]
            bad_f = torch.nonzero(~(ef <= tol)).flatten() + self.f_lo
            bad_b = torch.nonzero(~(eb <= tol)).flatten()
            n_relay_f += int(bad_f.numel())
            n_relay_b += int(bad_b.numel())
            host_bad_f.append(bad_f)
            host_bad_b.append(bad_b)
            if bad_f.numel():
                idl = bad_f.to(device=self.dev)
                self.halo_state[idl] = self._truth_fwd(idl, compact)          # carry snapshot = new "estimate"
                fwd(mode=1, ids=idl.to(torch.int32))
            self._exchange_fwd(compact, nxt)
            if bad_b.numel():
                idb = bad_b.to(device=self.dev)
                self.beta_halo[idb] = self.beta_end[idb + 1]
            both = torch.unique(torch.cat([bad_f, bad_b]))
            if both.numel() and not forward_only:
                bwd(mode=1, ids=both.to(device=self.dev, dtype=torch.int32))
            if not forward_only:
                self._exchange_bwd(nxt)
            self.tail.zero_()
            self._check_fwd(compact, self.f_lo, False, self.tail[T_FAIL_F:T_FAIL_F + 1])
            if not forward_only:
                self._check_bwd(self.b_hi, False, self.tail[T_FAIL_B:T_FAIL_B + 1])
            self._lml_to_tail(lmr)
            err, any_f, any_b = self._verdict()
        if any_f or any_b:
            raise RuntimeError("E-step: seams still above seam_tol=%g after %d repair sweeps (worst forward %.3g, "
                               "backward %.3g)" % (tol, self.max_sweeps, float(err[self.f_lo:S].max()) if S > self.f_lo
                                                   else 0.0, float(err[S:S + self.b_hi].max()) if self.b_hi else 0.0))
        if seams and not forward_only:           # (a forward-only pass leaves no backward warm starts behind)
            self.warm_cur, self.warm_valid = nxt, True
        halo_used = self.halos[0]
        if self.adaptive:
            n_fail = n_fix_f + n_fix_b + n_left
            if n_fail:
                e1 = self.err1_host
                fail_f = torch.cat([torch.nonzero(~(e1[1:S] <= tol)).flatten() + 1] + host_bad_f).unique().numpy()
                fail_b = torch.cat([torch.nonzero(~(e1[S:2 * S - 1] <= tol)).flatten()] + host_bad_b).unique().numpy()
            else:
                fail_f = fail_b = _NO_CHAINS
            self._adapt(n_fail, fail_f, fail_b)
        else:
            self._adapt(0, None, None)

        c = self.core
        ef, eb = err[self.f_lo:S], err[S:S + self.b_hi]
        res = EStepResult()
        res.core = c
        res.ll, res.lmr = self.ll[c], lmr[c]
        res.alpha = None if compact else self.alpha[c]
        res.alpha_ext, res.r_ext = (None if compact else self.alpha), r
        res.xi16 = xi16
        res.gamma = gamma[c] if gamma is not None else None
        res.gamma_lat = gamma_lat[c] if gamma_lat is not None else None
        res.dyn_marg = dyn[c] if dyn is not None else None
        res.r = r[c] if r is not None else None
        # local sums; the caller all-reduces them together with the spike-weighted statistics
        # (the compact path leaves sum_t gamma to the statistics GEMM: ones column of the fp16 counts)
        res.tw = (None if (compact or forward_only or not want_tw)
                  else self.tw_partial.sum(dim=0, dtype=torch.float64).to(torch.float32))
        # global (summed over ranks) log marginal: a host scalar, it came with the verdict
        res.log_marginal = self.tail_host[T_LML].clone()
        res.n_relay_fwd, res.n_relay_bwd = n_relay_f, n_relay_b
        res.n_fix_fwd, res.n_fix_bwd = n_fix_f, n_fix_b               # repaired on the device (all ranks)
        res.repaired = repaired                                       # host sweeps ran; same verdict on every rank
        res.seam_err_fwd = float(ef.max()) if ef.numel() else 0.0
        res.seam_err_bwd = float(eb.max()) if eb.numel() else 0.0
        res.plan = self.plan
        res.halo = halo_used
        return res
