"""Time sharding over ranks (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Rank r owns a contiguous block of time bins.  The only exchanges of the EM hot path are
  * once per fit: ``halo`` rows of the spike matrix from each neighbour (warm-up bins of the scan),
  * per pass: one boundary message (4K floats) to each neighbour (one all-gather of fixed-size buffers),
  * per EM iteration: one all-reduce of the packed sufficient statistics (K*N + K + 1 floats),
  * per seam-repair sweep: one scalar all-reduce (does any rank still have a failing seam?).
The reference has no multi-device code; this layer is new.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class TimeShard:
    def __init__(self, group=None, single=False):
        self.group = group
        self.active = (not single) and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1
        # gloo moves host memory only: device tensors are staged through the host.  That is how several ranks
        # share ONE GPU (the driver's single-GPU test box, tests/test_gpu_multirank.py); on NCCL nothing is staged.
        self.staged = self.active and dist.get_backend(group) == "gloo"

    def _out(self, t):
        """tensor handed to the backend for `t` (host copy of a device tensor under gloo)"""
        return t.detach().cpu().contiguous() if (self.staged and t.is_cuda) else t.contiguous()

    def _recv_like(self, like):
        return torch.empty(like.shape, dtype=like.dtype, device="cpu" if self.staged else like.device)

    def _in(self, t, like):
        return t.to(like.device) if (t is not None and t.device != like.device) else t

    @property
    def is_first(self):
        return self.rank == 0

    @property
    def is_last(self):
        return self.rank == self.world - 1

    def _peer(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _exchange(self, to_left, to_right, like_left, like_right):
        """Send `to_left` to rank-1 and `to_right` to rank+1; receive into fresh tensors shaped like
        `like_left` (from rank-1) and `like_right` (from rank+1).  Any of them may be None."""
        ops, from_left, from_right = [], None, None
        if not self.is_first:
            if to_left is not None:
                ops.append(dist.P2POp(dist.isend, self._out(to_left), self._peer(self.rank - 1), self.group))
            if like_left is not None:
                from_left = self._recv_like(like_left)
                ops.append(dist.P2POp(dist.irecv, from_left, self._peer(self.rank - 1), self.group))
        if not self.is_last:
            if to_right is not None:
                ops.append(dist.P2POp(dist.isend, self._out(to_right), self._peer(self.rank + 1), self.group))
            if like_right is not None:
                from_right = self._recv_like(like_right)
                ops.append(dist.P2POp(dist.irecv, from_right, self._peer(self.rank + 1), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if like_left is not None:
            from_left = self._in(from_left, like_left)
        if like_right is not None:
            from_right = self._in(from_right, like_right)
        return from_left, from_right

    def halo_exchange(self, x, halo):
        """x: [T_core, ...] block of this rank.  Returns (x_ext, h_left, h_right) where x_ext has the
        last `halo` rows of the left neighbour in front and the first `halo` rows of the right one behind."""
        if not self.active or halo <= 0:
            return x, 0, 0
        if x.shape[0] < halo:
            raise ValueError("time block of %d bins is shorter than the halo %d" % (x.shape[0], halo))
        head, tail = x[:halo], x[-halo:]
        from_left, from_right = self._exchange(head, tail, tail, head)
        parts, hl, hr = [], 0, 0
        if from_left is not None:
            parts.append(from_left); hl = halo
        parts.append(x)
        if from_right is not None:
            parts.append(from_right); hr = halo
        return torch.cat(parts, dim=0).contiguous(), hl, hr

    def boundary(self, to_left, to_right):
        """One message to each neighbour (same shape everywhere); returns (from_left, from_right)."""
        if not self.active:
            return None, None
        # every rank calls with the same pattern: something arrives from the left only if ranks send
        # rightwards (to_right given), and from the right only if ranks send leftwards
        like = to_left if to_left is not None else to_right
        return self._exchange(to_left, to_right, like if to_right is not None else None,
                              like if to_left is not None else None)

    def neighbour_gather(self, buf):
        """All ranks' fixed-size message buffers, [world, n] on buf's device: ONE collective per pass -- its host cost
        (a single enqueue) is a fraction of a batched send/recv group, and the payload (8K floats per rank) is noise on
        NVSwitch.  Rank r reads rows r-1 and r+1."""
        n = buf.numel()
        key = (n, buf.device, buf.dtype)
        out = self._gather_out.get(key) if hasattr(self, "_gather_out") else None
        if out is None:
            if not hasattr(self, "_gather_out"):
                self._gather_out = {}
            out = self._gather_out[key] = torch.empty((self.world, n), dtype=buf.dtype,
                                                       device="cpu" if self.staged else buf.device)
        dist.all_gather_into_tensor(out.view(-1), self._out(buf.reshape(-1)), group=self.group)
        return out.to(buf.device) if self.staged else out

    def block_offset(self, T):
        """(first global bin of this rank's block, total number of bins) for a local block of T bins."""
        if not self.active:
            return 0, int(T)
        sizes = [None] * self.world
        dist.all_gather_object(sizes, int(T), group=self.group)
        return int(sum(sizes[:self.rank])), int(sum(sizes))

    def allreduce_sum_(self, *tensors):
        """In-place sum over ranks of several tensors packed into one collective."""
        if not self.active:
            return
        flat = torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])
        flat = self._allreduce(flat, dist.ReduceOp.SUM)
        o = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[o:o + n].reshape(t.shape).to(t.dtype))
            o += n

    def broadcast_(self, t, src=0):
        """In-place broadcast from rank `src` of the group."""
        if self.active:
            if self.staged and t.is_cuda:
                h = t.detach().cpu()
                dist.broadcast(h, src=self._peer(src), group=self.group)
                t.copy_(h)
            else:
                dist.broadcast(t, src=self._peer(src), group=self.group)

    def allreduce_max_(self, t):
        """In-place element-wise maximum over ranks."""
        if self.active:
            t.copy_(self._allreduce(t, dist.ReduceOp.MAX))

    def _allreduce(self, t, op):
        """all-reduce of one tensor; returns the reduced tensor on t's device (t itself unless staged)"""
        if self.staged and t.is_cuda:
            h = t.detach().cpu()
            dist.all_reduce(h, op=op, group=self.group)
            return h.to(t.device)
        dist.all_reduce(t, op=op, group=self.group)
        return t

    def allreduce_flat_sum_(self, flat):
        """In-place sum over ranks of ONE contiguous buffer (no packing copies): the per-iteration collective of
        the EM loop -- statistics, log marginal and seam verdict travel together."""
        if self.active:
            r = self._allreduce(flat, dist.ReduceOp.SUM)
            if r is not flat:
                flat.copy_(r)

    def max_int(self, v, device):
        if not self.active:
            return int(v)
        t = torch.tensor([int(v)], dtype=torch.int64, device="cpu" if self.staged else device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())
