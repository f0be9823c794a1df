"""Host <-> device transfer helpers for the T-sized result arrays.

The reference returns a mix of host arrays (``np.exp(log_posterior_all)`` and its marginals,
reference core.py:475-477, :688-690) and device-resident jax arrays (``log_posterior_final``, the
saved snapshots, ...).  Here the former become real NumPy arrays, copied through a small ring of
pinned staging buffers with the host-side memcpy (and its first-touch page faults) spread over a
thread pool; the latter become :class:`LazyHostArray` objects that stay on the GPU until something
reads them (``np.asarray(x)``, ``x[...]``, arithmetic through NumPy), like a jax array would.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_CHUNK = 32 << 20
_NBUF = 6
_lock = threading.Lock()
_ring = None
_pool = None
_stream = {}


def host_threads(want):
    """Worker threads this process may use for host-side copies / page population: `want`, scaled down when several
    ranks share the host (8 ranks x (6 staging + 8 populate threads) on one 32-core socket was round 1's e2e collapse
    at 8 GPUs)."""
    try:
        import torch.distributed as dist
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    except Exception:
        world = 1
    local = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    cores = os.cpu_count() or 8
    return max(2, min(int(want), cores // (2 * max(1, local))))


def _staging():
    global _ring, _pool
    with _lock:
        if _ring is None:
            bufs = [torch.empty(_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(_NBUF)]
            _ring = [(b, b.numpy()) for b in bufs]
            _pool = ThreadPoolExecutor(max_workers=host_threads(_NBUF))
    return _ring, _pool


def _copy_stream(dev):
    cs = _stream.get(dev.index)
    if cs is None:
        cs = _stream[dev.index] = torch.cuda.Stream(device=dev)
    return cs


def to_device(a, device):
    """NumPy array -> device tensor of the same dtype.  Large pageable arrays go through the pinned staging
    ring: worker threads copy 32 MB pieces into pinned buffers while earlier pieces are in flight on a copy
    stream (a plain pageable cudaMemcpy reaches ~11 GB/s on the B200 box; the ring several times that)."""
    a = np.ascontiguousarray(a)
    if a.nbytes < (8 << 20) or not a.flags.c_contiguous or os.environ.get("PMG_H2D_RING", "1") == "0":
        return torch.as_tensor(a).to(device)
    device = torch.device(device)
    tdtype = torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype
    out = torch.empty(a.shape, dtype=tdtype, device=device)
    src = a.reshape(-1).view(np.uint8)
    dst = out.view(-1).view(torch.uint8)
    nbytes = a.nbytes
    ring, pool = _staging()
    cs = _copy_stream(device)
    cs.wait_stream(torch.cuda.current_stream(device))
    chunks = list(range(0, nbytes, _CHUNK))
    fut = [None] * _NBUF
    ev = [None] * _NBUF

    def fill(b, lo, n):
        if ev[b] is not None:
            ev[b].synchronize()                  # the previous copy out of this staging buffer is done
        np.copyto(ring[b][1][:n], src[lo:lo + n])

    for i in range(len(chunks) + _NBUF - 1):
        if i < len(chunks):
            b, lo = i % _NBUF, chunks[i]
            n = min(_CHUNK, nbytes - lo)
            fut[b] = (pool.submit(fill, b, lo, n), lo, n)
        j = i - (_NBUF - 1)
        if 0 <= j < len(chunks):
            b = j % _NBUF
            f, lo, n = fut[b]
            f.result()
            with torch.cuda.stream(cs):
                dst[lo:lo + n].copy_(ring[b][0][:n], non_blocking=True)
                e = torch.cuda.Event()
                e.record(cs)
            ev[b] = e
    torch.cuda.current_stream(device).wait_stream(cs)
    for e in ev:
        if e is not None:
            e.synchronize()                      # the staging ring may be reused by the next call
    return out


_libc = None
_MADV_HUGEPAGE = 14
_MADV_POPULATE_WRITE = 23                        # Linux >= 5.14: fault the pages in (writable) without touching data
_TOUCH_PIECE = int(os.environ.get("PMG_TOUCH_PIECE_MB", "4")) << 20


def _touch(view):
    """First-touch the pages of a uint8 view on this thread, with the GIL released (ctypes foreign calls).
    The populate calls are issued in small pieces: each holds the process's memory-map lock for reading, and a
    long hold stalls every mmap/munmap of the main thread (allocator traffic of the EM loop) behind it
    (measured on the B200 box, fit_em at the headline size: 64 MB pieces 0.45 s, 4 MB pieces 0.31 s; transparent
    huge pages make the populate slower, not faster)."""
    global _libc
    import ctypes
    if _libc is None:
        _libc = ctypes.CDLL(None, use_errno=True)
    addr = view.ctypes.data
    end = addr + view.size
    lo = addr & ~4095
    if os.environ.get("PMG_TOUCH_HUGEPAGE", "0") != "0":
        _libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(end - lo), ctypes.c_int(_MADV_HUGEPAGE))
    while lo < end:
        n = min(_TOUCH_PIECE, end - lo)
        rc = _libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(n), ctypes.c_int(_MADV_POPULATE_WRITE))
        if rc != 0:                              # older kernel / unsupported mapping: write the bytes instead
            a = max(lo, addr)
            ctypes.memset(ctypes.c_void_p(a), 0, lo + n - a)
        lo += n


class HostBuffers:
    """Result arrays allocated up front with their pages touched on background threads, so that the
    first-touch page faults of the (multi-GB) outputs overlap the EM iterations instead of the final
    device->host copies."""

    def __init__(self, specs, threads=None):
        threads = threads or host_threads(int(os.environ.get("PMG_TOUCH_THREADS", "8")))
        self._pool = ThreadPoolExecutor(max_workers=threads)
        self._arrays, self._futs = {}, {}
        for name, shape, dtype in specs:
            arr = np.empty(shape, dtype=dtype)
            flat = arr.reshape(-1).view(np.uint8)
            self._arrays[name] = arr
            self._futs[name] = [self._pool.submit(_touch, flat[lo:lo + (64 << 20)])
                                for lo in range(0, flat.size, 64 << 20)]

    def take(self, name, shape=None):
        """The prefaulted array (waits for its pages), or None if it was not planned with this shape."""
        arr = self._arrays.pop(name, None)
        if arr is None or (shape is not None and tuple(arr.shape) != tuple(shape)):
            return None
        for f in self._futs.pop(name):
            f.result()
        return arr

    def close(self):
        self._pool.shutdown(wait=False)


def to_numpy(t, out=None):
    """Device tensor -> NumPy array (pipelined through pinned staging for large tensors).
    out: optional preallocated (ideally prefaulted) destination of the same shape and dtype."""
    if not isinstance(t, torch.Tensor):
        return np.asarray(t)
    if not t.is_cuda:
        return t.detach().numpy().copy()          # always a fresh array, like the device path
    t = t.detach().contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (8 << 20):
        return t.cpu().numpy()
    ring, pool = _staging()
    np_dtype = torch.empty(0, dtype=t.dtype).numpy().dtype
    if out is None or tuple(out.shape) != tuple(t.shape) or out.dtype != np_dtype or not out.flags.c_contiguous:
        out = np.empty(tuple(t.shape), dtype=np_dtype)
    dst = out.reshape(-1).view(np.uint8)
    src = t.view(-1).view(torch.uint8)
    dev = t.device
    cs = _copy_stream(dev)
    cs.wait_stream(torch.cuda.current_stream(dev))
    pending = [None] * _NBUF

    def drain(ev, buf_np, lo, n):
        ev.synchronize()
        np.copyto(dst[lo:lo + n], buf_np[:n])

    for i, lo in enumerate(range(0, nbytes, _CHUNK)):
        b = i % _NBUF
        if pending[b] is not None:
            pending[b].result()
        n = min(_CHUNK, nbytes - lo)
        buf, buf_np = ring[b]
        with torch.cuda.stream(cs):
            buf[:n].copy_(src[lo:lo + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        pending[b] = pool.submit(drain, ev, buf_np, lo, n)
    for f in pending:
        if f is not None:
            f.result()
    t.record_stream(cs)
    return out


class LazyHostArray(np.lib.mixins.NDArrayOperatorsMixin):
    """Array-like view of a device tensor that is copied to the host on first use.

    Supports ``np.asarray``, indexing, ``shape/dtype/ndim/size``, ``len`` and NumPy ufuncs (through
    ``__array__``).  ``transform`` (e.g. ``torch.log``) is applied on the device right before the
    copy, so derived quantities cost no device memory until someone asks for them.
    """

    def __init__(self, tensor, transform=None, shape=None, producer=None, device=None):
        """tensor: device tensor, or None with `producer` (a callable returning the device tensor, run on
        first use), `shape` and the `device` the producer launches on."""
        self._tensor = tensor
        self._producer = producer
        self._shape = tuple(shape) if shape is not None else None
        self._f = transform
        self._host = None
        self._device = tensor.device if tensor is not None else device

    def _guard(self):
        """kernels behind the producer / transform launch on the tensor's device, whatever device is current"""
        import contextlib
        d = self._device
        if d is not None and torch.device(d).type == "cuda":
            return torch.cuda.device(d)
        return contextlib.nullcontext()

    @property
    def _t(self):
        if self._tensor is None:
            with self._guard():
                self._tensor = self._producer()
        return self._tensor

    def device_tensor(self):
        if self._f is None:
            return self._t
        t = self._t
        with self._guard():
            return self._f(t)

    def __reduce__(self):
        """Pickles as the materialised NumPy array (the reference's jax arrays pickle to host data too); a live
        device tensor or a local producer would tie the pickle to this process and its GPU."""
        return (np.asarray, (np.array(self.numpy()),))

    def numpy(self):
        if self._host is None:
            self._host = to_numpy(self.device_tensor())
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    @property
    def shape(self):
        return self._shape if self._shape is not None else tuple(self._t.shape)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def dtype(self):
        if self._tensor is None:
            return np.dtype(np.float32)
        return torch.empty(0, dtype=self._t.dtype).numpy().dtype

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if self._host is not None:
            return self._host[idx]
        try:
            sub = self._t[idx]
        except (TypeError, IndexError, RuntimeError):
            return self.numpy()[idx]
        return to_numpy(sub if self._f is None else self._f(sub))

    def item(self):
        return self.numpy().item()

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(x.numpy() if isinstance(x, LazyHostArray) else x for x in inputs)
        return getattr(ufunc, method)(*inputs, **kwargs)

    def __getattr__(self, name):
        # anything else (sum, max, argmax, reshape, T, ...) behaves like the materialised NumPy array
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.numpy(), name)

    def __repr__(self):
        where = self._tensor.device if self._tensor is not None else "device (not generated yet)"
        return "LazyHostArray(shape=%s, dtype=%s, on %s)" % (self.shape, self.dtype, where)
