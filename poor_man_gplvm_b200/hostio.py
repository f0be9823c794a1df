"""Host <-> device transfer helpers for the T-sized result arrays.

The reference returns a mix of host arrays (``np.exp(log_posterior_all)`` and its marginals,
reference core.py:475-477, :688-690) and device-resident jax arrays (``log_posterior_final``, the
saved snapshots, ...).  Here the former become real NumPy arrays, copied through a small ring of
pinned staging buffers with the host-side memcpy (and its first-touch page faults) spread over a
thread pool; the latter become :class:`LazyHostArray` objects that stay on the GPU until something
reads them (``np.asarray(x)``, ``x[...]``, arithmetic through NumPy), like a jax array would.
"""
from __future__ import annotations

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


def _staging():
    global _ring, _pool
    with _lock:
        if _ring is None:
            bufs = [torch.empty(_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(_NBUF)]
            _ring = [(b, b.numpy()) for b in bufs]
            _pool = ThreadPoolExecutor(max_workers=_NBUF)
    return _ring, _pool


def to_numpy(t):
    """Device tensor -> fresh NumPy array (pipelined through pinned staging for large tensors)."""
    if not isinstance(t, torch.Tensor):
        return np.asarray(t)
    if not t.is_cuda:
        return t.detach().numpy()
    t = t.detach().contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (8 << 20):
        return t.cpu().numpy()
    ring, pool = _staging()
    out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    dst = out.reshape(-1).view(np.uint8)
    src = t.view(-1).view(torch.uint8)
    dev = t.device
    cs = _stream.get(dev.index)
    if cs is None:
        cs = _stream[dev.index] = torch.cuda.Stream(device=dev)
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

    def __init__(self, tensor, transform=None, shape=None, producer=None):
        """tensor: device tensor, or None with `producer` (a callable returning the device tensor, run on
        first use) and `shape`."""
        self._tensor = tensor
        self._producer = producer
        self._shape = tuple(shape) if shape is not None else None
        self._f = transform
        self._host = None

    @property
    def _t(self):
        if self._tensor is None:
            self._tensor = self._producer()
        return self._tensor

    def device_tensor(self):
        return self._t if self._f is None else self._f(self._t)

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
