from __future__ import annotations

import os

import numpy as np
import torch

X64 = bool(int(os.environ.get("JAXSHIM_X64", "0")))
FLOAT = torch.float64 if X64 else torch.float32
INT = torch.int64 if X64 else torch.int32
torch.set_default_dtype(FLOAT)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = self.arr.clone()
        out[self.idx] = asarray(v).to(out.dtype) if not isinstance(v, (int, float)) else v
        return out

    def add(self, v):
        out = self.arr.clone()
        out[self.idx] += asarray(v).to(out.dtype) if not isinstance(v, (int, float)) else v
        return out


class Array(torch.Tensor):
    """torch.Tensor with the handful of jax.Array / ndarray spellings the reference uses."""

    @property
    def at(self):
        return _At(self)

    def astype(self, dtype):
        return self.to(_dtype(dtype))

    def dot(self, other):
        return self.__matmul__(other)

    def __matmul__(self, other):
        a, b = self.as_subclass(torch.Tensor), asarray(other).as_subclass(torch.Tensor)
        dt = torch.promote_types(a.dtype, b.dtype)
        return wrap(torch.matmul(a.to(dt), b.to(dt)))

    def __rmatmul__(self, other):
        return asarray(other).__matmul__(self)

    def block_until_ready(self):
        return self

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def item(self):
        return self.detach().as_subclass(torch.Tensor).item()

    def tolist(self):
        return self.detach().as_subclass(torch.Tensor).tolist()

    def __format__(self, spec):
        return format(self.item(), spec) if self.dim() == 0 else repr(self)

    def __hash__(self):
        return id(self)


def _dtype(d):
    if d is None:
        return None
    if isinstance(d, torch.dtype):
        return d
    if d in (float, "float"):
        return FLOAT
    if d in (int, "int"):
        return INT
    if d in (bool, "bool"):
        return torch.bool
    nd = np.dtype(d)
    if nd == np.float64 and not X64:
        return torch.float32
    if nd == np.int64 and not X64:
        return torch.int32
    return getattr(torch, nd.name)


def wrap(t):
    if isinstance(t, Array):
        return t
    return t.as_subclass(Array)


def asarray(x, dtype=None):
    """Anything -> Array with jax's default-dtype rules (float64 -> float32, int64 -> int32 unless X64)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, (bool, np.bool_)):
        t = torch.tensor(bool(x))
    elif isinstance(x, (int, np.integer)):
        t = torch.tensor(int(x), dtype=INT)
    elif isinstance(x, (float, np.floating)):
        with np.errstate(over="ignore"):
            v = np.float64(x) if X64 else np.float32(x)      # jax: python scalars are weakly typed -> overflow to inf
        t = torch.tensor(float(v), dtype=FLOAT)
    else:
        a = np.asarray(x)
        if a.dtype == object:
            a = np.asarray([np.asarray(e) for e in x])
        if a.dtype == np.float64 and not X64:
            a = a.astype(np.float32)
        elif a.dtype == np.float32 and X64:
            a = a.astype(np.float64)
        elif a.dtype == np.int64 and not X64:
            a = a.astype(np.int32)
        elif a.dtype == np.uint64:
            a = a.astype(np.int64)
        elif a.dtype == np.uint32:
            a = a.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(a)).clone()
    if dtype is not None:
        t = t.to(_dtype(dtype))
    return wrap(t)
