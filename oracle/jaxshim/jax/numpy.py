"""jax.numpy subset on torch CPU tensors."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _core
from ._core import Array, asarray, wrap, FLOAT, INT

ndarray = Array
inf = math.inf
pi = math.pi
nan = math.nan
newaxis = None
float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool_ = np.bool_


def _t(x):
    return asarray(x)


def _pair(a, b):
    a, b = _t(a), _t(b)
    if a.dtype != b.dtype:
        dt = torch.promote_types(a.dtype, b.dtype)
        # weak python scalars do not promote fp32 arrays (handled by asarray); ints with floats -> float
        a, b = a.to(dt), b.to(dt)
    return a, b


def array(x, dtype=None):
    return asarray(x, dtype).clone()


def ones(shape, dtype=None):
    return wrap(torch.ones(shape if not isinstance(shape, int) else (shape,), dtype=_core._dtype(dtype) or FLOAT))


def zeros(shape, dtype=None):
    return wrap(torch.zeros(shape if not isinstance(shape, int) else (shape,), dtype=_core._dtype(dtype) or FLOAT))


def full(shape, v, dtype=None):
    return wrap(torch.full(shape if not isinstance(shape, int) else (shape,), v, dtype=_core._dtype(dtype) or FLOAT))


def zeros_like(x):
    return wrap(torch.zeros_like(_t(x)))


def ones_like(x):
    return wrap(torch.ones_like(_t(x)))


def eye(n, dtype=None):
    return wrap(torch.eye(n, dtype=_core._dtype(dtype) or FLOAT))


def arange(*a, dtype=None):
    if all(isinstance(v, (int, np.integer)) for v in a):
        return wrap(torch.arange(*[int(v) for v in a], dtype=_core._dtype(dtype) or INT))
    return wrap(torch.arange(*a, dtype=_core._dtype(dtype) or FLOAT))


def linspace(a, b, n):
    return wrap(torch.linspace(a, b, n, dtype=FLOAT))


def _un(f):
    def g(x):
        x = _t(x)
        if not x.dtype.is_floating_point:
            x = x.to(FLOAT)
        return wrap(f(x))
    return g


exp = _un(torch.exp)
log = _un(torch.log)
log1p = _un(torch.log1p)
sqrt = _un(torch.sqrt)
ceil = _un(torch.ceil)
floor = _un(torch.floor)
tanh = _un(torch.tanh)
isnan = lambda x: wrap(torch.isnan(_t(x)))
isinf = lambda x: wrap(torch.isinf(_t(x)))
isfinite = lambda x: wrap(torch.isfinite(_t(x)))
abs = lambda x: wrap(torch.abs(_t(x)))
absolute = abs
square = lambda x: wrap(torch.square(_t(x)))


def logaddexp(a, b):
    a, b = _pair(a, b)
    return wrap(torch.logaddexp(a, b))


def maximum(a, b):
    a, b = _pair(a, b)
    return wrap(torch.maximum(a, b))


def minimum(a, b):
    a, b = _pair(a, b)
    return wrap(torch.minimum(a, b))


def add(a, b):
    a, b = _pair(a, b)
    return wrap(a + b)


def multiply(a, b):
    a, b = _pair(a, b)
    return wrap(a * b)


def _axis_kw(axis, keepdims):
    kw = {}
    if axis is not None:
        kw["dim"] = tuple(axis) if isinstance(axis, (tuple, list)) else axis
    if keepdims:
        kw["keepdim"] = True
    return kw


def sum(x, axis=None, keepdims=False):
    x = _t(x)
    if x.dtype == torch.bool:
        x = x.to(INT)
    return wrap(torch.sum(x, **_axis_kw(axis, keepdims)))


def mean(x, axis=None, keepdims=False):
    return wrap(torch.mean(_t(x), **_axis_kw(axis, keepdims)))


def max(x, axis=None, keepdims=False):
    x = _t(x)
    return wrap(torch.amax(x, **_axis_kw(axis, keepdims))) if axis is not None else wrap(x.max())


def min(x, axis=None, keepdims=False):
    x = _t(x)
    return wrap(torch.amin(x, **_axis_kw(axis, keepdims))) if axis is not None else wrap(x.min())


def argmax(x, axis=None):
    return wrap(torch.argmax(_t(x), dim=axis))


def cumsum(x, axis=None):
    x = _t(x)
    return wrap(torch.cumsum(x.reshape(-1) if axis is None else x, dim=0 if axis is None else axis))


def where(c, a, b):
    c = _t(c)
    if c.dtype != torch.bool:
        c = c != 0
    a, b = _pair(a, b)
    return wrap(torch.where(c, a, b))


def concatenate(xs, axis=0):
    ts = [_t(x) for x in xs]
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = torch.promote_types(dt, t.dtype)
    return wrap(torch.cat([t.to(dt) for t in ts], dim=axis))


def stack(xs, axis=0):
    return wrap(torch.stack([_t(x) for x in xs], dim=axis))


def broadcast_to(x, shape):
    return wrap(torch.broadcast_to(_t(x), (shape,) if isinstance(shape, int) else tuple(shape)))


def einsum(spec, *ops):
    ts = [_t(o) for o in ops]
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = torch.promote_types(dt, t.dtype)
    return wrap(torch.einsum(spec, *[t.to(dt) for t in ts]))


def dot(a, b):
    a, b = _pair(a, b)
    return wrap(torch.matmul(a, b))


matmul = dot


def outer(a, b):
    a, b = _pair(a, b)
    return wrap(torch.outer(a, b))


def reshape(x, shape):
    return wrap(_t(x).reshape(shape))


def transpose(x, axes=None):
    x = _t(x)
    return wrap(x.permute(*axes) if axes is not None else x.T)


def expand_dims(x, axis):
    return wrap(_t(x).unsqueeze(axis))


def squeeze(x, axis=None):
    x = _t(x)
    return wrap(x.squeeze() if axis is None else x.squeeze(axis))


def clip(x, lo=None, hi=None):
    return wrap(torch.clamp(_t(x), lo, hi))


def diag(x):
    return wrap(torch.diag(_t(x)))


def tile(x, reps):
    return wrap(_t(x).repeat(*((reps,) if isinstance(reps, int) else reps)))


def allclose(a, b, rtol=1e-5, atol=1e-8):
    a, b = _pair(a, b)
    return bool(torch.allclose(a, b, rtol=rtol, atol=atol))


class linalg:
    @staticmethod
    def svd(a, full_matrices=True):
        # LAPACK gesdd on the host, like jaxlib's CPU backend
        u, s, vh = torch.linalg.svd(_t(a), full_matrices=full_matrices)
        return wrap(u), wrap(s), wrap(vh)

    @staticmethod
    def solve(a, b):
        a, b = _pair(a, b)
        return wrap(torch.linalg.solve(a, b))

    @staticmethod
    def norm(x, ord=None, axis=None):
        x = _t(x)
        if not x.dtype.is_floating_point:
            x = x.to(FLOAT)
        if x.dim() == 0:
            return wrap(torch.abs(x))
        return wrap(torch.linalg.norm(x, ord=ord, dim=axis))

    @staticmethod
    def inv(a):
        return wrap(torch.linalg.inv(_t(a)))

