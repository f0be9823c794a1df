from __future__ import annotations

import torch

from .._core import asarray, wrap, FLOAT


def logsumexp(a, axis=None, keepdims=False, b=None):
    """jax.scipy.special.logsumexp: max-shifted, non-finite max replaced by 0, all -inf -> -inf."""
    a = asarray(a)
    if not a.dtype.is_floating_point:
        a = a.to(FLOAT)
    dims = tuple(range(a.dim())) if axis is None else (tuple(axis) if isinstance(axis, (tuple, list)) else (axis,))
    m = torch.amax(a, dim=dims, keepdim=True)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    s = torch.sum(torch.exp(a - m), dim=dims, keepdim=True)
    out = torch.log(s) + m
    if not keepdims:
        out = out.squeeze(dims) if dims else out
    return wrap(out)


def xlogy(x, y):
    x, y = asarray(x), asarray(y)
    if not x.dtype.is_floating_point:
        x = x.to(y.dtype if y.dtype.is_floating_point else FLOAT)
    return wrap(torch.xlogy(x, y.to(x.dtype) if y.dtype != x.dtype else y))


def gammaln(x):
    x = asarray(x)
    if not x.dtype.is_floating_point:
        x = x.to(FLOAT)
    return wrap(torch.lgamma(x))
