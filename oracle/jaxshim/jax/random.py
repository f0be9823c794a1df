"""jax.random stand-in: keys are uint32[2] like JAX's, but the streams come from NumPy generators
(NOT threefry).  Golden runs inject every random array explicitly, so nothing pinned depends on this."""
from __future__ import annotations

import numpy as np

from ._core import asarray, FLOAT, X64


def PRNGKey(seed):
    return asarray(np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.int64))


key = PRNGKey


def _rng(k):
    a = np.asarray(k).astype(np.uint64).ravel()
    return np.random.default_rng([int(v) for v in a])


def split(k, num=2):
    r = _rng(k)
    return asarray(r.integers(0, 2 ** 32, size=(num, 2), dtype=np.int64))


def _fd():
    return np.float64 if X64 else np.float32


def uniform(k, shape=(), dtype=None, minval=0., maxval=1.):
    u = _rng(k).random(size=tuple(shape)).astype(_fd())
    return asarray(u * (maxval - minval) + minval)


def normal(k, shape=(), dtype=None):
    return asarray(_rng(k).standard_normal(size=tuple(shape)).astype(_fd()))


def choice(k, a, shape=(), replace=True, p=None):
    a_np = np.arange(a) if isinstance(a, (int, np.integer)) else np.asarray(a)
    pp = None
    if p is not None:
        pp = np.asarray(p, dtype=np.float64)
        pp = pp / pp.sum()
    return asarray(_rng(k).choice(a_np, size=tuple(shape) if shape else None, replace=replace, p=pp))


def poisson(k, lam, shape=None):
    lam = np.asarray(lam, dtype=np.float64)
    return asarray(_rng(k).poisson(lam, size=shape if shape is not None else lam.shape))


def randint(k, shape, minval, maxval):
    return asarray(_rng(k).integers(minval, maxval, size=tuple(shape)))
