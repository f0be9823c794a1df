"""JAX-compatible counter-based PRNG (threefry2x32, 20 rounds) on NumPy.

Restates the bit streams of ``jax.random`` as of jax 0.4.26 with its defaults
(``jax_default_prng_impl = threefry2x32``, ``jax_threefry_partitionable = False``) so that the same
``key`` gives the same draws as the reference's ``init_latent_posterior`` (core.py:579),
``initialize_params`` (:430) and ``sample*`` (:526-569, :794-800) — SURVEY.md section 8(f) F1.

  PRNGKey(seed)        -> uint32[2] = [seed >> 32, seed & 0xffffffff]
  split(key, n)        -> threefry(key, iota(2n)) reshaped [n, 2]
  random_bits(key, n)  -> counts iota(n), padded to even, first half / second half fed as the two input
                          words of a block; the first output words of all blocks, then the second ones
  uniform float32      -> bitcast((bits >> 9) | 0x3f800000) - 1, then * (hi - lo) + lo, clamped below at lo
  normal float32       -> sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1)), erfinv = XLA's single-precision
                          polynomial (Giles 2010)

The threefry core is checked against the Random123 known-answer vectors (tests/test_host_cpu.py).  JAX itself
is not installable in this image, so the layers above the core (counter layout, bits -> float, erfinv) are
restated from the upstream sources and could not be compared with a live JAX; ``log`` / ``erfinv`` may differ
from XLA's by an ulp.  The device version of ``uniform`` for the [T, K] posterior initialisation is
``pmg_threefry_posterior_init`` (csrc/pmg_prng.cu), tested against this module.
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """20-round threefry2x32 on uint32 arrays (broadcasting); returns the two output words."""
    with np.errstate(over="ignore"):
        k0, k1 = np.uint32(k0), np.uint32(k1)
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = np.asarray(x0, dtype=np.uint32) + ks[0]
        x1 = np.asarray(x1, dtype=np.uint32) + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + np.uint32(i + 1)
    return x0, x1


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def as_key(key):
    """int seed, uint32[2] key (NumPy / jax-like / torch) -> uint32[2]."""
    if key is None:
        return PRNGKey(0)
    a = np.asarray(key)
    if a.ndim == 0:
        return PRNGKey(int(a))
    a = a.astype(np.uint64).ravel()
    if a.size != 2:
        raise ValueError("a PRNG key is an int seed or a uint32[2] array, got shape %s" % (np.asarray(key).shape,))
    return a.astype(np.uint32)


def random_bits(key, n):
    """n uint32 words in jax's (non-partitionable) threefry order."""
    key = as_key(key)
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype=np.uint32)
    half = (n + 1) // 2
    c0 = np.arange(half, dtype=np.uint32)
    c1 = np.arange(half, 2 * half, dtype=np.uint32)
    if 2 * half > n:
        c1[-1] = 0            # the pad element of an odd-sized count array
    y0, y1 = threefry2x32(key[0], key[1], c0, c1)
    return np.concatenate([y0, y1])[:n]


def split(key, num=2):
    return random_bits(key, 2 * int(num)).reshape(int(num), 2)


def bits_to_unit_float(bits):
    """uint32 -> float32 in [0, 1) from the top 23 bits."""
    return ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    f = bits_to_unit_float(random_bits(key, n))
    lo, hi = np.float32(minval), np.float32(maxval)
    out = np.maximum(lo, f * (hi - lo) + lo)
    return out.reshape(shape)


def erfinv_f32(x):
    """XLA's float32 erf_inv (Giles, "Approximating the erfinv function", 2010)."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log((np.float32(1.0) - x) * (np.float32(1.0) + x)).astype(np.float32)
    small = w < np.float32(5.0)
    ws = (w - np.float32(2.5)).astype(np.float32)
    wl = (np.sqrt(np.maximum(w, np.float32(5.0))) - np.float32(3.0)).astype(np.float32)
    cs = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
          -0.00417768164, 0.246640727, 1.50140941)
    cl = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
          0.00943887047, 1.00167406, 2.83297682)
    ps = np.full_like(ws, np.float32(cs[0]))
    for c in cs[1:]:
        ps = (np.float32(c) + ps * ws).astype(np.float32)
    pl = np.full_like(wl, np.float32(cl[0]))
    for c in cl[1:]:
        pl = (np.float32(c) + pl * wl).astype(np.float32)
    out = (np.where(small, ps, pl) * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.float32(np.inf) * x, out)


def normal(key, shape=()):
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = uniform(key, shape, minval=lo, maxval=1.0)
    return (np.float32(np.sqrt(2)) * erfinv_f32(u)).astype(np.float32)
