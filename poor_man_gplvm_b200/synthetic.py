"""Synthetic spike-count generators shared by tests and bench.py.

Mirrors the generative story of the reference's ``sample_latent`` / ``sample_y``
(reference core.py:526-569, :794-800) without its JAX PRNG: a +-1 random walk on
``0..K-1`` that jumps to a uniform bin with probability ``p_jump``, Gaussian-bump
tuning curves, Poisson counts.  Spikes are an *input* of the hot path, so parity
does not depend on the sampler.
"""
from __future__ import annotations

import numpy as np


def bump_tuning(n_latent_bin, n_neuron, rng, peak=(0.5, 2.0), floor=0.02, width_frac=0.1):
    K, N = n_latent_bin, n_neuron
    centres = rng.uniform(0, K, size=N)
    peaks = rng.uniform(peak[0], peak[1], size=N)
    x = np.arange(K)[:, None]
    w = max(1.0, width_frac * K)
    return (floor + peaks[None, :] * np.exp(-0.5 * ((x - centres[None, :]) / w) ** 2)).astype(np.float32)


def walk_latent(T, n_latent_bin, rng, p_jump=0.01):
    steps = rng.integers(-1, 2, size=T)
    jumps = rng.random(T) < p_jump
    targets = rng.integers(0, n_latent_bin, size=T)
    lat = np.empty(T, dtype=np.int64)
    cur = int(rng.integers(0, n_latent_bin))
    for t in range(T):
        cur = int(targets[t]) if jumps[t] else min(n_latent_bin - 1, max(0, cur + int(steps[t])))
        lat[t] = cur
    return lat, jumps


def make_dataset(T, n_neuron, n_latent_bin, seed=0, p_jump=0.01):
    """Returns dict(y[T,N] float32 counts, latent[T], tuning_true[K,N])."""
    rng = np.random.default_rng(seed)
    tuning = bump_tuning(n_latent_bin, n_neuron, rng)
    lat, jumps = walk_latent(T, n_latent_bin, rng, p_jump)
    y = rng.poisson(tuning[lat]).astype(np.float32)
    return {"y": y, "latent": lat, "jump": jumps, "tuning_true": tuning}


def make_dataset_torch(T, n_neuron, n_latent_bin, device, seed=0, p_jump=0.01, tuning_seed=None):
    """Device-side generator for the large bench configs (different PRNG stream
    from ``make_dataset``; same distribution).  The walk is generated as a
    cumulative sum of +-1 steps reflected into ``0..K-1`` between jump times.
    tuning_seed: seed of the tuning curves (default: ``seed``).  Time-sharded ranks pass the same
    ``tuning_seed`` and different ``seed``s: blocks of ONE recording (same neurons, same tuning), each
    with its own latent trajectory and spikes."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    K = n_latent_bin
    rng = np.random.default_rng(seed if tuning_seed is None else tuning_seed)
    tuning = torch.from_numpy(bump_tuning(K, n_neuron, rng)).to(device)
    steps = torch.randint(-1, 2, (T,), generator=g, device=device)
    jumps = torch.rand(T, generator=g, device=device) < p_jump
    targets = torch.randint(0, K, (T,), generator=g, device=device)
    # segment-wise cumulative sum restarted at each jump, folded by reflection
    seg = torch.cumsum(jumps.to(torch.int64), 0)
    csum = torch.cumsum(steps.to(torch.int64), 0)
    base_idx = torch.where(jumps, torch.arange(T, device=device), torch.zeros((), dtype=torch.int64, device=device))
    base_idx = torch.cummax(base_idx, 0).values
    start_val = torch.where(seg > 0, targets[base_idx], torch.full((), K // 2, dtype=torch.int64, device=device))
    rel = csum - csum[base_idx] + torch.where(seg > 0, torch.zeros((), dtype=torch.int64, device=device), steps[0].to(torch.int64) * 0)
    pos = start_val + rel
    period = 2 * (K - 1) if K > 1 else 1
    pos = torch.remainder(pos, period)
    lat = torch.where(pos >= K, period - pos, pos)
    y = torch.poisson(tuning[lat], generator=g)
    return {"y": y, "latent": lat, "tuning_true": tuning}
