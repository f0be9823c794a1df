"""Transition kernels and the GP basis (host side, built once per fit / model).

Mirrors the reference's ``poor_man_gplvm/gp_kernel.py:42-89``
(``create_transition_prob_1d``) and ``poor_man_gplvm/core.py:41-73``
(``generate_basis``) with the same names and return order.  These are K x K
objects built once; they are computed in NumPy fp32 and uploaded.
"""
from __future__ import annotations

import numpy as np

_TAP_FLOOR = 1e-37   # taps below this (fp32 denormal range) are outside the band


def rbf_kernel_matrix(n, lengthscale, var=1.0):
    """reference gp_kernel.py:14-20 on the integer grid: exp(-(x-y)^2 / ls^2) * var (no 1/2)."""
    x = np.arange(n, dtype=np.float32)
    d2 = (x[:, None] - x[None, :]) ** 2
    ls = np.float32(lengthscale)
    lin = np.exp(-d2 / ls ** 2) * np.float32(var)
    log = -d2 / ls ** 2 + np.log(np.float32(var))
    return lin.astype(np.float32), log.astype(np.float32)


def create_transition_prob_1d(possible_latent_bin, possible_dynamics=None, movement_variance=1.,
                              p_move_to_jump=0.01, p_jump_to_move=0.01, custom_kernel=None):
    """reference gp_kernel.py:42-89.  Returns (latent_transition_kernel_l [2,K,K],
    log_latent_transition_kernel_l [2,K,K], dynamics_transition_kernel [2,2],
    log_dynamics_transition_kernel [2,2]) as float32 NumPy arrays."""
    K = len(possible_latent_bin)
    if custom_kernel is None:
        lin0, log0 = rbf_kernel_matrix(K, movement_variance, 1.0)
    else:
        lin0 = np.asarray(custom_kernel, dtype=np.float32)
        with np.errstate(divide="ignore"):
            log0 = np.log(lin0)
        log0 = np.where(log0 == np.inf, np.float32(-10000.0), log0).astype(np.float32)
    lin1 = np.full((K, K), np.float32(1.0) / np.float32(K), dtype=np.float32)
    log1 = np.log(lin1)
    P, logP = [], []
    for lin, lg in ((lin0, log0), (lin1, log1)):
        z = lin.sum(axis=1, keepdims=True)
        P.append(lin / z)
        logP.append(lg - np.log(z))
    M = np.array([[1 - p_move_to_jump, p_move_to_jump],
                  [p_jump_to_move, 1 - p_jump_to_move]], dtype=np.float32)
    with np.errstate(divide="ignore"):
        logM = np.log(M)
    return np.stack(P).astype(np.float32), np.stack(logP).astype(np.float32), M, logM.astype(np.float32)


def generate_basis(lengthscale, n_latent_bin, explained_variance_threshold_basis=0.999, include_bias=True,
                   basis_type='rbf', custom_kernel=None):
    """reference core.py:41-73: RBF Gram -> SVD -> truncate at the explained singular-value
    fraction -> scale columns by S^(1/4) -> prepend a ones column."""
    if custom_kernel is not None:
        basis_type = 'custom_kernel'
    if basis_type == 'rbf':
        gram, _ = rbf_kernel_matrix(n_latent_bin, lengthscale, 1.0)
    elif basis_type == 'custom_kernel':
        assert custom_kernel is not None, "custom_kernel must be provided when basis_type is custom_kernel"
        gram = np.asarray(custom_kernel, dtype=np.float32)
    else:
        raise ValueError("unsupported basis_type %r" % (basis_type,))
    U, S, _ = np.linalg.svd(gram)
    n_basis = int((np.cumsum(S / S.sum()) < explained_variance_threshold_basis).sum()) + 1
    basis = U[:, :n_basis] * np.sqrt(np.sqrt(S))[:n_basis][None, :]
    if include_bias:
        basis = np.concatenate([np.ones((n_latent_bin, 1), dtype=np.float32), basis], axis=1)
    return np.ascontiguousarray(basis, dtype=np.float32)


def tap_floor(n_latent_bin, p_move_to_jump):
    """Entries of the move kernel below this value (relative to its diagonal) are treated as zero.

    Every joint transition row carries at least p_move_to_jump/K of mass on every (jump, x') state,
    so a move-kernel entry below 1e-6 of that floor perturbs no posterior entry by more than fp32
    rounding; the floor is additionally capped at 1e-11.  With p_move_to_jump == 0 nothing is
    dropped above the fp32 normal range."""
    if p_move_to_jump is None or p_move_to_jump <= 0:
        return _TAP_FLOOR
    return max(_TAP_FLOOR, min(1e-11, 1e-6 * float(p_move_to_jump) / float(n_latent_bin)))


def move_operator_host(n_latent_bin, movement_variance=1., custom_kernel=None, p_move_to_jump=None):
    """Factor the "move" transition P0 for the scan kernels.

    Default RBF kernel: P0[x,x'] = taps[|x-x'|] * inv_z[x] (Toeplitz numerator, row
    normaliser), band half width W = last tap above the fp32 normal range.
    Custom kernel: band storage of the row-normalised matrix.
    Returns a dict of host arrays; see pmg_transition in include/pmgplvm_b200.h.
    """
    K = int(n_latent_bin)
    if custom_kernel is None:
        d = np.arange(K, dtype=np.float32)
        taps_full = np.exp(-(d ** 2) / np.float32(movement_variance) ** 2).astype(np.float32)
        lin0, _ = rbf_kernel_matrix(K, movement_variance, 1.0)
        z = lin0.sum(axis=1)
        nz = np.nonzero(taps_full >= tap_floor(K, p_move_to_jump))[0]
        W = int(nz.max()) if nz.size else 0
        return {"kind": 0, "W": W, "taps": np.ascontiguousarray(taps_full[:W + 1]),
                "inv_z": (np.float32(1.0) / z).astype(np.float32)}
    lin0 = np.asarray(custom_kernel, dtype=np.float32)
    P0 = lin0 / lin0.sum(axis=1, keepdims=True)
    ii, jj = np.nonzero(P0 >= tap_floor(K, p_move_to_jump) * max(float(P0.max()), 1e-30))
    W = int(np.abs(ii - jj).max()) if ii.size else 0
    band_fwd = np.zeros((2 * W + 1, K), dtype=np.float32)
    band_bwd = np.zeros((2 * W + 1, K), dtype=np.float32)
    xs = np.arange(K)
    for j in range(2 * W + 1):
        src = xs - W + j
        ok = (src >= 0) & (src < K)
        band_fwd[j, ok] = P0[src[ok], xs[ok]]     # P0[x'-W+j, x']
        band_bwd[j, ok] = P0[xs[ok], src[ok]]     # P0[x, x-W+j]
    return {"kind": 1, "W": W, "band_fwd": band_fwd, "band_bwd": band_bwd}
