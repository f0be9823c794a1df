"""optax 0.2.2 subset: adam = scale_by_adam(b1=.9, b2=.999, eps=1e-8, eps_root=0) then scale(-lr)."""
from __future__ import annotations

from collections import namedtuple

import torch

from jax import tree_util
from jax._core import asarray, wrap, INT

ScaleByAdamState = namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
EmptyState = namedtuple("EmptyState", [])
GradientTransformation = namedtuple("GradientTransformation", ["init", "update"])


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        z = lambda p: wrap(torch.zeros_like(asarray(p)))
        return (ScaleByAdamState(count=wrap(torch.zeros((), dtype=INT)), mu=tree_util.tree_map(z, params),
                                 nu=tree_util.tree_map(z, params)), EmptyState())

    def update(grads, state, params=None):
        st = state[0]
        mu = tree_util.tree_map(lambda g, m: wrap((1 - b1) * asarray(g) + b1 * asarray(m)), grads, st.mu)
        nu = tree_util.tree_map(lambda g, v: wrap((1 - b2) * asarray(g) * asarray(g) + b2 * asarray(v)), grads, st.nu)
        count = wrap(asarray(st.count) + 1)
        c = float(count.item())
        fdt = asarray(tree_util.tree_leaves(grads)[0]).dtype
        bc1 = 1 - torch.tensor(b1, dtype=fdt) ** c
        bc2 = 1 - torch.tensor(b2, dtype=fdt) ** c
        mu_hat = tree_util.tree_map(lambda m: asarray(m) / bc1, mu)
        nu_hat = tree_util.tree_map(lambda v: asarray(v) / bc2, nu)
        upd = tree_util.tree_map(lambda m, v: wrap(-learning_rate * (m / (torch.sqrt(v + eps_root) + eps))), mu_hat, nu_hat)
        return upd, (ScaleByAdamState(count=count, mu=mu, nu=nu), EmptyState())

    return GradientTransformation(init, update)


def apply_updates(params, updates):
    return tree_util.tree_map(lambda p, u: wrap(asarray(p) + asarray(u).to(asarray(p).dtype)), params, updates)
