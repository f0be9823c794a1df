"""jax.lax control flow as plain Python loops."""
from __future__ import annotations

import torch

from . import tree_util
from ._core import asarray, wrap


def scan(f, init, xs=None, length=None, reverse=False):
    if xs is None:
        n = length
    else:
        n = asarray(tree_util.tree_leaves(xs)[0]).shape[0]
    carry = init
    ys = []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        x = None if xs is None else tree_util.tree_map(lambda a: asarray(a)[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    if reverse:
        ys = ys[::-1]
    if not ys:
        # zero-length scan: jax still returns (0, ...)-shaped outputs; probe f once for the structure
        if xs is None:
            return carry, None
        probe = tree_util.tree_map(lambda a: torch.zeros(asarray(a).shape[1:], dtype=asarray(a).dtype), xs)
        _, y0 = f(init, probe)
        return carry, tree_util.tree_map(lambda a: wrap(torch.zeros((0,) + tuple(asarray(a).shape),
                                                                    dtype=asarray(a).dtype)), y0)
    flat0, treedef = tree_util.tree_flatten(ys[0])
    flats = [tree_util.tree_flatten(y)[0] for y in ys]
    stacked = [wrap(torch.stack([asarray(fl[j]) for fl in flats], dim=0)) for j in range(len(flat0))]
    return carry, tree_util.tree_unflatten(treedef, stacked)


def while_loop(cond_fun, body_fun, init):
    c = init
    while bool(cond_fun(c)):
        c = body_fun(c)
    return c


def fori_loop(lo, hi, body, init):
    c = init
    for i in range(int(lo), int(hi)):
        c = body(i, c)
    return c


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def stop_gradient(x):
    return wrap(asarray(x).detach())
