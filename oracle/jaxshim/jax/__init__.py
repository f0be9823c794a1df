"""Torch-CPU-backed stand-in for the `jax` import name (test infrastructure; see ../README.md)."""
from __future__ import annotations

import functools

import torch

from . import _core
from ._core import Array  # noqa: F401
from . import numpy  # noqa: F401
from . import lax  # noqa: F401
from . import nn  # noqa: F401
from . import random  # noqa: F401
from . import scipy  # noqa: F401
from . import tree_util  # noqa: F401

__version__ = "0.0-shim"


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def _leaves_map(f, tree):
    return tree_util.tree_map(f, tree)


def vmap(fun, in_axes=0, out_axes=0):
    """Python loop over the mapped axis, outputs stacked along out_axes (pytrees of tuples/lists/dicts)."""

    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap: in_axes length %d != number of arguments %d" % (len(axes), len(args)))
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                leaves = tree_util.tree_leaves(a)
                n = _core.asarray(leaves[0]).shape[ax]
                break
        if n is None:
            raise ValueError("vmap: no mapped argument")

        def take(a, ax, i):
            if ax is None:
                return a
            return _leaves_map(lambda x: _core.asarray(x).select(ax, i), a)

        outs = [fun(*[take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        first = outs[0]
        flat0, treedef = tree_util.tree_flatten(first)
        oaxes = out_axes if isinstance(out_axes, (tuple, list)) else (out_axes,) * len(flat0)
        if len(oaxes) != len(flat0):
            # out_axes given per top-level output
            oaxes = (out_axes,) * len(flat0) if not isinstance(out_axes, (tuple, list)) else tuple(oaxes)
        stacked = []
        flats = [tree_util.tree_flatten(o)[0] for o in outs]
        for j in range(len(flat0)):
            stacked.append(_core.wrap(torch.stack([_core.asarray(f[j]) for f in flats], dim=oaxes[j])))
        return tree_util.tree_unflatten(treedef, stacked)

    return wrapped


def value_and_grad(fun, argnums=0, has_aux=False):
    def wrapped(*args, **kw):
        args = list(args)
        p = args[argnums]
        leaves, treedef = tree_util.tree_flatten(p)
        req = [_core.asarray(x).detach().clone().requires_grad_(True) for x in leaves]
        args[argnums] = tree_util.tree_unflatten(treedef, [_core.wrap(r) for r in req])
        with torch.enable_grad():
            out = fun(*args, **kw)
            val = out[0] if has_aux else out
            grads = torch.autograd.grad(_core.asarray(val), req)
        g = tree_util.tree_unflatten(treedef, [_core.wrap(x.detach()) for x in grads])
        val_d = _core.wrap(_core.asarray(val).detach())
        return ((val_d, out[1]), g) if has_aux else (val_d, g)

    return wrapped


def grad(fun, argnums=0):
    vg = value_and_grad(fun, argnums)
    return lambda *a, **k: vg(*a, **k)[1]


def device_get(x):
    return x


def block_until_ready(x):
    return x
