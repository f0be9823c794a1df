"""Pytrees: tuples, lists, dicts (sorted keys), namedtuples; everything else is a leaf."""
from __future__ import annotations

import functools


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(tree):
    leaves = []

    def rec(t):
        if t is None:
            return ("none",)
        if _is_namedtuple(t):
            return ("nt", type(t), [rec(v) for v in t])
        if isinstance(t, (tuple, list)):
            return ("seq", type(t), [rec(v) for v in t])
        if isinstance(t, dict):
            ks = sorted(t.keys())
            return ("dict", ks, [rec(t[k]) for k in ks])
        leaves.append(t)
        return ("leaf",)

    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        if d[0] == "none":
            return None
        if d[0] == "leaf":
            return next(it)
        if d[0] == "nt":
            return d[1](*[rec(c) for c in d[2]])
        if d[0] == "seq":
            return d[1](rec(c) for c in d[2])
        return {k: rec(c) for k, c in zip(d[1], d[2])}

    return rec(treedef)


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def tree_map(f, tree, *rest):
    leaves, td = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


def tree_reduce(f, tree, *init):
    return functools.reduce(f, tree_leaves(tree), *init)
