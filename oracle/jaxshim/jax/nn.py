from __future__ import annotations

import torch

from ._core import asarray, wrap


def softplus(x):
    """jax.nn.softplus(x) = logaddexp(x, 0)."""
    x = asarray(x)
    return wrap(torch.logaddexp(x, torch.zeros((), dtype=x.dtype)))


def sigmoid(x):
    return wrap(torch.sigmoid(asarray(x)))


def softmax(x, axis=-1):
    return wrap(torch.softmax(asarray(x), dim=axis))
