from __future__ import annotations

import math

import torch

from .._core import asarray, wrap, FLOAT


class norm:
    @staticmethod
    def logpdf(x, loc=0., scale=1.):
        """-(log(2 pi scale^2) + ((x-loc)/scale)^2)/2"""
        x = asarray(x)
        loc = asarray(loc).to(x.dtype)
        scale = asarray(scale).to(x.dtype)
        z = (x - loc) / scale
        return wrap(-(torch.log(2 * math.pi * scale * scale) + z * z) / 2)


class poisson:
    @staticmethod
    def logpmf(k, mu):
        k, mu = asarray(k), asarray(mu)
        kf = k.to(mu.dtype)
        return wrap(torch.xlogy(kf, mu) - mu - torch.lgamma(kf + 1))
