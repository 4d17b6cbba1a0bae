"""Sparse (Gauss-Bernoulli) belief (reference tramp/beliefs/sparse.py:5-27)."""
import numpy as np
from . import _dev
from .. import ops, _lib


def _F(eta):
    if np.ndim(eta) != 0:
        raise ValueError("eta must be a scalar")
    return ops.make_factor(_lib.GAUSS_BERNOULLI_PRIOR, 0.0, 0.0, float(eta), 0.0)


def A(a, b, eta):
    return _dev.elementwise(_F(eta), a, b, None, "A")


def r(a, b, eta):
    return _dev.elementwise(_F(eta), a, b, None, "r")


def v(a, b, eta):
    return _dev.elementwise(_F(eta), a, b, None, "v")


def p(a, b, eta):
    """Weight of the Gaussian component, expit(normal.A(a, b) - eta) (reference sparse.py:9-12),
    evaluated by the device moment routine (`sparse_weight`, trb_moments.cuh)."""
    return _dev.elementwise(_F(eta), a, b, None, "p")


def tau(a, b, eta):
    return v(a, b, eta) + r(a, b, eta)**2
