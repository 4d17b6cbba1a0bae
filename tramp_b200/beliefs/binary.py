"""Binary belief (reference tramp/beliefs/binary.py:4-17)."""
from . import _dev
from .. import ops, _lib

_F = lambda: ops.make_factor(_lib.BINARY_PRIOR)  # noqa: E731


def A(b):
    # binary_prior's scalar log-partition is A(b + b0) - A(b0) - a/2 with
    # b0 = 0: A(0) = log 2 is added back and a = 0 drops the last term
    import numpy as np
    return _dev.elementwise(_F(), 0.0, b, None, "A") + np.log(2.0)


def r(b):
    return _dev.elementwise(_F(), 1.0, b, None, "r")


def v(b):
    return _dev.elementwise(_F(), 1.0, b, None, "v")


def tau(b):
    return 1.
