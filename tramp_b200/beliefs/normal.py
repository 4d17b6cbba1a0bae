"""Normal belief (reference tramp/beliefs/normal.py)."""
import numpy as np
from . import _dev
from .. import ops, _lib


def A(a, b):
    # log-partition of N(b/a, 1/a): the truncated-normal logZ on (-inf, +inf)
    from ..utils.truncated_normal import truncated_normal_logZ
    a, b = np.broadcast_arrays(np.asarray(a, float), np.asarray(b, float))
    return truncated_normal_logZ(b / a, 1 / a, -np.inf, np.inf)


def r(a, b):
    return _dev.elementwise(ops.make_factor(_lib.GAUSSIAN_PRIOR), a, b, None, "r")


def v(a, b):
    return _dev.elementwise(ops.make_factor(_lib.GAUSSIAN_PRIOR), a, b, None, "v")


def tau(a, b):
    return v(a, b) + r(a, b)**2
