"""Normal belief p(x) ~ exp(-a x^2 / 2 + b x) = N(b / a, 1 / a) (reference
tramp/beliefs/normal.py): mean and variance from the moment routine of a flat Gaussian
prior factor, log-partition 0.5 (b^2 / a + ln 2 pi / a) as the log-partition of the
normal truncated to the whole line (tramp_b200/csrc/trb_moments.cuh)."""
import numpy as np

from . import _dev
from .. import ops, _lib
from ..utils.truncated_normal import truncated_normal_logZ


def _flat_prior(what, a, b):
    return _dev.elementwise(ops.make_factor(_lib.GAUSSIAN_PRIOR), a, b, None, what)


def r(a, b):
    return _flat_prior("r", a, b)


def v(a, b):
    return _flat_prior("v", a, b)


def tau(a, b):
    mean, var = r(a, b), v(a, b)
    return var + mean**2


def A(a, b):
    a, b = np.broadcast_arrays(np.asarray(a, float), np.asarray(b, float))
    return truncated_normal_logZ(b / a, 1 / a, -np.inf, np.inf)
