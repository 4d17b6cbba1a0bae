"""Truncated-normal belief on [xmin, xmax] (reference tramp/beliefs/truncated.py:7-25).

Natural parameters (a, b) <-> moment parameters r0 = b / a, v0 = 1 / a of the
untruncated Gaussian.  Every function is one launch of the device routine
`truncated_normal` (tramp_b200/csrc/trb_moments.cuh), which produces mean,
variance, log-partition and mass together; `tau` uses that instead of two
separate evaluations."""
from ..utils.truncated_normal import truncated_normal_moments

MEAN, VAR, LOGZ, PROBA = range(4)


def _moment(which, a, b, xmin, xmax):
    return truncated_normal_moments(b / a, 1 / a, xmin, xmax, only=which)


def A(a, b, xmin, xmax):
    "log-partition"
    return _moment(LOGZ, a, b, xmin, xmax)


def r(a, b, xmin, xmax):
    "mean"
    return _moment(MEAN, a, b, xmin, xmax)


def v(a, b, xmin, xmax):
    "variance"
    return _moment(VAR, a, b, xmin, xmax)


def p(a, b, xmin, xmax):
    "mass of N(b / a, 1 / a) inside [xmin, xmax]"
    return _moment(PROBA, a, b, xmin, xmax)


def tau(a, b, xmin, xmax):
    "second moment, mean^2 + variance, from a single evaluation"
    mean, var = truncated_normal_moments(b / a, 1 / a, xmin, xmax)[:2]
    return mean**2 + var
