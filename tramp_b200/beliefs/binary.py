"""Binary belief p(x) ~ exp(b x) on x = +-1 (reference tramp/beliefs/binary.py:4-17):
log-partition ln 2cosh(b), mean tanh(b), variance 1 - tanh(b)^2, second moment 1.

All three are evaluated by the moment routine of the BinaryPrior factor
(tramp_b200/csrc/trb_moments.cuh) with a symmetric prior (b0 = 0): its scalar
log-partition is A(b + b0) - A(b0) - a/2, so with a = 0 the belief's log-partition is
that value plus A(0) = ln 2."""
import math

from . import _dev
from .. import ops, _lib

_LN2 = math.log(2.0)


def _symmetric_prior(what, b, a):
    return _dev.elementwise(ops.make_factor(_lib.BINARY_PRIOR), a, b, None, what)


def A(b):
    return _symmetric_prior("A", b, a=0.0) + _LN2


def r(b):
    return _symmetric_prior("r", b, a=1.0)


def v(b):
    return _symmetric_prior("v", b, a=1.0)


def tau(b):
    return 1.
