"""Positive belief: the truncated normal on the half line [0, inf) (reference
tramp/beliefs/positive.py:8-26), i.e. `truncated` with fixed bounds; the device
routine takes its erfcx fast path for a half-infinite interval."""
import functools
import math

from . import truncated

_HALF_LINE = dict(xmin=0, xmax=math.inf)

A = functools.partial(truncated.A, **_HALF_LINE)
r = functools.partial(truncated.r, **_HALF_LINE)
v = functools.partial(truncated.v, **_HALF_LINE)
p = functools.partial(truncated.p, **_HALF_LINE)
tau = functools.partial(truncated.tau, **_HALF_LINE)
