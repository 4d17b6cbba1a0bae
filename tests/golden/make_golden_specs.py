"""Parameter tables shared by make_golden.py (which needs the reference) and
the tests (which must not)."""
import numpy as np

PRIOR_SPECS = [
    dict(kind="gauss_bernoulli", rho=0.1, mean=0, var=1),
    dict(kind="gauss_bernoulli", rho=0.3, mean=0.5, var=2.0),
    dict(kind="gauss_bernoulli", rho=0.5, mean=0, var=1),
    dict(kind="binary", p_pos=0.5),
    dict(kind="binary", p_pos=0.6),
    dict(kind="gaussian", mean=0, var=1),
    dict(kind="gaussian", mean=0.3, var=2.0),
]
LIK_SPECS = [
    dict(kind="gaussian", var=1),
    dict(kind="gaussian", var=0.01),
    dict(kind="sgn"),
    dict(kind="abs"),
]
TRUNC_CASES = [
    # (a, xmin, xmax) -- together they reach every branch of F0/F1/F2
    (1.0, -1.0, 1.0),          # other
    (1.0, 0.0, np.inf),        # half-infinite (+)
    (2.0, -np.inf, 0.5),       # half-infinite (-)
    (1.0, 2.0, 3.0),           # pos
    (1.0, -3.0, -2.0),         # neg
    (1.0, 1.0, 1.0 + 5e-8),    # close (Taylor)
    (0.25, -0.5, 4.0),         # other, asymmetric
    (9.0, 0.2, 0.9),           # pos, large a
]
