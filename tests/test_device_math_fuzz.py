"""Differential fuzz of the device moment routines (tramp_b200/csrc/trb_moments.cuh,
compiled for the host by tests/_device_math_host.py) against the oracle's restatement
of the reference (numpy / scipy) on 10^5 random points per factor: precisions a over 14
decades, natural means b on the scale EP produces (sqrt(a) times a few standard
deviations).  The golden grids pin chosen points; this looks for a branch or threshold
that differs anywhere in between."""
import numpy as np
import pytest

from tests import _device_math_host as H
from oracle import tramp_oracle as O

N_POINTS = 100_000
PRIORS = [dict(kind="gauss_bernoulli", rho=0.1), dict(kind="gauss_bernoulli", rho=0.7, mean=0.3, var=2.0),
          dict(kind="binary", p_pos=0.6), dict(kind="gaussian", mean=0.2, var=0.5)]
LIKELIHOODS = [dict(kind="sgn"), dict(kind="abs"), dict(kind="gaussian", var=0.3)]


@pytest.fixture(scope="module")
def host_math():
    if H.load() is None:
        pytest.skip("nvcc not available")


def _points(seed):
    rng = np.random.RandomState(seed)
    a = 10 ** rng.uniform(-6, 8, N_POINTS)
    b = np.sqrt(a) * rng.randn(N_POINTS) * rng.choice([1.0, 3.0, 10.0], N_POINTS)
    return rng, a, b


def _check(r, v, A, r_ref, v_ref, A_ref, a, b, unit):
    assert not np.isnan(r).any() and not np.isnan(v).any() and not np.isnan(A).any()
    # means: relative to the larger of the value and the Gaussian scale |b| / a
    assert np.all(np.abs(r - r_ref) <= 1e-11 * np.maximum(np.abs(r_ref), np.abs(b) / a))
    # variances: the reference's own formulas cancel against 1 (1 - tanh^2, 1 + g2 - g1^2 with
    # g1^2 ~ 100 ten standard deviations out), so they are compared in the natural unit of the
    # variable: 1 (or y^2) for +-1 / +-|y| variables, the incoming variance 1 / a otherwise
    assert np.all(np.abs(v - v_ref) <= 1e-11 * v_ref + 1e-12 * unit)
    assert np.all(np.abs(A - A_ref) <= 1e-11 * np.abs(A_ref) + 1e-12)


@pytest.mark.parametrize("spec", PRIORS, ids=lambda s: s["kind"] + str(s.get("rho", "")))
def test_prior_moments_fuzz(host_math, spec):
    from tramp_b200 import ops
    _, a, b = _points(1)
    with np.errstate(all="ignore"):
        r, v, A = H.factor_elementwise(ops.factor_from_spec(spec), a, b)
        r_ref, v_ref = O.prior_forward_posterior(dict(spec, isotropic=False), a, b)
        if spec["kind"] == "gauss_bernoulli":
            a0, b0, eta = O._gb_nat(spec)
            A_ref = O.sparse_A(a + a0, b + b0, eta) - O.sparse_A(a0, b0, eta)
        elif spec["kind"] == "binary":
            b0 = 0.5 * np.log(spec["p_pos"] / (1 - spec["p_pos"]))
            A_ref = O.binary_A(b + b0) - O.binary_A(b0) - 0.5 * a
        else:
            a0, b0 = 1 / spec["var"], spec["mean"] / spec["var"]
            A_ref = O.normal_A(a + a0, b + b0) - O.normal_A(a0, b0)
    _check(r, v, A, r_ref, np.broadcast_to(v_ref, r.shape), A_ref, a, b,
           unit=np.ones_like(a) if spec["kind"] == "binary" else 1 / a)


@pytest.mark.parametrize("spec", LIKELIHOODS, ids=lambda s: s["kind"])
def test_likelihood_moments_fuzz(host_math, spec):
    from tramp_b200 import ops
    rng, a, b = _points(2)
    kind = spec["kind"]
    y = rng.choice([-1.0, 1.0], N_POINTS) if kind == "sgn" else \
        (np.abs(rng.randn(N_POINTS)) if kind == "abs" else rng.randn(N_POINTS))
    with np.errstate(all="ignore"):
        r, v, A = H.factor_elementwise(ops.factor_from_spec(dict(spec, role="likelihood")), a, b, y)
        r_ref, v_ref = O.likelihood_backward_posterior(dict(spec, isotropic=False), a, b, y)
        if kind == "sgn":
            A_ref = O.positive_A(a, b * y)
        elif kind == "abs":
            A_ref = -0.5 * a * y**2 + O.binary_A(b * y)
        else:
            ay = 1 / spec["var"]
            A_ref = O.normal_A(a + ay, b + ay * y) - O.normal_A(ay, ay * y)
    _check(r, v, A, r_ref, np.broadcast_to(v_ref, r.shape), A_ref, a, b,
           unit=np.maximum(1.0, y**2) if kind == "abs" else 1 / a)


@pytest.mark.parametrize("shape", ["half_line_up", "half_line_down", "narrow", "wide"])
def test_truncated_normal_fuzz(host_math, shape):
    """utils/truncated_normal.py:14-298 on random intervals: the erfcx half-line path and
    the five finite-interval branches (inf / close / neg / pos / other) must be taken for
    the same arguments as in the reference -- no NaN or infinity on one side only -- and
    agree where the reference's own formulas are well conditioned.  Narrow intervals
    (width down to 1e-9, the Taylor branch) lose digits in the variance in both."""
    import ctypes as C
    lib = H.load()
    rng = np.random.RandomState({"half_line_up": 0, "half_line_down": 1, "narrow": 2, "wide": 3}[shape])

    def p(x):
        return x.ctypes.data_as(C.POINTER(C.c_double))
    n = 20_000
    for _ in range(8):
        if shape == "half_line_up":
            lo, hi = float(rng.randn()), np.inf
        elif shape == "half_line_down":
            lo, hi = -np.inf, float(rng.randn())
        elif shape == "narrow":
            lo = float(2 * rng.randn())
            hi = lo + float(10 ** rng.uniform(-9, 1))
        else:
            lo, hi = float(-10 ** rng.uniform(-2, 1)), float(10 ** rng.uniform(-2, 1))
        v0 = 10 ** rng.uniform(-4, 3, n)
        edge = rng.choice([0.0, lo if np.isfinite(lo) else 0.0, hi if np.isfinite(hi) else 0.0], n)
        r0 = edge + np.sqrt(v0) * rng.randn(n) * rng.choice([1.0, 3.0, 8.0], n)
        got = [np.empty(n) for _ in range(4)]
        lib.hm_truncated_normal(n, p(r0), p(v0), lo, hi, *[p(o) for o in got])
        with np.errstate(all="ignore"):
            want = [O.truncated_normal_mean(r0, v0, lo, hi), O.truncated_normal_var(r0, v0, lo, hi),
                    O.truncated_normal_logZ(r0, v0, lo, hi), O.truncated_normal_proba(r0, v0, lo, hi)]
        units = (np.sqrt(v0) + np.abs(r0), v0, np.ones(n), np.ones(n))
        tols = (1e-6, 1e-2 if shape == "narrow" else 1e-6, 1e-6, 1e-11)
        for g, w, unit, tol in zip(got, want, units, tols):
            assert np.array_equal(np.isnan(g), np.isnan(w)) and np.array_equal(np.isinf(g), np.isinf(w))
            ok = np.isfinite(w)
            assert np.all(np.abs(g[ok] - w[ok]) <= tol * np.maximum(np.abs(w[ok]), 1e-3 * unit[ok]))
