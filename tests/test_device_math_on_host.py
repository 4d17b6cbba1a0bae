"""The device routines of tramp_b200/csrc/trb_moments.cuh (prior / likelihood
moments, log-partitions, truncated normal: SURVEY 8a rows a8-a16) compiled as HOST
functions and checked against the reference's golden vectors, so that the formulas
the kernels evaluate are covered in the build container, which has no GPU.

The host build is tests/_device_math_host.py.  Host and device special functions
differ by an ulp or two, which is what the tolerances of tests/test_gpu_primitives.py
(repeated here) already allow.
"""
import ctypes as C
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests import _device_math_host as H
from tests.golden.make_golden_specs import PRIOR_SPECS, LIK_SPECS, TRUNC_CASES

RTOL = 1e-11


@pytest.fixture(scope="module")
def host_math():
    lib = H.load()
    if lib is None:
        pytest.skip("nvcc not available")
    return lib


def _p(x):
    return x.ctypes.data_as(C.POINTER(C.c_double))


def _factor(lib, spec, a, b, y=None):
    from tramp_b200 import ops
    return H.factor_elementwise(ops.factor_from_spec(spec), a, np.asarray(b, float), y)


@pytest.fixture(scope="module")
def el(golden_dir):
    return np.load(os.path.join(golden_dir, "elementwise.npz"))


@pytest.mark.parametrize("i", range(len(PRIOR_SPECS)))
def test_prior_moments(host_math, el, i):
    """reference priors/*.py compute_forward_posterior / scalar_log_partition on the
    golden grid; isotropic case: v and A are the means over the components."""
    r, v, A = _factor(host_math, PRIOR_SPECS[i], el["grid_a"], el["grid_b"])
    assert_allclose(r, el[f"prior{i}_r"], rtol=RTOL, atol=1e-300)
    assert_allclose(v, el[f"prior{i}_v"], rtol=RTOL, atol=1e-15)
    assert_allclose(A, el[f"prior{i}_A"], rtol=RTOL, atol=1e-13)
    for j, a_s in enumerate(el["iso_a"]):
        r, v, A = _factor(host_math, PRIOR_SPECS[i], a_s, el["grid_bnorm"] * np.sqrt(a_s))
        assert_allclose(r, el[f"prior{i}_iso{j}_r"], rtol=RTOL, atol=1e-300)
        assert_allclose(v.mean(), el[f"prior{i}_iso{j}_v"], rtol=RTOL)
        assert_allclose(A.mean(), el[f"prior{i}_iso{j}_A"], rtol=1e-10)


@pytest.mark.parametrize("i", range(len(LIK_SPECS)))
def test_likelihood_moments(host_math, el, i):
    """reference likelihoods/*.py compute_backward_posterior / scalar_log_partition."""
    spec = dict(LIK_SPECS[i], role="likelihood")
    y = el[f"lik{i}_y"]
    r, v, A = _factor(host_math, spec, el["grid_a"], el["grid_b"], y)
    assert_allclose(r, el[f"lik{i}_r"], rtol=RTOL, atol=1e-300)
    assert_allclose(v[:100], el[f"lik{i}_v"][:100], rtol=1e-11, atol=1e-15)
    assert_allclose(v, el[f"lik{i}_v"], rtol=1e-7, atol=1e-15 * max(1.0, np.abs(y).max()**2))
    assert_allclose(A, el[f"lik{i}_A"], rtol=RTOL, atol=1e-13)
    for j, a_s in enumerate(el["iso_a"]):
        b = el["grid_bnorm"] * np.sqrt(a_s)
        r, v, A = _factor(host_math, spec, a_s, b, y)
        assert_allclose(r, el[f"lik{i}_iso{j}_r"], rtol=RTOL,
                        atol=16 * np.finfo(float).eps * np.abs(b).max() / a_s)
        assert_allclose(v.mean(), el[f"lik{i}_iso{j}_v"], rtol=RTOL)
        assert_allclose(A.mean(), el[f"lik{i}_iso{j}_A"], rtol=1e-10)


@pytest.mark.parametrize("i", range(len(TRUNC_CASES)))
def test_truncated_normal(host_math, el, i):
    """reference utils/truncated_normal.py:234-298, all five finite-interval branches
    and the erfcx half-line path (cases of tests/golden/make_golden_specs.py)."""
    a, lo, hi = TRUNC_CASES[i]
    b = el["trunc_b"]
    r0, v0 = np.ascontiguousarray(b / a), np.full_like(b, 1 / a)
    mean, var, logZ, proba = (np.empty_like(b) for _ in range(4))
    with np.errstate(all="ignore"):
        host_math.hm_truncated_normal(b.size, _p(r0), _p(v0), lo, hi, _p(mean), _p(var), _p(logZ), _p(proba))
    kw = dict(rtol=1e-9, equal_nan=True)
    assert_allclose(mean, el[f"trunc{i}_r"], atol=1e-12, **kw)
    assert_allclose(var, el[f"trunc{i}_v"], atol=1e-12, **kw)
    assert_allclose(logZ, el[f"trunc{i}_A"], atol=1e-12, **kw)
    assert_allclose(proba, el[f"trunc{i}_p"], atol=1e-15, **kw)


def test_constant_message_factors(host_math):
    """Gaussian prior and Gaussian likelihood send constant messages (reference
    gaussian_prior.py:86-89, gaussian_likelihood.py:68-71): the kernels skip their moments."""
    from tramp_b200 import _lib
    const = {k for k in range(6) if host_math.hm_is_constant_message(k)}
    assert const == {_lib.GAUSSIAN_PRIOR, _lib.GAUSSIAN_LIKELIHOOD}
