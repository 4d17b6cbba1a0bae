"""CPU: a subset of the reference's published tables (examples/glm/data/*.csv,
tests/golden/make_reference_examples.py) through the public API, the two
State-Evolution kernels and the EP sweep emulated by the oracle
(tests/_emulated_device.py).  tests/test_gpu_se_reference_examples.py runs every
row on the GPU."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests._emulated_device import emulated_device  # noqa: F401  (fixture)
from tests import reference_examples as R


@pytest.fixture(scope="module")
def ref():
    return R.load()


def test_sgn_retrieval_mse_curves_subset(emulated_device, ref):  # noqa: F811
    """Short runs (3-5 iterations of the 2-D measure) of both initialisations and both
    sparsities, one row at the uninformative fixed point where today's reference raises."""
    rows = ref["sgn_mse"]
    R.check_sgn_mse_rows(rows[[24, 84, 162, 176, 178, 235]], batched=False)
    R.check_sgn_mse_rows(rows[[173, 175, 239, 120]], batched=True)


@pytest.mark.parametrize("k", [0, 9, 18])
def test_cs_critical_lines_subset(emulated_device, ref, k):  # noqa: F811
    rho, alpha = ref["cs_critical"][k]
    assert abs(R.cs_critical_alpha(float(rho)) - alpha) < 1e-9


def test_compressed_sensing_ep_vs_se_subset(emulated_device, ref):  # noqa: F811
    """compressed_sensing_ep_vs_se.csv (N = 1000, unseeded).  With output_var = 1e-11
    every non-zero singular value saturates n_eff, so the State-Evolution column does
    not depend on the draw of W: it is a golden vector.  EP's variance and the
    empirical mse are statistical (a few per cent from instance to instance)."""
    t = ref["cs_ep_vs_se"]
    for k, seed in ((8, 42), (61, 7)):
        rho, alpha, se_v, se_n, ep_v, ep_n, mse = t[k]
        by = R.run_all(R.cs_scenario(float(rho), float(alpha), seed), metrics=["mse"])
        assert_allclose(by["SE"]["v"], se_v, rtol=1e-9)
        assert by["SE"]["n_iter"] == int(se_n)
        assert_allclose(by["EP"]["v"], ep_v, rtol=0.2)
        assert_allclose(by["mse"]["v"], mse, rtol=0.25)


def test_perceptron_ep_vs_se_subset(emulated_device, ref):  # noqa: F811
    """perceptron_ep_vs_se.csv: Binary(p_pos = 0.25) / Sgn at N = 1000; here the SE
    column depends on the spectrum of the drawn W (1e-3 relative)."""
    t = ref["perceptron_ep_vs_se"]
    k = int(np.argmin(np.abs(t[:, 0] - 0.25) + np.abs(t[:, 1] - 0.42)))
    p_pos, alpha, se_v, se_n, ep_v, ep_n, mse = t[k]
    by = R.run_all(R.perceptron_scenario(float(p_pos), float(alpha), seed=42))
    assert_allclose(by["SE"]["v"], se_v, rtol=0.01)
    assert by["SE"]["n_iter"] == int(se_n)
    assert_allclose(by["EP"]["v"], ep_v, rtol=0.1)
    assert_allclose(by["mse"]["v"], mse, rtol=0.3)
