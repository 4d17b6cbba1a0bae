"""GPU parity at the BASELINE.json north-star shape (configs[2]: N = 4096,
alpha = 0.5, own W per instance), where the CPU oracle cannot be run for every
instance and iteration: size-independent properties of the sweep
(tests/full_size_properties.py) plus one instance against the oracle.  The same
property functions run on CPU at a small shape in tests/test_ep_host_logic_cpu.py."""
import os

import pytest

from tests import full_size_properties as P

pytestmark = pytest.mark.gpu

# 80 instances put the sweep in the regime of the headline benchmark (operator traffic far
# above the CUDA-graph threshold, one CTA per instance in the update kernels);
# TRB_TEST_INSTANCES=4 runs the same checks in the launch-bound regime (CUDA-graph replay,
# update kernels split over 2-CTA clusters).
B = int(os.environ.get("TRB_TEST_INSTANCES", "80"))
N, M, N_ITER = 4096, 2048, 60
SUB = (2, B - 2) if B >= 8 else (1, 3)


@pytest.fixture(scope="module")
def data():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return P.make_batch(B, N, M, seed=2026)


@pytest.fixture(scope="module")
def general(data):
    """(r_x, v_x, mse trajectory) of the general 4-pass sweep; the exact 3-pass and
    2-pass schedules must reproduce it."""
    return P.schedules_agree(data, N_ITER)


def test_full_size_schedules_agree(general):
    assert general[0].shape == (B, N) and general[2].shape == (N_ITER, B)


def test_full_size_instance_against_oracle(data, general):
    P.oracle_sample(data, general, b=1, n_iter=N_ITER)


def test_full_size_instances_are_independent(data, general):
    P.instances_are_independent(data, general, SUB[0], SUB[1], N_ITER)


def test_full_size_bayes_optimal_consistency(general):
    P.bayes_optimal_consistency(general)


def test_full_size_linear_gaussian_closed_form(data):
    P.linear_gaussian_closed_form(data)


@pytest.mark.parametrize("config", ["sparse_regression_n1000", "sign_perceptron_n2000"])
def test_baseline_single_instance_configs_against_oracle(config):
    """BASELINE.json configs[0] and configs[1] at their full sizes as single instances
    (default path: the persistent whole-sweep kernel) against the oracle on the same W, y:
      0: GaussBernoulliPrior(N=1000, rho=0.1) @ LinearChannel(M=500) @ GaussianLikelihood(1e-2), 100 it
      1: GaussianPrior(N=2000) @ LinearChannel(alpha=2) @ SgnLikelihood, damping 0.5, 50 it
    (same construction as tools/bench_small_configs.py, which measured 1e-13 / 7e-13)."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from numpy.testing import assert_allclose
    from oracle import tramp_oracle as orc
    from tramp_b200.priors import GaussBernoulliPrior, GaussianPrior
    from tramp_b200.likelihoods import GaussianLikelihood, SgnLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    rng = np.random.RandomState(42)
    if config == "sparse_regression_n1000":
        N, M, n_iter, damping = 1000, 500, 100, None
        W = rng.randn(M, N) / np.sqrt(N)
        x = rng.randn(N) * (rng.rand(N) < 0.1)
        y = W @ x + 0.1 * rng.randn(M)
        prior, lik = GaussBernoulliPrior(size=N, rho=0.1), GaussianLikelihood(y=y, var=1e-2)
        pspec, lspec = dict(kind="gauss_bernoulli", rho=0.1), dict(kind="gaussian", var=1e-2, y=y)
    else:
        N, M, n_iter, damping = 2000, 4000, 50, 0.5
        W = rng.randn(M, N) / np.sqrt(N)
        x = rng.randn(N)
        y = np.where(W @ x >= 0, 1.0, -1.0)
        prior, lik = GaussianPrior(size=N), SgnLikelihood(y=y)
        pspec, lspec = dict(kind="gaussian"), dict(kind="sgn", y=y)
    model = (prior @ V("x") @ LinearChannel(W) @ V("z") @ lik).to_model()
    ep = ExpectationPropagation(model)
    track = TrackErrors({"x": x})
    ep.iterate(max_iter=n_iter, callback=track, damping=damping)
    got = ep.get_variables_data(["x"])
    with np.errstate(all="ignore"):
        ref = orc.ep_glm(pspec, W, lspec, n_iter, damping=damping, x_true=x)
    assert_allclose(got["x"]["r"], ref["r_x"], rtol=0, atol=1e-9 * np.abs(ref["r_x"]).max())
    assert_allclose(got["x"]["v"], ref["v_x"], rtol=1e-9)
    assert_allclose(np.array([e["mse"] for e in track.errors]), np.array(ref["traj"]["mse_x"]), rtol=1e-9)
