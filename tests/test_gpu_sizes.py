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
