"""The bodies of the GPU State-Evolution tests (tests/test_gpu_se.py) through the
emulated device: `trb_se_run` / `trb_se_measure` evaluated by the oracle with the
kernels' quadrature rule.  Covers the factor-level SE API (beliefs measures, errors,
free energies, the node-by-node protocol), both iteration paths, warm starts, error
behaviour, scenarios and the grid / critical-alpha helpers on CPU; the launch-count
assertion of the batched-grid test needs the real library and stays GPU-only."""
import os

import numpy as np
import pytest

from tests._emulated_device import emulated_device  # noqa: F401  (fixture)
from tests import test_gpu_se as G
from tests.golden.se_specs import SE_PRIOR_SPECS, SE_LIK_SPECS
from oracle import se_oracle as S


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "se.npz"))


@pytest.fixture(scope="module")
def gl():
    return S.Integrator("gl")


@pytest.mark.parametrize("i", range(len(SE_PRIOR_SPECS)))
def test_prior_beliefs_measures(emulated_device, gold, gl, i):  # noqa: F811
    G.test_prior_beliefs_measures(gold, gl, i)


@pytest.mark.parametrize("i", range(len(SE_LIK_SPECS)))
def test_likelihood_beliefs_measures(emulated_device, gold, gl, i):  # noqa: F811
    G.test_likelihood_beliefs_measures(gold, gl, i)


def test_node_protocol_walk_equals_kernel_iteration(emulated_device):  # noqa: F811
    G.test_node_protocol_walk_equals_kernel_iteration()


@pytest.mark.parametrize("name", ["test_synchronous_callback_path_is_bitwise_identical",
                                  "test_warm_start_continues", "test_error_behaviour_mirrors_reference"])
def test_driver_paths(emulated_device, gold, name):  # noqa: F811
    getattr(G, name)(gold)


def test_scenario_run_all_se_and_ep(emulated_device):  # noqa: F811
    G.test_scenario_run_all_se_and_ep()


def test_grid_helpers_and_critical_alpha(emulated_device):  # noqa: F811
    G.test_grid_helpers_and_critical_alpha()
