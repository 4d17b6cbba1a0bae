"""CPU coverage of the Python side of State Evolution: the same
public-API calls as tests/test_gpu_se.py, with the two SE kernels emulated by
the oracle (tests/_emulated_device.py).  What is under test here is the glue --
initialisers, damping maps, record replay into callbacks, per-problem stopping,
snapshots, scenarios and grid helpers -- not the kernels."""
import os
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests._emulated_device import emulated_device  # noqa: F401  (fixture)
from tests.golden.se_specs import SE_RUNS, SE_ENTROPY_RUNS
from tests.test_gpu_se import make_model, run_device


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "se.npz"))


# the 2-D measure of AbsLikelihood is slow in numpy; everything else runs here, the
# empirical-spectrum channels (LinearChannel factorised with torch on the CPU) included
CPU_RUNS = [n for n in sorted(SE_RUNS) if SE_RUNS[n]["lik"]["kind"] != "abs"]


@pytest.mark.parametrize("name", CPU_RUNS)
def test_public_api_reproduces_reference_runs(emulated_device, gold, name):  # noqa: F811
    case = SE_RUNS[name]
    se, vx, vz = run_device(case)
    assert emulated_device.calls["trb_se_run"] == 1          # whole recursion in one call
    assert se.n_iter == int(gold[f"{name}_n_iter"])
    assert_allclose(vx, gold[f"{name}_vx"], rtol=1e-6)
    assert_allclose(vz, gold[f"{name}_vz"], rtol=1e-6)
    data = se.get_variables_data()
    assert_allclose([data["x"]["v"], data["z"]["v"]], gold[f"{name}_v_final"], rtol=1e-6)
    assert_allclose([data["x"]["tau"], data["z"]["tau"]], gold[f"{name}_tau"], rtol=1e-9)
    edges = se.get_edges_data(["a", "direction", "tau"])
    assert [e["direction"] for e in edges] == ["fwd"] * 4 + ["bwd"] * 4
    assert [e["x_id"] for e in edges] == ["x", "x", "z", "z", "z", "z", "x", "x"]
    assert_allclose([e["a"] for e in edges], gold[f"{name}_a"], rtol=1e-6, atol=1e-5)
    if name in SE_ENTROPY_RUNS:
        assert_allclose(se.entropy(), gold[f"{name}_entropy"], rtol=1e-6, atol=1e-7)
        nodes = se.get_nodes_data(["A", "v"])
        assert [n["type"] for n in nodes] == ["factor", "variable", "factor", "variable", "factor"]
        assert all(n["A"] is not None for n in nodes)


def test_synchronous_callbacks_snapshots_and_warm_start(emulated_device, gold):  # noqa: F811
    from tramp_b200.algos import StateEvolution, Callback, PassCallback, EarlyStopping, JoinCallback
    case = SE_RUNS["cs_a05"]

    class Spy(Callback):
        def __init__(self):
            self.seen = []

        def __call__(self, algo, i, max_iter):
            self.seen.append((i, algo.n_iter, algo.get_variable_data("x")["v"]))
    spy = Spy()
    se = StateEvolution(make_model(case))
    se.iterate(max_iter=200, callback=JoinCallback([spy, EarlyStopping()]))    # host-side EarlyStopping
    assert se.n_iter == int(gold["cs_a05_n_iter"]) == len(spy.seen)
    assert emulated_device.calls["trb_se_run"] == se.n_iter                    # one call per iteration
    assert_allclose([s[2] for s in spy.seen], gold["cs_a05_vx"], rtol=1e-6)
    snap = se.snapshot()
    v_before = se.get_variable_data("x")["v"]
    se.iterate(max_iter=3, callback=PassCallback(), warm_start=True)
    assert se.n_iter == int(gold["cs_a05_n_iter"]) + 3
    se.reset_message_dag(snap)
    assert se.get_variable_data("x")["v"] == v_before
    with pytest.raises(ValueError, match="not in variables"):
        se.get_variable_data("w")


def test_batched_models_stop_independently(emulated_device):  # noqa: F811
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.algos import StateEvolution, EarlyStopping, TrackEvolution, JoinCallback
    from tramp_b200.experiments import run_state_evolution, run_state_evolution_grid
    kw = dict(prior_type="gauss_bernoulli", output_type="gaussian", prior_rho=0.3, output_var=1e-3)
    alphas = [0.2, 0.5, 0.8, 1.1]
    models = [glm_state_evolution(alpha=a, **kw) for a in alphas]
    se = StateEvolution(models)
    evo = TrackEvolution(ids=["x"])
    se.iterate(max_iter=200, callback=JoinCallback([evo, EarlyStopping()]))
    assert emulated_device.calls["trb_se_run"] == 1
    v = se.get_variable_data("x")["v"]
    assert v.shape == (4,) and np.all(np.diff(v) < 0)
    assert se.n_iter == se.n_iter_per_problem.max() and len(set(se.n_iter_per_problem.tolist())) > 1
    df = evo.get_dataframe()
    assert len(df) == se.n_iter and df.v.iloc[0].shape == (4,)
    grid = run_state_evolution_grid(["x"], [glm_state_evolution(alpha=a, **kw) for a in alphas],
                                    max_iter=200)
    for g, a in enumerate(alphas):
        one = run_state_evolution(["x"], glm_state_evolution(alpha=a, **kw), max_iter=200)
        assert grid[g] == one
        assert one[0]["v"] == v[g] and one[0]["n_iter"] == se.n_iter_per_problem[g]


def test_grid_keeps_its_valid_problems_when_one_leaves_the_domain(emulated_device):  # noqa: F811
    """One run that violates az > 1/tau_z raises like the reference
    (sgn_likelihood.py:80-81); inside a grid launched at once it reads v = NaN and
    keeps its flag, the other problems are unaffected."""
    from tramp_b200 import _lib
    from tramp_b200.priors import GaussianPrior, BinaryPrior
    from tramp_b200.likelihoods import SgnLikelihood
    from tramp_b200.channels import MarchenkoPasturChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import StateEvolution
    from tramp_b200.experiments import run_state_evolution_grid

    def build(prior):
        return (prior @ V(id="x") @ MarchenkoPasturChannel(alpha=2.0) @ V(id="z") @ SgnLikelihood(y=None)).to_model()
    se = StateEvolution([build(GaussianPrior(size=None)), build(BinaryPrior(size=None, p_pos=0.6))])
    se.iterate(max_iter=30)
    v = se.get_variable_data("x")["v"]
    assert np.isnan(v[0]) and 0 < v[1] < 1
    assert se.flags[0] & _lib.FLAG_SE_DOMAIN and not se.flags[1] & _lib.FLAG_SE_DOMAIN
    alone = StateEvolution(build(BinaryPrior(size=None, p_pos=0.6)))
    alone.iterate(max_iter=30)
    assert alone.get_variable_data("x")["v"] == v[1] and alone.n_iter == se.n_iter_per_problem[1]
    grid = run_state_evolution_grid(["x"], [build(GaussianPrior(size=None)), build(BinaryPrior(size=None, p_pos=0.6))],
                                    max_iter=30)
    assert np.isnan(grid[0][0]["v"]) and grid[1][0]["v"] == v[1]
    for models in ([build(GaussianPrior(size=None))] * 2, [build(GaussianPrior(size=None))]):
        with pytest.raises(AssertionError, match="az must be greater"):
            StateEvolution(models).iterate(max_iter=5)


def test_errors_and_factor_level_api(emulated_device):  # noqa: F811
    from tramp_b200.priors import GaussianPrior, GaussBernoulliPrior
    from tramp_b200.likelihoods import SgnLikelihood
    from tramp_b200.channels import MarchenkoPasturChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import StateEvolution
    model = (GaussianPrior(size=None) @ V(id="x") @ MarchenkoPasturChannel(alpha=2.0) @ V(id="z")
             @ SgnLikelihood(y=None)).to_model()
    with pytest.raises(AssertionError, match="az must be greater"):
        StateEvolution(model).iterate(max_iter=5)
    lk = SgnLikelihood(y=None)
    with pytest.raises(AssertionError, match="az must be greater"):
        lk.compute_backward_error(1.0, 1.0)
    p = GaussBernoulliPrior(size=None, rho=0.2)
    ax = np.array([0.1, 1.0, 5.0])
    v = p.compute_forward_error(ax)
    assert v.shape == (3,) and np.all(np.diff(v) < 0) and isinstance(p.compute_forward_error(1.0), float)
    assert_allclose(p.compute_forward_state_evolution(1.0), 1 / v[1] - 1.0, rtol=1e-12)
    assert_allclose(p.compute_forward_overlap(1.0), p.second_moment() - v[1], rtol=1e-12)
    with pytest.raises(NotImplementedError):
        p.beliefs_measure(1.0, lambda b: b)


def test_critical_alpha_search(emulated_device):  # noqa: F811
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.experiments import find_critical_alpha
    kw = dict(prior_type="gauss_bernoulli", output_type="gaussian", prior_rho=0.3, output_var=1e-10)
    crit = dict(id="x", a0=0.0, mse_criterion="perfect", alpha_min=0.3, alpha_max=0.95,
                model_builder=glm_state_evolution, alpha_tol=2e-2, vtol=1e-3, **kw)
    a_bisect = find_critical_alpha(**crit)
    a_grid = find_critical_alpha(grid=7, **crit)
    assert abs(a_bisect - a_grid) < 4e-2 and 0.3 < a_bisect < 0.95
    random = find_critical_alpha(id="x", a0=0.0, mse_criterion="random", alpha_min=1e-4, alpha_max=0.5,
                                 model_builder=glm_state_evolution, alpha_tol=5e-2, vtol=0.05, **kw)
    assert 1e-4 < random < 0.5


def test_scenario_state_evolution_reproduces_the_reference_run(emulated_device):  # noqa: F811
    """`BayesOptimalScenario.run_all(source="SE")` on `glm_generative` with the
    reference's seed protocol: the same script on the unmodified reference printed
    SE v = 0.009469020882561508 after 14 iterations (tests/test_gpu_se.py checks the
    EP half of the same run on the GPU)."""
    from tramp_b200.models import glm_generative
    from tramp_b200.experiments import BayesOptimalScenario
    from tramp_b200.algos import EarlyStopping
    np.random.seed(5)
    model = glm_generative(N=400, alpha=0.7, ensemble_type="gaussian", prior_type="gauss_bernoulli",
                           output_type="gaussian", prior_rho=0.2, output_var=1e-2)
    scenario = BayesOptimalScenario(model, x_ids=["x"])
    records = scenario.run_all(source="SE", metrics=["mse"], max_iter=100, callback=EarlyStopping())
    assert records == [dict(source="SE", x_id="x", v=records[0]["v"], n_iter=14)]
    assert_allclose(records[0]["v"], 0.009469020882561508, rtol=1e-9)
    df = scenario.se_convergence(max_iter=30)
    assert list(df.columns) == ["id", "v", "iter"] and len(df) == 30
    assert np.all(np.diff(df.v.values) < 1e-12)
    assert scenario.se.analytical is False and scenario.se.linear.rank == 280


def test_tracking_callbacks_on_state_evolution(emulated_device):  # noqa: F811
    """TrackObjective / TrackMessages (reference callbacks.py:49-85) work on the SE
    driver through its update_objective / get_edges_data / get_nodes_data."""
    from tramp_b200.algos import StateEvolution, TrackObjective, TrackMessages, JoinCallback
    case = SE_RUNS["cs_damped"]
    se = StateEvolution(make_model(case))
    obj, msgs = TrackObjective(), TrackMessages(keys=["a", "n_iter", "direction", "damping"])
    se.iterate(max_iter=4, callback=JoinCallback([obj, msgs]), damping=0.5)
    edges, nodes, model = obj.get_dataframe()
    assert len(model) == 4 and list(model.n_iter) == [1, 2, 3, 4]
    assert np.all(np.isfinite(model.A.values)) and np.all(np.diff(model.A.values) != 0)
    assert len(edges) == 4 * 8 and len(nodes) == 4 * 5
    assert set(nodes.type) == {"factor", "variable"} and nodes.A.notna().all()
    df = msgs.get_dataframe()
    assert len(df) == 4 * 8 and set(df.direction) == {"fwd", "bwd"}
    # constant damping sits on the four factor -> variable edges only
    last = df.tail(8)
    assert list(last.damping.fillna(0.0)) == [0.5, 0.0, 0.5, 0.0, 0.5, 0.0, 0.5, 0.0]
    assert_allclose(-model.A.values[-1], se.entropy(), rtol=1e-12)
