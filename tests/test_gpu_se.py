"""GPU parity of State Evolution: factor-level beliefs measures,
whole SE runs, the batched (grid) launch and the experiment helpers, against
the golden vectors of the reference (tests/golden/se.npz) and the SE oracle.

Tolerances.  The reference integrates with scipy quad / dblquad at their default
epsabs = epsrel = 1.49e-8, so its own values carry that error: against the
golden vectors the bar is 1e-7 relative + 2e-8 absolute on an integral and 1e-6
on a trajectory (a_new = 1/v - a amplifies it).  Against the oracle evaluated
with the SAME quadrature rule as the kernels ("gl") only rounding is left: 1e-9
relative (the moments cancel, 1 - tanh^2 or 1 + g2 - g1^2, so single evaluations
differ by more than an ulp between CUDA and numpy/scipy libm).
"""
import os
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import se_oracle as S
from tests.golden.se_specs import (
    SE_PRIOR_SPECS, SE_PRIOR_AX, SE_LIK_SPECS, SE_LIK_POINTS, SE_ABS_POINTS, SE_RUNS,
    SE_ENTROPY_RUNS, damping_dict, a_init_dict, spectrum_W,
)
from tests.test_se_oracle_golden import oracle_channel, run_oracle

pytestmark = pytest.mark.gpu

RTOL_REF, ATOL_REF = 1e-7, 2e-8
RTOL_SAME_RULE = 1e-9


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "se.npz"))


@pytest.fixture(scope="module")
def gl():
    return S.Integrator("gl")


def make_prior(spec, **extra):
    from tramp_b200.priors import GaussBernoulliPrior, BinaryPrior, GaussianPrior
    kw = {k: v for k, v in spec.items() if k != "kind"}
    kw.update(extra)
    return dict(gauss_bernoulli=GaussBernoulliPrior, binary=BinaryPrior,
                gaussian=GaussianPrior)[spec["kind"]](size=None, **kw)


def make_lik(spec):
    from tramp_b200.likelihoods import GaussianLikelihood, SgnLikelihood, AbsLikelihood
    kw = {k: v for k, v in spec.items() if k != "kind"}
    return dict(gaussian=GaussianLikelihood, sgn=SgnLikelihood,
                abs=AbsLikelihood)[spec["kind"]](y=None, **kw)


def make_channel(spec):
    from tramp_b200.channels import LinearChannel, MarchenkoPasturChannel
    if spec["kind"] == "marchenko":
        return MarchenkoPasturChannel(alpha=spec["alpha"])
    return LinearChannel(spectrum_W(spec))


def make_model(case):
    from tramp_b200.variables import SISOVariable as V
    return (make_prior(case["prior"]) @ V(id="x") @ make_channel(case["channel"]) @ V(id="z")
            @ make_lik(case["lik"])).to_model()


def run_device(case, callbacks=()):
    from tramp_b200.algos import (StateEvolution, EarlyStopping, TrackEvolution, JoinCallback,
                                  CustomInit, ConstantInit)
    se = StateEvolution(make_model(case))
    evo = TrackEvolution()
    cbs = [evo] + ([EarlyStopping(**case["early"])] if case.get("early") else []) + list(callbacks)
    init = CustomInit(a_init=case["a_init"]) if case.get("a_init") else ConstantInit(a=0, b=0)
    se.iterate(max_iter=case["max_iter"], callback=JoinCallback(cbs), initializer=init,
               damping=case.get("damping"))
    df = evo.get_dataframe()
    return se, df[df.id == "x"].v.values.astype(float), df[df.id == "z"].v.values.astype(float)


@pytest.mark.parametrize("i", range(len(SE_PRIOR_SPECS)))
def test_prior_beliefs_measures(gold, gl, i):
    spec = SE_PRIOR_SPECS[i]
    p = make_prior(spec)
    assert_allclose(p.second_moment(), gold[f"prior{i}_tau"], rtol=1e-15)
    v = p.compute_forward_error(SE_PRIOR_AX)               # one launch for the whole grid
    A = p.compute_free_energy(SE_PRIOR_AX)
    assert_allclose(v, gold[f"prior{i}_v"], rtol=RTOL_REF, atol=ATOL_REF)
    assert_allclose(A, gold[f"prior{i}_A"], rtol=RTOL_REF, atol=ATOL_REF)
    assert_allclose(v, [S.prior_forward_error(spec, ax, gl) for ax in SE_PRIOR_AX],
                    rtol=RTOL_SAME_RULE, atol=1e-15)
    assert_allclose(A, [S.prior_free_energy(spec, ax, gl) for ax in SE_PRIOR_AX],
                    rtol=RTOL_SAME_RULE, atol=1e-13)
    # scalar in -> float out, like the reference
    for k in (1, 4, 7):
        assert isinstance(p.compute_forward_error(float(SE_PRIOR_AX[k])), float)
        an = p.compute_forward_state_evolution(float(SE_PRIOR_AX[k]))
        assert_allclose(an, S.prior_forward_se(spec, SE_PRIOR_AX[k], gl), rtol=1e-9)
        tol = RTOL_REF * gold[f"prior{i}_anew"][k] + 2 * ATOL_REF / gold[f"prior{i}_v"][k]**2
        assert abs(an - gold[f"prior{i}_anew"][k]) <= tol
    assert_allclose(p.compute_forward_overlap(1.0), p.second_moment() - p.compute_forward_error(1.0))
    assert_allclose(p.compute_mutual_information(1.0),
                    0.5 * p.second_moment() - p.compute_free_energy(1.0))
    if spec["kind"] != "gaussian":
        with pytest.raises(NotImplementedError):
            p.beliefs_measure(1.0, lambda bx: bx)


@pytest.mark.parametrize("i", range(len(SE_LIK_SPECS)))
def test_likelihood_beliefs_measures(gold, gl, i):
    spec = SE_LIK_SPECS[i]
    lk = make_lik(spec)
    pts = SE_ABS_POINTS if spec["kind"] == "abs" else SE_LIK_POINTS
    az, tau = pts[:, 0], pts[:, 1]
    v = np.array([lk.compute_backward_error(a, t) for a, t in pts])
    A = np.array([lk.compute_free_energy(a, t) for a, t in pts])
    assert_allclose(v, gold[f"lik{i}_v"], rtol=RTOL_REF, atol=ATOL_REF)
    assert_allclose(A, gold[f"lik{i}_A"], rtol=RTOL_REF, atol=ATOL_REF)
    assert_allclose(v, [S.lik_backward_error(spec, a, t, gl) for a, t in pts], rtol=RTOL_SAME_RULE)
    assert_allclose(A, [S.lik_free_energy(spec, a, t, gl) for a, t in pts], rtol=RTOL_SAME_RULE,
                    atol=1e-13)
    # vectorised over az at one tau
    same = tau == tau[0]
    assert_allclose(lk.compute_backward_error(az[same], tau[0]), v[same], rtol=1e-15)
    for k in range(len(pts)):
        an = lk.compute_backward_state_evolution(az[k], tau[k])
        tol = RTOL_REF * gold[f"lik{i}_anew"][k] + 2 * ATOL_REF / gold[f"lik{i}_v"][k]**2
        assert abs(an - gold[f"lik{i}_anew"][k]) <= tol
    if spec["kind"] != "gaussian":
        with pytest.raises(AssertionError, match="az must be greater"):
            lk.compute_backward_error(1.0, 1.0)            # mz_hat = 0 (sgn_likelihood.py:80-81)


@pytest.mark.parametrize("name", sorted(SE_RUNS))
def test_runs_match_reference(gold, gl, name):
    case = SE_RUNS[name]
    se, vx, vz = run_device(case)
    assert se.n_iter == int(gold[f"{name}_n_iter"])
    assert_allclose(vx, gold[f"{name}_vx"], rtol=1e-6)
    assert_allclose(vz, gold[f"{name}_vz"], rtol=1e-6)
    data = se.get_variables_data()
    assert_allclose([data["x"]["v"], data["z"]["v"]], gold[f"{name}_v_final"], rtol=1e-6)
    assert_allclose([data["x"]["tau"], data["z"]["tau"]], gold[f"{name}_tau"], rtol=1e-9)
    a = np.array([r["a"] for r in se.get_edges_data(["a"])])
    assert_allclose(a, gold[f"{name}_a"], rtol=1e-6, atol=8 * np.finfo(float).eps * a.max())
    # same rule on the CPU: rounding only
    r = run_oracle(case, gl)
    assert r["n_iter"] == se.n_iter
    assert_allclose(vx, r["vx"], rtol=1e-9)
    assert_allclose(vz, r["vz"], rtol=1e-9)
    # a_new = 1/v - a_other cancels when the other precision is huge (noiseless
    # likelihood, a = 1e10): the result carries eps * max(a) in ANY implementation
    assert_allclose(a, r["a"], rtol=1e-9, atol=8 * np.finfo(float).eps * a.max())
    if name in SE_ENTROPY_RUNS:
        assert_allclose(se.entropy(), gold[f"{name}_entropy"], rtol=1e-6, atol=1e-7)
        assert_allclose(se.entropy(), S.se_entropy(case["prior"], r["channel"], case["lik"], a, gl),
                        rtol=1e-9, atol=1e-11)


def test_node_protocol_walk_equals_kernel_iteration():
    """The reference's node-by-node protocol (`forward_state_evolution` /
    `backward_state_evolution` on factors and variables, base.py:209-233, 377-410,
    sub_variables.py:33-43) walked by hand for two iterations gives the kernel's
    trajectory: the factor-level API and `trb_se_run` share their device routines."""
    from tramp_b200.algos import StateEvolution, PassCallback
    case = SE_RUNS["perceptron_gb"]
    model = make_model(case)
    model.init_second_moments()
    prior, x, lin, z, lik = model.forward_ordering
    tau = model.get_second_moments()
    a = {k: 0.0 for k in ("e1", "e2", "e3", "e4", "e5", "e6", "e7", "e8")}

    def msg(source, target, key, direction, var):
        return (source, target, dict(a=a[key], direction=direction, tau=tau[var]))
    vs = []
    for _ in range(2):
        (_, _, d), = prior.forward_state_evolution([msg(x, prior, "e8", "bwd", "x")])
        a["e1"] = d["a"]
        (_, _, d), = x.forward_state_evolution([msg(prior, x, "e1", "fwd", "x"), msg(lin, x, "e7", "bwd", "x")])
        a["e2"] = d["a"]
        (_, _, d), = lin.forward_state_evolution([msg(x, lin, "e2", "fwd", "x"), msg(z, lin, "e6", "bwd", "z")])
        a["e3"] = d["a"]
        (_, _, d), = z.forward_state_evolution([msg(lin, z, "e3", "fwd", "z"), msg(lik, z, "e5", "bwd", "z")])
        a["e4"] = d["a"]
        assert lik.forward_state_evolution([msg(z, lik, "e4", "fwd", "z")]) == []
        (_, _, d), = lik.backward_state_evolution([msg(z, lik, "e4", "fwd", "z")])
        a["e5"] = d["a"]
        (_, _, d), = z.backward_state_evolution([msg(lin, z, "e3", "fwd", "z"), msg(lik, z, "e5", "bwd", "z")])
        a["e6"] = d["a"]
        (_, _, d), = lin.backward_state_evolution([msg(x, lin, "e2", "fwd", "x"), msg(z, lin, "e6", "bwd", "z")])
        a["e7"] = d["a"]
        (_, _, d), = x.backward_state_evolution([msg(prior, x, "e1", "fwd", "x"), msg(lin, x, "e7", "bwd", "x")])
        a["e8"] = d["a"]
        assert prior.backward_state_evolution([msg(x, prior, "e8", "bwd", "x")]) == []
        vs.append((x.posterior_v([msg(prior, x, "e1", "fwd", "x"), msg(lin, x, "e7", "bwd", "x")]),
                   z.posterior_v([msg(lin, z, "e3", "fwd", "z"), msg(lik, z, "e5", "bwd", "z")])))
    se = StateEvolution(make_model(case))
    se.iterate(max_iter=2, callback=PassCallback())
    assert_allclose(vs, np.stack([se.records["vx"][:, 0], se.records["vz"][:, 0]], axis=1), rtol=1e-12)
    assert_allclose([a[f"e{k}"] for k in range(1, 9)], [r["a"] for r in se.get_edges_data(["a"])],
                    rtol=1e-12)
    # free energies through the same protocol (base.py:172-178, 412-419)
    se.update_objective()
    A_x = x.free_energy([msg(prior, x, "e1", "fwd", "x"), msg(lin, x, "e7", "bwd", "x")])
    A_lin = lin.free_energy([msg(x, lin, "e2", "fwd", "x"), msg(z, lin, "e6", "bwd", "z")])
    assert_allclose([A_x, A_lin], [se.A_nodes["x"], se.A_nodes[lin.id]], rtol=1e-12)


def test_synchronous_callback_path_is_bitwise_identical(gold):
    """A callback the kernel cannot replay forces one launch per iteration; the
    trajectory must not change."""
    from tramp_b200.algos import Callback
    case = SE_RUNS["cs_a05"]

    class Spy(Callback):
        def __init__(self):
            self.seen = []

        def __call__(self, algo, i, max_iter):
            self.seen.append((i, algo.get_variable_data("x")["v"], algo.n_iter))
    spy = Spy()
    se1, vx1, vz1 = run_device(case)
    se2, vx2, vz2 = run_device(case, callbacks=[spy])
    assert se1.n_iter == se2.n_iter == len(spy.seen)
    assert np.array_equal(vx1, vx2) and np.array_equal(vz1, vz2)
    assert [s[2] for s in spy.seen] == list(range(1, se2.n_iter + 1))
    assert np.array_equal([s[1] for s in spy.seen], vx2)
    a1 = [r["a"] for r in se1.get_edges_data(["a"])]
    a2 = [r["a"] for r in se2.get_edges_data(["a"])]
    assert a1 == a2


def test_warm_start_continues(gold):
    from tramp_b200.algos import StateEvolution, PassCallback
    case = SE_RUNS["cs_damped"]
    full = StateEvolution(make_model(case))
    full.iterate(max_iter=40, callback=PassCallback(), damping=0.5)
    split = StateEvolution(make_model(case))
    split.iterate(max_iter=15, callback=PassCallback(), damping=0.5)
    split.iterate(max_iter=25, callback=PassCallback(), damping=0.5, warm_start=True)
    assert split.n_iter == full.n_iter == 40
    assert full.get_variables_data() == split.get_variables_data()
    assert_allclose(full.get_variable_data("x")["v"], gold["cs_damped_v_final"][0], rtol=1e-6)


def test_batched_grid_equals_single_runs(gl):
    """A list of models runs in ONE launch (one CTA per model), each with its own
    early stopping; every entry equals the single-model run bit for bit."""
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.algos import StateEvolution, EarlyStopping, TrackEvolution, JoinCallback
    from tramp_b200 import _lib
    alphas = np.linspace(0.05, 1.5, 30)
    rhos = [0.1, 0.4]
    build = [dict(alpha=float(al), prior_rho=rho) for rho in rhos for al in alphas]
    models = [glm_state_evolution(prior_type="gauss_bernoulli", output_type="gaussian",
                                  output_var=1e-3, **kw) for kw in build]
    lib = _lib.load()
    lib.trb_profile_reset(0)
    se = StateEvolution(models)
    evo = TrackEvolution(ids=["x"])
    se.iterate(max_iter=200, callback=JoinCallback([evo, EarlyStopping()]))
    assert lib.trb_profile_launches(-1) == 1
    v = se.get_variable_data("x")["v"]
    assert v.shape == (60,) and se.n_iter == se.n_iter_per_problem.max()
    assert len(set(se.n_iter_per_problem.tolist())) > 3          # they stop at different times
    df = evo.get_dataframe()
    assert len(df) == se.n_iter and df.v.iloc[0].shape == (60,)
    for g in (0, 7, 29, 30, 44, 59):
        one = StateEvolution(models[g])
        one.iterate(max_iter=200, callback=EarlyStopping())
        assert one.n_iter == se.n_iter_per_problem[g]
        assert one.get_variable_data("x")["v"] == v[g]
        r = S.se_glm(dict(kind="gauss_bernoulli", rho=build[g]["prior_rho"], mean=0, var=1),
                     dict(kind="marchenko", alpha=build[g]["alpha"]), dict(kind="gaussian", var=1e-3),
                     200, early=dict(tol=1e-6), integ=gl)
        assert r["n_iter"] == one.n_iter
        assert_allclose(v[g], r["v"][0], rtol=1e-9)
    # more measurements never hurt: v decreases along alpha at fixed rho
    assert np.all(np.diff(v[:30]) < 1e-5)


def test_error_behaviour_mirrors_reference(gold):
    from tramp_b200.priors import GaussianPrior
    from tramp_b200.likelihoods import SgnLikelihood
    from tramp_b200.channels import MarchenkoPasturChannel, GaussianChannel
    from tramp_b200.variables import SISOVariable as V, SILeafVariable as O
    from tramp_b200.algos import StateEvolution
    assert int(gold["gauss_sgn_raises"]) == 1
    model = (GaussianPrior(size=None) @ V(id="x") @ MarchenkoPasturChannel(alpha=2.0) @ V(id="z")
             @ SgnLikelihood(y=None)).to_model()
    with pytest.raises(AssertionError, match="az must be greater"):
        StateEvolution(model).iterate(max_iter=5)
    gen = (GaussianPrior(size=None) @ V(id="x") @ MarchenkoPasturChannel(alpha=2.0) @ V(id="z")
           @ GaussianChannel(var=1.0) @ O(id="y")).to_model()
    with pytest.raises(NotImplementedError):
        StateEvolution(gen)
    se = StateEvolution(make_model(SE_RUNS["cs_a05"]))
    with pytest.raises(NotImplementedError):
        se.iterate(max_iter=2, damping="adaptive")
    with pytest.raises(ValueError, match="not in variables"):
        se.iterate(max_iter=2)
        se.get_variable_data("w")


def test_scenario_run_all_se_and_ep():
    """TeacherStudentScenario.run_all / run_se / se_convergence and
    run_state_evolution (reference teacher_student_scenario.py:54-89, 117-130, 158-178)."""
    from tramp_b200.models import glm_generative, glm_state_evolution
    from tramp_b200.experiments import BayesOptimalScenario, run_state_evolution, run_experiments
    from tramp_b200.algos import EarlyStopping
    np.random.seed(5)
    model = glm_generative(N=400, alpha=0.7, ensemble_type="gaussian", prior_type="gauss_bernoulli",
                           output_type="gaussian", prior_rho=0.2, output_var=1e-2)
    scenario = BayesOptimalScenario(model, x_ids=["x"])
    records = scenario.run_all(metrics=["mse"], max_iter=100, callback=EarlyStopping())
    by = {r["source"]: r for r in records}
    assert set(by) == {"SE", "EP", "mse"}
    # the same script run on the unmodified reference (same seed, hence same W, x, y) printed:
    #   SE v = 0.009469020882561508 (n_iter 14), EP v = 0.01149781034627002 (n_iter 17),
    #   mse = 0.012888367874476584
    assert_allclose(by["SE"]["v"], 0.009469020882561508, rtol=1e-6)
    assert_allclose(by["EP"]["v"], 0.01149781034627002, rtol=1e-6)
    assert_allclose(by["mse"]["v"], 0.012888367874476584, rtol=1e-6)
    assert by["SE"]["n_iter"] == 14 and by["EP"]["n_iter"] == 17
    df = scenario.se_convergence(max_iter=30)
    assert list(df.columns) == ["id", "v", "iter"] and len(df) == scenario.se.n_iter
    assert np.all(np.diff(df.v.values) < 1e-12)
    # Marchenko-Pastur prediction for the same parameters is close to the empirical-spectrum one
    mp = glm_state_evolution(alpha=0.7, prior_type="gauss_bernoulli", output_type="gaussian",
                             prior_rho=0.2, output_var=1e-2)
    rec = run_state_evolution(["x"], mp, max_iter=100)
    assert rec[0]["x_id"] == "x" and abs(rec[0]["v"] - by["SE"]["v"]) < 0.2 * by["SE"]["v"]

    def run(alpha, prior_rho):
        m = glm_state_evolution(alpha=alpha, prior_type="gauss_bernoulli", output_type="gaussian",
                                prior_rho=prior_rho, output_var=1e-2)
        return run_state_evolution(["x"], m, max_iter=100)
    df = run_experiments(run, alpha=[0.3, 0.7], prior_rho=[0.2, 0.4])
    assert len(df) == 4 and set(df.columns) == {"x_id", "v", "n_iter", "alpha", "prior_rho"}
    assert_allclose(df[(df.alpha == 0.7) & (df.prior_rho == 0.2)].v.values[0], rec[0]["v"], rtol=1e-15)


def test_grid_helpers_and_critical_alpha():
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.experiments import (run_state_evolution, run_state_evolution_grid,
                                        find_critical_alpha, find_state_evolution_mse)
    kw = dict(prior_type="gauss_bernoulli", output_type="gaussian", prior_rho=0.3, output_var=1e-10)
    alphas = [0.2, 0.5, 0.8]
    grid = run_state_evolution_grid(["x"], [glm_state_evolution(alpha=a, **kw) for a in alphas],
                                    max_iter=200)
    for a, recs in zip(alphas, grid):
        one = run_state_evolution(["x"], glm_state_evolution(alpha=a, **kw), max_iter=200)
        assert recs == one
    v = find_state_evolution_mse("x", 0.0, np.array(alphas), glm_state_evolution, **kw)
    assert_allclose(v, [g[0]["v"] for g in grid], rtol=1e-15)
    # noiseless compressed sensing from an uninformed start: the recovery threshold of SE
    crit = dict(id="x", a0=0.0, mse_criterion="perfect", alpha_min=0.3, alpha_max=0.95,
                model_builder=glm_state_evolution, alpha_tol=1e-3, vtol=1e-3, **kw)
    a_bisect = find_critical_alpha(**crit)
    a_grid = find_critical_alpha(grid=31, **crit)
    assert abs(a_bisect - a_grid) < 2e-3
    assert 0.3 < a_bisect < 0.95
    # the threshold separates failure from recovery
    lo, hi = find_state_evolution_mse("x", 0.0, np.array([a_bisect - 0.02, a_bisect + 0.02]),
                                      glm_state_evolution, **kw)
    assert lo > 1e-3 > hi
