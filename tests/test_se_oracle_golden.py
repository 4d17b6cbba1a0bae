"""CPU tests of the State-Evolution row:

* the SE oracle (oracle/se_oracle.py) is pinned against the golden vectors that
  tests/golden/make_golden_se.py produced from the unmodified reference;
* the quadrature rule the CUDA kernels use (restated in the oracle, integrator
  "gl") agrees with the reference's scipy quad to quad's own tolerance;
* the host side: grid runner, Marchenko-Pastur ensemble / channel closed forms,
  model second moments, argument and error behaviour that needs no GPU.
"""
import os
import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_allclose

from oracle import se_oracle as S
from tests.golden.se_specs import (
    SE_PRIOR_SPECS, SE_PRIOR_AX, SE_LIK_SPECS, SE_LIK_POINTS, SE_ABS_POINTS, SE_MP_ALPHAS,
    SE_MP_POINTS, SE_RUNS, SE_ENTROPY_RUNS, damping_dict, a_init_dict, spectrum_W,
)

# the oracle's "quad" integrator makes the reference's own scipy calls
RTOL_QUAD = 1e-11
# reference quad/dblquad stop at epsabs = epsrel = 1.49e-8 (scipy default): a
# different rule can only agree to that
RTOL_GL, ATOL_GL = 1e-7, 2e-8


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "se.npz"))


@pytest.fixture(scope="module")
def gl():
    return S.Integrator("gl")


def oracle_channel(ch):
    if ch["kind"] == "marchenko":
        return dict(ch)
    W = spectrum_W(ch)
    s = np.linalg.svd(W, compute_uv=False)
    spectrum = np.zeros(W.shape[1])
    spectrum[:len(s)] = s**2
    return dict(kind="spectrum", spectrum=spectrum, Nx=W.shape[0], rank=np.linalg.matrix_rank(W))


def run_oracle(case, integ):
    return S.se_glm(case["prior"], oracle_channel(case["channel"]), case["lik"], case["max_iter"],
                    damping_dict(case.get("damping")), a_init_dict(case.get("a_init")),
                    case.get("early"), integ)


@pytest.mark.parametrize("i", range(len(SE_PRIOR_SPECS)))
def test_prior_measures(gold, gl, i):
    spec = SE_PRIOR_SPECS[i]
    assert_allclose(S.prior_second_moment(spec), gold[f"prior{i}_tau"], rtol=1e-15)
    for integ, rtol, atol in ((S.QUAD, RTOL_QUAD, 0), (gl, RTOL_GL, ATOL_GL)):
        v = [S.prior_forward_error(spec, ax, integ) for ax in SE_PRIOR_AX]
        an = [S.prior_forward_se(spec, ax, integ) for ax in SE_PRIOR_AX]
        A = [S.prior_free_energy(spec, ax, integ) for ax in SE_PRIOR_AX]
        assert_allclose(v, gold[f"prior{i}_v"], rtol=rtol, atol=atol)
        assert_allclose(A, gold[f"prior{i}_A"], rtol=rtol, atol=atol)
        # a_new = 1/v - a amplifies the quadrature tolerance by 1/v^2
        v_ref = gold[f"prior{i}_v"]
        assert np.all(np.abs(np.array(an) - gold[f"prior{i}_anew"])
                      <= rtol * gold[f"prior{i}_anew"] + 2 * atol / v_ref**2 + 1e-300)


@pytest.mark.parametrize("i", range(len(SE_LIK_SPECS)))
def test_likelihood_measures(gold, gl, i):
    spec = SE_LIK_SPECS[i]
    pts = SE_ABS_POINTS if spec["kind"] == "abs" else SE_LIK_POINTS
    for integ, rtol, atol in ((S.QUAD, RTOL_QUAD, 0), (gl, RTOL_GL, ATOL_GL)):
        use = pts[:1] if (spec["kind"] == "abs" and integ is S.QUAD) else pts   # dblquad is slow
        n = len(use)
        v = [S.lik_backward_error(spec, az, tau, integ) for az, tau in use]
        an = [S.lik_backward_se(spec, az, tau, integ) for az, tau in use]
        A = [S.lik_free_energy(spec, az, tau, integ) for az, tau in use]
        assert_allclose(v, gold[f"lik{i}_v"][:n], rtol=rtol, atol=atol)
        assert_allclose(A, gold[f"lik{i}_A"][:n], rtol=rtol, atol=atol)
        v_ref = gold[f"lik{i}_v"][:n]
        assert np.all(np.abs(np.array(an) - gold[f"lik{i}_anew"][:n])
                      <= rtol * gold[f"lik{i}_anew"][:n] + 2 * atol / v_ref**2)


def test_likelihood_domain_assertion():
    """sgn_likelihood.py:80-81 / abs_likelihood.py:57-58."""
    for kind in ("sgn", "abs"):
        with pytest.raises(AssertionError):
            S.lik_backward_error(dict(kind=kind), 1.0, 1.0)


@pytest.mark.parametrize("i", range(len(SE_MP_ALPHAS)))
def test_marchenko_pastur_channel(gold, i):
    alpha = SE_MP_ALPHAS[i]
    ch = dict(kind="marchenko", alpha=alpha, mean_spectrum=S.mp_mean_spectrum(alpha))
    assert_allclose(ch["mean_spectrum"], gold[f"mp{i}_mean_spectrum"], rtol=1e-13)
    vx = [S.channel_forward_error(ch, az, ax) for az, ax in SE_MP_POINTS]
    vz = [S.channel_backward_error(ch, az, ax) for az, ax in SE_MP_POINTS]
    ok = [(az, ax) for az, ax in SE_MP_POINTS if az > 0 and ax > 0]
    A = [S.channel_free_energy(ch, az, ax, 0.7) for az, ax in ok]
    assert_allclose(vx, gold[f"mp{i}_vx"], rtol=1e-13)
    assert_allclose(vz, gold[f"mp{i}_vz"], rtol=1e-13)
    assert_allclose(A, gold[f"mp{i}_A"], rtol=1e-13)
    # the product's closed forms (host scalars; the kernels restate them)
    from tramp_b200.channels import MarchenkoPasturChannel
    mp = MarchenkoPasturChannel(alpha=alpha)
    assert_allclose(mp.ensemble.mean_spectrum, gold[f"mp{i}_mean_spectrum"], rtol=1e-9)
    assert_allclose([mp.compute_forward_error(az, ax, 1.0) for az, ax in SE_MP_POINTS],
                    gold[f"mp{i}_vx"], rtol=1e-9)
    assert_allclose([mp.compute_backward_error(az, ax, 1.0) for az, ax in SE_MP_POINTS],
                    gold[f"mp{i}_vz"], rtol=1e-9)
    assert_allclose([mp.compute_free_energy(az, ax, 0.7) for az, ax in ok], gold[f"mp{i}_A"], rtol=1e-9)
    assert_allclose(mp.ensemble.measure(lambda z: z), alpha, rtol=1e-12)


@pytest.mark.parametrize("name", sorted(SE_RUNS))
def test_runs_quad(gold, name):
    case = SE_RUNS[name]
    if case["lik"]["kind"] == "abs" and name != "phase_bin":
        pytest.skip("dblquad restatement is slow; phase_bin covers the branch")
    r = run_oracle(case, S.QUAD)
    assert r["n_iter"] == int(gold[f"{name}_n_iter"])
    assert_allclose(r["vx"], gold[f"{name}_vx"], rtol=1e-10)
    assert_allclose(r["vz"], gold[f"{name}_vz"], rtol=1e-10)
    assert_allclose(r["a"], gold[f"{name}_a"], rtol=1e-10)
    assert_allclose(r["v"], gold[f"{name}_v_final"], rtol=1e-10)
    assert_allclose(r["tau"], gold[f"{name}_tau"], rtol=1e-13)
    if name in SE_ENTROPY_RUNS:
        H = S.se_entropy(case["prior"], r["channel"], case["lik"], r["a"], S.QUAD)
        assert_allclose(H, gold[f"{name}_entropy"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("name", sorted(SE_RUNS))
def test_runs_device_rule(gold, gl, name):
    """The rule the kernels use reproduces the reference trajectories to the
    reference's own quadrature tolerance (amplified through 1/v - a)."""
    case = SE_RUNS[name]
    r = run_oracle(case, gl)
    assert r["n_iter"] == int(gold[f"{name}_n_iter"])
    assert_allclose(r["vx"], gold[f"{name}_vx"], rtol=1e-6)
    assert_allclose(r["vz"], gold[f"{name}_vz"], rtol=1e-6)
    assert_allclose(r["a"], gold[f"{name}_a"], rtol=1e-6)
    if name in SE_ENTROPY_RUNS:
        H = S.se_entropy(case["prior"], r["channel"], case["lik"], r["a"], gl)
        assert_allclose(H, gold[f"{name}_entropy"], rtol=1e-6, atol=1e-7)


def test_mapped_rule_integrates_the_normal_density():
    for c in (0.0, -0.01, 3.3, -10.0, 25.0, float("nan")):
        for rule in ((160, 32, 1e-7), (48, 16, 1e-4)):
            t, w = S.mapped_rule(*rule, c)
            assert t.min() > -10 and t.max() < 10
            assert_allclose(w.sum(), 1.0, rtol=1e-13)
            assert_allclose((w * t * t).sum(), 1.0, rtol=1e-12)


# ---------------------------------------------------------------- host logic
def test_grid_runner():
    from tramp_b200.experiments import (run_experiments, simple_run_experiments, save_experiments,
                                        get_experiments_from_kwargs)
    pts = get_experiments_from_kwargs(a=[1, 2], b=np.array([3.0, 4.0]), c="x")
    assert pts == [dict(a=1, b=3.0, c="x"), dict(a=1, b=4.0, c="x"),
                   dict(a=2, b=3.0, c="x"), dict(a=2, b=4.0, c="x")]
    seen = []

    def run(a, b, c):
        if a == 2 and b == 4.0:
            raise RuntimeError("boom")
        return [dict(s=a + b), dict(s=-(a + b))] if a == 1 else dict(s=a * b)
    df = run_experiments(run, on_progress=lambda i, n: seen.append((i, n)), a=[1, 2],
                         b=np.array([3.0, 4.0]), c="x")
    assert seen == [(1, 4), (2, 4), (3, 4), (4, 4)]
    assert list(df.columns) == ["s", "a", "b", "c"]
    assert df.s.tolist() == [4.0, -4.0, 5.0, -5.0, 6.0]           # the failing point is skipped
    with pytest.raises(RuntimeError):
        simple_run_experiments(run, a=[2], b=[4.0], c="x")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.csv")
        save_experiments(lambda a: dict(v=a * a), path, a=[1, 2, 3])
        assert pd.read_csv(path).v.tolist() == [1, 4, 9]


def test_binary_and_grid_search():
    from tramp_b200.experiments.critical_alpha import binary_search, grid_search
    xc = 0.637241
    r = binary_search(lambda x: x > xc, 0.0, 2.0, 1e-6)
    assert abs(r["xmid"] - xc) < 1e-6 and r["xerr"] < 1e-6
    g = grid_search(lambda xs: np.asarray(xs) > xc, 0.0, 2.0, 1e-6, grid=15)
    assert abs(g["xmid"] - xc) < 1e-6 and g["n_iter"] <= 6
    with pytest.raises(ValueError):
        binary_search(lambda x: x > xc, 1.0, 2.0, 1e-6)
    with pytest.raises(ValueError):
        grid_search(lambda xs: np.asarray(xs) > xc, 1.0, 2.0, 1e-6, grid=15)


def test_se_model_and_second_moments():
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.channels import MarchenkoPasturChannel
    m = glm_state_evolution(alpha=0.5, prior_type="gauss_bernoulli", output_type="gaussian",
                            prior_rho=0.1, output_var=1e-2)
    kinds = [type(n).__name__ for n in m.forward_ordering]
    assert kinds == ["GaussBernoulliPrior", "SISOVariable", "MarchenkoPasturChannel", "SISOVariable",
                     "GaussianLikelihood"]
    m.init_second_moments()
    tau = m.get_second_moments()
    assert_allclose([tau["x"], tau["z"]], [0.1, 0.1], rtol=1e-12)
    assert isinstance(m.forward_ordering[2], MarchenkoPasturChannel)


def test_state_evolution_host_errors():
    """What can be rejected without a GPU is rejected like the reference does."""
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.algos import StateEvolution
    m = glm_state_evolution(alpha=0.5, prior_type="binary", output_type="sgn")
    with pytest.raises(ValueError, match="not a Model"):
        StateEvolution("nope")
    se = StateEvolution(m)
    with pytest.raises(ValueError, match="damping must be"):
        se.iterate(max_iter=1, damping=3)
    with pytest.raises(ValueError, match="no factor->variable edge"):
        se.configure_damping([("w", "fwd", 0.5)])
    se.configure_damping([("x", "bwd", 0.3), ("z", "fwd", 0.2)])
    assert se.damp == dict(e1=0.0, e3=0.2, e5=0.0, e7=0.3)
    se.configure_damping(0.5)
    assert se.damp == dict(e1=0.5, e3=0.5, e5=0.5, e7=0.5)
    with pytest.raises(ValueError, match="never initialized"):
        StateEvolution(m).iterate(max_iter=1, warm_start=True)
    import torch
    if not torch.cuda.is_available():
        from tramp_b200._lib import TrbError
        with pytest.raises(TrbError, match="no CPU fallback"):
            StateEvolution(m).iterate(max_iter=1)


def test_early_stopping_callback_logic():
    """EarlyStopping.__call__ (reference callbacks.py:206-243) on a scripted trajectory."""
    from tramp_b200.algos import EarlyStopping

    class Fake:
        def __init__(self, vs):
            self.vs, self.i, self.restored = vs, 0, None

        def get_variables_data(self, ids):
            return {"x": dict(v=self.vs[self.i])}

        def snapshot(self):
            return self.i

        def reset_message_dag(self, snap):
            self.restored = snap

    def drive(vs, **kw):
        algo, es = Fake(vs), EarlyStopping(**kw)
        for i in range(len(vs)):
            algo.i = i
            if es(algo, i, len(vs)):
                return i, algo.restored
        return None, algo.restored
    assert drive([1.0, 0.5, 0.5 + 5e-7, 0.1]) == (2, None)                     # tolerance
    assert drive([1.0, 0.5, 0.2, float("nan")]) == (3, 2)                       # nan -> restore
    assert drive([1, .9, .8, .7, .6, .5, .4, .9]) == (7, 6)                     # increase after wait
    assert drive([1, .9, .8, .7, 1.2, .5, .4, .3, .2, .1])[0] is None           # early increase tolerated
    assert drive([1.0, 0.5, 0.05, 0.01], min_variance=0.1) == (2, None)
    # a batched algorithm hands one array per variable: every instance must satisfy the test
    arr = lambda *rows: [np.array(r) for r in rows]
    assert drive(arr([1.0, 1.0], [0.5, 0.5], [0.5 + 5e-7, 0.4], [0.5 + 6e-7, 0.4 + 1e-7])) == (3, None)
    assert drive(arr([1.0, 1.0], [0.5, 0.5], [0.2, float("nan")])) == (2, 1)
