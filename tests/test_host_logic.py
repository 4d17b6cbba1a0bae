"""CPU tests: the C-ABI library loads and exports every declared symbol, and
the host-side mirror of the reference API (DAG algebra, Model, initial
conditions, damping configuration, callbacks) behaves like the reference."""
import ctypes
import os
import re
import numpy as np
import pytest
from numpy.testing import assert_allclose
from tests._emulated_device import emulated_device  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tramp_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "tramp_b200.h")).read()
    declared = set(re.findall(r"\b(trb_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tramp_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in tramp_b200/_lib.py"
    assert lib.trb_version() >= 100
    assert lib.trb_sizeof_factor() == ctypes.sizeof(_lib.TrbFactor)
    assert lib.trb_sizeof_sweep() == ctypes.sizeof(_lib.TrbSweep)


def test_argument_errors_do_not_need_a_gpu():
    """Bad arguments are rejected before any launch, with a message."""
    from tramp_b200 import _lib
    lib = _lib.load()
    rc = lib.trb_lin_project(None, 0, 4, 4, 4, 1, None, 4, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.trb_last_error()
    f = _lib.TrbFactor(kind=99)
    rc = lib.trb_factor_posterior(ctypes.byref(f), 1, 4, 4, 1, 0, 1, None, 1, 1, 0, None)
    assert rc == -1 and b"unknown factor kind" in lib.trb_last_error()
    rc = lib.trb_truncated_normal(4, 1, 1, 2.0, 1.0, None, None, None, None, None)
    assert rc == -1 and b"zmin" in lib.trb_last_error()
    assert lib.trb_lin_expand_slots(0, 4) == -1


def test_kernel_choice_switches_and_timeline_without_a_gpu():
    """The switches of trb_sweep_run (rescale inside the projections, chunked updates) and the
    launch timeline are plain host state: callable without a device, and empty before any launch."""
    from tramp_b200 import _lib
    lib = _lib.load()
    for value in (0, 1):
        lib.trb_set_fused_rescale(value)
    for mask in (0, 1, 2, 3, -1):
        lib.trb_set_update_kernels(mask)
    lib.trb_profile_reset(0)
    ms, kinds = (ctypes.c_double * 4)(), (ctypes.c_int * 4)()
    assert lib.trb_profile_timeline(ms, kinds, 4) == 0
    assert lib.trb_profile_launches(-1) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200._lib import TrbError
    with pytest.raises(TrbError, match="no CPU fallback"):
        GaussBernoulliPrior(size=4).compute_forward_posterior(1.0, np.zeros(4))


def _glm(N=12, M=6, batch=None, seed=0):
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.channels import LinearChannel, GaussianChannel
    from tramp_b200.variables import SISOVariable as V, SILeafVariable as O
    rng = np.random.RandomState(seed)
    W = rng.randn(*((batch,) if batch else ()), M, N)
    return (GaussBernoulliPrior(size=N, rho=0.3, batch=batch) @ V("x") @ LinearChannel(W) @ V("z")
            @ GaussianChannel(var=0.1) @ O("y")).to_model()


def test_dag_algebra_and_model():
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.models import Model
    m = _glm()
    names = [type(n).__name__ for n in m.forward_ordering]
    assert names == ["GaussBernoulliPrior", "SISOVariable", "LinearChannel", "SISOVariable",
                     "GaussianChannel", "SILeafVariable"]
    assert m.variable_ids == ["x", "z", "y"] and m.factor_ids == ["f_0", "f_1", "f_2"]
    s = m.sample(seed=3)
    assert s["x"].shape == (12,) and s["z"].shape == (6,) and s["y"].shape == (6,)
    # seed != 0 reseeds the global RNG (base_model.py:73-74): same draw twice
    assert np.array_equal(m.sample(seed=3)["y"], s["y"])
    obs = m.to_observed({"y": s["y"]})
    assert isinstance(obs, Model)
    assert isinstance(obs.forward_ordering[-1], GaussianLikelihood)
    assert obs.forward_ordering[-1].var == 0.1 and obs.variable_ids == ["x", "z"]
    obs.init_shapes()
    assert obs.get_shapes() == {"x": (12,), "z": (6,)}
    with pytest.raises(NotImplementedError):
        m.model_dag + m.model_dag


def test_arity_is_checked():
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.models.dag_algebra import DAG
    dag = GaussBernoulliPrior(size=4) @ V("x")
    with pytest.raises(ValueError):
        dag.to_model()          # dangling placeholder is not a Factor/Variable
    assert isinstance(dag, DAG)


def test_batched_model_sampling_and_shapes():
    m = _glm(batch=3)
    s = m.sample(seed=5)
    assert s["x"].shape == (3, 12) and s["y"].shape == (3, 6)
    obs = m.to_observed({"y": s["y"]})
    obs.init_shapes()
    assert obs.get_shapes() == {"x": (3, 12), "z": (3, 6)}


def test_sampling_matches_reference_rng_order():
    """GaussBernoulliPrior.sample / GaussianChannel.sample consume numpy's global
    RNG exactly as the reference does (gauss_bernoulli_prior.py:38-42,
    gaussian_channel.py:12-15)."""
    m = _glm(N=10, M=5)
    W = m.forward_ordering[2].W
    s = m.sample(seed=11)
    np.random.seed(11)
    xg = np.random.standard_normal(10)
    xb = np.random.binomial(n=1, size=10, p=0.3)
    x = xg * xb
    z = W @ x
    y = z + np.sqrt(0.1) * np.random.standard_normal(5)
    assert np.array_equal(s["x"], x) and np.allclose(s["z"], z) and np.allclose(s["y"], y)


def test_glm_generative_draws_W_first():
    from tramp_b200.models import glm_generative
    np.random.seed(7)
    m = glm_generative(N=8, alpha=0.5, ensemble_type="gaussian", prior_type="binary",
                       output_type="sgn", prior_p_pos=0.6)
    np.random.seed(7)
    W = (1 / np.sqrt(8)) * np.random.randn(4, 8)      # gaussian_ensemble.py:19-20
    assert np.array_equal(m.forward_ordering[2].W, W)
    y = m.sample()["y"]
    assert set(np.unique(y)) <= {-1.0, 1.0}


def test_initial_conditions():
    from tramp_b200.algos import ConstantInit, NoisyInit, CustomInit
    assert ConstantInit(a=2, b=3).init("a", (4,), "x", "fwd") == 2
    assert np.array_equal(ConstantInit(a=2, b=3).init("b", (4,), "x", "fwd"), 3 * np.ones(4))
    np.random.seed(0)
    b = NoisyInit(b_var=4).init("b", (1000,), "x", "fwd")
    assert abs(b.std() - 2) < 0.2
    c = CustomInit(a_init=[("x", "bwd", 5.0)], b_init=[("x", "bwd", np.arange(4.))], a=1, b=0)
    assert c.init("a", (4,), "x", "bwd") == 5.0 and c.init("a", (4,), "x", "fwd") == 1
    assert np.array_equal(c.init("b", (4,), "x", "bwd"), np.arange(4.))
    assert np.array_equal(c.init("b", (4,), "z", "bwd"), np.zeros(4))


def test_damping_configuration_and_chain_validation():
    from tramp_b200.algos import ExpectationPropagation
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.channels import LinearChannel
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.variables import SISOVariable as V
    W, y = np.random.randn(5, 10), np.random.randn(5)
    m = (GaussBernoulliPrior(size=10) @ V("x") @ LinearChannel(W) @ V("z")
         @ GaussianLikelihood(y=y, var=0.1)).to_model()
    ep = ExpectationPropagation(m)
    assert (ep.N, ep.M, ep.B, ep.batched) == (10, 5, 1, False)
    ep.configure_damping(None)
    assert ep.damp == dict(e1=0.0, e3=0.0, e5=0.0, e7=0.0)
    ep.configure_damping(0.5)
    assert ep.damp == dict(e1=0.5, e3=0.5, e5=0.5, e7=0.5)
    ep.configure_damping([("x", "fwd", 0.1), ("z", "bwd", 0.9)])   # message_passing.py:100-105
    assert ep.damp == dict(e1=0.1, e3=0.0, e5=0.9, e7=0.0)
    with pytest.raises(ValueError, match="damping must be"):
        ep.configure_damping(1)
    ep.configure_damping("adaptive")               # message_passing.py:84-88: host-driven schedule
    assert ep.damping and ep.adaptive_damping
    ep.configure_damping(0.5)
    assert not ep.adaptive_damping
    with pytest.raises(ValueError):
        ep.iterate(max_iter=1, warm_start=True)        # message dag was never initialized
    bad = (GaussBernoulliPrior(size=9) @ V("x") @ LinearChannel(W) @ V("z")
           @ GaussianLikelihood(y=y, var=0.1)).to_model()
    with pytest.raises(ValueError, match="does not match"):
        ExpectationPropagation(bad)


def test_callbacks_replay_records_like_the_reference():
    from tramp_b200.algos import TrackErrors, TrackEvolution, JoinCallback, EarlyStoppingEP, TrackEstimate

    class Algo:
        batched, x_id, z_id, variable_ids = False, "x", "z", ["x", "z"]
    rec = dict(mse=np.array([[0.5], [0.25]]), smse=np.array([[0.4], [0.2]]),
               vx=np.array([[1.0], [0.5]]), vz=np.array([[2.0], [1.0]]))
    track = TrackErrors({"x": np.zeros(3)}, metrics=["mse", "sign_mse"])
    evo = TrackEvolution()
    join = JoinCallback([track, evo])
    assert join.device_replayable(Algo)
    for i in range(2):
        join.replay(Algo, i, 2, rec)
    assert track.get_dataframe().to_dict("list") == {"id": ["x", "x"], "iter": [0, 1],
                                                     "mse": [0.5, 0.25], "sign_mse": [0.4, 0.2]}
    df = evo.get_dataframe()
    assert list(df[df.id == "z"].v) == [2.0, 1.0]
    assert not JoinCallback([track, TrackEstimate()]).device_replayable(Algo)
    assert EarlyStoppingEP()._var_mask(Algo) == 3 and EarlyStoppingEP(ids=["x"])._var_mask(Algo) == 1
    assert not TrackErrors({"z": np.zeros(3)}).device_replayable(Algo)


def test_wishart_singular_values_follow_the_gaussian_ensemble():
    """The bidiagonal model used for synthetic data reproduces the singular-value
    law of a Gaussian matrix (compared with direct SVDs, Marchenko-Pastur edges)."""
    from tramp_b200.synthetic import gaussian_singular_values
    M, N = 100, 200
    s = gaussian_singular_values(40, M, N, seed=1, workers=2)
    assert s.shape == (40, M) and np.all(np.diff(s, axis=1) <= 0)
    rng = np.random.RandomState(2)
    d = np.stack([np.linalg.svd(rng.randn(M, N) / np.sqrt(N), compute_uv=False) for _ in range(40)])
    assert abs((s**2).mean() - (d**2).mean()) < 0.01          # E tr(W W^T)/M = 1
    assert abs(s[:, 0].mean() - d[:, 0].mean()) < 0.02 and abs(s[:, -1].mean() - d[:, -1].mean()) < 0.02
    assert abs((s**4).mean() - (d**4).mean()) < 0.05


def test_thin_svd_methods_on_cpu_tensors(emulated_device):
    """Routing of `thin_svd_device`: "auto" takes the hand-written block-Jacobi set-up through the
    Gram matrix when cond(W)^2 <= 1e4 (orthogonality is lost like eps * cond(W)^2 there), on W
    itself otherwise, and the library SVD only for a rank-deficient W.  (Kernels emulated; the
    library baselines "svd" / "gram" are plain tensor algebra.)"""
    import torch
    from tramp_b200.channels.linear_channel import thin_svd_device, LAST_SETUP_STATS
    rng = np.random.RandomState(0)

    def quality(W, method):
        Ut, s, Vt = thin_svd_device(torch.tensor(W)[None], method)
        R = s.shape[1]
        eye = torch.eye(R, dtype=torch.float64)
        rec = (Ut.transpose(1, 2) * s[:, None, :]) @ Vt
        return (s[0].numpy(), float((rec[0] - torch.tensor(W)).abs().max()),
                float((Vt[0] @ Vt[0].T - eye).abs().max()), float((Ut[0] @ Ut[0].T - eye).abs().max()))
    for M, N in ((60, 120), (120, 60)):
        W = rng.randn(M, N) / np.sqrt(N)
        s_ref = np.linalg.svd(W, compute_uv=False)
        for method in ("svd", "gram", "auto", "jacobi", "jacobi_direct"):
            s, rec, orth_v, orth_u = quality(W, method)
            np.testing.assert_allclose(s, s_ref, rtol=1e-11)
            assert rec < 1e-12 and orth_v < 1e-11 and orth_u < 1e-11
        quality(W, "auto")
        assert LAST_SETUP_STATS["route"] == "gram"
    W = rng.randn(80, 80) / np.sqrt(80)             # square Gaussian: cond^2 ~ 1e5 ... 1e7
    s_svd = quality(W, "svd")
    s_auto = quality(W, "auto")
    assert (s_svd[0][0] / s_svd[0][-1])**2 > 1e4 and LAST_SETUP_STATS["route"] == "direct"
    np.testing.assert_allclose(s_auto[0], s_svd[0], rtol=1e-9)
    assert s_auto[1] < 1e-12 and s_auto[2] < 1e-11             # V from the rotations, U = W V / s
    Ww = rng.randn(60, 120) / np.sqrt(120)          # rank deficient: both Jacobi routes rejected
    Ww[-1] = Ww[0]
    np.testing.assert_array_equal(quality(Ww, "auto")[0], quality(Ww, "svd")[0])
    W[-1] = W[0]                                    # rank deficient
    s_auto, s_svd = quality(W, "auto"), quality(W, "svd")
    np.testing.assert_array_equal(s_auto[0], s_svd[0])
    assert s_auto[0][-1] < 1e-12
    with pytest.raises(ValueError):
        thin_svd_device(torch.tensor(W)[None], "qr")


def test_c_abi_from_plain_c(tmp_path):
    """include/tramp_b200.h is valid C99 (no C++-isms, no torch types) and a C
    program links against the shared library with nothing else."""
    import shutil
    import subprocess
    from tramp_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib.load()
    exe = str(tmp_path / "c_abi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi_client.c"), "-o", exe, "-L", libdir, "-ltramp_b200",
           "-Wl,-rpath," + libdir]
    build = subprocess.run(cmd, capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.startswith("ok"), (run.returncode, run.stdout, run.stderr)


def test_factories_name_what_this_build_covers():
    """get_prior / get_likelihood / get_channel / get_ensemble (reference */__init__.py):
    an unknown name is a KeyError as there, with the supported names in the message."""
    from tramp_b200.priors import get_prior
    from tramp_b200.likelihoods import get_likelihood
    from tramp_b200.channels import get_channel
    from tramp_b200.ensembles import get_ensemble
    assert type(get_prior(size=3, prior_type="binary")).__name__ == "BinaryPrior"
    assert type(get_channel("abs")).__name__ == "AbsChannel"
    for call, kw in ((get_prior, dict(size=3, prior_type="exponential")),
                     (get_likelihood, dict(y=None, likelihood_type="modulus")),
                     (get_channel, dict(channel_type="relu")),
                     (get_ensemble, dict(ensemble_type="binary"))):
        with pytest.raises(KeyError, match="not part of tramp_b200"):
            call(**kw)


@pytest.mark.parametrize("shape,method", [((2, 64, 128), "jacobi"), ((1, 50, 120), "jacobi"), ((3, 48, 48), "jacobi_direct"),
                                          ((2, 120, 60), "jacobi"), ((1, 5, 9), "auto"), ((1, 1, 4), "auto"),
                                          ((1, 40, 41), "auto"), ((2, 33, 70), "jacobi_direct")])
def test_jacobi_thin_svd_host_logic(emulated_device, shape, method):
    """The hand-written set-up (svd_method "auto" / "jacobi" / "jacobi_direct"): padding of the work
    matrix, sweep loop and stopping rule, sorting, back-multiplication, tall / wide / odd shapes,
    against LAPACK.  The kernels are emulated here (tests/_emulated_device.py); the same body runs
    on the device in tests/test_gpu_setup.py."""
    from tests.setup_properties import check_thin_svd
    check_thin_svd(shape, method)
