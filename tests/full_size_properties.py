"""Size-independent properties of the EP sweep, written once and run twice:

* on the GPU at the BASELINE.json north-star shape (N = 4096, M = 2048, own W per
  instance) by tests/test_gpu_sizes.py, where running the CPU oracle for every
  instance and iteration would take minutes;
* on CPU tensors at a small shape by tests/test_ep_host_logic_cpu.py, with the sweep
  emulated by the oracle (tests/_emulated_device.py), which pins the expected
  values and tolerances of these checks themselves.

Nothing here reads /root/reference or imports oracle/ except `oracle_sample`.
"""
import numpy as np
from numpy.testing import assert_allclose


def make_batch(B, N, M, rho=0.1, var_noise=1e-2, seed=0):
    """B teacher instances with exactly Gaussian W in factored form (tramp_b200/synthetic.py)."""
    from tramp_b200 import synthetic
    return synthetic.gaussian_glm_batch(B, N, M, rho, var_noise, seed=seed, workers=8)


def _linear(data, lo=None, hi=None):
    from tramp_b200.channels import LinearChannel
    sl = slice(lo, hi)
    Ut, s, Vt = data["Ut"][sl].contiguous(), data["s"][sl].contiguous(), data["Vt"][sl].contiguous()
    M, N = data["y"].shape[1], data["x"].shape[1]
    return LinearChannel.from_factors(Ut, s, Vt, Nx=M, Nz=N, rank=min(M, N))


def sparse_glm_ep(data, var_noise=1e-2, rho=0.1, lo=None, hi=None, schedule="general"):
    """The north-star model on instances [lo, hi) of `data`."""
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation
    sl = slice(lo, hi)
    y = data["y"][sl].contiguous()
    B, N = y.shape[0], data["x"].shape[1]
    model = (GaussBernoulliPrior(size=N, rho=rho, batch=B) @ V("x") @ _linear(data, lo, hi) @ V("z")
             @ GaussianLikelihood(y=y, var=var_noise)).to_model()
    ep = ExpectationPropagation(model)
    ep.schedule = schedule
    return ep


def _run(ep, x_true, n_iter):
    from tramp_b200.algos import TrackErrors, TrackEvolution, JoinCallback
    track, evo = TrackErrors({"x": x_true}), TrackEvolution(ids=["x"])
    ep.iterate(max_iter=n_iter, callback=JoinCallback([track, evo]))
    out = ep.get_variables_data(["x"])
    mse = np.array([np.atleast_1d(e["mse"]) for e in track.errors])       # [n_iter, B]
    return out["x"]["r"], np.atleast_1d(out["x"]["v"]), mse


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


# ---------------------------------------------------------------------------
def linear_gaussian_closed_form(data, var_noise=1e-2, var_prior=1.0, n_iter=3, rtol=1e-10):
    """GaussianPrior @ LinearChannel @ GaussianLikelihood: both outer messages are the
    exact isotropic factors (reference gaussian_prior.py:86-89, gaussian_likelihood.py:
    68-71), so from the first iteration on the posterior of x is the ridge solution
        r_x = V_R diag(s / (s^2 + var_noise / var_prior)) U_R^T y,
        v_x = mean_i 1 / (1 / var_prior + s_i^2 / var_noise)   (s_i = 0 for i >= R)
    and from the second on r_z = W r_x, v_z = mean_j s_j^2 / (...) (linear_channel.py:69-105).
    The expected values are torch FP64 products of the factors, not our kernels."""
    import torch
    from tramp_b200.priors import GaussianPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, PassCallback
    y = data["y"]
    B, M = y.shape
    N = data["x"].shape[1]
    model = (GaussianPrior(size=N, mean=0, var=var_prior, batch=B) @ V("x") @ _linear(data) @ V("z")
             @ GaussianLikelihood(y=y, var=var_noise)).to_model()
    Ut, s, Vt = data["Ut"][:, :, :M], data["s"], data["Vt"][:, :, :N]
    ty = torch.bmm(Ut, y[:, :, None])[:, :, 0]                       # U_R^T y
    precision = 1 / var_prior + s * s / var_noise
    coef = (s / var_noise) / precision * ty
    want_rx = torch.bmm(Vt.transpose(1, 2), coef[:, :, None])[:, :, 0].cpu().numpy()
    want_rz = torch.bmm(Ut.transpose(1, 2), (s * coef)[:, :, None])[:, :, 0].cpu().numpy()
    R = s.shape[1]
    want_vx = (((1 / precision).sum(1) + (N - R) * var_prior) / N).cpu().numpy()
    want_vz = ((s * s / precision).sum(1) / M).cpu().numpy()
    for schedule in ("general", "auto"):
        ep = ExpectationPropagation(model)
        ep.schedule = schedule
        ep.iterate(max_iter=n_iter, callback=PassCallback())
        got = ep.get_variables_data(["x", "z"])
        for b in range(B):
            assert_allclose(got["x"]["r"][b], want_rx[b], rtol=0, atol=rtol * np.abs(want_rx[b]).max())
            assert_allclose(got["z"]["r"][b], want_rz[b], rtol=0, atol=rtol * np.abs(want_rz[b]).max())
        assert_allclose(got["x"]["v"], want_vx, rtol=rtol)
        assert_allclose(got["z"]["v"], want_vz, rtol=rtol)


def schedules_agree(data, n_iter, rtol=1e-10):
    """The 3-pass and 2-pass Gaussian-likelihood schedules are exact rewritings of
    the general 4-pass sweep: same posterior and same MSE trajectory."""
    x_true = data["x"]
    ref = _run(sparse_glm_ep(data, schedule="general"), x_true, n_iter)
    for schedule in ("gauss3", "gauss2", "auto"):
        got = _run(sparse_glm_ep(data, schedule=schedule), x_true, n_iter)
        assert _rel(got[0], ref[0]) <= rtol, schedule
        assert_allclose(got[1], ref[1], rtol=rtol, err_msg=schedule)
        assert_allclose(got[2], ref[2], rtol=rtol, err_msg=schedule)
    return ref


def instances_are_independent(data, ref, lo, hi, n_iter, rtol=1e-10):
    """Instances share nothing: a sub-batch run on its own reproduces its
    rows of the full batch.  Not bitwise: the expansions' partial sums are split
    over CTAs by position in the batch (tramp_b200/csrc/trb_linear.cu)."""
    x_true = data["x"][lo:hi].contiguous()
    got = _run(sparse_glm_ep(data, lo=lo, hi=hi), x_true, n_iter)
    assert _rel(got[0], ref[0][lo:hi]) <= rtol
    assert_allclose(got[1], ref[1][lo:hi], rtol=rtol)
    assert_allclose(got[2], ref[2][:, lo:hi], rtol=rtol)


def bayes_optimal_consistency(ref, rtol=0.25, gain=4.0):
    """Teacher = student: at the fixed point the variance EP reports is the error it
    makes, up to O(N^-1/2) fluctuations (7 % per instance at N = 4096 on the CPU
    oracle), and the error fell by more than `gain` since the first iteration."""
    _, vx, mse = ref
    assert abs(vx.mean() / mse[-1].mean() - 1) < rtol
    assert np.all(mse[-1] * gain < mse[0])


def oracle_sample(data, ref, b, n_iter, var_noise=1e-2, rho=0.1, rtol=1e-9):
    """Instance b against the CPU restatement of the reference (oracle/, full SVD +
    nine GEMVs per iteration) on the same W, y: r_x, v_x and the MSE trajectory to
    the north star's 1e-9."""
    from oracle import tramp_oracle as orc
    from tramp_b200 import synthetic
    W = synthetic.dense_W(data, b)
    x_b, y_b = data["x"][b].cpu().numpy(), data["y"][b].cpu().numpy()
    want = orc.ep_glm(dict(kind="gauss_bernoulli", rho=rho), W, dict(kind="gaussian", var=var_noise, y=y_b),
                      n_iter, x_true=x_b)
    assert _rel(ref[0][b], want["r_x"]) <= rtol
    assert_allclose(ref[1][b], want["v_x"], rtol=rtol)
    assert_allclose(ref[2][:, b], np.array(want["traj"]["mse_x"]), rtol=rtol)
