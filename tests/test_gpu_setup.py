"""GPU tests of the hand-written LinearChannel set-up (tramp_b200/csrc/trb_setup.cu): thin SVD
by block Jacobi on the FP64 tensor cores against LAPACK (reference: np.linalg.svd in
channels/linear/linear_channel.py:8-15), and an EP sweep on real Gaussian W at the north-star
shape against the oracle.  The shape / stopping logic also runs in the CPU suite with the
kernels emulated (tests/test_host_logic.py)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.setup_properties import check_thin_svd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("shape,method", [
    ((2, 64, 128), "jacobi"), ((1, 50, 120), "jacobi"), ((3, 48, 48), "jacobi_direct"),
    ((2, 120, 60), "jacobi"), ((1, 5, 9), "auto"), ((1, 1, 4), "auto"), ((1, 40, 41), "auto"),
    ((2, 33, 70), "jacobi_direct"), ((3, 500, 1000), "auto"), ((1, 4000, 2000), "auto"),
    ((2, 1000, 1000), "auto"), ((20, 256, 512), "jacobi"), ((1, 96, 3000), "jacobi_direct")])
def test_thin_svd_against_lapack(shape, method):
    check_thin_svd(shape, method)


def test_jacobi_kernels_against_numpy_one_round():
    """One sweep of trb_jacobi_sweep keeps A^T A (the rows are only rotated), lowers the
    off-diagonal mass, and reports the largest cosine of the pairs it met."""
    import torch
    from tramp_b200 import ops
    rng = np.random.RandomState(0)
    B, n_rows, ld = 3, 96, 128
    A0 = rng.randn(B, n_rows, ld)
    A0[:, 90:] = 0.0                                    # padding rows
    A = ops.to_dev(A0).clone()
    work = ops.jacobi_workspace(B, n_rows, ld, A.device)
    off = ops.jacobi_sweep(A, work, skip_tol=1e-15, max_inner=2).cpu().numpy()
    A1 = A.cpu().numpy()
    for b in range(B):
        assert_allclose(A1[b].T @ A1[b], A0[b].T @ A0[b], atol=1e-11)      # Q orthogonal
        assert np.all(A1[b, 90:] == 0.0)
        g0, g1 = A0[b] @ A0[b].T, A1[b] @ A1[b].T
        off0 = np.linalg.norm(g0 - np.diag(np.diag(g0)))
        off1 = np.linalg.norm(g1 - np.diag(np.diag(g1)))
        assert off1 < 0.7 * off0
        assert 0.05 < off[b] < 1.0


def test_north_star_shape_real_W_setup_and_sweep():
    """SURVEY 8d inputs: W = randn(M, N) / sqrt(N) with N = 4096, M = 2048, factorised by
    LinearChannel(W) (default svd_method: the hand-written set-up), singular values against
    LAPACK on the host, then 20 EP iterations against the oracle on the same W, y."""
    import torch
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.channels import LinearChannel
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    from oracle import tramp_oracle as orc
    B, N, M, n_iter = 2, 4096, 2048, 20
    Ws, xs, ys = [], [], []
    for i in range(B):
        np.random.seed(100 + i)
        W = np.random.randn(M, N) / np.sqrt(N)
        x = np.random.randn(N) * (np.random.rand(N) < 0.1)
        Ws.append(W), xs.append(x), ys.append(W @ x + 0.1 * np.random.randn(M))
    W, x, y = np.stack(Ws), np.stack(xs), np.stack(ys)
    lin = LinearChannel(W)
    model = (GaussBernoulliPrior(size=N, rho=0.1, batch=B) @ V("x") @ lin @ V("z")
             @ GaussianLikelihood(y=y, var=1e-2)).to_model()
    ep = ExpectationPropagation(model)
    track = TrackErrors({"x": x})
    ep.iterate(max_iter=n_iter, callback=track)
    s = lin.s.cpu().numpy()
    got = ep.get_variables_data()
    s_ref = np.linalg.svd(W[0], compute_uv=False)
    assert_allclose(s[0], s_ref, rtol=1e-12)
    ref = orc.ep_glm(dict(kind="gauss_bernoulli", rho=0.1), W[0], dict(kind="gaussian", var=1e-2, y=y[0]),
                     n_iter, x_true=x[0])
    assert_allclose(got["x"]["r"][0], ref["r_x"], rtol=1e-9, atol=1e-9 * np.abs(ref["r_x"]).max())
    assert_allclose(got["x"]["v"][0], ref["v_x"], rtol=1e-9)
    mse = np.array([e["mse"][0] for e in track.errors])
    assert_allclose(mse, np.array(ref["traj"]["mse_x"]), rtol=1e-9)


@pytest.mark.parametrize("shape", [(5, 300, 700), (40, 200, 300), (3, 1100, 2300)])
def test_fused_kernels_match_three_kernel_path(shape):
    """The kernel fusions of the set-up (trb_jacobi_set_fused: one kernel per round for short rows
    and few pairs; Gram + eigenvectors in one kernel, with one or several CTAs per pair) against
    the plain three-launches-per-round path on the same matrices: same factorisation."""
    from tramp_b200 import _lib, ops
    from tramp_b200.channels.linear_channel import thin_svd_device, LAST_SETUP_STATS
    lib = _lib.load()
    W = np.random.RandomState(11).randn(*shape) / np.sqrt(shape[2])
    out = {}
    try:
        for mask in (3, 2, 1, 0):
            lib.trb_jacobi_set_fused(mask)
            Ut, s, Vt = thin_svd_device(ops.to_dev(W), "jacobi")
            out[mask] = (Ut.cpu().numpy(), s.cpu().numpy(), Vt.cpu().numpy(), LAST_SETUP_STATS["sweeps"])
    finally:
        lib.trb_jacobi_set_fused(1)
    s_ref = np.linalg.svd(W, compute_uv=False)
    for mask in (3, 2, 1):
        assert_allclose(out[mask][1], s_ref, rtol=1e-11)
        assert abs(out[mask][3] - out[0][3]) <= 1
        # same subspaces: |<u_i, u_i'>| = 1 up to the sign convention of each run
        dots = np.abs(np.einsum("brm,brm->br", out[mask][0], out[0][0]))
        assert_allclose(dots, 1.0, atol=1e-8)
    # Gram + eigenvectors fused is the same arithmetic in the same order as the two launches
    assert np.array_equal(out[2][1], out[0][1])


def test_rank_deficient_and_ill_conditioned_matrices():
    """`svd_method="auto"`: a wide matrix with two equal rows (rank deficient: the Gram route is
    rejected on its condition number, the direct route on its smallest singular value, the library
    SVD takes over) and a nearly square one (direct route) still give the reference's rank
    (np.linalg.matrix_rank, linear_channel.py:37) and a usable factorisation."""
    from tramp_b200.channels import LinearChannel
    from tramp_b200.channels.linear_channel import LAST_SETUP_STATS
    rng = np.random.RandomState(21)
    W = rng.randn(60, 120) / np.sqrt(120)
    W[-1] = W[0]
    lin = LinearChannel(W)
    lin._setup()
    assert lin.rank == np.linalg.matrix_rank(W) == 59
    s = lin.s.cpu().numpy()[0]
    assert_allclose(s[:59], np.linalg.svd(W, compute_uv=False)[:59], rtol=1e-10)
    rec = (lin.Ut[0, :, :60].T * lin.s[0]) @ lin.Vt[0, :, :120]
    assert_allclose(rec.cpu().numpy(), W, atol=1e-12)
    Wsq = rng.randn(200, 210) / np.sqrt(210)                  # aspect 0.95: cond^2 > 1e4
    lin2 = LinearChannel(Wsq)
    lin2._setup()
    assert LAST_SETUP_STATS["route"] == "direct" and lin2.rank == 200
    assert_allclose(lin2.s.cpu().numpy()[0], np.linalg.svd(Wsq, compute_uv=False), rtol=1e-10)
