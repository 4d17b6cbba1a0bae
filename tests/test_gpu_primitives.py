"""GPU parity of the C-ABI primitives against the golden vectors of the
reference and against the CPU oracle (bit-level tolerance noted per test)."""
import os
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.golden.make_golden_specs import PRIOR_SPECS, LIK_SPECS, TRUNC_CASES

pytestmark = pytest.mark.gpu

# FP64 special functions (erfcx, tanh, exp, log) of CUDA and of scipy agree to a
# few ulp; 1e-11 relative leaves two decades of margin below the 1e-9 bar.
RTOL = 1e-11


@pytest.fixture(scope="module")
def ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tramp_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def el(golden_dir):
    return np.load(os.path.join(golden_dir, "elementwise.npz"))


@pytest.fixture(scope="module")
def lin(golden_dir):
    return np.load(os.path.join(golden_dir, "linear.npz"))


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("i", range(len(PRIOR_SPECS)))
def test_prior_posterior_and_log_partition(ops, el, i):
    f = ops.factor_from_spec(PRIOR_SPECS[i])
    a, b = el["grid_a"], el["grid_b"]
    n = a.size
    A = ops.padded(a[None, :])
    Bv = ops.padded(b[None, :])
    r, v = ops.factor_posterior(f, A, Bv, None, n, True, True)
    assert_allclose(_np(r)[0, :n], el[f"prior{i}_r"], rtol=RTOL, atol=1e-300)
    # 1 - tanh^2 cancels near |b| >> 1: a 1-ulp difference in tanh is an absolute
    # 2e-16 on v, so v is compared to a few ulp of 1.0 there
    assert_allclose(_np(v)[0, :n], el[f"prior{i}_v"], rtol=RTOL, atol=1e-15)
    Ael = ops.factor_log_partition(f, A, Bv, None, n, True, True)
    assert_allclose(_np(Ael)[0, :n], el[f"prior{i}_A"], rtol=RTOL, atol=1e-13)
    for j, a_s in enumerate(el["iso_a"]):
        a1 = ops.to_dev(np.array([a_s]))
        Bv = ops.padded((el["grid_bnorm"] * np.sqrt(a_s))[None, :])
        r, v = ops.factor_posterior(f, a1, Bv, None, n, False, False)
        assert_allclose(_np(r)[0, :n], el[f"prior{i}_iso{j}_r"], rtol=RTOL, atol=1e-300)
        assert_allclose(_np(v)[0], el[f"prior{i}_iso{j}_v"], rtol=RTOL)
        Am = ops.factor_log_partition(f, a1, Bv, None, n, False, False)
        assert_allclose(_np(Am)[0], el[f"prior{i}_iso{j}_A"], rtol=1e-10)
        a_io = ops.zeros(1)
        b_io = ops.zeros(1, Bv.shape[1])
        ops.factor_message(f, a1, Bv, None, n, a_io, b_io)
        assert_allclose(_np(a_io)[0], el[f"prior{i}_iso{j}_anew"], rtol=RTOL)
        bn = el[f"prior{i}_iso{j}_bnew"]
        assert_allclose(_np(b_io)[0, :n], bn, rtol=1e-10, atol=1e-10 * np.abs(bn).max())


@pytest.mark.parametrize("i", range(len(LIK_SPECS)))
def test_likelihood_posterior_and_log_partition(ops, el, i):
    f = ops.factor_from_spec(dict(LIK_SPECS[i], role="likelihood"))
    a, b, y = el["grid_a"], el["grid_b"], el[f"lik{i}_y"]
    n = a.size
    A, Bv, Y = ops.padded(a[None, :]), ops.padded(b[None, :]), ops.padded(y[None, :])
    r, v = ops.factor_posterior(f, A, Bv, Y, n, True, True)
    assert_allclose(_np(r)[0, :n], el[f"lik{i}_r"], rtol=RTOL, atol=1e-300)
    # strict on the reference's own test inputs (first 100 grid points); on the
    # stress grid v0*(1 + g2 - g1^2) cancels by up to 1e7, in the reference too
    assert_allclose(_np(v)[0, :100], el[f"lik{i}_v"][:100], rtol=1e-11, atol=1e-15)
    assert_allclose(_np(v)[0, :n], el[f"lik{i}_v"], rtol=1e-7, atol=1e-15 * max(1.0, np.abs(y).max()**2))
    Ael = ops.factor_log_partition(f, A, Bv, Y, n, True, True)
    assert_allclose(_np(Ael)[0, :n], el[f"lik{i}_A"], rtol=RTOL, atol=1e-13)
    for j, a_s in enumerate(el["iso_a"]):
        a1 = ops.to_dev(np.array([a_s]))
        b = el["grid_bnorm"] * np.sqrt(a_s)
        Bv = ops.padded(b[None, :])
        r, v = ops.factor_posterior(f, a1, Bv, Y, n, False, False)
        # sgn: r = r0 + s0*g1 cancels for strongly negative b*y (|r0| up to 2e10 on the
        # stress grid); the reference's own value is only good to eps*|r0| there
        assert_allclose(_np(r)[0, :n], el[f"lik{i}_iso{j}_r"], rtol=RTOL,
                        atol=16 * np.finfo(float).eps * np.abs(b).max() / a_s)
        assert_allclose(_np(r)[0, :100], el[f"lik{i}_iso{j}_r"][:100], rtol=RTOL, atol=1e-13)
        assert_allclose(_np(v)[0], el[f"lik{i}_iso{j}_v"], rtol=RTOL)
        Am = ops.factor_log_partition(f, a1, Bv, Y, n, False, False)
        assert_allclose(_np(Am)[0], el[f"lik{i}_iso{j}_A"], rtol=1e-10)
        a_io = ops.zeros(1)
        b_io = ops.zeros(1, Bv.shape[1])
        ops.factor_message(f, a1, Bv, Y, n, a_io, b_io)
        assert_allclose(_np(a_io)[0], el[f"lik{i}_iso{j}_anew"], rtol=RTOL)
        bn = el[f"lik{i}_iso{j}_bnew"]
        assert_allclose(_np(b_io)[0, :n], bn, rtol=1e-10, atol=1e-10 * np.abs(bn).max())


@pytest.mark.parametrize("i", range(len(TRUNC_CASES)))
def test_truncated_normal(ops, el, i):
    a, lo, hi = TRUNC_CASES[i]
    b = el["trunc_b"]
    r0 = ops.to_dev(b / a)
    v0 = ops.to_dev(np.full_like(b, 1 / a))
    mean, var, logZ, proba = ops.truncated_normal(r0, v0, lo, hi)
    # the "close" Taylor branch and near-cancelling tails lose digits in the
    # reference itself; compare at 1e-9 there, nan == nan
    kw = dict(rtol=1e-9, equal_nan=True)
    assert_allclose(_np(mean), el[f"trunc{i}_r"], atol=1e-12, **kw)
    assert_allclose(_np(var), el[f"trunc{i}_v"], atol=1e-12, **kw)
    assert_allclose(_np(logZ), el[f"trunc{i}_A"], atol=1e-12, **kw)
    # proba = Phi(ymax) - Phi(ymin) (utils/misc.py:50-52) cancels in the tails:
    # both implementations carry an absolute error of a few ulp of 1.0
    assert_allclose(_np(proba), el[f"trunc{i}_p"], atol=1e-15, **kw)


def _thin_svd(W):
    U, s, Vt = np.linalg.svd(W, full_matrices=False)
    return U, s, Vt


@pytest.mark.parametrize("impl", [1, 2])
def test_linear_channel_primitives(ops, lin, impl):
    """rz, vz, rx, vx of LinearChannel through project -> rescale -> expand."""
    for i in range(int(lin["lin_nW"])):
        W = lin[f"lin{i}_W"]
        M, N = W.shape
        rank = int(lin[f"lin{i}_rank"])
        U, s, Vt = _thin_svd(W)
        R = s.size
        Vt_d = ops.padded(Vt).unsqueeze(0).contiguous()      # [1, R, ldn]
        Ut_d = ops.padded(U.T.copy()).unsqueeze(0).contiguous()  # [1, R, ldm]
        s_d = ops.to_dev(s[None, :])
        s2_d = ops.to_dev((s**2)[None, :])
        bz, bx = lin[f"lin{i}_bz"], lin[f"lin{i}_bx"]
        bz_d, bx_d = ops.padded(bz[None, :]), ops.padded(bx[None, :])
        tz = ops.lin_project(Vt_d, R, N, bz_d, 1, impl)
        tx = ops.lin_project(Ut_d, R, M, bx_d, 1, impl)
        assert_allclose(_np(tz)[0], Vt @ bz, rtol=1e-12, atol=1e-13)
        assert_allclose(_np(tx)[0], U.T @ bx, rtol=1e-12, atol=1e-13)
        for j, (az, ax) in enumerate(lin["lin_ab"]):
            if az == 0:
                continue
            az_d, ax_d = ops.to_dev(np.array([az])), ops.to_dev(np.array([ax]))
            coef, vx = ops.lin_rescale(0, 1, R, N, M, rank, s_d, s2_d, az_d, ax_d, tz, tx)
            rx = ops.lin_expand(Ut_d, R, M, coef, 1, impl)
            assert_allclose(_np(vx)[0], lin[f"lin{i}_{j}_vx"], rtol=1e-12)
            ref = lin[f"lin{i}_{j}_rx"]
            # the reference forms rx = W @ rz (linear_channel.py:88); when az << ax the
            # null-space part of rz is ~bz/az and W annihilates it only to roundoff, so
            # its rx carries an absolute error ~eps*|rz|; the thin-SVD form has no such term
            noise = 64 * np.finfo(float).eps * np.abs(lin[f"lin{i}_{j}_rz"]).max()
            assert_allclose(_np(rx)[0, :M], ref, rtol=1e-9, atol=max(noise, 1e-12 * max(1.0, np.abs(ref).max())))
            coef, vz = ops.lin_rescale(1, 1, R, N, M, rank, s_d, s2_d, az_d, ax_d, tz, tx)
            add = bz_d if R < N else None
            rz = ops.lin_expand(Vt_d, R, N, coef, 1, impl, add=add, add_div=az_d if R < N else None)
            # vz = (1 - n_eff)/az: n_eff -> rank/Nz when az << ax, so 1 - n_eff is only
            # good to eps in BOTH implementations (square W: it is ~1e-12 itself)
            assert_allclose(_np(vz)[0], lin[f"lin{i}_{j}_vz"], rtol=1e-12,
                            atol=16 * np.finfo(float).eps / max(az, 1e-11))
            ref = lin[f"lin{i}_{j}_rz"]
            assert_allclose(_np(rz)[0, :N], ref, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("shape", [(3, 40, 100), (5, 64, 1000), (2, 300, 2048), (2, 17, 4096),
                                    (1, 9, 6000), (7, 33, 130)])
def test_gemv_shapes_vs_numpy(ops, impl, shape):
    """project / expand on ragged shapes (odd n, rows not a multiple of the
    chunk, several instances per CTA and several CTAs per instance)."""
    B, R, n = shape
    rng = np.random.RandomState(B * 1000 + R)
    A = rng.randn(B, R, n)
    x = rng.randn(B, n)
    c = rng.randn(B, R)
    ld = ops.pad_ld(n)
    import torch
    A_d = torch.zeros((B, R, ld), dtype=torch.float64, device="cuda")
    A_d[:, :, :n] = torch.as_tensor(A, device="cuda")
    x_d = ops.padded(x)
    t = ops.lin_project(A_d, R, n, x_d, B, impl)
    assert_allclose(_np(t), np.einsum("brn,bn->br", A, x), rtol=1e-11, atol=1e-11)
    out = ops.lin_expand(A_d, R, n, ops.to_dev(c), B, impl)
    assert_allclose(_np(out)[:, :n], np.einsum("brn,br->bn", A, c), rtol=1e-11, atol=1e-11)
    # shared operator (stride 0), config-4 style
    t = ops.lin_project(A_d[:1].contiguous(), R, n, x_d, B, impl)
    assert_allclose(_np(t), np.einsum("rn,bn->br", A[0], x), rtol=1e-11, atol=1e-11)
    # masked instances are skipped
    active = torch.ones(B, dtype=torch.int32, device="cuda")
    active[0] = 0
    t = ops.lin_project(A_d, R, n, x_d, B, impl, active=active)
    assert np.all(_np(t)[0] == 0)
    assert_allclose(_np(t)[1:], np.einsum("brn,bn->br", A, x)[1:], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("shape", [(1, 5, 7), (16, 40, 100), (5, 64, 1000), (70, 129, 131), (130, 17, 257),
                                    (200, 300, 2048), (33, 255, 77), (128, 128, 128), (300, 130, 272),
                                    (64, 256, 160)])
@pytest.mark.parametrize("variant", [0, 1])
def test_shared_operator_dmma_gemm_vs_numpy(ops, shape, variant):
    """The FP64 tensor-core (DMMA m8n8k4) GEMMs of a batch that shares one
    operator: ragged B / R / n (odd sizes, tile tails, 8-byte-aligned coef rows),
    padding columns holding garbage on input and untouched on output.  variant 0 =
    TMA/mbarrier kernel (cp.async one when R is odd), 1 = cp.async kernel."""
    from tramp_b200 import _lib
    _lib.load().trb_gemm_set_variant(variant)
    try:
        _check_dmma_gemm(ops, shape)
    finally:
        _lib.load().trb_gemm_set_variant(0)


def _check_dmma_gemm(ops, shape):
    B, R, n = shape
    rng = np.random.RandomState(B * 1000 + R)
    A = rng.randn(R, n)
    x = rng.randn(B, n)
    c = rng.randn(B, R)
    ld = ops.pad_ld(n)
    import torch
    A_d = torch.full((1, R, ld), float("nan"), dtype=torch.float64, device="cuda")
    A_d[0, :, :n] = torch.as_tensor(A, device="cuda")
    x_d = torch.full((B, ld), float("nan"), dtype=torch.float64, device="cuda")
    x_d[:, :n] = torch.as_tensor(x, device="cuda")
    t = ops.lin_project_gemm(A_d, R, n, x_d, B)
    assert_allclose(_np(t), x @ A.T, rtol=1e-11, atol=1e-11)
    out = torch.full((B, ld), 7.0, dtype=torch.float64, device="cuda")
    ops.lin_expand_gemm(A_d, R, n, ops.to_dev(c), B, out=out)
    assert_allclose(_np(out)[:, :n], c @ A, rtol=1e-11, atol=1e-11)
    assert np.all(_np(out)[:, n:] == 7.0)
    # same numbers as the GEMV path on the shared operator, to summation-order round-off
    A0 = A_d.clone()
    A0[0, :, n:] = 0
    t2 = ops.lin_project(A0, R, n, ops.padded(x), B, 2)
    assert_allclose(_np(t), _np(t2), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape", [(2, 21, 9000), (1, 40, 20000), (3, 10, 16390)])
def test_gemv_wide_operators_use_column_panels(ops, shape):
    """ld > 8192 doubles exceeds one TMA ring stage: the operator is processed as
    equal column panels (BASELINE config 5 has rows of 65536 doubles)."""
    B, R, n = shape
    rng = np.random.RandomState(R)
    A = rng.randn(B, R, n)
    x = rng.randn(B, n)
    c = rng.randn(B, R)
    ld = ops.pad_ld(n)
    import torch
    A_d = torch.zeros((B, R, ld), dtype=torch.float64, device="cuda")
    A_d[:, :, :n] = torch.as_tensor(A, device="cuda")
    t = ops.lin_project(A_d, R, n, ops.padded(x), B)
    assert_allclose(_np(t), np.einsum("brn,bn->br", A, x), rtol=1e-11, atol=1e-10)
    out = ops.lin_expand(A_d, R, n, ops.to_dev(c), B)
    assert_allclose(_np(out)[:, :n], np.einsum("brn,br->bn", A, c), rtol=1e-11, atol=1e-10)


@pytest.mark.parametrize("shape", [(1, 37, 21), (5, 300, 140), (3, 64, 64), (2, 50, 90)])
def test_schedule_primitives_vs_numpy(ops, shape):
    """The device building blocks of the factor-by-factor schedule (trb_adaptive.cu) against the
    reference's own expressions: Variable.compute_log_partition (base.py:146-155),
    LinearChannel.compute_log_partition (linear_channel.py:127-132, dense W), compute_ab_new
    (base.py:250-255), the trial message of the adaptive damping (message_passing.py:169-171)."""
    import torch
    B, N, M = shape
    rng = np.random.RandomState(B * 1000 + N)
    W = rng.randn(B, M, N) / np.sqrt(N)
    a1, a2 = rng.rand(B) + 0.2, rng.rand(B) * 3 + 0.1
    b1, b2 = rng.randn(B, N), rng.randn(B, N)
    bx = rng.randn(B, M)
    d = lambda x: ops.padded(x) if np.ndim(x) == 2 else ops.to_dev(x)      # noqa: E731
    # variable objective (and +inf for a non-positive precision)
    A = _np(ops.variable_log_partition(d(a1), d(b1), d(a2), d(b2), N))
    ref = 0.5 * np.sum((b1 + b2)**2 / (a1 + a2)[:, None] + np.log(2 * np.pi / (a1 + a2))[:, None], axis=1)
    assert_allclose(A, ref, rtol=1e-12)
    assert np.all(np.isinf(_np(ops.variable_log_partition(d(-a1), d(b1), d(0 * a2), d(b2), N))))
    # channel objective from the singular-basis vectors, against the dense formula
    from tramp_b200.channels import LinearChannel
    lin = LinearChannel(W)
    lin._setup()
    tz = ops.lin_project(lin.Vt, lin.R, N, ops.padded(b1, lin.ldn), B)
    tx = ops.lin_project(lin.Ut, lin.R, M, ops.padded(bx, lin.ldm), B)
    bz2 = ops.row_dot(d(b1), d(b1), N)
    assert_allclose(_np(bz2), np.sum(b1 * b1, axis=1), rtol=1e-13)
    A_lin = _np(ops.lin_log_partition(lin.s, lin.s2, N, d(a1), d(a2), tz, tx, bz2 if lin.R < N else None))
    for i in range(B):
        C = W[i].T @ W[i]
        b = b1[i] + W[i].T @ bx[i]
        rz = np.linalg.solve(a1[i] * np.identity(N) + a2[i] * C, b)
        spectrum = np.clip(np.linalg.eigvalsh(C), 0, None)
        ref_i = 0.5 * np.sum(b * rz) + 0.5 * np.sum(np.log(2 * np.pi / (a1[i] + a2[i] * spectrum)))
        assert_allclose(A_lin[i], ref_i, rtol=1e-10)
    # trial message, per-instance and scalar step size
    beta = rng.rand(B)
    a_out, b_out = torch.empty_like(d(a1)), torch.zeros_like(d(b1))
    ops.message_trial(d(a1), d(b1), d(a2), d(b2), N, d(beta), a_out, b_out)
    assert_allclose(_np(b_out)[:, :N], b1 + beta[:, None] * (b2 - b1), rtol=1e-14, atol=1e-15)
    assert_allclose(_np(a_out), a1 + beta * (a2 - a1), rtol=1e-14)
    ops.message_trial(d(a1), d(b1), d(a2), d(b2), N, 0.25, a_out, b_out)
    assert_allclose(_np(b_out)[:, :N], b1 + 0.25 * (b2 - b1), rtol=1e-14, atol=1e-15)
    # compute_ab_new with the clip
    r, v = rng.randn(B, N), np.array([1e-25, 0.5, 1e13, 0.3, 2.0])[:B]
    a_new, b_new = ops.message_from_posterior(d(r), d(v), d(a1), d(b1), N)
    an_ref = np.clip(1 / np.maximum(v, 1e-20) - a1, 1e-11, 1e11)
    assert_allclose(_np(a_new), an_ref, rtol=1e-14)
    assert_allclose(_np(b_new)[:, :N], r * (a1 + an_ref)[:, None] - b1, rtol=1e-13, atol=1e-13)
    # masked row copy
    mask = (np.arange(B) % 2 == 0)
    dst_a, dst_b = d(a2).clone(), d(b2).clone()
    ops.rows_select(ops.to_dev(mask.astype(np.float64)).to(torch.int32), d(a1), d(b1), dst_a, dst_b, N)
    assert_allclose(_np(dst_b)[:, :N], np.where(mask[:, None], b1, b2))
    assert_allclose(_np(dst_a), np.where(mask, a1, a2))
