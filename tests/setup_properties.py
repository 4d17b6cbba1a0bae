"""Checks of the LinearChannel set-up shared by the CPU suite (kernels emulated) and the GPU suite."""
import numpy as np
from numpy.testing import assert_allclose


def check_thin_svd(shape, method, seed=3, rtol_s=1e-11, atol_orth=1e-11):
    """thin_svd_device(W, method) against LAPACK: singular values, ordering, reconstruction,
    orthonormal factors (reference: np.linalg.svd in channels/linear/linear_channel.py:8-15)."""
    import torch
    from tramp_b200 import ops
    from tramp_b200.channels.linear_channel import thin_svd_device
    B, M, N = shape
    W = np.random.RandomState(seed).randn(B, M, N) / np.sqrt(N)
    Ut, s, Vt = thin_svd_device(ops.to_dev(W), method)
    R = min(M, N)
    assert Ut.shape == (B, R, M) and s.shape == (B, R) and Vt.shape == (B, R, N)
    Ut, s, Vt = Ut.cpu(), s.cpu(), Vt.cpu()
    s_ref = np.linalg.svd(W, compute_uv=False)
    assert_allclose(s.numpy(), s_ref, rtol=rtol_s, atol=1e-13 * s_ref.max())
    assert np.all(np.diff(s.numpy(), axis=-1) <= 0)
    assert_allclose(torch.einsum("brm,br,brn->bmn", Ut, s, Vt).numpy(), W, atol=1e-12)
    # a factor obtained as W V / s (or its transpose) is orthonormal to eps * cond(W)
    atol_orth = max(atol_orth, 64 * np.finfo(float).eps * float((s_ref[:, 0] / s_ref[:, -1]).max()) * R**0.5)
    eye = np.broadcast_to(np.eye(R), (B, R, R))
    assert_allclose((Ut @ Ut.transpose(1, 2)).numpy(), eye, atol=atol_orth)
    assert_allclose((Vt @ Vt.transpose(1, 2)).numpy(), eye, atol=atol_orth)
