"""Synthetic Gaussian sensing matrices generated directly in thin-SVD form.

A matrix W with iid N(0, 1/N) entries (reference ensembles/gaussian_ensemble.py:
11-21) is bi-orthogonally invariant: W = U diag(s) V^T with U Haar on O(M), the
columns of V Haar on the Stiefel manifold, and s distributed as the singular
values of such a matrix, all independent.  Drawing the three factors directly
therefore yields EXACTLY Gaussian-ensemble instances while skipping the
O(M^2 N) dense SVD per instance (1.3 s per 2048x4096 instance with cuSOLVER on
B200 -- 11 minutes for the 512 instances of one benchmark shard):

  * Haar factors: Gaussian matrix -> Cholesky-QR applied twice (BLAS-3 only);
    Q = A R^-1 with R's diagonal positive is the Haar-distributed QR factor.
  * singular values: Dumitriu-Edelman bidiagonal model -- G G^T for G in
    R^{M x N} iid N(0,1) has the spectrum of B B^T, B lower bidiagonal with
    diag chi_N, chi_{N-1}, ..., chi_{N-M+1} and sub-diagonal chi_{M-1}, ..., chi_1.

This is benchmark / test data plumbing (setup is outside the EP hot path, and
the reference reports its SVD time separately, compute_benchmark.py:27).
"""
import numpy as np

from . import ops


def haar_rows(B, R, n, generator, chunk=16, ld=None):
    """[B, R, ld] device tensor whose rows (length n, zero padded to ld) are
    orthonormal and Haar distributed."""
    t = ops.torch()
    ld = ld or ops.pad_ld(n)
    dev = ops.device()
    out = t.zeros((B, R, ld), dtype=t.float64, device=dev)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        At = t.randn((b1 - b0, R, n), dtype=t.float64, device=dev, generator=generator)
        for _ in range(2):   # CholeskyQR2: second pass restores orthogonality to ~eps
            G = At @ At.transpose(1, 2)
            L = t.linalg.cholesky(G)
            At = t.linalg.solve_triangular(L, At, upper=False)
        out[b0:b1, :, :n] = At
    return out


def _wishart_singular_values(args):
    M, N, seed = args
    from scipy.linalg import eigvalsh_tridiagonal
    rng = np.random.RandomState(seed)
    R = min(M, N)
    big = max(M, N)
    d = np.sqrt(rng.chisquare(big - np.arange(R)))
    e = np.sqrt(rng.chisquare(np.arange(R - 1, 0, -1))) if R > 1 else np.zeros(0)
    Td = d**2
    Td[1:] += e**2
    ev = eigvalsh_tridiagonal(Td, d[:-1] * e) if R > 1 else Td
    return np.sqrt(np.maximum(ev[::-1], 0.0))


def gaussian_singular_values(B, M, N, seed, workers=8):
    """[B, R] singular values (descending) of B independent M x N matrices with
    iid N(0, 1/N) entries."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = [(M, N, seed + 7919 * i) for i in range(B)]
    with ThreadPoolExecutor(max_workers=workers) as ex:
        sv = list(ex.map(_wishart_singular_values, jobs))
    return np.stack(sv) / np.sqrt(N)


def gaussian_glm_batch(B, N, M, rho=0.1, var_noise=1e-2, seed=0, chunk=16, workers=8):
    """B independent sparse-GLM teacher instances with Gaussian W in factored form.

    Returns dict(Ut [B,R,ldm], s [B,R], Vt [B,R,ldn], x [B,N], y [B,M]) on the
    current device; x ~ GaussBernoulli(rho, 0, 1), y = W x + sqrt(var_noise) noise.
    """
    t = ops.torch()
    dev = ops.device()
    gen = t.Generator(device=dev)
    gen.manual_seed(seed)
    R = min(M, N)
    Vt = haar_rows(B, R, N, gen, chunk)
    Ut = haar_rows(B, R, M, gen, chunk)
    s = t.as_tensor(gaussian_singular_values(B, M, N, seed, workers), device=dev)
    x = t.randn((B, N), dtype=t.float64, device=dev, generator=gen)
    x = x * (t.rand((B, N), dtype=t.float64, device=dev, generator=gen) < rho)
    tz = t.bmm(Vt[:, :, :N], x[:, :, None])[:, :, 0]
    z = t.bmm(Ut[:, :, :M].transpose(1, 2), (s * tz)[:, :, None])[:, :, 0]
    y = z + np.sqrt(var_noise) * t.randn((B, M), dtype=t.float64, device=dev, generator=gen)
    return dict(Ut=Ut, s=s, Vt=Vt, x=x, y=y, z=z)


def dense_W(batch, b):
    """Reconstruct instance b's dense W (host numpy), e.g. for the CPU oracle."""
    t = ops.torch()
    Ut, s, Vt = batch["Ut"][b], batch["s"][b], batch["Vt"][b]
    M, N = batch["y"].shape[1], batch["x"].shape[1]
    W = (Ut[:, :M].transpose(0, 1) * s[None, :]) @ Vt[:, :N]
    return W.cpu().numpy()
