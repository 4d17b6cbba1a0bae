"""Host-side wrappers over the C ABI: device tensors in, device tensors out.

PyTorch is used for device memory and streams only; all arithmetic on the EP
path happens in libtramp_b200.so.  Vectors are batch-major float64 tensors
`[B, ld]` with the leading dimension padded to a multiple of 16 doubles.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import TrbFactor, ptr, check, current_stream

AMIN = 1e-11  # Factor.AMIN, reference base.py:239
AMAX = 1e+11  # Factor.AMAX, reference base.py:238
LD_ALIGN = 16


def torch():
    import torch as _t
    return _t


def device():
    _lib.require_cuda()
    return torch().device("cuda", torch().cuda.current_device())


def pad_ld(n):
    return ((int(n) + LD_ALIGN - 1) // LD_ALIGN) * LD_ALIGN


def is_tensor(x):
    t = torch()
    return isinstance(x, t.Tensor)


def to_dev(x, dtype=None):
    """numpy / scalar / tensor -> contiguous device tensor (float64 by default)."""
    t = torch()
    dtype = dtype or t.float64
    if is_tensor(x):
        return x.to(device=device(), dtype=dtype).contiguous()
    return t.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)),
                       device=device()).to(dtype).contiguous()


def padded(x2d, ld=None):
    """[B, n] (host or device) -> zero-padded device tensor [B, ld]."""
    t = torch()
    x = to_dev(x2d)
    assert x.dim() == 2
    B, n = x.shape
    ld = ld or pad_ld(n)
    if ld == n:
        return x
    out = t.zeros((B, ld), dtype=t.float64, device=x.device)
    out[:, :n] = x
    return out


def zeros(*shape, dtype=None):
    t = torch()
    return t.zeros(*shape, dtype=dtype or t.float64, device=device())


def make_factor(kind, p0=0.0, p1=0.0, p2=0.0, p3=0.0, amin=AMIN, amax=AMAX):
    return TrbFactor(kind=kind, _pad=0, p0=float(p0), p1=float(p1), p2=float(p2), p3=float(p3),
                     amin=float(amin), amax=float(amax))


# ---------------------------------------------------------------------------
# elementwise factors
# ---------------------------------------------------------------------------
def factor_posterior(f, a, b, y, n, a_elementwise, v_elementwise):
    """a: [B] or [B, ld]; b, y: [B, ld] device tensors.  Returns r [B, ld], v [B] or [B, ld]."""
    t = torch()
    lib = _lib.load()
    B, ld = b.shape
    r = t.zeros_like(b)
    v = t.zeros_like(b) if v_elementwise else t.zeros(B, dtype=t.float64, device=b.device)
    check(lib.trb_factor_posterior(C.byref(f), B, n, ld, ptr(a), int(a_elementwise), ptr(b), ptr(y),
                                   ptr(r), ptr(v), int(v_elementwise), current_stream()))
    return r, v


def factor_log_partition(f, a, b, y, n, a_elementwise, A_elementwise):
    t = torch()
    lib = _lib.load()
    B, ld = b.shape
    A = t.zeros_like(b) if A_elementwise else t.zeros(B, dtype=t.float64, device=b.device)
    check(lib.trb_factor_log_partition(C.byref(f), B, n, ld, ptr(a), int(a_elementwise), ptr(b),
                                       ptr(y), ptr(A), int(A_elementwise), current_stream()))
    return A


def factor_message(f, a_in, b_in, y, n, a_io, b_io, damping=0.0, a_copy=None, scratch=None,
                   flags=None, active=None):
    t = torch()
    lib = _lib.load()
    B, ld = b_in.shape
    if scratch is None:
        scratch = t.empty_like(b_in)
    check(lib.trb_factor_message(C.byref(f), B, n, ld, ptr(a_in), ptr(b_in), ptr(y), ptr(a_io),
                                 ptr(b_io), ptr(a_copy), float(damping or 0.0), ptr(scratch),
                                 ptr(flags), ptr(active), current_stream()))
    return a_io, b_io


def truncated_normal(r0, v0, zmin, zmax):
    """Elementwise truncated-normal (mean, var, logZ, proba) on device tensors."""
    t = torch()
    lib = _lib.load()
    r0 = r0.contiguous()
    v0 = v0.contiguous()
    n = r0.numel()
    outs = [t.empty_like(r0) for _ in range(4)]
    check(lib.trb_truncated_normal(n, ptr(r0), ptr(v0), float(zmin), float(zmax),
                                   *[ptr(o) for o in outs], current_stream()))
    return outs


def posterior_rv(a1, b1, a2, b2, n):
    t = torch()
    lib = _lib.load()
    B, ld = b1.shape
    r = t.zeros_like(b1)
    v = t.zeros(B, dtype=t.float64, device=b1.device)
    check(lib.trb_posterior_rv(B, n, ld, ptr(a1), ptr(b1), ptr(a2), ptr(b2), ptr(r), ptr(v),
                               current_stream()))
    return r, v


# ---------------------------------------------------------------------------
# linear channel primitives
# ---------------------------------------------------------------------------
def lin_project(A, R, n, vec, B, impl=0, active=None, out=None):
    """A: [Bop, R, ld] device; vec: [B, ldvec].  Returns t [B, R]."""
    t_ = torch()
    lib = _lib.load()
    ld = A.shape[-1]
    stride = 0 if A.shape[0] == 1 else A.stride(0)
    if out is None:
        out = t_.zeros((B, R), dtype=t_.float64, device=A.device)
    check(lib.trb_lin_project(ptr(A), stride, R, n, ld, B, ptr(vec), vec.shape[-1], ptr(out),
                              ptr(active), impl, current_stream()))
    return out


def lin_expand_slots(B, R):
    return _lib.load().trb_lin_expand_slots(B, R)


def lin_expand(A, R, n, coef, B, impl=0, active=None, add=None, add_div=None, out=None, part=None):
    """Returns out[B, ld] = sum_i coef[b, i] A[b, i, :] (+ add / add_div[:, None])."""
    t_ = torch()
    lib = _lib.load()
    ld = A.shape[-1]
    stride = 0 if A.shape[0] == 1 else A.stride(0)
    ns = lin_expand_slots(B, R)
    if part is None:
        part = t_.empty((B, ns, ld), dtype=t_.float64, device=A.device)
    check(lib.trb_lin_expand(ptr(A), stride, R, n, ld, B, ptr(coef), ptr(part), ptr(active), impl,
                             current_stream()))
    if out is None:
        out = t_.zeros((B, ld), dtype=t_.float64, device=A.device)
    check(lib.trb_lin_reduce_slots(B, R, n, ld, ptr(part), ptr(add), ptr(add_div), ptr(out),
                                   current_stream()))
    return out


def lin_project_gemm(A, R, n, vec, B, out=None):
    """Shared operator A: [1, R, ld] or [R, ld]; vec: [B, ldvec].  t[B, R] = vec @ A^T (DMMA GEMM)."""
    t_ = torch()
    if out is None:
        out = t_.zeros((B, R), dtype=t_.float64, device=A.device)
    check(_lib.load().trb_lin_project_gemm(ptr(A), R, n, A.shape[-1], B, ptr(vec), vec.shape[-1],
                                           ptr(out), current_stream()))
    return out


def lin_expand_gemm(A, R, n, coef, B, out=None):
    """out[B, ld] = coef[B, R] @ A[R, ld] (DMMA GEMM); columns >= n are not written."""
    t_ = torch()
    ld = A.shape[-1]
    if out is None:
        out = t_.zeros((B, ld), dtype=t_.float64, device=A.device)
    check(_lib.load().trb_lin_expand_gemm(ptr(A), R, n, ld, B, ptr(coef), ptr(out), out.shape[-1],
                                          current_stream()))
    return out


def lin_rescale(direction, B, R, Nz, Nx, rank, s, s2, az, ax, tz, tx, active=None,
                null_space=None, want_coef=True, want_v=True, coef=None, v=None):
    """Spectrum rescale; null_space defaults to R < Nz (pass it explicitly when R
    is a row shard)."""
    t_ = torch()
    lib = _lib.load()
    stride = 0 if s.shape[0] == 1 else s.stride(0)
    if want_coef and coef is None:
        coef = t_.zeros((B, R), dtype=t_.float64, device=s.device)
    if want_v and v is None:
        v = t_.zeros(B, dtype=t_.float64, device=s.device)
    if null_space is None:
        null_space = R < Nz
    check(lib.trb_lin_rescale(direction, B, R, Nz, Nx, rank, int(null_space), ptr(s), ptr(s2), stride,
                              ptr(az), ptr(ax), ptr(tz), ptr(tx), ptr(coef if want_coef else None),
                              ptr(v if want_v else None), ptr(active), current_stream()))
    return coef, v


# ---------------------------------------------------------------------------
# factor-by-factor schedule on the device (trb_adaptive.cu)
# ---------------------------------------------------------------------------
def message_trial(a_old, b_old, a_new, b_new, n, beta, a_out, b_out):
    """(a_out, b_out) = old + beta (new - old); beta a float or a device tensor [B]."""
    B, ld = b_old.shape
    per_instance = is_tensor(beta)
    check(_lib.load().trb_message_trial(B, n, ld, ptr(a_old), ptr(b_old), ptr(a_new), ptr(b_new),
                                        ptr(beta) if per_instance else None, 0.0 if per_instance else float(beta),
                                        ptr(a_out), ptr(b_out), current_stream()))
    return a_out, b_out


def variable_log_partition(a1, b1, a2, b2, n):
    """A[B] of the variable on which the messages (a1, b1) and (a2, b2) meet (a SUM over components)."""
    t = torch()
    B, ld = b1.shape
    A = t.empty(B, dtype=t.float64, device=b1.device)
    check(_lib.load().trb_variable_log_partition(B, n, ld, ptr(a1), ptr(b1), ptr(a2), ptr(b2), ptr(A),
                                                 current_stream()))
    return A


def lin_log_partition(s, s2, Nz, az, ax, tz, tx, bz2):
    t = torch()
    B, R = tz.shape
    stride = 0 if s.shape[0] == 1 else s.stride(0)
    A = t.empty(B, dtype=t.float64, device=tz.device)
    check(_lib.load().trb_lin_log_partition(B, R, Nz, ptr(s), ptr(s2), stride, ptr(az), ptr(ax), ptr(tz), ptr(tx),
                                            ptr(bz2), ptr(A), current_stream()))
    return A


def row_dot(x, y, n):
    t = torch()
    B, ld = x.shape
    out = t.empty(B, dtype=t.float64, device=x.device)
    check(_lib.load().trb_row_dot(B, n, ld, ptr(x), ptr(y), ptr(out), current_stream()))
    return out


def message_from_posterior(r, v, a_in, b_in, n, amin=AMIN, amax=AMAX):
    """compute_ab_new: the (a_new [B], b_new [B, ld]) a channel sends, from its posterior (r, v)."""
    t = torch()
    B, ld = b_in.shape
    a_new = t.empty(B, dtype=t.float64, device=b_in.device)
    b_new = t.zeros_like(b_in)
    check(_lib.load().trb_message_from_posterior(B, n, ld, ptr(r), ptr(v), ptr(a_in), ptr(b_in), float(amin),
                                                 float(amax), ptr(a_new), ptr(b_new), current_stream()))
    return a_new, b_new


def rows_select(mask, src_a, src_b, dst_a, dst_b, n):
    """dst[b] = src[b] for the instances with mask[b] != 0 (mask: int32 device tensor [B])."""
    B, ld = src_b.shape
    check(_lib.load().trb_rows_select(B, n, ld, ptr(mask), ptr(src_a), ptr(src_b), ptr(dst_a), ptr(dst_b),
                                      current_stream()))


# ---------------------------------------------------------------------------
# LinearChannel set-up: block-Jacobi orthogonalisation of rows (trb_setup.cu)
# ---------------------------------------------------------------------------
JACOBI_ROWS = 32    # rows of a block pair: the row count of the work matrix is a multiple of it
JACOBI_COLS = 64    # positions per ring stage: the leading dimension is a multiple of it


def jacobi_workspace(B, n_rows, ld, dev):
    """Scratch of trb_jacobi_sweep: partial Grams, rotations, rotate flags, convergence measure."""
    t = torch()
    lib = _lib.load()
    pairs = n_rows // JACOBI_ROWS
    z = lib.trb_jacobi_zsplit(B, n_rows, ld)
    return dict(S=t.empty((B, pairs, z, JACOBI_ROWS * JACOBI_ROWS), dtype=t.float64, device=dev),
                J=t.empty((B, pairs, JACOBI_ROWS * JACOBI_ROWS), dtype=t.float64, device=dev),
                flag=t.zeros((2, B, pairs), dtype=t.int32, device=dev),     # rotate flags + arrival counters
                off=t.zeros(B, dtype=t.float64, device=dev))


def jacobi_sweep(A, work, skip_tol, max_inner=2):
    """One sweep (every pair of 16-row blocks once) of A[b] <- Q^T A[b], in place.
    Returns the device tensor [B] of the largest |cos| between two rows seen BEFORE their
    rotation in this sweep."""
    lib = _lib.load()
    B, n_rows, ld = A.shape
    check(lib.trb_jacobi_sweep(ptr(A), A.stride(0), B, n_rows, ld, ptr(work["S"]), ptr(work["J"]),
                               ptr(work["flag"]), ptr(work["off"]), float(skip_tol), int(max_inner),
                               current_stream()))
    return work["off"]


def row_norms(A, n):
    """A [B, rows, ld] -> euclidean norms of A[b, i, :n], [B, rows]."""
    t = torch()
    B, rows, ld = A.shape
    out = t.empty((B, rows), dtype=t.float64, device=A.device)
    check(_lib.load().trb_row_norms(ptr(A), A.stride(0), B, rows, n, ld, ptr(out), current_stream()))
    return out


def rows_gather_scale(src, n, perm=None, scale=None, R=None, ld_dst=None, out=None):
    """out[b, i, :n] = scale[b, i] * src[b, perm[b, i], :n] (perm / scale optional), zero padded
    to ld_dst columns.  `out` may be `src` itself when perm is None (in-place row scaling)."""
    t = torch()
    B, rows, ld_src = src.shape
    R = int(R if R is not None else (perm.shape[1] if perm is not None else rows))
    ld_dst = int(ld_dst or n)
    if out is None:
        out = t.empty((B, R, ld_dst), dtype=t.float64, device=src.device)
    if perm is not None:
        perm = perm.to(t.int64).contiguous()
    if scale is not None:
        scale = scale.contiguous()
    check(_lib.load().trb_rows_gather_scale(ptr(src), src.stride(0), ld_src, ptr(perm), ptr(scale), B, R, n,
                                            ptr(out), out.stride(0), out.shape[-1], current_stream()))
    return out


# ---------------------------------------------------------------------------
# natural parameters of the separable factors (host scalars, numpy)
# ---------------------------------------------------------------------------
def gauss_bernoulli_factor(rho, mean, var, amin=AMIN, amax=AMAX):
    """reference priors/gauss_bernoulli_prior.py:33-36 (+ the constant of :82)."""
    a0 = 1 / var
    b0 = mean / var
    normal_A0 = 0.5 * (b0**2 / a0 + np.log(2 * np.pi / a0))
    eta = normal_A0 - np.log(rho / (1 - rho))
    A0 = np.logaddexp(eta, normal_A0)
    return make_factor(_lib.GAUSS_BERNOULLI_PRIOR, a0, b0, eta, A0, amin, amax)


def binary_factor(p_pos, amin=AMIN, amax=AMAX):
    """reference priors/binary_prior.py:26-28."""
    b0 = 0.5 * np.log(p_pos / (1 - p_pos))
    return make_factor(_lib.BINARY_PRIOR, b0, amin=amin, amax=amax)


def gaussian_prior_factor(mean, var, amin=AMIN, amax=AMAX):
    """reference priors/gaussian_prior.py:30-32."""
    return make_factor(_lib.GAUSSIAN_PRIOR, 1 / var, mean / var, amin=amin, amax=amax)


def gaussian_likelihood_factor(var, amin=AMIN, amax=AMAX):
    """reference likelihoods/gaussian_likelihood.py:16."""
    return make_factor(_lib.GAUSSIAN_LIKELIHOOD, 1 / var, amin=amin, amax=amax)


def sgn_factor(amin=AMIN, amax=AMAX):
    return make_factor(_lib.SGN_LIKELIHOOD, amin=amin, amax=amax)


def abs_factor(amin=AMIN, amax=AMAX):
    return make_factor(_lib.ABS_LIKELIHOOD, amin=amin, amax=amax)


def factor_from_spec(spec):
    """dict(kind=..., **params) -> TrbFactor (the spec format of oracle/ and tests/golden)."""
    kind = spec["kind"]
    amin, amax = spec.get("AMIN", AMIN), spec.get("AMAX", AMAX)
    if kind == "gauss_bernoulli":
        return gauss_bernoulli_factor(spec.get("rho", 0.5), spec.get("mean", 0),
                                      spec.get("var", 1), amin, amax)
    if kind == "binary":
        return binary_factor(spec.get("p_pos", 0.5), amin, amax)
    if kind == "gaussian":
        if "y" in spec or spec.get("role") == "likelihood":
            return gaussian_likelihood_factor(spec.get("var", 1), amin, amax)
        return gaussian_prior_factor(spec.get("mean", 0), spec.get("var", 1), amin, amax)
    if kind == "sgn":
        return sgn_factor(amin, amax)
    if kind == "abs":
        return abs_factor(amin, amax)
    raise ValueError(f"unknown factor kind {kind!r}")


# ---------------------------------------------------------------------------
# State Evolution: quadrature rule and beliefs measures
# ---------------------------------------------------------------------------
QUAD_LIMIT = 10.0           # reference utils/integration.py:27, 45: quad(..., -10, 10)
# (panels, order, kappa) of the sinh-mapped composite Gauss-Legendre rule, see
# trb_quadrature in include/tramp_b200.h
QUAD_1D = (160, 32, 1e-7)   # 5120 nodes per 1-D integral
QUAD_2D = (48, 16, 1e-4)    # 768 x 768 nodes per 2-D integral

_quad_cache = {}


def quadrature(rule1=None, rule2=None):
    """Device copy of the Gauss-Legendre templates + the C descriptor of the rule
    (cached per device and size)."""
    rule1, rule2 = tuple(rule1 or QUAD_1D), tuple(rule2 or QUAD_2D)
    dev = device()
    key = (str(dev), rule1, rule2)
    if key not in _quad_cache:
        x1, w1 = np.polynomial.legendre.leggauss(rule1[1])
        x2, w2 = np.polynomial.legendre.leggauss(rule2[1])
        tens = [to_dev(x) for x in (x1, w1, x2, w2)]
        q = _lib.TrbQuadrature(x=ptr(tens[0]), w=ptr(tens[1]), Q=rule1[1], P=rule1[0], kappa=rule1[2],
                               x2=ptr(tens[2]), w2=ptr(tens[3]), Q2=rule2[1], P2=rule2[0],
                               kappa2=rule2[2])
        _quad_cache[key] = (q, tens)
    return _quad_cache[key][0]


def factors_to_dev(factors):
    """list of TrbFactor -> device byte tensor holding the C array."""
    t = torch()
    arr = (TrbFactor * len(factors))(*factors)
    host = t.frombuffer(bytearray(bytes(arr)), dtype=t.uint8)
    return host.to(device())


def se_measure(factors, what, a, tau=None, quad=None):
    """beliefs_measure of every factor at precision a[b] (and second moment tau[b]).
    factors: one TrbFactor (shared) or a list of B.  Returns (out [B], flags [B]) on the device."""
    t = torch()
    lib = _lib.load()
    a = to_dev(a).reshape(-1)
    B = a.numel()
    shared = isinstance(factors, TrbFactor)
    fdev = factors_to_dev([factors] if shared else list(factors))
    if not shared and len(factors) != B:
        raise ValueError("one factor per precision expected")
    tau_d = None if tau is None else to_dev(tau).reshape(-1).expand(B).contiguous()
    out = t.zeros(B, dtype=t.float64, device=a.device)
    flags = t.zeros(B, dtype=t.int32, device=a.device)
    q = quad or quadrature()
    check(lib.trb_se_measure(ptr(fdev), 0 if shared else 1, int(what), B, ptr(a), ptr(tau_d),
                             C.byref(q), ptr(out), ptr(flags), current_stream()))
    return out, flags
