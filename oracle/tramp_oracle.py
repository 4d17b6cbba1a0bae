"""CPU oracle for the Tree-AMP expectation-propagation (EP) sweep.

TEST INFRASTRUCTURE ONLY.  This module is a plain numpy/scipy restatement of
the reference algorithm (sphinxteam/tramp, paths below are relative to
/root/reference/tramp).  It exists to CHECK the CUDA path and to serve as the
timed CPU baseline in bench.py (`cpu_baseline`, `--impl reference`, kind
"port").  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import it; nothing under tramp_b200/ does, and the product path has no CPU
fallback.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against tests/golden/*.npz, which tests/golden/make_golden.py produced by
running the unmodified reference (under a networkx-1.x veneer) in the build
container on seeded inputs.

The arithmetic follows the reference formula by formula (same operand order,
full SVD with the dense S matrix, nine GEMVs per iteration) so that its timing
is representative of the reference's CPU path; the restructured thin-SVD
schedule lives only in the CUDA product.

Third-party arithmetic on the path (not vendored by the reference, unpinned in
its setup.py:11-14): numpy (OpenBLAS dgemv, LAPACK gesdd) and scipy.special
(erfcx, erf, erfc, expit).  The same libraries are used here.
"""
import numpy as np
from scipy.special import erf, erfc, erfcx, expit  # noqa: F401

AMIN = 1e-11  # base.py:239
AMAX = 1e+11  # base.py:238


# --------------------------------------------------------------------------
# base.py
# --------------------------------------------------------------------------
def safe_inv(v):
    """base.py:44-46 `inv`."""
    return 1 / np.maximum(v, 1e-20)


def ab_new(r, v, a, b, amin=AMIN, amax=AMAX):
    """base.py:250-255 `Factor.compute_ab_new`."""
    a_new = np.clip(safe_inv(v) - a, amin, amax)
    b_new = r * (a + a_new) - b
    return a_new, b_new


def variable_log_partition(ax, bx):
    """base.py:146-150 `Variable.compute_log_partition` (a SUM, inf if ax<=0)."""
    if ax <= 0:
        return np.inf
    return 0.5 * np.sum(bx**2 / ax + np.log(2 * np.pi / ax))


# --------------------------------------------------------------------------
# beliefs/normal.py, binary.py, sparse.py
# --------------------------------------------------------------------------
def normal_A(a, b):
    """beliefs/normal.py:3-4."""
    return 0.5 * (b**2 / a + np.log(2 * np.pi / a))


def binary_A(b):
    """beliefs/binary.py:4-5."""
    return np.logaddexp(b, -b)


def binary_r(b):
    """beliefs/binary.py:8-9."""
    return np.tanh(b)


def binary_v(b):
    """beliefs/binary.py:12-13."""
    return 1 - np.tanh(b)**2


def sparse_A(a, b, eta):
    """beliefs/sparse.py:5-6."""
    return np.logaddexp(eta, normal_A(a, b))


def sparse_p(a, b, eta):
    """beliefs/sparse.py:9-12."""
    return expit(normal_A(a, b) - eta)


def sparse_r(a, b, eta):
    """beliefs/sparse.py:15-17."""
    return sparse_p(a, b, eta) * (b / a)


def sparse_v(a, b, eta):
    """beliefs/sparse.py:20-22."""
    s = sparse_p(a, b, eta)
    return s / a + s * (1 - s) * (b / a)**2


# --------------------------------------------------------------------------
# utils/truncated_normal.py
# --------------------------------------------------------------------------
_SQRT2 = np.sqrt(2)


def log_Phi(x):
    """utils/truncated_normal.py:21-29 (zero for x>=30)."""
    x = np.asarray(x, dtype=float)
    y = np.zeros_like(x)
    m = x < 30
    xm = x[m]
    y[m] = np.log(0.5 * erfcx(-xm / _SQRT2)) - 0.5 * xm**2
    return y


def _branches(x, y, thresh):
    """Branch masks shared by F0/F1/F2 (utils/truncated_normal.py:68-89)."""
    swap = np.abs(x) > np.abs(y)                       # `switch`, :14-18
    x, y = np.where(swap, y, x), np.where(swap, x, y)
    inf = np.isinf(y)
    close = ~inf & (np.abs(x - y) <= thresh)
    neg = (x < 0) & (y < 0)
    pos = (x > 0) & (y > 0)
    other = ~(neg | pos)
    neg = ~inf & ~close & neg
    pos = ~inf & ~close & pos
    other = ~inf & ~close & other
    return x, y, inf, close, neg, pos, other


def F0(x, y, thresh=1e-7):
    """log|erf(y)-erf(x)|, utils/truncated_normal.py:32-89."""
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    x, y, inf, close, neg, pos, other = _branches(x, y, thresh)
    F = np.zeros_like(x)
    xi, yi = x[inf], y[inf]
    F[inf] = np.log(erfcx(np.sign(yi) * xi)) - xi**2                    # :32-34
    xc, e = x[close], y[close] - x[close]
    F[close] = (-xc * e + (1 / 6) * (xc**2 - 2) * e**2                  # :37-45
                - (1 / 180) * (xc**4 + 2 * xc**2 - 8)
                + np.log(2 * e / np.sqrt(np.pi))) - xc**2
    xn, yn = x[neg], y[neg]
    D = np.exp(xn**2 - yn**2)
    F[neg] = np.log(np.abs(D * erfcx(-yn) - erfcx(-xn))) - xn**2        # :48-53
    xp, yp = x[pos], y[pos]
    D = np.exp(xp**2 - yp**2)
    F[pos] = np.log(np.abs(erfcx(xp) - D * erfcx(yp))) - xp**2          # :56-61
    xo, yo = x[other], y[other]
    F[other] = np.log(np.abs(erf(yo) - erf(xo)))                        # :64-65
    return F


def F1(x, y, thresh=1e-7):
    """(exp(-x^2)-exp(-y^2))/(erf(y)-erf(x)), utils/truncated_normal.py:92-144."""
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    x, y, inf, close, neg, pos, other = _branches(x, y, thresh)
    F = np.zeros_like(x)
    xi, yi = x[inf], y[inf]
    F[inf] = np.sign(yi) / erfcx(np.sign(yi) * xi)                      # :92-94
    xc, e = x[close], y[close] - x[close]
    F[close] = np.sqrt(np.pi) * (xc + (1 / 2) * e - (1 / 6) * e**2      # :97-105
                                 - (1 / 12) * e**3
                                 + (1 / 90) * xc * (xc**2 + 1.) * e**4)
    xn, yn = x[neg], y[neg]
    D = np.exp(xn**2 - yn**2)
    F[neg] = (1 - D) / (D * erfcx(-yn) - erfcx(-xn))                    # :108-110
    xp, yp = x[pos], y[pos]
    D = np.exp(xp**2 - yp**2)
    F[pos] = (1 - D) / (erfcx(xp) - D * erfcx(yp))                      # :113-115
    xo, yo = x[other], y[other]
    D = np.exp(xo**2 - yo**2)
    F[other] = np.exp(-xo**2) * (1 - D) / (erf(yo) - erf(xo))           # :118-120
    return F


def F2(x, y, thresh=1e-7):
    """(x exp(-x^2)-y exp(-y^2))/(erf(y)-erf(x)), utils/truncated_normal.py:147-200."""
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    x, y, inf, close, neg, pos, other = _branches(x, y, thresh)
    F = np.zeros_like(x)
    xi, yi = x[inf], y[inf]
    F[inf] = np.sign(yi) * xi / erfcx(np.sign(yi) * xi)                 # :147-149
    xc, e = x[close], y[close] - x[close]
    F[close] = np.sqrt(np.pi) * (xc**2 - 1 / 2 + xc * e                 # :152-161
                                 - (1 / 3) * (xc**2 - 1) * e**2
                                 - (1 / 3) * xc * e**3
                                 + (1 / 90) * (2 * xc**4 + 3 * xc**2 - 8) * e**4)
    xn, yn = x[neg], y[neg]
    D = np.exp(xn**2 - yn**2)
    F[neg] = (xn - D * yn) / (D * erfcx(-yn) - erfcx(-xn))              # :164-166
    xp, yp = x[pos], y[pos]
    D = np.exp(xp**2 - yp**2)
    F[pos] = (xp - D * yp) / (erfcx(xp) - D * erfcx(yp))                # :169-171
    xo, yo = x[other], y[other]
    D = np.exp(xo**2 - yo**2)
    F[other] = np.exp(-xo**2) * (xo - D * yo) / (erf(yo) - erf(xo))     # :174-176
    return F


def _G(ymin, ymax, zmin, zmax, which):
    """G0/G1/G2 dispatch incl. the half-infinite fast path (:203-231)."""
    if zmax == +np.inf:
        x, s = ymin, +1.0
    elif zmin == -np.inf:
        x, s = ymax, -1.0
    else:
        if which == 0:
            return np.log(0.5) + F0(ymin / _SQRT2, ymax / _SQRT2)        # :203-205
        if which == 1:
            return np.sqrt(2 / np.pi) * F1(ymin / _SQRT2, ymax / _SQRT2)  # :208-210
        return (2 / np.sqrt(np.pi)) * F2(ymin / _SQRT2, ymax / _SQRT2)   # :213-215
    if which == 0:
        return log_Phi(-s * x)                                          # :218-221
    u = x / _SQRT2
    if which == 1:
        return np.sqrt(2 / np.pi) * (s / erfcx(s * u))                  # :224-226
    return (2 / np.sqrt(np.pi)) * (s * u / erfcx(s * u))                # :229-231


def truncated_normal_mean(r0, v0, zmin, zmax):
    """utils/truncated_normal.py:234-246."""
    assert zmin < zmax
    s0 = np.sqrt(v0)
    g1 = _G((zmin - r0) / s0, (zmax - r0) / s0, zmin, zmax, 1)
    return r0 + s0 * g1


def truncated_normal_var(r0, v0, zmin, zmax):
    """utils/truncated_normal.py:249-266."""
    assert zmin < zmax
    s0 = np.sqrt(v0)
    ymin, ymax = (zmin - r0) / s0, (zmax - r0) / s0
    g1 = _G(ymin, ymax, zmin, zmax, 1)
    g2 = _G(ymin, ymax, zmin, zmax, 2)
    return v0 * (1. + g2 - g1**2)


def truncated_normal_logZ(r0, v0, zmin, zmax):
    """utils/truncated_normal.py:269-298."""
    assert zmin < zmax
    s0 = np.sqrt(v0)
    g0 = _G((zmin - r0) / s0, (zmax - r0) / s0, zmin, zmax, 0)
    return 0.5 * np.log(2 * np.pi * v0) + 0.5 * r0**2 / v0 + g0


def truncated_normal_proba(r0, v0, zmin, zmax):
    """utils/truncated_normal.py:284-291 with utils/misc.py:50-52 `norm_cdf`."""
    assert zmin < zmax
    s0 = np.sqrt(v0)
    ymin = -np.inf if zmin == -np.inf else (zmin - r0) / s0
    ymax = +np.inf if zmax == +np.inf else (zmax - r0) / s0
    cdf = lambda t: 0.5 * (1 + erf(t / _SQRT2))  # noqa: E731
    return cdf(ymax) - cdf(ymin)


# beliefs/positive.py:8-17 and beliefs/truncated.py:7-25
def positive_A(a, b):
    return truncated_normal_logZ(b / a, 1 / a, 0, np.inf)


def positive_r(a, b):
    return truncated_normal_mean(b / a, 1 / a, 0, np.inf)


def positive_v(a, b):
    return truncated_normal_var(b / a, 1 / a, 0, np.inf)


def truncated_A(a, b, xmin, xmax):
    return truncated_normal_logZ(b / a, 1 / a, xmin, xmax)


def truncated_r(a, b, xmin, xmax):
    return truncated_normal_mean(b / a, 1 / a, xmin, xmax)


def truncated_v(a, b, xmin, xmax):
    return truncated_normal_var(b / a, 1 / a, xmin, xmax)


def truncated_p(a, b, xmin, xmax):
    return truncated_normal_proba(b / a, 1 / a, xmin, xmax)


# --------------------------------------------------------------------------
# priors  (spec = dict(kind=..., size=N, isotropic=True, **params))
# --------------------------------------------------------------------------
def _gb_nat(spec):
    """priors/gauss_bernoulli_prior.py:33-36 natural parameters."""
    a0 = 1 / spec.get("var", 1)
    b0 = spec.get("mean", 0) / spec.get("var", 1)
    rho = spec.get("rho", 0.5)
    eta = normal_A(a0, b0) - np.log(rho / (1 - rho))
    return a0, b0, eta


def prior_forward_posterior(spec, ax, bx):
    """compute_forward_posterior of the three in-scope priors.

    gauss_bernoulli: priors/gauss_bernoulli_prior.py:70-77
    binary:          priors/binary_prior.py:57-63
    gaussian:        priors/gaussian_prior.py:63-68
    """
    kind = spec["kind"]
    iso = spec.get("isotropic", True)
    if kind == "gauss_bernoulli":
        a0, b0, eta = _gb_nat(spec)
        a, b = ax + a0, bx + b0
        rx, vx = sparse_r(a, b, eta), sparse_v(a, b, eta)
        if iso:
            vx = vx.mean()
        return rx, vx
    if kind == "binary":
        p_pos = spec.get("p_pos", 0.5)
        b = bx + 0.5 * np.log(p_pos / (1 - p_pos))
        rx, vx = binary_r(b), binary_v(b)
        if iso:
            vx = vx.mean()
        return rx, vx
    if kind == "gaussian":
        a = ax + 1 / spec.get("var", 1)
        b = bx + spec.get("mean", 0) / spec.get("var", 1)
        return b / a, 1 / a
    raise ValueError(kind)


def prior_forward_message(spec, ax, bx):
    """priors/base_prior.py:13-16; Gaussian override priors/gaussian_prior.py:86-89
    (constant, unclipped)."""
    if spec["kind"] == "gaussian":
        a0 = 1 / spec.get("var", 1)
        b0 = spec.get("mean", 0) / spec.get("var", 1)
        return a0 * np.ones_like(ax), b0 * np.ones_like(bx)
    rx, vx = prior_forward_posterior(spec, ax, bx)
    return ab_new(rx, vx, ax, bx, spec.get("AMIN", AMIN), spec.get("AMAX", AMAX))


def prior_log_partition(spec, ax, bx):
    """compute_log_partition (per-component MEAN): gauss_bernoulli_prior.py:79-83,
    binary_prior.py:65-68, gaussian_prior.py:70-74."""
    kind = spec["kind"]
    if kind == "gauss_bernoulli":
        a0, b0, eta = _gb_nat(spec)
        A = sparse_A(ax + a0, bx + b0, eta) - sparse_A(a0, b0, eta)
        return A.mean()
    if kind == "binary":
        p_pos = spec.get("p_pos", 0.5)
        b0 = 0.5 * np.log(p_pos / (1 - p_pos))
        A = binary_A(bx + b0) - binary_A(b0) - 0.5 * ax
        return A.mean()
    if kind == "gaussian":
        a0 = 1 / spec.get("var", 1)
        b0 = spec.get("mean", 0) / spec.get("var", 1)
        A = normal_A(ax + a0, bx + b0) - normal_A(a0, b0)
        return A.mean()
    raise ValueError(kind)


# --------------------------------------------------------------------------
# likelihoods  (spec = dict(kind=..., y=..., isotropic=True, **params))
# --------------------------------------------------------------------------
def likelihood_backward_posterior(spec, az, bz, y=None):
    """compute_backward_posterior.

    gaussian: likelihoods/gaussian_likelihood.py:43-49
    sgn:      likelihoods/sgn_likelihood.py:32-37 (beliefs/positive.py)
    abs:      likelihoods/abs_likelihood.py:31-36 (beliefs/binary.py)
    """
    kind = spec["kind"]
    y = spec["y"] if y is None else y
    iso = spec.get("isotropic", True)
    if kind == "gaussian":
        ay = 1 / spec.get("var", 1)
        by = ay * y
        a, b = az + ay, bz + by
        return b / a, 1 / a
    if kind == "sgn":
        rz = y * positive_r(az, bz * y)
        vz = positive_v(az, bz * y)
        if iso:
            vz = vz.mean()
        return rz, vz
    if kind == "abs":
        rz = y * binary_r(bz * y)
        vz = (y**2) * binary_v(bz * y)
        if iso:
            vz = vz.mean()
        return rz, vz
    raise ValueError(kind)


def likelihood_backward_message(spec, az, bz):
    """likelihoods/base_likelihood.py:25-28; Gaussian override
    likelihoods/gaussian_likelihood.py:68-71 (constant, unclipped)."""
    if spec["kind"] == "gaussian":
        var = spec.get("var", 1)
        return 1 / var, spec["y"] / var
    rz, vz = likelihood_backward_posterior(spec, az, bz)
    return ab_new(rz, vz, az, bz, spec.get("AMIN", AMIN), spec.get("AMAX", AMAX))


def likelihood_log_partition(spec, az, bz, y=None):
    """compute_log_partition (per-component MEAN): gaussian_likelihood.py:51-56,
    sgn_likelihood.py:39-41, abs_likelihood.py:38-40."""
    kind = spec["kind"]
    y = spec["y"] if y is None else y
    if kind == "gaussian":
        ay = 1 / spec.get("var", 1)
        by = ay * y
        A = normal_A(az + ay, bz + by) - normal_A(ay, by)
        return A.mean()
    if kind == "sgn":
        return positive_A(az, bz * y).mean()
    if kind == "abs":
        return (-0.5 * az * (y**2) + binary_A(bz * y)).mean()
    raise ValueError(kind)


# --------------------------------------------------------------------------
# channels/linear/linear_channel.py
# --------------------------------------------------------------------------
class LinearOp:
    """State of `LinearChannel.__init__` (linear_channel.py:30-46): rank by
    matrix_rank, FULL svd with the dense rectangular S (:8-15)."""

    AMIN, AMAX = AMIN, AMAX          # Factor.AMIN / AMAX (base.py:238-243), per instance after
                                     # reset_precision_bounds

    def __init__(self, W):
        W = np.asarray(W, dtype=float)
        self.W = W
        self.Nx, self.Nz = W.shape
        self.rank = np.linalg.matrix_rank(W)
        self.alpha = self.Nx / self.Nz
        U, s, VT = np.linalg.svd(W, full_matrices=True)
        self.U, self.V = U, VT.T
        S = np.zeros((U.shape[0], self.V.shape[0]))
        S[:len(s), :len(s)] = np.diag(s)
        self.S = S
        self.spectrum = np.diag(S.T @ S)
        self.singular = self.spectrum[:self.rank]


def lin_n_eff(op, az, ax):
    """linear_channel.py:58-67."""
    if ax == 0:
        return 0.
    if az / ax == 0:
        return op.rank / op.Nz
    return np.sum(op.singular / (az / ax + op.singular)) / op.Nz


def lin_backward_mean(op, az, bz, ax, bx):
    """linear_channel.py:69-83 (precompute_svd=True branch)."""
    bx_svd = op.U.T @ bx
    bz_svd = op.V.T @ bz
    resolvent = 1 / (az + ax * op.spectrum)
    rz_svd = resolvent * (bz_svd + op.S.T @ bx_svd)
    return op.V @ rz_svd


def lin_forward_mean(op, az, bz, ax, bx):
    """linear_channel.py:85-89."""
    return op.W @ lin_backward_mean(op, az, bz, ax, bx)


def lin_backward_variance(op, az, ax):
    """linear_channel.py:91-97."""
    az = np.maximum(1e-11, az)
    return (1 - lin_n_eff(op, az, ax)) / az


def lin_forward_variance(op, az, ax):
    """linear_channel.py:99-105."""
    if ax == 0:
        return np.mean(op.singular) * op.rank / (op.Nx * az)
    return lin_n_eff(op, az, ax) / (op.alpha * ax)


def lin_backward_posterior(op, az, bz, ax, bx):
    """linear_channel.py:107-111."""
    return lin_backward_mean(op, az, bz, ax, bx), lin_backward_variance(op, az, ax)


def lin_forward_posterior(op, az, bz, ax, bx):
    """linear_channel.py:113-117."""
    return lin_forward_mean(op, az, bz, ax, bx), lin_forward_variance(op, az, ax)


def lin_log_partition(op, az, bz, ax, bx):
    """linear_channel.py:127-132 (a SUM)."""
    rz = lin_backward_mean(op, az, bz, ax, bx)
    b = bz + op.W.T @ bx
    a = az + ax * op.spectrum
    return 0.5 * np.sum(b * rz) + 0.5 * np.sum(np.log(2 * np.pi / a))


# --------------------------------------------------------------------------
# algos/metrics.py, algos/callbacks.py
# --------------------------------------------------------------------------
def mean_squared_error(x_true, x_pred):
    """algos/metrics.py:5-6."""
    return np.mean((x_true - x_pred)**2)


def sign_symmetric_mse(x_true, x_pred):
    """algos/metrics.py:9-14."""
    return min(np.mean((x_true - x_pred)**2), np.mean((x_true + x_pred)**2))


def _rms(x):
    """algos/callbacks.py:246-247."""
    return np.sqrt(np.mean(x**2))


# --------------------------------------------------------------------------
# algos/message_passing.py + expectation_propagation.py on the chain
#   prior -> x -> linear -> z -> likelihood
# Edges: e1 prior->x, e2 x->lin, e3 lin->z, e4 z->lik (fwd);
#                     e5 lik->z, e6 z->lin, e7 lin->x, e8 x->prior (bwd).
# --------------------------------------------------------------------------
def _damp(d, old, new):
    """message_passing.py:119-127 `compute_constant_damping` (falsy d: no-op)."""
    if not d:
        return new
    return d * old + (1 - d) * new


def ep_glm(prior, W, lik, max_iter, damping=None, init=None, x_true=None,
           early_stopping=None, op=None, record_r=False):
    """Run EP on the observed GLM exactly as `ExpectationPropagation(model)
    .iterate(max_iter, callback, initializer, damping)` does
    (message_passing.py:330-357 with :249-269; sub_variables.py:16-31).

    damping : None | float | dict edge-name -> float  (edges "e1","e3","e5","e7";
              message_passing.py:100-105 only damps factor->variable edges)
    init    : dict edge-name -> (a, b) initial messages; default ConstantInit(0,0)
              (initial_conditions.py:13-24)
    early_stopping : None | dict(tol=1e-6, wait_increase=5, max_increase=0.2)
              = EarlyStoppingEP(ids="all") (callbacks.py:250-286)
    Returns a dict of final edges, posteriors and per-iteration trajectories.
    """
    op = op or LinearOp(W)
    N, M = op.Nz, op.Nx
    sizes = dict(e1=N, e2=N, e7=N, e8=N, e3=M, e4=M, e5=M, e6=M)
    E = {}
    for k, n in sizes.items():
        if init and k in init:
            a0, b0 = init[k]
            E[k] = [a0, np.array(b0, dtype=float) * np.ones(n)]
        else:
            E[k] = [0, np.zeros(n)]
    if isinstance(damping, dict):
        damp = {k: damping.get(k) for k in ("e1", "e3", "e5", "e7")}
    else:
        damp = {k: damping for k in ("e1", "e3", "e5", "e7")}

    def check(name, a, b):
        """message_passing.py:187-209."""
        if np.isnan(a):
            raise ValueError(f"{name} a is nan")
        if np.isnan(b).any():
            raise ValueError(f"{name} b is nan")

    traj = dict(mse_x=[], v_x=[], v_z=[], tol=[], r_x=[], r_z=[])
    old_rs = None
    old_E = None
    n_iter = 0
    status = "max_iter"
    for i in range(max_iter):
        # ---- forward pass (message_passing.py:249-255), Gauss-Seidel ----
        a, b = prior_forward_message(prior, E["e8"][0], E["e8"][1])      # F1
        check("e1", a, b)
        E["e1"] = [_damp(damp["e1"], E["e1"][0], a), _damp(damp["e1"], E["e1"][1], b)]
        E["e2"] = [E["e1"][0], E["e1"][1]]                                # F2
        az, bz, ax, bx = E["e2"][0], E["e2"][1], E["e6"][0], E["e6"][1]
        rx, vx = lin_forward_posterior(op, az, bz, ax, bx)               # F3
        a, b = ab_new(rx, vx, ax, bx, op.AMIN, op.AMAX)
        check("e3", a, b)
        E["e3"] = [_damp(damp["e3"], E["e3"][0], a), _damp(damp["e3"], E["e3"][1], b)]
        E["e4"] = [E["e3"][0], E["e3"][1]]                                # F4
        # ---- backward pass (message_passing.py:257-263) ----
        a, b = likelihood_backward_message(lik, E["e4"][0], E["e4"][1])  # B1
        check("e5", a, b)
        E["e5"] = [_damp(damp["e5"], E["e5"][0], a), _damp(damp["e5"], E["e5"][1], b)]
        E["e6"] = [E["e5"][0], E["e5"][1]]                                # B2
        az, bz, ax, bx = E["e2"][0], E["e2"][1], E["e6"][0], E["e6"][1]
        rz, vz = lin_backward_posterior(op, az, bz, ax, bx)              # B3
        a, b = ab_new(rz, vz, az, bz, op.AMIN, op.AMAX)
        check("e7", a, b)
        E["e7"] = [_damp(damp["e7"], E["e7"][0], a), _damp(damp["e7"], E["e7"][1], b)]
        E["e8"] = [E["e7"][0], E["e7"][1]]                                # B4
        # ---- update_variables (message_passing.py:265-269; base.py:152-161) ----
        r_x = (E["e1"][1] + E["e7"][1]) / (E["e1"][0] + E["e7"][0])
        v_x = 1. / (E["e1"][0] + E["e7"][0])
        r_z = (E["e3"][1] + E["e5"][1]) / (E["e3"][0] + E["e5"][0])
        v_z = 1. / (E["e3"][0] + E["e5"][0])
        n_iter += 1
        traj["v_x"].append(float(v_x))
        traj["v_z"].append(float(v_z))
        if x_true is not None:
            traj["mse_x"].append(mean_squared_error(r_x, x_true))
        if record_r:
            traj["r_x"].append(r_x.copy())
            traj["r_z"].append(r_z.copy())
        # ---- EarlyStoppingEP (callbacks.py:258-286) ----
        if early_stopping is not None:
            new_rs = [r_x, r_z]
            if old_rs is not None:
                tols = [_rms(n - o) / _rms(n) for o, n in zip(old_rs, new_rs)]
                traj["tol"].append(max(tols))
                if max(tols) < early_stopping.get("tol", 1e-6):
                    status = "converged"
                    break
                if (i > early_stopping.get("wait_increase", 5)
                        and max(tols) > early_stopping.get("max_increase", 0.2)):
                    # reset_message_dag(old_message_dag), callbacks.py:281-283: edges AND
                    # the variables' r, v go back to the end of the previous iteration
                    E = old_E
                    r_x, v_x, r_z, v_z = old_post
                    status = "diverged"
                    break
            else:
                traj["tol"].append(np.nan)
            old_rs = new_rs
            old_E = {k: [v[0], v[1].copy()] for k, v in E.items()}
            old_post = (r_x, v_x, r_z, v_z)
    return dict(edges=E, r_x=r_x, v_x=v_x, r_z=r_z, v_z=v_z, n_iter=n_iter,
                status=status, traj=traj, op=op)


def ep_log_evidence(prior, op, lik, E):
    """`ExpectationPropagation.log_evidence()` = message_passing.py:306-328:
    sum of node log-partitions minus sum over fwd edges of the variable
    log-partition of (fwd + bwd) message."""
    A_prior = prior_log_partition(prior, E["e8"][0], E["e8"][1])
    A_x = variable_log_partition(E["e1"][0] + E["e7"][0], E["e1"][1] + E["e7"][1])
    A_lin = lin_log_partition(op, E["e2"][0], E["e2"][1], E["e6"][0], E["e6"][1])
    A_z = variable_log_partition(E["e3"][0] + E["e5"][0], E["e3"][1] + E["e5"][1])
    A_lik = likelihood_log_partition(lik, E["e4"][0], E["e4"][1])
    A_nodes = A_prior + A_x + A_lin + A_z + A_lik
    A_edges = (
        variable_log_partition(E["e1"][0] + E["e8"][0], E["e1"][1] + E["e8"][1])
        + variable_log_partition(E["e2"][0] + E["e7"][0], E["e2"][1] + E["e7"][1])
        + variable_log_partition(E["e3"][0] + E["e6"][0], E["e3"][1] + E["e6"][1])
        + variable_log_partition(E["e4"][0] + E["e5"][0], E["e4"][1] + E["e5"][1])
    )
    return A_nodes - A_edges, dict(prior=A_prior, x=A_x, linear=A_lin, z=A_z,
                                   likelihood=A_lik)
