"""CPU oracle for Tree-AMP State Evolution (SE) on the generalized linear model.

TEST INFRASTRUCTURE ONLY.  A plain numpy/scipy restatement of the reference's
scalar State Evolution (sphinxteam/tramp; paths below are relative to
/root/reference/tramp) for the chain prior -> x -> linear -> z -> likelihood.
It exists to CHECK the CUDA path (tramp_b200/csrc/trb_se.cu); only tests/ and
bench tooling may import it, nothing under tramp_b200/ does.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against tests/golden/se.npz, produced by tests/golden/make_golden_se.py running
the unmodified reference in the build container.

Two integrators:
  * "quad" (default): scipy.integrate.quad / dblquad on [-10, 10], exactly the
    calls of utils/integration.py:13-46 -- this is what the golden vectors pin.
  * "gl": the sinh-mapped composite Gauss-Legendre rule the CUDA kernels use
    (restated here, nothing is imported from tramp_b200).  It separates quadrature error from implementation error: the device
    must agree with "gl" to rounding, and "gl" with "quad" to quad's tolerance.

Third-party arithmetic (unpinned in the reference's setup.py): scipy.integrate
(QUADPACK qagse through quad/dblquad, default epsabs = epsrel = 1.49e-8),
scipy.special.
"""
import numpy as np
from scipy.integrate import quad, dblquad

from . import tramp_oracle as O

LIMIT = 10.0


def norm_pdf(x):
    """utils/misc.py:46-47."""
    return np.exp(-0.5 * x**2) / np.sqrt(2 * np.pi)


# --------------------------------------------------------------------------
# integrators
# --------------------------------------------------------------------------
def mapped_rule(panels, order, kappa, c, limit=LIMIT):
    """Nodes t and weights (times the normal density) of the rule the CUDA kernels
    use: composite Gauss-Legendre in u with t = c + kappa sinh(u), the centre c
    clamped into [-limit, limit] (include/tramp_b200.h, trb_quadrature)."""
    c = min(max(float(c), -limit), limit) if c == c else 0.0
    x, w = np.polynomial.legendre.leggauss(order)
    u_lo = np.arcsinh((-limit - c) / kappa)
    du = (np.arcsinh((limit - c) / kappa) - u_lo) / panels
    u = (u_lo + (np.arange(panels)[:, None] + 0.5 * (x[None, :] + 1.0)) * du).ravel()
    t = c + kappa * np.sinh(u)
    wt = (0.5 * du * np.tile(w, panels)) * (kappa * np.cosh(u)) * norm_pdf(t)
    return t, wt


class Integrator:
    """`centre` arguments say where the integrand has its structure (the image
    of b = -b0); only the "gl" rule uses them, quad adapts by itself."""

    def __init__(self, kind="quad", rule1=(160, 32, 1e-7), rule2=(48, 16, 1e-4)):
        self.kind = kind
        self.rule1, self.rule2 = rule1, rule2

    def gaussian_measure(self, m, s, f, centre=0.0):
        """utils/integration.py:13-28."""
        if self.kind == "gl":
            t, w = mapped_rule(*self.rule1, centre)
            return float(np.sum(w * f(m + s * t)))
        return quad(lambda x: norm_pdf(x) * f(m + s * x), -LIMIT, LIMIT)[0]

    def gaussian_measure_2d(self, m1, s1, m2, s2, f, centre2=None):
        """utils/integration.py:31-46.  centre2(x1-argument of f) -> centre of the inner rule."""
        if self.kind == "gl":
            t1, w1 = mapped_rule(*self.rule2, 0.0)
            total = 0.0
            for x1, wx in zip(t1, w1):
                z = m1 + s1 * x1
                t2, w2 = mapped_rule(*self.rule2, centre2(z) if centre2 else 0.0)
                total += wx * np.sum(w2 * f(z, m2 + s2 * t2))
            return float(total)
        return dblquad(lambda x2, x1: norm_pdf(x1) * norm_pdf(x2) * f(m1 + s1 * x1, m2 + s2 * x2),
                       -LIMIT, LIMIT, -LIMIT, LIMIT)[0]


QUAD = Integrator("quad")


# --------------------------------------------------------------------------
# priors (spec as in tramp_oracle: dict(kind=..., **params))
# --------------------------------------------------------------------------
def prior_second_moment(spec):
    """gauss_bernoulli_prior.py:47-48, binary_prior.py:38-39, gaussian_prior.py:40-41."""
    kind = spec["kind"]
    if kind == "gauss_bernoulli":
        return spec.get("rho", 0.5) * (spec.get("mean", 0)**2 + spec.get("var", 1))
    if kind == "binary":
        return 1.
    return spec.get("mean", 0)**2 + spec.get("var", 1)


def prior_scalar(spec, what, ax, bx):
    """scalar_forward_variance ("v") / scalar_log_partition ("A"):
    gauss_bernoulli_prior.py:59-68, binary_prior.py:48-55."""
    kind = spec["kind"]
    if kind == "gauss_bernoulli":
        a0, b0, eta = O._gb_nat(spec)
        a, b = ax + a0, bx + b0
        if what == "v":
            return O.sparse_v(a, b, eta)
        return O.sparse_A(a, b, eta) - O.sparse_A(a0, b0, eta)
    if kind == "binary":
        p_pos = spec.get("p_pos", 0.5)
        b0 = 0.5 * np.log(p_pos / (1 - p_pos))
        if what == "v":
            return O.binary_v(bx + b0)
        return O.binary_A(bx + b0) - O.binary_A(b0) - 0.5 * ax
    raise ValueError(kind)


def prior_beliefs_measure(spec, ax, what, integ=QUAD):
    """gauss_bernoulli_prior.py:112-118, binary_prior.py:80-84."""
    kind = spec["kind"]

    def f(bx):
        return prior_scalar(spec, what, ax, bx)
    if kind == "gauss_bernoulli":
        rho, mean, var = spec.get("rho", 0.5), spec.get("mean", 0), spec.get("var", 1)
        b0 = mean / var
        s0, m1, s1 = np.sqrt(ax), ax * mean, np.sqrt(ax + (ax**2) * var)
        mu_0 = integ.gaussian_measure(0, s0, f, -b0 / s0 if s0 > 0 else 0.0)
        mu_1 = integ.gaussian_measure(m1, s1, f, (-b0 - m1) / s1 if s1 > 0 else 0.0)
        return (1 - rho) * mu_0 + rho * mu_1
    if kind == "binary":
        p_pos = spec.get("p_pos", 0.5)
        b0, s0 = 0.5 * np.log(p_pos / (1 - p_pos)), np.sqrt(ax)
        mu_pos = integ.gaussian_measure(+ax, s0, f, (-b0 - ax) / s0 if s0 > 0 else 0.0)
        mu_neg = integ.gaussian_measure(-ax, s0, f, (-b0 + ax) / s0 if s0 > 0 else 0.0)
        return p_pos * mu_pos + (1 - p_pos) * mu_neg
    raise ValueError(kind)


def prior_forward_error(spec, ax, integ=QUAD):
    """base_prior.py:71-74; gaussian_prior.py:92-95."""
    if spec["kind"] == "gaussian":
        return 1 / (ax + 1 / spec.get("var", 1))
    return prior_beliefs_measure(spec, ax, "v", integ)


def a_new(v, a, amin=O.AMIN, amax=O.AMAX):
    """base.py:245-248 `compute_a_new`."""
    return np.clip(O.safe_inv(v) - a, amin, amax)


def prior_forward_se(spec, ax, integ=QUAD):
    """base_prior.py:66-69; gaussian_prior.py:102-104 (constant)."""
    if spec["kind"] == "gaussian":
        return 1 / spec.get("var", 1)
    return a_new(prior_forward_error(spec, ax, integ), ax,
                 spec.get("AMIN", O.AMIN), spec.get("AMAX", O.AMAX))


def prior_free_energy(spec, ax, integ=QUAD):
    """base_prior.py:82-85; gaussian_prior.py:129-138."""
    if spec["kind"] == "gaussian":
        var = spec.get("var", 1)
        I = 0.5 * np.log((ax + 1 / var) * var)
        return 0.5 * ax * prior_second_moment(spec) - I
    return prior_beliefs_measure(spec, ax, "A", integ)


# --------------------------------------------------------------------------
# likelihoods
# --------------------------------------------------------------------------
def lik_scalar(spec, what, az, bz, y):
    """scalar_backward_variance / compute_log_partition of one component:
    sgn_likelihood.py:27-30, 39-41; abs_likelihood.py:26-29, 38-40."""
    kind = spec["kind"]
    if kind == "sgn":
        return O.positive_v(az, bz * y) if what == "v" else O.positive_A(az, bz * y)
    if kind == "abs":
        if what == "v":
            return (y**2) * O.binary_v(bz * y)
        return -0.5 * az * (y**2) + O.binary_A(bz * y)
    raise ValueError(kind)


def positive_p(a, b):
    """beliefs/positive.py:24-26."""
    return O.truncated_normal_proba(b / a, 1 / a, 0, np.inf)


def lik_beliefs_measure(spec, az, tau_z, what, integ=QUAD):
    """sgn_likelihood.py:79-92, abs_likelihood.py:56-65."""
    kind = spec["kind"]
    mz_hat = az - 1 / tau_z
    assert mz_hat > 0, "az must be greater than 1/ tau_z"
    if kind == "sgn":
        sz_eff = np.sqrt(mz_hat + (mz_hat**2) * tau_z)
        mu_pos = integ.gaussian_measure(
            0, sz_eff, lambda bz: positive_p(az, +bz) * lik_scalar(spec, what, az, bz, +1))
        mu_neg = integ.gaussian_measure(
            0, sz_eff, lambda bz: positive_p(az, -bz) * lik_scalar(spec, what, az, bz, -1))
        return mu_pos + mu_neg
    if kind == "abs":
        def integrand(z, xi_b):
            bz = mz_hat * z + np.sqrt(mz_hat) * xi_b
            return lik_scalar(spec, what, az, bz, np.abs(z))
        return integ.gaussian_measure_2d(0, np.sqrt(tau_z), 0, 1, integrand,
                                         centre2=lambda z: -np.sqrt(mz_hat) * z)
    raise ValueError(kind)


def lik_backward_error(spec, az, tau_z, integ=QUAD):
    """base_likelihood.py:78-81; gaussian_likelihood.py:57-60."""
    if spec["kind"] == "gaussian":
        return 1 / (az + 1 / spec.get("var", 1))
    return lik_beliefs_measure(spec, az, tau_z, "v", integ)


def lik_backward_se(spec, az, tau_z, integ=QUAD):
    """base_likelihood.py:73-76; gaussian_likelihood.py:66-68 (constant)."""
    if spec["kind"] == "gaussian":
        return 1 / spec.get("var", 1)
    return a_new(lik_backward_error(spec, az, tau_z, integ), az,
                 spec.get("AMIN", O.AMIN), spec.get("AMAX", O.AMAX))


def lik_free_energy(spec, az, tau_z, integ=QUAD):
    """base_likelihood.py:88-92; gaussian_likelihood.py:129-132."""
    if spec["kind"] == "gaussian":
        var = spec.get("var", 1)
        return 0.5 * az * tau_z - 1 - 0.5 * np.log((az + 1 / var) * var)
    return lik_beliefs_measure(spec, az, tau_z, "A", integ)


# --------------------------------------------------------------------------
# linear channels.  channel = dict(kind="marchenko", alpha=...) or
# dict(kind="spectrum", spectrum=[Nz], Nx=..., rank=...)
# --------------------------------------------------------------------------
def mp_edges(alpha):
    """ensembles/marchenko_pastur_ensemble.py:10-12."""
    return (1 - np.sqrt(alpha))**2, (1 + np.sqrt(alpha))**2


def mp_mean_spectrum(alpha):
    """ensembles/marchenko_pastur_ensemble.py:13, 28-38 (scipy quad of z * bulk density)."""
    z_min, z_max = mp_edges(alpha)

    def integrand(z):
        return z * np.sqrt((z - z_min) * (z_max - z)) / (2 * np.pi * z)
    return max(0, 1 - alpha) * 0 + quad(integrand, z_min, z_max)[0]


def mp_F(alpha, gamma):
    """marchenko_pastur_ensemble.py:40-42."""
    z_min, z_max = mp_edges(alpha)
    return (np.sqrt(gamma * z_max + 1) - np.sqrt(gamma * z_min + 1))**2


def channel_alpha(ch):
    return ch["alpha"] if ch["kind"] == "marchenko" else ch["Nx"] / len(ch["spectrum"])


def channel_second_moment(ch, tau_z):
    """analytical_linear_channel.py:21-23; linear_channel.py:55-56."""
    if ch["kind"] == "marchenko":
        return tau_z * (ch["mean_spectrum"] / ch["alpha"])
    return tau_z * np.sum(ch["spectrum"]) / ch["Nx"]


def channel_n_eff(ch, az, ax):
    """analytical_linear_channel.py:25-36; linear_channel.py:58-67."""
    if ax == 0:
        return 0.
    if ch["kind"] == "marchenko":
        if az / ax == 0:
            return min(1, ch["alpha"])
        gamma = ax / az
        eta = 1 - mp_F(ch["alpha"], gamma) / (4 * gamma)
        return 1 - eta
    Nz = len(ch["spectrum"])
    if az / ax == 0:
        return ch["rank"] / Nz
    singular = ch["spectrum"][:ch["rank"]]
    return np.sum(singular / (az / ax + singular)) / Nz


def channel_backward_error(ch, az, ax):
    """analytical_linear_channel.py:38-44; linear_channel.py:91-97."""
    az = np.maximum(1e-11, az)
    return (1 - channel_n_eff(ch, az, ax)) / az


def channel_forward_error(ch, az, ax):
    """analytical_linear_channel.py:46-51; linear_channel.py:99-105."""
    alpha = channel_alpha(ch)
    if ax == 0:
        if ch["kind"] == "marchenko":
            return ch["mean_spectrum"] / (alpha * az)
        singular = ch["spectrum"][:ch["rank"]]
        return np.mean(singular) * ch["rank"] / (ch["Nx"] * az)
    return channel_n_eff(ch, az, ax) / (alpha * ax)


def channel_mutual_information(ch, az, ax, tau_z):
    """analytical_linear_channel.py:53-57 (Shannon transform,
    marchenko_pastur_ensemble.py:48-54); linear_channel.py:134-137."""
    if ch["kind"] == "marchenko":
        alpha, gamma = ch["alpha"], ax / az
        F = mp_F(alpha, gamma)
        S = np.log(1 + alpha * gamma - F / 4) + alpha * np.log(1 + gamma - F / 4) - F / (4 * gamma)
        return 0.5 * np.log(az * tau_z) + 0.5 * S
    return np.mean(0.5 * np.log((az + ax * ch["spectrum"]) * tau_z))


def channel_free_energy(ch, az, ax, tau_z):
    """analytical_linear_channel.py:59-63; linear_channel.py:139-143."""
    tau_x = channel_second_moment(ch, tau_z)
    I = channel_mutual_information(ch, az, ax, tau_z)
    return 0.5 * (az * tau_z + channel_alpha(ch) * ax * tau_x) - I + 0.5 * np.log(2 * np.pi * tau_z / np.e)


def variable_free_energy(ax, tau_x):
    """base.py:126-133."""
    I = 0.5 * np.log(ax * tau_x)
    return 0.5 * ax * tau_x - I + 0.5 * np.log(2 * np.pi * tau_x / np.e)


# --------------------------------------------------------------------------
# the recursion
# --------------------------------------------------------------------------
def _damp(d, old, new):
    """message_passing.py:119-127 (`if not damping: return data`)."""
    return d * old + (1 - d) * new if d else new


def se_glm(prior, channel, lik, max_iter, damping=None, a_init=None, early=None,
           integ=QUAD):
    """State Evolution of prior -> x -> channel -> z -> likelihood.

    algos/state_evolution.py:5-27 on the schedule of message_passing.py:249-269,
    330-357; SISO variables pass `a` through (sub_variables.py:33-43).

    damping: dict(e1=, e3=, e5=, e7=) for the factor->variable edges.
    a_init:  dict(edge name -> initial a), default 0 (ConstantInit).
    early:   None or dict(tol, min_variance, wait_increase, max_increase, ids)
             = EarlyStopping (callbacks.py:195-243), ids a subset of ("x", "z").
    Returns dict(vx=[n_iter], vz=[n_iter] trajectories, a=[8] final, n_iter,
    v=(vx, vz) final, tau=(tau_x, tau_z)).
    """
    if channel["kind"] == "marchenko" and "mean_spectrum" not in channel:
        channel = dict(channel, mean_spectrum=mp_mean_spectrum(channel["alpha"]))
    d = dict(e1=0., e3=0., e5=0., e7=0.)
    d.update(damping or {})
    a = {f"e{k}": 0. for k in range(1, 9)}
    a.update(a_init or {})
    tau_x = prior_second_moment(prior)
    tau_z = channel_second_moment(channel, tau_x)
    lin_amin, lin_amax = channel.get("AMIN", O.AMIN), channel.get("AMAX", O.AMAX)
    vx_t, vz_t = [], []
    vx = vz = None
    old_vs, old_state, n_iter = None, None, 0
    for i in range(max_iter):
        # forward pass
        a["e1"] = _damp(d["e1"], a["e1"], prior_forward_se(prior, a["e8"], integ))
        a["e2"] = a["e1"]
        v = channel_forward_error(channel, a["e2"], a["e6"])
        a["e3"] = _damp(d["e3"], a["e3"], a_new(v, a["e6"], lin_amin, lin_amax))
        a["e4"] = a["e3"]
        # backward pass
        a["e5"] = _damp(d["e5"], a["e5"], lik_backward_se(lik, a["e4"], tau_z, integ))
        a["e6"] = a["e5"]
        v = channel_backward_error(channel, a["e2"], a["e6"])
        a["e7"] = _damp(d["e7"], a["e7"], a_new(v, a["e2"], lin_amin, lin_amax))
        a["e8"] = a["e7"]
        # update_variables: base.py:167-170
        vx = 1. / (a["e1"] + a["e7"])
        vz = 1. / (a["e3"] + a["e5"])
        vx_t.append(vx)
        vz_t.append(vz)
        n_iter += 1
        if early is not None:
            ids = early.get("ids", ("x", "z"))
            new_vs = [v_ for k, v_ in (("x", vx), ("z", vz)) if k in ids]
            if any(v_ < early.get("min_variance", -1) for v_ in new_vs):
                break
            if any(np.isnan(v_) for v_ in new_vs):
                a, vx, vz = old_state
                break
            if old_vs:
                tols = [abs(o - n) for o, n in zip(old_vs, new_vs)]
                if max(tols) < early.get("tol", 1e-6):
                    break
                increase = [n - o for o, n in zip(old_vs, new_vs)]
                if i > early.get("wait_increase", 5) and max(increase) > early.get("max_increase", 0.2):
                    a, vx, vz = old_state
                    break
            old_vs = new_vs
            old_state = (dict(a), vx, vz)
    return dict(vx=np.array(vx_t), vz=np.array(vz_t),
                a=np.array([a[f"e{k}"] for k in range(1, 9)]), n_iter=n_iter,
                v=(vx, vz), tau=(tau_x, tau_z), channel=channel)


def se_entropy(prior, channel, lik, a, integ=QUAD):
    """-A_model with node free energies (message_passing.py:306-328,
    state_evolution.py:22-28).  a: [8] edge precisions e1..e8."""
    if channel["kind"] == "marchenko" and "mean_spectrum" not in channel:
        channel = dict(channel, mean_spectrum=mp_mean_spectrum(channel["alpha"]))
    e = {f"e{k}": a[k - 1] for k in range(1, 9)}
    tau_x = prior_second_moment(prior)
    tau_z = channel_second_moment(channel, tau_x)
    A_nodes = (prior_free_energy(prior, e["e8"], integ)
               + variable_free_energy(e["e1"] + e["e7"], tau_x)
               + channel_free_energy(channel, e["e2"], e["e6"], tau_x)
               + variable_free_energy(e["e3"] + e["e5"], tau_z)
               + lik_free_energy(lik, e["e4"], tau_z, integ))
    A_edges = (variable_free_energy(e["e1"] + e["e8"], tau_x)
               + variable_free_energy(e["e2"] + e["e7"], tau_x)
               + variable_free_energy(e["e3"] + e["e6"], tau_z)
               + variable_free_energy(e["e4"] + e["e5"], tau_z))
    return -(A_nodes - A_edges)
