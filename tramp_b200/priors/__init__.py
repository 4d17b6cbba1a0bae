"""Priors on the EP hot path (reference tramp/priors/).

Same constructors, attributes and EP methods as the reference classes; the
moments run on the GPU (tramp_b200/csrc/trb_moments.cuh).  `batch=B` is the one
extension: B independent instances share the prior's parameters, messages
become `a: (B,)`, `b: (B, N)`.
"""
import numpy as np

from ..base import Factor, Registry, _Arg, measure_out
from .. import ops, _lib


class Prior(Factor):
    """reference priors/base_prior.py:6-16."""
    n_next = 1
    n_prev = 0

    def _trb_factor(self):
        raise NotImplementedError

    def _sample_shape(self):
        size = self.size if isinstance(self.size, tuple) else (self.size,)
        return size if self.batch is None else (self.batch,) + size

    def infer_shape(self):
        """Batched priors (an extension with no reference RNG stream to preserve)
        report their shape directly; un-batched ones return None so that
        Model.init_shapes calls sample() exactly as the reference does
        (base_model.py:96-109 advances the global RNG)."""
        return None if self.batch is None else [self._sample_shape()]

    def compute_forward_posterior(self, ax, bx):
        """(rx, vx); vx is the mean over components when isotropic."""
        arg = _Arg(ax, bx)
        r, v = ops.factor_posterior(self._trb_factor(), arg.a, arg.b, None, arg.n,
                                    arg.a_elementwise, not self.isotropic)
        return arg.vec_out(r), (arg.scalar_out(v) if self.isotropic else arg.vec_out(v))

    def compute_forward_message(self, ax, bx):
        """reference priors/base_prior.py:13-16."""
        rx, vx = self.compute_forward_posterior(ax, bx)
        return self.compute_ab_new(rx, vx, ax, bx)

    def compute_log_partition(self, ax, bx):
        """Mean over components of the scalar log-partition."""
        arg = _Arg(ax, bx)
        A = ops.factor_log_partition(self._trb_factor(), arg.a, arg.b, None, arg.n,
                                     arg.a_elementwise, False)
        return arg.scalar_out(A)

    # scalar_* : elementwise versions used by the reference's tests
    def _scalar(self, ax, bx, what):
        ax_, bx_ = np.atleast_1d(np.asarray(ax, float)), np.atleast_1d(np.asarray(bx, float))
        ax_, bx_ = np.broadcast_arrays(ax_, bx_)
        arg = _Arg(np.ascontiguousarray(ax_), np.ascontiguousarray(bx_))
        f = self._trb_factor()
        if what == "A":
            out = ops.factor_log_partition(f, arg.a, arg.b, None, arg.n, True, True)
        else:
            r, v = ops.factor_posterior(f, arg.a, arg.b, None, arg.n, True, True)
            out = r if what == "r" else v
        out = arg.vec_out(out)
        return float(out[0]) if np.ndim(ax) == 0 and np.ndim(bx) == 0 else out

    def scalar_forward_mean(self, ax, bx):
        return self._scalar(ax, bx, "r")

    def scalar_forward_variance(self, ax, bx):
        return self._scalar(ax, bx, "v")

    def scalar_log_partition(self, ax, bx):
        return self._scalar(ax, bx, "A")

    # ---- State Evolution (reference priors/base_prior.py:66-90) ---------------
    def beliefs_measure(self, ax, f):
        """Average of f over the law of the incoming belief b_x at precision ax
        (gauss_bernoulli_prior.py:112-118, binary_prior.py:80-84).  The reference
        takes any Python callable and integrates it with scipy quad; on the device
        f is one of the two functions SE needs: "v" (scalar_forward_variance) or
        "A" (scalar_log_partition).  ax: scalar or array (one problem per entry)."""
        what = {"v": _lib.MEASURE_V, "A": _lib.MEASURE_A}.get(f)
        if what is None:
            raise NotImplementedError('beliefs_measure runs on the device for f = "v" or "A" only')
        ax_ = np.atleast_1d(np.asarray(ax, float))
        tau = np.full(ax_.shape, float(self.second_moment()))
        out, _ = ops.se_measure(self._trb_factor(), what, ax_, tau)
        return measure_out(out, ax)

    def compute_forward_error(self, ax):
        return self.beliefs_measure(ax, "v")

    def compute_forward_state_evolution(self, ax):
        vx = self.compute_forward_error(ax)
        return self.compute_a_new(vx, ax)

    def compute_forward_overlap(self, ax):
        return self.second_moment() - self.compute_forward_error(ax)

    def compute_free_energy(self, ax):
        return self.beliefs_measure(ax, "A")

    def compute_mutual_information(self, ax):
        return 0.5 * ax * self.second_moment() - self.compute_free_energy(ax)


class GaussBernoulliPrior(Prior):
    r"""Gauss-Bernoulli prior $p(x)=[1-\rho]\delta(x)+\rho\mathcal{N}(x|r,v)$
    (reference priors/gauss_bernoulli_prior.py:8-83)."""

    def __init__(self, size, rho=0.5, mean=0, var=1, isotropic=True, batch=None):
        self.size = size
        self.rho = rho
        self.mean = mean
        self.var = var
        self.isotropic = isotropic
        self.repr_init()
        self.batch = batch
        self.sigma = np.sqrt(var)
        self.a = 1 / var
        self.b = mean / var
        self.eta = 0.5 * (self.b**2 / self.a + np.log(2 * np.pi / self.a)) - np.log(rho / (1 - rho))

    def _trb_factor(self):
        return ops.gauss_bernoulli_factor(self.rho, self.mean, self.var, self.AMIN, self.AMAX)

    def sample(self):
        """reference gauss_bernoulli_prior.py:38-42 (numpy global RNG, same draw order)."""
        shape = self._sample_shape()
        X_gauss = self.mean + self.sigma * np.random.standard_normal(shape)
        X_bernoulli = np.random.binomial(n=1, size=shape, p=self.rho)
        return X_gauss * X_bernoulli

    def math(self):
        return r"$\mathcal{N}_\rho$"

    def second_moment(self):
        return self.rho * (self.mean**2 + self.var)


class BinaryPrior(Prior):
    r"""Binary prior $p(x) = p_+ \delta_+(x) + p_- \delta_-(x)$
    (reference priors/binary_prior.py:8-68)."""

    def __init__(self, size, p_pos=0.5, isotropic=True, batch=None):
        self.size = size
        self.p_pos = p_pos
        self.isotropic = isotropic
        self.repr_init()
        self.batch = batch
        self.p_neg = 1 - p_pos
        self.b = 0.5 * np.log(self.p_pos / self.p_neg)

    def _trb_factor(self):
        return ops.binary_factor(self.p_pos, self.AMIN, self.AMAX)

    def sample(self):
        """reference binary_prior.py:30-33."""
        p = [self.p_neg, self.p_pos]
        return np.random.choice([-1, +1], size=self._sample_shape(), replace=True, p=p)

    def math(self):
        return r"$p_\pm$"

    def second_moment(self):
        return 1.


class GaussianPrior(Prior):
    r"""Gaussian prior $p(x)=\mathcal{N}(x|r, v)$ (reference priors/gaussian_prior.py:8-89)."""

    def __init__(self, size, mean=0, var=1, isotropic=True, batch=None):
        self.size = size
        self.mean = mean
        self.var = var
        self.isotropic = isotropic
        self.repr_init()
        self.batch = batch
        self.sigma = np.sqrt(var)
        self.a = 1 / var
        self.b = mean / var

    def _trb_factor(self):
        return ops.gaussian_prior_factor(self.mean, self.var, self.AMIN, self.AMAX)

    def sample(self):
        return self.mean + self.sigma * np.random.standard_normal(self._sample_shape())

    def math(self):
        return r"$\mathcal{N}$"

    def second_moment(self):
        return self.mean**2 + self.var

    def compute_forward_posterior(self, ax, bx):
        """reference gaussian_prior.py:63-68: vx = 1/a is NOT averaged (it already has ax's shape)."""
        arg = _Arg(ax, bx)
        r, v = ops.factor_posterior(self._trb_factor(), arg.a, arg.b, None, arg.n,
                                    arg.a_elementwise, arg.a_elementwise)
        return arg.vec_out(r), (arg.vec_out(v) if arg.a_elementwise else arg.scalar_out(v))

    def compute_forward_message(self, ax, bx):
        """Constant, unclipped message (reference gaussian_prior.py:86-89)."""
        if ops.is_tensor(bx):
            t = ops.torch()
            return self.a * t.ones_like(ops.to_dev(ax)), self.b * t.ones_like(bx)
        return self.a * np.ones_like(ax), self.b * np.ones_like(bx)

    def compute_forward_state_evolution(self, ax):
        """Constant (reference gaussian_prior.py:102-104)."""
        return self.a


PRIOR_CLASSES = Registry("prior", {
    "gaussian": GaussianPrior,
    "gauss_bernoulli": GaussBernoulliPrior,
    "binary": BinaryPrior,
})


def get_prior(size, prior_type, **kwargs):
    """reference priors/__init__.py:25-27."""
    return PRIOR_CLASSES[prior_type](size=size, **kwargs)
