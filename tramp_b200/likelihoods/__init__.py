"""Likelihoods on the EP hot path (reference tramp/likelihoods/).

`y` of shape (M,) is one instance, (B, M) a batch of B instances.
"""
import numpy as np

from ..base import Factor, Registry, _Arg, measure_out, se_domain_error
from .. import ops, _lib


class Likelihood(Factor):
    """reference likelihoods/base_likelihood.py:6-28."""
    n_next = 0
    n_prev = 1

    def get_size(self, y):
        if y is None:
            return None
        if len(y.shape) == 1:
            return y.shape[0]
        return y.shape

    def _trb_factor(self):
        raise NotImplementedError

    def infer_shape(self, z_shape):
        return None if self.batch is None else []

    @property
    def batch(self):
        y = self.y
        return None if (y is None or len(np.shape(y)) < 2) else int(np.shape(y)[0])

    def compute_backward_posterior(self, az, bz, y):
        arg = _Arg(az, bz, y)
        r, v = ops.factor_posterior(self._trb_factor(), arg.a, arg.b, arg.y, arg.n,
                                    arg.a_elementwise, not self.isotropic)
        return arg.vec_out(r), (arg.scalar_out(v) if self.isotropic else arg.vec_out(v))

    def compute_backward_message(self, az, bz):
        """reference likelihoods/base_likelihood.py:25-28."""
        rz, vz = self.compute_backward_posterior(az, bz, self.y)
        return self.compute_ab_new(rz, vz, az, bz)

    def compute_log_partition(self, az, bz, y):
        arg = _Arg(az, bz, y)
        A = ops.factor_log_partition(self._trb_factor(), arg.a, arg.b, arg.y, arg.n,
                                     arg.a_elementwise, False)
        return arg.scalar_out(A)

    def _scalar(self, az, bz, y, what):
        scalar = np.ndim(az) == 0 and np.ndim(bz) == 0 and np.ndim(y) == 0
        az_, bz_, y_ = np.broadcast_arrays(np.atleast_1d(np.asarray(az, float)),
                                           np.atleast_1d(np.asarray(bz, float)),
                                           np.atleast_1d(np.asarray(y, float)))
        arg = _Arg(np.ascontiguousarray(az_), np.ascontiguousarray(bz_), np.ascontiguousarray(y_))
        f = self._trb_factor()
        if what == "A":
            out = ops.factor_log_partition(f, arg.a, arg.b, arg.y, arg.n, True, True)
        else:
            r, v = ops.factor_posterior(f, arg.a, arg.b, arg.y, arg.n, True, True)
            out = r if what == "r" else v
        out = arg.vec_out(out)
        return float(out[0]) if scalar else out

    def scalar_backward_mean(self, az, bz, y):
        return self._scalar(az, bz, y, "r")

    def scalar_backward_variance(self, az, bz, y):
        return self._scalar(az, bz, y, "v")

    def scalar_log_partition(self, az, bz, y):
        return self._scalar(az, bz, y, "A")

    # ---- State Evolution (reference likelihoods/base_likelihood.py:73-98) ------
    def beliefs_measure(self, az, tau_z, f):
        """Average of f over the joint law of (b_z, y) at precision az and second
        moment tau_z (sgn_likelihood.py:79-92, abs_likelihood.py:56-65); f = "v"
        (scalar_backward_variance) or "A" (compute_log_partition), evaluated by
        the device quadrature.  Raises AssertionError when az <= 1/tau_z, like
        the reference's `assert mz_hat > 0`."""
        what = {"v": _lib.MEASURE_V, "A": _lib.MEASURE_A}.get(f)
        if what is None:
            raise NotImplementedError('beliefs_measure runs on the device for f = "v" or "A" only')
        az_ = np.atleast_1d(np.asarray(az, float))
        tau = np.array(np.broadcast_to(np.asarray(tau_z, float), az_.shape))
        out, flags = ops.se_measure(self._trb_factor(), what, az_, tau)
        se_domain_error(flags)
        return measure_out(out, az)

    def compute_backward_error(self, az, tau_z):
        return self.beliefs_measure(az, tau_z, "v")

    def compute_backward_state_evolution(self, az, tau_z):
        vz = self.compute_backward_error(az, tau_z)
        return self.compute_a_new(vz, az)

    def compute_backward_overlap(self, az, tau_z):
        return tau_z - self.compute_backward_error(az, tau_z)

    def compute_free_energy(self, az, tau_z):
        return self.beliefs_measure(az, tau_z, "A")

    def compute_mutual_information(self, az, tau_z):
        "Note: returns H = mutual information I + noise entropy N (reference :94-98)"
        A = self.compute_free_energy(az, tau_z)
        return 0.5 * az * tau_z - A + 0.5 * np.log(2 * np.pi * tau_z / np.e)


class GaussianLikelihood(Likelihood):
    """reference likelihoods/gaussian_likelihood.py:7-71."""

    def __init__(self, y, var=1, y_name="y", isotropic=True):
        self.y_name = y_name
        self.size = self.get_size(y)
        self.var = var
        self.isotropic = isotropic
        self.repr_init()
        self.y = y
        self.sigma = np.sqrt(var)
        self.a = 1 / var
        self.b = None if y is None else y / var

    def _trb_factor(self):
        return ops.gaussian_likelihood_factor(self.var, self.AMIN, self.AMAX)

    def sample(self, X):
        return X + self.sigma * np.random.standard_normal(X.shape)

    def math(self):
        return r"$\mathcal{N}$"

    def compute_backward_posterior(self, az, bz, y):
        """reference gaussian_likelihood.py:43-49: vz = 1/a keeps az's shape (no mean)."""
        arg = _Arg(az, bz, y)
        r, v = ops.factor_posterior(self._trb_factor(), arg.a, arg.b, arg.y, arg.n,
                                    arg.a_elementwise, arg.a_elementwise)
        return arg.vec_out(r), (arg.vec_out(v) if arg.a_elementwise else arg.scalar_out(v))

    def compute_backward_message(self, az, bz):
        """Constant, unclipped message (reference gaussian_likelihood.py:68-71)."""
        return self.a, self.b

    def compute_backward_state_evolution(self, az, tau_z):
        """Constant (reference gaussian_likelihood.py:66-68)."""
        return self.a


class SgnLikelihood(Likelihood):
    """reference likelihoods/sgn_likelihood.py:9-41."""

    def __init__(self, y, y_name="y", isotropic=True):
        self.y_name = y_name
        self.size = self.get_size(y)
        self.isotropic = isotropic
        self.repr_init()
        self.y = y

    def _trb_factor(self):
        return ops.sgn_factor(self.AMIN, self.AMAX)

    def sample(self, X):
        return np.sign(X)

    def math(self):
        return r"$\mathrm{sgn}$"


class AbsLikelihood(Likelihood):
    """reference likelihoods/abs_likelihood.py:8-40."""

    def __init__(self, y, y_name="y", isotropic=True):
        self.y_name = y_name
        self.size = self.get_size(y)
        self.isotropic = isotropic
        self.repr_init()
        self.y = y

    def _trb_factor(self):
        return ops.abs_factor(self.AMIN, self.AMAX)

    def sample(self, X):
        return np.abs(X)

    def math(self):
        return r"$\mathrm{abs}$"


LIKELIHOOD_CLASSES = Registry("likelihood", {
    "gaussian": GaussianLikelihood,
    "abs": AbsLikelihood,
    "sgn": SgnLikelihood,
})


def get_likelihood(y, likelihood_type, **kwargs):
    """reference likelihoods/__init__.py:25-27."""
    return LIKELIHOOD_CLASSES[likelihood_type](y=y, **kwargs)
