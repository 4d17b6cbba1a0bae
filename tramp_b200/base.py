"""Base classes of the model/plugin API (mirrors reference tramp/base.py).

`Variable` and `Factor` keep the reference's names, arities and message
conventions (`message = [(source, target, data)]`, data = dict(a, b, direction)),
so that user-written factors and the reference's own tests read the same.  The
arithmetic behind the in-scope factors runs on the GPU through tramp_b200.ops.
"""
import logging
import numpy as np

from . import ops

logger = logging.getLogger(__name__)


class ReprMixin():
    """reference base.py:10-32."""
    _repr_initialized = False

    def repr_init(self, pad=None, reinit=False):
        if reinit or not self._repr_initialized:
            self._repr_kwargs = self.__dict__.copy()
            self._repr_pad = pad
            self._repr_initialized = True

    def __repr__(self):
        pad = f"\n{self._repr_pad}" if self._repr_pad else ""
        args = ",".join(f"{pad}{key}={val}" for key, val in self._repr_kwargs.items())
        if self._repr_pad:
            args += "\n"
        return f"{self.__class__.__name__}({args})"



class Registry(dict):
    """name -> class table behind get_prior / get_likelihood / get_channel /
    get_ensemble.  An unknown name is a KeyError as in the reference (a plain dict
    lookup there); the message says what this build covers, because the reference
    knows many more names (DESIGN 1 "Out of scope")."""

    def __init__(self, what, classes):
        super().__init__(classes)
        self.what = what

    def __missing__(self, name):
        raise KeyError(f"{self.what} type {name!r} is not part of tramp_b200 "
                       f"(the GLM expectation-propagation path: {', '.join(sorted(self))})")


def filter_message(message, direction):
    """reference base.py:35-41."""
    return [(s, t, d) for s, t, d in message if d["direction"] == direction]


def inv(v):
    """Numerically safe inverse (reference base.py:44-46)."""
    if ops.is_tensor(v):
        return 1 / v.clamp_min(1e-20)
    return 1 / np.maximum(v, 1e-20)


def _clip(x, lo, hi):
    if ops.is_tensor(x):
        return x.clamp(lo, hi)
    return np.clip(x, lo, hi)


class Variable(ReprMixin):
    """reference base.py:49-233 (EP part)."""

    def __init__(self, id, n_prev, n_next):
        self.id = id
        self.n_prev = n_prev
        self.n_next = n_next
        self.repr_init()

    def __add__(self, other):
        from .models.dag_algebra import DAG
        return DAG(self) + other

    def __matmul__(self, other):
        from .models.dag_algebra import DAG
        return DAG(self) @ other

    def math(self):
        return r"$" + self.id + r"$"

    def posterior_ab(self, message):
        """reference base.py:152-155."""
        a_hat = sum(data["a"] for source, target, data in message)
        b_hat = sum(data["b"] for source, target, data in message)
        return a_hat, b_hat

    def posterior_rv(self, message):
        """reference base.py:157-161 (a batch: one precision per instance, a_hat [B], b_hat [B, n])."""
        a_hat, b_hat = self.posterior_ab(message)
        a_col = np.asarray(a_hat)[..., None] if np.ndim(a_hat) and np.ndim(b_hat) > np.ndim(a_hat) else a_hat
        return b_hat / a_col, 1. / a_hat

    def compute_log_partition(self, ax, bx):
        """reference base.py:146-150 (a SUM over components; inf if ax <= 0), per instance of a batch."""
        if np.ndim(ax) == 0:
            if ax <= 0:
                return np.inf
            return 0.5 * np.sum(bx**2 / ax + np.log(2 * np.pi / ax))
        ax = np.asarray(ax, dtype=float)
        with np.errstate(all="ignore"):
            logZ = 0.5 * np.sum(bx**2 / ax[:, None] + np.log(2 * np.pi / ax[:, None]), axis=-1)
        return np.where(ax <= 0, np.inf, logZ)

    def log_partition(self, message):
        ax, bx = self.posterior_ab(message)
        return self.compute_log_partition(ax, bx)

    # ---- State Evolution: precisions only (reference base.py:109-144, 163-178, 209-233)
    def posterior_a(self, message):
        return sum(data["a"] for source, target, data in message)

    def posterior_v(self, message):
        return 1. / self.posterior_a(message)

    def _parse_tau(self, message):
        return message[0][2]["tau"]

    def compute_mutual_information(self, ax, tau_x):
        return 0.5 * np.log(ax * tau_x)

    def compute_free_energy(self, ax, tau_x):
        I = self.compute_mutual_information(ax, tau_x)
        return 0.5 * ax * tau_x - I + 0.5 * np.log(2 * np.pi * tau_x / np.e)

    def compute_dual_mutual_information(self, vx, tau_x):
        return 0.5 * np.log(tau_x / vx) - 0.5

    def compute_dual_free_energy(self, mx, tau_x):
        return 0.5 * np.log(2 * np.pi * (tau_x - mx))

    def free_energy(self, message):
        return self.compute_free_energy(self.posterior_a(message), self._parse_tau(message))

    def _state_evolution(self, message, incoming, outgoing):
        a_hat = self.posterior_a(message)
        return [(target, source, dict(a=a_hat - data["a"], direction=outgoing))
                for source, target, data in filter_message(message, incoming)]

    def forward_state_evolution(self, message):
        """to every next factor: the total precision minus what it sent (reference base.py:209-220)."""
        return [] if self.n_next == 0 else self._state_evolution(message, "bwd", "fwd")

    def backward_state_evolution(self, message):
        """reference base.py:222-233."""
        return [] if self.n_prev == 0 else self._state_evolution(message, "fwd", "bwd")


class Factor(ReprMixin):
    """reference base.py:236-365 (EP part)."""

    AMAX = ops.AMAX
    AMIN = ops.AMIN

    def reset_precision_bounds(self, AMIN, AMAX):
        """reference base.py:241-243."""
        self.AMIN = AMIN
        self.AMAX = AMAX

    def compute_a_new(self, v, a):
        return _clip(inv(v) - a, self.AMIN, self.AMAX)

    def compute_ab_new(self, r, v, a, b):
        """a_new clipped to [AMIN, AMAX]; b_new = r (a + a_new) - b (reference base.py:250-255)."""
        a_new = _clip(inv(v) - a, self.AMIN, self.AMAX)
        v_inv = (a + a_new)
        if ops.is_tensor(r) and r.dim() == 2 and ops.is_tensor(v_inv) and v_inv.dim() == 1:
            v_inv = v_inv[:, None]
        elif isinstance(r, np.ndarray) and r.ndim == 2 and np.ndim(v_inv) == 1:
            v_inv = np.asarray(v_inv)[:, None]
        b_new = r * v_inv - b
        return a_new, b_new

    def __add__(self, other):
        from .models.dag_algebra import DAG
        return DAG(self) + other

    def __matmul__(self, other):
        from .models.dag_algebra import DAG
        return DAG(self) @ other

    def _parse_message_ab(self, message):
        """reference base.py:285-306."""
        z_message = filter_message(message, "fwd")
        assert len(z_message) == self.n_prev
        az = [data["a"] for source, target, data in z_message]
        bz = [data["b"] for source, target, data in z_message]
        z_source = [source for source, target, data in z_message]
        if self.n_prev == 1:
            az, bz, z_source = az[0], bz[0], z_source[0]
        x_message = filter_message(message, "bwd")
        assert len(x_message) == self.n_next
        ax = [data["a"] for source, target, data in x_message]
        bx = [data["b"] for source, target, data in x_message]
        x_source = [source for source, target, data in x_message]
        if self.n_next == 1:
            ax, bx, x_source = ax[0], bx[0], x_source[0]
        return z_source, x_source, az, bz, ax, bx

    def forward_message(self, message):
        """reference base.py:329-346 (single next variable)."""
        if self.n_next == 0:
            return []
        z_source, x_source, az, bz, ax, bx = self._parse_message_ab(message)
        if self.n_prev == 0:
            ax_new, bx_new = self.compute_forward_message(ax, bx)
        else:
            ax_new, bx_new = self.compute_forward_message(az, bz, ax, bx)
        if self.n_next != 1:
            raise NotImplementedError("multi-edge factors are outside the EP hot path")
        return [(self, x_source, dict(a=ax_new, b=bx_new, direction="fwd"))]

    def backward_message(self, message):
        """reference base.py:348-365 (single previous variable)."""
        if self.n_prev == 0:
            return []
        z_source, x_source, az, bz, ax, bx = self._parse_message_ab(message)
        if self.n_next == 0:
            az_new, bz_new = self.compute_backward_message(az, bz)
        else:
            az_new, bz_new = self.compute_backward_message(az, bz, ax, bx)
        if self.n_prev != 1:
            raise NotImplementedError("multi-edge factors are outside the EP hot path")
        return [(self, z_source, dict(a=az_new, b=bz_new, direction="bwd"))]

    def log_partition(self, message):
        """reference base.py:367-375."""
        z_source, x_source, az, bz, ax, bx = self._parse_message_ab(message)
        if self.n_prev == 0:
            return self.compute_log_partition(ax, bx)
        if self.n_next == 0:
            return self.compute_log_partition(az, bz, self.y)
        return self.compute_log_partition(az, bz, ax, bx)


    # ---- State Evolution (reference base.py:308-327, 377-419) -----------------
    def _parse_message_a(self, message):
        z_message = filter_message(message, "fwd")
        assert len(z_message) == self.n_prev
        az = [data["a"] for source, target, data in z_message]
        tau_z = [data["tau"] for source, target, data in z_message]
        z_source = [source for source, target, data in z_message]
        if self.n_prev == 1:
            az, tau_z, z_source = az[0], tau_z[0], z_source[0]
        x_message = filter_message(message, "bwd")
        assert len(x_message) == self.n_next
        ax = [data["a"] for source, target, data in x_message]
        x_source = [source for source, target, data in x_message]
        if self.n_next == 1:
            ax, x_source = ax[0], x_source[0]
        return z_source, x_source, az, ax, tau_z

    def forward_state_evolution(self, message):
        if self.n_next == 0:
            return []
        z_source, x_source, az, ax, tau_z = self._parse_message_a(message)
        if self.n_prev == 0:
            ax_new = self.compute_forward_state_evolution(ax)
        else:
            ax_new = self.compute_forward_state_evolution(az, ax, tau_z)
        if self.n_next != 1:
            raise NotImplementedError("multi-edge factors are outside the hot path")
        return [(self, x_source, dict(a=ax_new, direction="fwd"))]

    def backward_state_evolution(self, message):
        if self.n_prev == 0:
            return []
        z_source, x_source, az, ax, tau_z = self._parse_message_a(message)
        if self.n_next == 0:
            az_new = self.compute_backward_state_evolution(az, tau_z)
        else:
            az_new = self.compute_backward_state_evolution(az, ax, tau_z)
        if self.n_prev != 1:
            raise NotImplementedError("multi-edge factors are outside the hot path")
        return [(self, z_source, dict(a=az_new, direction="bwd"))]

    def free_energy(self, message):
        z_source, x_source, az, ax, tau_z = self._parse_message_a(message)
        if self.n_prev == 0:
            return self.compute_free_energy(ax)
        if self.n_next == 0:
            return self.compute_free_energy(az, tau_z)
        return self.compute_free_energy(az, ax, tau_z)


def se_domain_error(flags):
    """The reference asserts `mz_hat > 0` inside Likelihood.beliefs_measure
    (sgn_likelihood.py:80-81, abs_likelihood.py:57-58); the kernels flag it."""
    from . import _lib
    if int(flags.max().item()) & _lib.FLAG_SE_DOMAIN:
        raise AssertionError("az must be greater than 1/ tau_z")


def measure_out(out, like):
    """Device [B] result -> float for scalar input, numpy array otherwise."""
    x = out.cpu().numpy()
    return float(x[0]) if np.ndim(like) == 0 else x.reshape(np.shape(like))


# ---------------------------------------------------------------------------
# array plumbing shared by the separable factors
# ---------------------------------------------------------------------------
class _Arg:
    """Normalises (a, b[, y]) arguments of the factor API to device tensors.

    b: (n,) or (B, n) numpy array / tensor.  a: scalar, (B,), or same shape as b
    (isotropic=False).  Results are returned in the caller's array type and
    shape: numpy in -> numpy out, device tensor in -> device tensor out."""

    def __init__(self, a, b, y=None):
        self.numpy_out = not ops.is_tensor(b)
        b_ = ops.to_dev(b)
        self.batched = (b_.dim() == 2)
        if b_.dim() > 2:
            raise ValueError("b must be 1-d (one instance) or 2-d (batch, n)")
        self.shape = tuple(b_.shape)
        b2 = b_ if self.batched else b_[None, :]
        self.B, self.n = b2.shape
        self.b = ops.padded(b2)
        self.ld = self.b.shape[1]
        a_ = ops.to_dev(a)
        if a_.numel() == 1 and not (a_.dim() >= 1 and self.n == 1 and not self.batched):
            self.a_elementwise = False                      # one precision for everything
            self.a = a_.reshape(1).expand(self.B).contiguous()
        elif self.batched and tuple(a_.shape) == (self.B,):
            self.a_elementwise = False                      # one precision per instance
            self.a = a_
        elif tuple(a_.shape) == self.shape:
            self.a_elementwise = True                       # isotropic=False: one per component
            self.a = ops.padded(a_ if self.batched else a_[None, :], self.ld)
        else:
            raise ValueError(f"a of shape {tuple(a_.shape)} does not match b of shape {self.shape}")
        self.y = None
        if y is not None:
            y_ = ops.to_dev(y)
            y2 = y_ if y_.dim() == 2 else y_[None, :]
            if y2.shape[0] == 1 and self.B > 1:
                y2 = y2.expand(self.B, -1)
            self.y = ops.padded(y2.contiguous(), self.ld)

    def vec_out(self, t):
        t = t[:, :self.n]
        if not self.batched:
            t = t[0]
        return t.cpu().numpy() if self.numpy_out else t.contiguous()

    def scalar_out(self, t):
        if self.numpy_out:
            x = t.cpu().numpy()
            return x if self.batched else float(x[0])
        return t if self.batched else t[0]
