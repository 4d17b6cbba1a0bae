import numpy as np
from .base_channel import Channel


class GaussianChannel(Channel):
    """Additive-noise channel (reference channels/noise/gaussian_channel.py:5-33).

    In the observed GLM it is replaced by GaussianLikelihood
    (models/dag_algebra.py:21-23), so only `sample`, `var` and the closed-form
    messages are kept."""

    def __init__(self, var=1):
        self.var = var
        self.repr_init()
        self.sigma = np.sqrt(var)
        self.a = 1 / var

    def sample(self, Z):
        noise = self.sigma * np.random.standard_normal(Z.shape)
        return Z + noise

    def math(self):
        return r"$\mathcal{N}$"

    def second_moment(self, tau_z):
        return tau_z + self.var

    def compute_forward_message(self, az, bz, ax, bx):
        kz = self.a / (self.a + az)
        return kz * az, kz * bz

    def compute_backward_message(self, az, bz, ax, bx):
        kx = self.a / (self.a + ax)
        return kx * ax, kx * bx
