"""Sgn / Abs channels: teacher-side sampling only.  As EP factors they are
always observed, i.e. replaced by SgnLikelihood / AbsLikelihood
(reference models/dag_algebra.py:24-29)."""
import numpy as np
from .base_channel import Channel


class SgnChannel(Channel):
    """reference channels/activation/piecewise_linear_channel.py:80-84; the
    piecewise sampler gives +1 at z == 0 (utils/linear_region.py:27-30)."""

    def __init__(self):
        self.repr_init()
        self.name = "sgn"

    def sample(self, Z):
        return np.where(Z >= 0, 1.0, 0.0) - np.where(Z < 0, 1.0, 0.0)

    def math(self):
        return r"$\textrm{sgn}$"

    def second_moment(self, tau_z):
        return 1.


class AbsChannel(Channel):
    """reference channels/activation/piecewise_linear_channel.py:87-91."""

    def __init__(self):
        self.repr_init()
        self.name = "abs"

    def sample(self, Z):
        return np.abs(Z)

    def math(self):
        return r"$\textrm{abs}$"

    def second_moment(self, tau_z):
        return tau_z
