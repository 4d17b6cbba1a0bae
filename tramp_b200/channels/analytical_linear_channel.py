"""Linear channels known only through their spectral law (reference
tramp/channels/linear/analytical_linear_channel.py:8-90).

State Evolution needs nothing of W but its eigenvalue distribution; for iid
Gaussian W that is the Marchenko-Pastur law, whose eta and Shannon transforms
are closed forms.  These are scalar formulas: inside a batched SE run they are
evaluated by `k_se_run` (tramp_b200/csrc/trb_se.cu); the methods below are the
reference's factor-level API for the same quantities.
"""
import logging
import numpy as np

from .base_channel import Channel
from ..ensembles import MarchenkoPasturEnsemble

logger = logging.getLogger(__name__)


class AnalyticalLinearChannel(Channel):
    """reference analytical_linear_channel.py:8-65."""

    def __init__(self, ensemble, name="W"):
        self.name = name
        self.alpha = ensemble.alpha
        self.repr_init()
        self.ensemble = ensemble

    def sample(self, Z):
        F = self.ensemble.generate(Z.shape[0])
        return F @ Z

    def math(self):
        return r"$" + self.name + "$"

    def second_moment(self, tau_z):
        return tau_z * (self.ensemble.mean_spectrum / self.alpha)

    def compute_n_eff(self, az, ax):
        "Effective number of parameters"
        if ax == 0:
            logger.info(f"ax=0 in {self} compute_n_eff")
            return 0.
        if az / ax == 0:
            logger.info(f"az/ax=0 in {self} compute_n_eff")
            return min(1, self.alpha)
        return 1 - self.ensemble.eta_transform(ax / az)

    def compute_backward_error(self, az, ax, tau_z):
        if az == 0:
            logger.info(f"az=0 in {self} compute_backward_error")
        az = np.maximum(1e-11, az)
        return (1 - self.compute_n_eff(az, ax)) / az

    def compute_forward_error(self, az, ax, tau_z):
        if ax == 0:
            return self.ensemble.mean_spectrum / (self.alpha * az)
        return self.compute_n_eff(az, ax) / (self.alpha * ax)

    def compute_mutual_information(self, az, ax, tau_z):
        S = self.ensemble.shannon_transform(ax / az)
        return 0.5 * np.log(az * tau_z) + 0.5 * S

    def compute_free_energy(self, az, ax, tau_z):
        tau_x = self.second_moment(tau_z)
        I = self.compute_mutual_information(az, ax, tau_z)
        return 0.5 * (az * tau_z + self.alpha * ax * tau_x) - I + 0.5 * np.log(2 * np.pi * tau_z / np.e)


class MarchenkoPasturChannel(AnalyticalLinearChannel):
    """reference analytical_linear_channel.py:68-90."""

    def __init__(self, alpha, name="W"):
        super().__init__(ensemble=MarchenkoPasturEnsemble(alpha=alpha), name=name)

    def compute_precision(self, vz, vx, tau_z):
        ax = 1 / vx - 1 / vz
        az = (1 - self.alpha * ax * vx) / vz
        return az, ax
