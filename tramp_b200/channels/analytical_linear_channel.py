"""Linear channels known only through their spectral law (reference
tramp/channels/linear/analytical_linear_channel.py:8-90).

State Evolution needs nothing of the matrix but the eigenvalue distribution of
F^T F: the "effective number of parameters" is one minus its eta transform, the
mutual information its Shannon transform.  For iid Gaussian entries the law is
Marchenko-Pastur and both transforms are closed forms, so these are a handful
of scalar operations; inside a batched run `k_se_run`
(tramp_b200/csrc/trb_se.cu, `channel_n_eff` and friends) evaluates the same
formulas on the device, and the methods below are the reference's factor-level
API for them.
"""
import logging
import numpy as np

from .base_channel import Channel
from ..ensembles import MarchenkoPasturEnsemble

logger = logging.getLogger(__name__)

AZ_FLOOR = 1e-11   # the backward error never divides by less (reference :41)


class AnalyticalLinearChannel(Channel):
    """x = F z with F drawn from `ensemble` (anything with alpha, mean_spectrum,
    eta_transform, shannon_transform and generate)."""

    def __init__(self, ensemble, name="W"):
        self.name = name
        self.alpha = ensemble.alpha
        self.repr_init()
        self.ensemble = ensemble

    def math(self):
        return r"$" + self.name + "$"

    def sample(self, Z):
        return self.ensemble.generate(Z.shape[0]) @ Z

    def second_moment(self, tau_z):
        return tau_z * (self.ensemble.mean_spectrum / self.alpha)

    # -- the three quantities of the State-Evolution update -------------------
    def compute_n_eff(self, az, ax):
        "Effective number of parameters, with the two degenerate limits of the reference (:27-33)"
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
        az = np.maximum(AZ_FLOOR, az)
        return (1 - self.compute_n_eff(az, ax)) / az

    def compute_forward_error(self, az, ax, tau_z):
        if ax == 0:      # nothing comes back from x yet: the prior variance of x = F z
            return self.ensemble.mean_spectrum / (self.alpha * az)
        return self.compute_n_eff(az, ax) / (self.alpha * ax)

    # -- free energy ------------------------------------------------------------
    def compute_mutual_information(self, az, ax, tau_z):
        return 0.5 * (np.log(az * tau_z) + self.ensemble.shannon_transform(ax / az))

    def compute_free_energy(self, az, ax, tau_z):
        overlap_terms = az * tau_z + self.alpha * ax * self.second_moment(tau_z)
        entropy_term = np.log(2 * np.pi * tau_z / np.e)
        return 0.5 * overlap_terms - self.compute_mutual_information(az, ax, tau_z) + 0.5 * entropy_term


class MarchenkoPasturChannel(AnalyticalLinearChannel):
    """F with iid N(0, 1/N) entries, alpha = M / N rows per column (reference :68-90)."""

    def __init__(self, alpha, name="W"):
        super().__init__(ensemble=MarchenkoPasturEnsemble(alpha=alpha), name=name)

    def compute_precision(self, vz, vx, tau_z):
        "The (az, ax) at which the channel's errors are (vz, vx) (reference :73-76)"
        ax = 1 / vx - 1 / vz
        return (1 - self.alpha * ax * vx) / vz, ax
