"""Random matrix ensembles (reference tramp/ensembles/): Gaussian only."""
import numpy as np
from ..base import ReprMixin, Registry


class Ensemble(ReprMixin):
    pass


class GaussianEnsemble(Ensemble):
    """iid N(0, 1/N) entries (reference ensembles/gaussian_ensemble.py:5-21);
    `batch=B` draws B matrices in sequence from the same numpy global RNG."""

    def __init__(self, M, N, batch=None):
        self.M = M
        self.N = N
        self.repr_init()
        self.batch = batch

    def generate(self):
        sigma_x = 1 / np.sqrt(self.N)
        if self.batch is None:
            return sigma_x * np.random.randn(self.M, self.N)
        return sigma_x * np.random.randn(self.batch, self.M, self.N)


class MarchenkoPasturEnsemble(Ensemble):
    """Spectral law of W^T W for iid N(0, 1/N) entries, M = alpha N (reference
    ensembles/marchenko_pastur_ensemble.py:6-54): bulk on [z_min, z_max] plus an
    atom of mass max(0, 1 - alpha) at 0."""

    def __init__(self, alpha):
        self.alpha = alpha
        self.repr_init()
        self.z_max = (1 + np.sqrt(alpha))**2
        self.z_min = (1 - np.sqrt(alpha))**2
        # reference :13 integrates z against the bulk with scipy quad; the integral
        # of sqrt((z - z_min)(z_max - z)) / (2 pi) is (z_max - z_min)^2 / 16 = alpha
        self.mean_spectrum = (self.z_max - self.z_min)**2 / 16

    def generate(self, N=1000):
        M = int(self.alpha * N)
        return np.random.randn(M, N) / np.sqrt(N)

    def bulk_density(self, z):
        return np.sqrt((z - self.z_min) * (self.z_max - z)) / (2 * np.pi * z)

    def measure(self, f, n=4096):
        """atomic + bulk part of the integral of f against the law (reference :31-38).
        The bulk carries a sqrt weight at both edges: Gauss-Chebyshev (second kind)
        nodes integrate it exactly instead of adaptive quad."""
        k = np.arange(1, n + 1)
        u = np.cos(k * np.pi / (n + 1))
        wts = np.pi / (n + 1) * np.sin(k * np.pi / (n + 1))**2
        c, h = 0.5 * (self.z_max + self.z_min), 0.5 * (self.z_max - self.z_min)
        z = c + h * u
        bulk = np.sum(wts * h * h * f(z) / (2 * np.pi * z))
        return max(0, 1 - self.alpha) * f(0) + bulk

    def compute_F(self, gamma):
        return (np.sqrt(gamma * self.z_max + 1) - np.sqrt(gamma * self.z_min + 1))**2

    def eta_transform(self, gamma):
        return 1 - self.compute_F(gamma) / (4 * gamma)

    def shannon_transform(self, gamma):
        F = self.compute_F(gamma)
        return (np.log(1 + self.alpha * gamma - F / 4)
                + self.alpha * np.log(1 + gamma - F / 4) - F / (4 * gamma))


ENSEMBLE_CLASSES = Registry("ensemble", {"gaussian": GaussianEnsemble, "marchenko_pastur": MarchenkoPasturEnsemble})


def get_ensemble(ensemble_type, **kwargs):
    return ENSEMBLE_CLASSES[ensemble_type](**kwargs)
