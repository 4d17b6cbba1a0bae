"""Random matrix ensembles (reference tramp/ensembles/): Gaussian only."""
import numpy as np
from ..base import ReprMixin


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


ENSEMBLE_CLASSES = {"gaussian": GaussianEnsemble}


def get_ensemble(ensemble_type, **kwargs):
    return ENSEMBLE_CLASSES[ensemble_type](**kwargs)
