"""Shared by the example scripts: a batch of teacher-student instances as ONE
generative model, and State-Evolution curves as one launch."""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tramp_b200.algos import CustomInit  # noqa: E402
from tramp_b200.experiments import BayesOptimalScenario, run_state_evolution_grid  # noqa: E402
from tramp_b200.models import glm_generative, glm_state_evolution  # noqa: E402


def batched_scenario(N, alpha, instances, seed, **glm):
    """`instances` independent (W, x, y) of the same shape in one model: the matrices
    are drawn first, then x, then the output noise, all from numpy's global RNG seeded
    once (the reference draws one instance per call, glm_generative :20-23)."""
    np.random.seed(seed)
    model = glm_generative(N=N, alpha=alpha, ensemble_type="gaussian", ensemble_batch=instances,
                           prior_batch=instances, **glm)
    scenario = BayesOptimalScenario(model, x_ids=["x"])
    scenario.setup()
    return scenario


def se_curve(alphas, source, a0=None, callback=None, **glm):
    """v_x at the SE fixed point for every alpha, one problem per CTA of one launch.
    a0: initial precision of the x -> prior edge, a number or one value per alpha."""
    models = [glm_state_evolution(alpha=float(alpha), **glm) for alpha in alphas]
    kwargs = dict(max_iter=200)
    if a0 is not None:
        kwargs["initializer"] = CustomInit(a_init=[("x", "bwd", a0)])
    if callback is not None:
        kwargs["callback"] = callback
    records = run_state_evolution_grid(["x"], models, **kwargs)
    return pd.DataFrame([dict(alpha=float(alpha), source=source, v=rec[0]["v"], n_iter=rec[0]["n_iter"])
                         for alpha, rec in zip(alphas, records)])
