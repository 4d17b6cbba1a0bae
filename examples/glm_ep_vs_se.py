"""EP variance, empirical mse and State-Evolution variance over a grid of alpha, for the
two Bayes-optimal GLMs whose tables the reference publishes under examples/glm/data/
(compressed_sensing_ep_vs_se.csv, perceptron_ep_vs_se.csv -- fixtures of
tests/test_gpu_se_reference_examples.py).  The protocol is the reference's: one
teacher-student instance per grid point, `BayesOptimalScenario.run_all` with the variance
based `EarlyStopping`, results stacked by `save_experiments` -- every call below exists
under the same name in `tramp`, which is what "drop-in" means for this path.
"""
import argparse
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tramp_b200.algos import EarlyStopping  # noqa: E402
from tramp_b200.experiments import BayesOptimalScenario, save_experiments  # noqa: E402
from tramp_b200.models import glm_generative  # noqa: E402

# table name -> (fixed model arguments, grid axes as functions of the number of alphas)
PROTOCOLS = {
    "compressed_sensing_ep_vs_se": (
        dict(prior_type="gauss_bernoulli", output_type="gaussian", output_var=1e-11, metrics=["mse"]),
        lambda n: dict(ensemble_type="gaussian", prior_rho=[0.25, 0.50, 0.75], alpha=np.linspace(0, 1, n)[1:])),
    # p_pos = 0.5 is left out: the symmetric perceptron violates az > 1/tau_z in State
    # Evolution (sgn_likelihood.py:80-81), in today's reference as well
    "perceptron_ep_vs_se": (
        dict(prior_type="binary", output_type="sgn", ensemble_type="gaussian", metrics=None),
        lambda n: dict(prior_p_pos=[0.25, 0.75], alpha=np.linspace(0, 2, 2 * n - 1)[1:])),
}


def make_runner(fixed):
    """run(N, alpha, **grid point) -> the records of one scenario (SE v, EP v, scores)."""
    fixed = dict(fixed)
    metrics = fixed.pop("metrics")

    def run(N, alpha, **point):
        model = glm_generative(N=N, alpha=alpha, **fixed, **point)
        scenario = BayesOptimalScenario(model, x_ids=["x"])
        score = dict(metrics=metrics) if metrics else {}
        return scenario.run_all(max_iter=200, callback=EarlyStopping(), **score)
    return run


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--alphas", type=int, default=50)
    ap.add_argument("--dir", default=os.path.dirname(os.path.abspath(__file__)))
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.WARNING)
    for table, (fixed, axes) in PROTOCOLS.items():
        save_experiments(make_runner(fixed), os.path.join(args.dir, table + ".csv"), N=args.n, **axes(args.alphas))


if __name__ == "__main__":
    main()
