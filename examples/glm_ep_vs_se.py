"""EP variance, empirical mse and State-Evolution variance over a grid of alpha: the
reference's examples/glm/data/compressed_sensing_ep_vs_se.py and perceptron_ep_vs_se.py
as they are, but for the import line -- the drop-in case.  (The published tables of
these two scripts are fixtures of tests/test_gpu_se_reference_examples.py.)"""
import argparse
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tramp_b200.models import glm_generative  # noqa: E402
from tramp_b200.experiments import save_experiments, BayesOptimalScenario  # noqa: E402
from tramp_b200.algos import EarlyStopping  # noqa: E402


def run_cs(N, alpha, ensemble_type, prior_rho):
    model = glm_generative(
        N=N, alpha=alpha, ensemble_type=ensemble_type,
        prior_type="gauss_bernoulli", output_type="gaussian",
        prior_rho=prior_rho, output_var=1e-11
    )
    scenario = BayesOptimalScenario(model, x_ids=["x"])
    early = EarlyStopping()
    return scenario.run_all(metrics=["mse"], max_iter=200, callback=early)


def run_perceptron(N, alpha, p_pos):
    model = glm_generative(
        N=N, alpha=alpha, ensemble_type="gaussian", prior_type="binary", output_type="sgn",
        prior_p_pos=p_pos
    )
    scenario = BayesOptimalScenario(model, x_ids=["x"])
    early = EarlyStopping()
    return scenario.run_all(max_iter=200, callback=early)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--alphas", type=int, default=50)
    ap.add_argument("--dir", default=os.path.dirname(os.path.abspath(__file__)))
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.WARNING)
    save_experiments(run_cs, os.path.join(args.dir, "compressed_sensing_ep_vs_se.csv"),
                     N=args.n, ensemble_type="gaussian", prior_rho=[0.25, 0.50, 0.75],
                     alpha=np.linspace(0, 1, args.alphas)[1:])
    # p_pos = 0.5 is left out: the symmetric perceptron violates az > 1/tau_z in State
    # Evolution (sgn_likelihood.py:80-81), in today's reference as well
    save_experiments(run_perceptron, os.path.join(args.dir, "perceptron_ep_vs_se.csv"),
                     N=args.n, p_pos=[0.25, 0.75], alpha=np.linspace(0, 2, 2 * args.alphas + 1)[1:])


if __name__ == "__main__":
    main()
