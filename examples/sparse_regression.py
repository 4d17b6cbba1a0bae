"""Sparse linear regression: EP against State Evolution and the Bayes-optimal error
(reference examples/figures/sparse_regression.py:45-103, without the plot).

    GaussBernoulliPrior(N, rho = 0.5) @ V @ LinearChannel(Gaussian W, M = alpha N) @ V
        @ GaussianChannel(var = 1e-10)

EP: the mse of `instances` teacher-student instances per alpha, averaged -- one batched
run per alpha instead of the reference's loop over seeds.  SE: uninformed
initialisation; BO: informed initialisation a0 = 10^(3 e^alpha) (reference :81-83),
one value per alpha.  Both curves are one launch each.
"""
import argparse
import logging

import numpy as np
import pandas as pd

from _common import batched_scenario, se_curve

GLM = dict(prior_type="gauss_bernoulli", output_type="gaussian", output_var=1e-10)


def run_EP(alpha, rho, N, instances, seed=0):
    scenario = batched_scenario(N, alpha, instances, seed, prior_rho=rho, **GLM)
    scenario.run_ep(max_iter=200)                       # default EarlyStoppingEP, per instance
    mse = scenario.compute_score(scenario.x_pred)["x"]["mse"]
    return dict(alpha=alpha, source="EP", v=float(np.mean(mse)), v_std=float(np.std(mse)),
                n_iter=int(scenario.ep.n_iter))


def run_SE(alphas, rho):
    return se_curve(alphas, "SE", prior_rho=rho, **GLM)


def run_BO(alphas, rho):
    return se_curve(alphas, "BO", a0=10 ** (3 * np.exp(np.asarray(alphas, float))), prior_rho=rho, **GLM)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--instances", type=int, default=25)
    ap.add_argument("--rho", type=float, default=0.5)
    ap.add_argument("--ep-alphas", type=int, default=33)
    ap.add_argument("--se-alphas", type=int, default=100)
    ap.add_argument("--csv", default=__file__.replace(".py", ".csv"))
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.WARNING)
    ep = pd.DataFrame([run_EP(float(alpha), args.rho, args.n, args.instances, seed=k)
                       for k, alpha in enumerate(np.linspace(0.03, 0.99, args.ep_alphas))])
    se_alphas = np.linspace(0.01, 1.0, args.se_alphas)
    df = pd.concat([ep, run_SE(se_alphas, args.rho), run_BO(se_alphas, args.rho)], ignore_index=True, sort=False)
    df["rho"] = args.rho
    df.to_csv(args.csv, index=False)
    return df


if __name__ == "__main__":
    print(main().groupby("source").v.describe())
