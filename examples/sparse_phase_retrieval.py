"""Sparse phase retrieval (real-valued): EP against State Evolution and the
Bayes-optimal error (reference examples/figures/sparse_phase_retrieval.py:47-112).

    GaussBernoulliPrior(N, rho = 0.6, mean = 0.01) @ V @ LinearChannel(Gaussian W) @ V @ AbsChannel

EP with damping 0.3 and EarlyStopping(wait_increase = 10), error up to a global sign;
SE from a0 = 0.1, BO from a0 = 1000 (2-D measure of AbsLikelihood on the device).
"""
import argparse
import logging

import numpy as np
import pandas as pd

from _common import batched_scenario, se_curve
from tramp_b200.algos import EarlyStopping

GLM = dict(prior_type="gauss_bernoulli", output_type="abs", prior_mean=0.01)


def run_EP(alpha, rho, N, instances, seed=0):
    scenario = batched_scenario(N, alpha, instances, seed, prior_rho=rho, **GLM)
    scenario.run_ep(max_iter=200, damping=0.3, callback=EarlyStopping(wait_increase=10))
    mse = scenario.compute_score(scenario.x_pred, metrics=["sign_mse"])["x"]["sign_mse"]
    return dict(alpha=alpha, source="EP", v=float(np.mean(mse)), v_std=float(np.std(mse)),
                n_iter=int(scenario.ep.n_iter))


def run_SE(alphas, rho):
    return se_curve(alphas, "SE", a0=0.1, callback=EarlyStopping(wait_increase=10), prior_rho=rho, **GLM)


def run_BO(alphas, rho):
    return se_curve(alphas, "BO", a0=10**3, callback=EarlyStopping(wait_increase=10), prior_rho=rho, **GLM)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--instances", type=int, default=25)
    ap.add_argument("--rho", type=float, default=0.6)
    ap.add_argument("--ep-alphas", type=int, default=40)
    ap.add_argument("--se-alphas", type=int, default=120)
    ap.add_argument("--csv", default=__file__.replace(".py", ".csv"))
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.WARNING)
    ep = pd.DataFrame([run_EP(float(alpha), args.rho, args.n, args.instances, seed=k)
                       for k, alpha in enumerate(np.linspace(0.03, 1.2, args.ep_alphas))])
    se_alphas = np.linspace(0.01, 1.2, args.se_alphas)
    df = pd.concat([ep, run_SE(se_alphas, args.rho), run_BO(se_alphas, args.rho)], ignore_index=True, sort=False)
    df["rho"] = args.rho
    df.to_csv(args.csv, index=False)
    return df


if __name__ == "__main__":
    print(main().groupby("source").v.describe())
