"""State Evolution over a grid of measurement densities: one batched launch on
the GPU (`StateEvolution([...])`, trb_se_run) next to the CPU port of the
reference's node-by-node scipy-quad recursion (oracle/se_oracle.py).

    python tools/bench_se.py [--grid 120] [--cpu-points 12] [--out profiles/r01_state_evolution.json]

This is the workload of the reference's examples (examples/glm/data/*_ep_vs_se.py,
examples/figures/sparse_regression.py: 100-120 values of alpha, max_iter 200).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=120)
    ap.add_argument("--cpu-points", type=int, default=12)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.algos import StateEvolution, EarlyStopping
    from oracle import se_oracle as S

    cases = {
        "compressed sensing (GaussBernoulli rho=0.5 / Gaussian var=1e-10)": dict(
            build=dict(prior_type="gauss_bernoulli", output_type="gaussian", prior_rho=0.5,
                       output_var=1e-10),
            oracle=(dict(kind="gauss_bernoulli", rho=0.5, mean=0, var=1), dict(kind="gaussian", var=1e-10)),
            alphas=np.linspace(0.01, 1.0, args.grid)),
        "perceptron (Binary p_pos=0.6 / Sgn)": dict(
            build=dict(prior_type="binary", output_type="sgn", prior_p_pos=0.6),
            oracle=(dict(kind="binary", p_pos=0.6), dict(kind="sgn")),
            alphas=np.linspace(0.02, 1.2, args.grid)),
        "phase retrieval (GaussBernoulli rho=0.6 mean=0.01 / Abs), 2-D quadrature": dict(
            build=dict(prior_type="gauss_bernoulli", output_type="abs", prior_rho=0.6, prior_mean=0.01),
            oracle=(dict(kind="gauss_bernoulli", rho=0.6, mean=0.01, var=1), dict(kind="abs")),
            alphas=np.linspace(0.05, 1.2, args.grid), max_iter=20, cpu_points=2),
    }
    report = dict(device=torch.cuda.get_device_name(0), grid=args.grid, cases={})
    for name, case in cases.items():
        max_iter = case.get("max_iter", 200)
        models = [glm_state_evolution(alpha=float(a), **case["build"]) for a in case["alphas"]]
        se = StateEvolution(models)
        se.iterate(max_iter=max_iter, callback=EarlyStopping())          # warm-up (module load, tables)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        se.iterate(max_iter=max_iter, callback=EarlyStopping())
        torch.cuda.synchronize()
        gpu_s = time.perf_counter() - t0
        v = se.get_variable_data("x")["v"]
        iters = int(se.n_iter_per_problem.sum())
        # CPU port on a subsample of the same grid
        idx = np.linspace(0, args.grid - 1, case.get("cpu_points", args.cpu_points)).astype(int)
        t0 = time.perf_counter()
        dev, cpu_iters = 0.0, 0
        for g in idx:
            r = S.se_glm(case["oracle"][0], dict(kind="marchenko", alpha=float(case["alphas"][g])),
                         case["oracle"][1], max_iter, early=dict(tol=1e-6))
            cpu_iters += r["n_iter"]
            dev = max(dev, abs(r["v"][0] - v[g]) / max(abs(r["v"][0]), 1e-12))
        cpu_s = time.perf_counter() - t0
        report["cases"][name] = dict(
            problems=args.grid, se_iterations_total=iters, gpu_seconds_whole_grid=gpu_s,
            gpu_se_iterations_per_s=iters / gpu_s,
            cpu_port_points=len(idx), cpu_port_seconds=cpu_s,
            cpu_port_se_iterations_per_s=cpu_iters / cpu_s,
            speedup_per_iteration=(iters / gpu_s) / (cpu_iters / cpu_s),
            max_rel_dev_v_vs_cpu_port=dev)
        print(name, json.dumps(report["cases"][name]))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
