"""The reference's OWN published benchmark, protocol of
examples/figures/compute_benchmark.py:16-27, 37-40, 66-70 (figure
examples/figures/benchmark.pdf; values recovered in BASELINE.md section 1):

    GaussBernoulliPrior(N=1000, rho=0.05) @ LinearChannel(Gaussian W, M=alpha N)
    @ GaussianChannel(var=1e-2),  BayesOptimalScenario.setup(seed),
    run_ep(max_iter=1000, damping=0.1) with the default EarlyStoppingEP,
    time = EP wall time + SVD precomputation time, median over seeds;
    run_se(max_iter=1000, damping=0.1) for the Bayes-optimal mse.

Run through tramp_b200's public API (same calls, same seeds) on the GPU, and
through the CPU port of the reference (oracle/) on the box's host.
-> gpurun_out/r02_published_protocol.json
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import tramp_oracle as orc
from tramp_b200.ensembles import GaussianEnsemble
from tramp_b200.priors import GaussBernoulliPrior
from tramp_b200.channels import GaussianChannel, LinearChannel
from tramp_b200.variables import SISOVariable as V, SILeafVariable as O
from tramp_b200.algos.metrics import mean_squared_error
from tramp_b200.experiments import BayesOptimalScenario

# (alpha, published EP wall time incl. SVD in seconds), BASELINE.md section 1
PUBLISHED = [(0.02, 0.19), (0.16, 0.70), (0.30, 0.65), (0.44, 0.85), (0.50, 0.78), (0.58, 0.86),
             (0.72, 1.19), (0.86, 1.45), (1.00, 1.70)]
PUBLISHED_MSE_OVER_RHO = {0.30: (0.133, 0.128), 0.44: (0.069, 0.067), 0.58: (0.0456, 0.0445),
                          1.00: (0.0220, 0.0212)}        # (EP, SE)
N, RHO, NOISE = 1000, 0.05, 1e-2
SEEDS = range(5)


def run_gpu(alpha, seed, svd_method="svd"):
    M = int(alpha * N)
    A = GaussianEnsemble(M=M, N=N).generate()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lin = LinearChannel(A, svd_method=svd_method)
    lin._setup()                      # the reference factorises in the constructor
    model = (GaussBernoulliPrior(size=N, rho=RHO) @ V("x") @ lin @ V("z")
             @ GaussianChannel(var=NOISE) @ O("y")).to_model()
    torch.cuda.synchronize()
    rec = dict(svd_time=time.perf_counter() - t0)
    scenario = BayesOptimalScenario(model, x_ids=["x"])
    scenario.setup(seed)
    t0 = time.perf_counter()
    x_data = scenario.run_ep(max_iter=1000, damping=0.1)
    rec["time"] = time.perf_counter() - t0
    rec["n_iter"] = x_data["n_iter"]
    rec["mse"] = float(mean_squared_error(x_data["x"]["r"], scenario.x_true["x"]))
    t0 = time.perf_counter()
    se = scenario.run_se(max_iter=1000, damping=0.1)
    rec["se_time"] = time.perf_counter() - t0
    rec["se_v"] = se["x"]["v"]
    rec["se_n_iter"] = se["n_iter"]
    return rec, A, scenario


def run_cpu(A, scenario):
    """The same run through the CPU port of the reference's algorithm."""
    t0 = time.perf_counter()
    op = orc.LinearOp(A)              # matrix_rank + full SVD, linear_channel.py:37-41
    svd = time.perf_counter() - t0
    y = scenario.observations["y"]
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        r = orc.ep_glm(dict(kind="gauss_bernoulli", rho=RHO), A, dict(kind="gaussian", var=NOISE, y=y),
                       1000, damping=0.1, early_stopping=dict(tol=1e-6, wait_increase=5, max_increase=0.2),
                       op=op)
    return dict(svd_time=svd, time=time.perf_counter() - t0, n_iter=len(r["traj"]["v_x"]),
                mse=float(mean_squared_error(r["r_x"], scenario.x_true["x"])))


def main():
    torch.cuda.set_device(0)
    np.random.seed(123)
    run_gpu(0.3, 0)                   # warm-up: CUDA context, cuSOLVER handles, module load
    run_gpu(0.3, 0, "auto")
    out = dict(protocol="examples/figures/compute_benchmark.py", N=N, rho=RHO, noise_var=NOISE,
               seeds=len(SEEDS), device=torch.cuda.get_device_name(0), host_cores=os.cpu_count(), rows=[])
    for alpha, published in PUBLISHED:
        gpu, cpu, auto, lib = [], [], [], []
        for seed in SEEDS:
            np.random.seed(1000 + seed)
            auto.append(run_gpu(alpha, seed, "auto")[0])      # svd_method="auto": the hand-written block-Jacobi set-up
            if alpha <= 0.92:
                np.random.seed(1000 + seed)
                lib.append(run_gpu(alpha, seed, "gram")[0])   # round-1 default: Gram + cuSOLVER eigh
            np.random.seed(1000 + seed)
            rec, A, scenario = run_gpu(alpha, seed)
            gpu.append(rec)
            if seed < 2:
                cpu.append(run_cpu(A, scenario))
        med = lambda rows, f: float(np.median([f(r) for r in rows]))
        row = dict(alpha=alpha, published_total_s=published,
                   gpu_total_s=med(gpu, lambda r: r["time"] + r["svd_time"]),
                   gpu_ep_s=med(gpu, lambda r: r["time"]), gpu_svd_s=med(gpu, lambda r: r["svd_time"]),
                   gpu_n_iter=med(gpu, lambda r: r["n_iter"]), gpu_se_s=med(gpu, lambda r: r["se_time"]),
                   mse_over_rho=med(gpu, lambda r: r["mse"]) / RHO, se_v_over_rho=med(gpu, lambda r: r["se_v"]) / RHO,
                   cpu_port_total_s=med(cpu, lambda r: r["time"] + r["svd_time"]),
                   cpu_port_ep_s=med(cpu, lambda r: r["time"]), cpu_port_n_iter=med(cpu, lambda r: r["n_iter"]),
                   cpu_port_mse_over_rho=med(cpu, lambda r: r["mse"]) / RHO)
        row.update(gpu_auto_total_s=med(auto, lambda r: r["time"] + r["svd_time"]),
                   gpu_auto_svd_s=med(auto, lambda r: r["svd_time"]),
                   gpu_auto_mse_over_rho=med(auto, lambda r: r["mse"]) / RHO,
                   gpu_auto_n_iter=med(auto, lambda r: r["n_iter"]))
        if lib:
            row.update(gpu_gram_eigh_total_s=med(lib, lambda r: r["time"] + r["svd_time"]),
                       gpu_gram_eigh_svd_s=med(lib, lambda r: r["svd_time"]))
        row["speedup_vs_published_auto"] = published / row["gpu_auto_total_s"]
        row["speedup_vs_published"] = published / row["gpu_total_s"]
        row["speedup_vs_cpu_port_same_host"] = row["cpu_port_total_s"] / row["gpu_total_s"]
        if alpha in PUBLISHED_MSE_OVER_RHO:
            row["published_mse_over_rho_ep_se"] = PUBLISHED_MSE_OVER_RHO[alpha]
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r02_published_protocol.json", "w"), indent=1)


if __name__ == "__main__":
    main()
