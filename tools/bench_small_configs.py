"""BASELINE configs[0] and configs[1] as SINGLE instances (launch-bound on a GPU):
  0: GaussBernoulliPrior(N=1000, rho=0.1) @ LinearChannel(M=500) @ GaussianLikelihood(1e-2), 100 it
  1: GaussianPrior(N=2000) @ LinearChannel(alpha=2) @ SgnLikelihood, damping 0.5, 50 it
Times iterate() through the public API (persistent whole-sweep kernel; and the
launch-per-stage path with the general / automatic schedule, CUDA-graph replay
on / off) next to the CPU oracle on the same W, y, and checks parity.
-> gpurun_out/r01_small_configs.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import tramp_oracle as orc
from tramp_b200 import _lib
from tramp_b200.priors import GaussBernoulliPrior, GaussianPrior
from tramp_b200.likelihoods import GaussianLikelihood, SgnLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation, TrackErrors

torch.cuda.set_device(0)
lib = _lib.load()
res = {}
for key, N, M, n_iter, damping in (("config0_sparse_regression", 1000, 500, 100, None),
                                   ("config1_sign_perceptron", 2000, 4000, 50, 0.5)):
    rng = np.random.RandomState(42)
    W = rng.randn(M, N) / np.sqrt(N)
    if key.startswith("config0"):
        x = rng.randn(N) * (rng.rand(N) < 0.1)
        y = W @ x + 0.1 * rng.randn(M)
        prior, lik = GaussBernoulliPrior(size=N, rho=0.1), GaussianLikelihood(y=y, var=1e-2)
        pspec, lspec = dict(kind="gauss_bernoulli", rho=0.1), dict(kind="gaussian", var=1e-2, y=y)
    else:
        x = rng.randn(N)
        y = np.where(W @ x >= 0, 1.0, -1.0)
        prior, lik = GaussianPrior(size=N), SgnLikelihood(y=y)
        pspec, lspec = dict(kind="gaussian"), dict(kind="sgn", y=y)
    t0 = time.perf_counter()
    lin = LinearChannel(W)
    lin._setup()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    model = (prior @ V("x") @ lin @ V("z") @ lik).to_model()
    out = dict(N=N, M=M, n_iter=n_iter, damping=damping, gpu_setup_svd_s=setup_s)
    variants = [("persistent_grid", "general", 1, 2), ("persistent_cluster", "general", 1, 3)] + [
        (f"{schedule}_graphs{graphs}", schedule, graphs, 0)
        for schedule in ("general", "auto") for graphs in (1, 0)] + [("persistent_auto", "auto", 1, 1)]
    for label, schedule, graphs, persistent in variants:
        for _ in (0,):
            lib.trb_set_cuda_graphs(graphs)
            lib.trb_set_persistent_sweep(persistent)
            ep = ExpectationPropagation(model)
            ep.schedule = schedule
            best = 1e30
            for rep in range(6):
                track = TrackErrors({"x": x})
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ep.iterate(max_iter=n_iter, callback=track, damping=damping)
                d = ep.get_variables_data(["x"])
                best = min(best, time.perf_counter() - t0)
            out[label] = dict(ms_per_sweep=best * 1e3, us_per_iter=best / n_iter * 1e6,
                              iterations_per_s=n_iter / best)
            if label.startswith("persistent") or label == "auto_graphs1":
                # marginal cost of an iteration: the same sweep 5x longer, host-side
                # fixed costs (descriptor, record copies, callback replay set-up) cancel
                long_best = 1e30
                for rep in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    ep.iterate(max_iter=5 * n_iter, callback=TrackErrors({"x": x}), damping=damping)
                    ep.get_variables_data(["x"])
                    long_best = min(long_best, time.perf_counter() - t0)
                out[label]["us_per_iter_marginal"] = (long_best - best) / (4 * n_iter) * 1e6
    lib.trb_set_cuda_graphs(1)
    lib.trb_set_persistent_sweep(-1)
    t0 = time.perf_counter()
    op = orc.LinearOp(W)
    cpu_setup = time.perf_counter() - t0
    with np.errstate(all="ignore"):
        orc.ep_glm(pspec, W, lspec, n_iter, damping=damping, x_true=x, op=op)
        t0 = time.perf_counter()
        ref = orc.ep_glm(pspec, W, lspec, n_iter, damping=damping, x_true=x, op=op)
        cpu_s = time.perf_counter() - t0
    out["cpu_oracle"] = dict(ms_per_sweep=cpu_s * 1e3, iterations_per_s=n_iter / cpu_s, setup_svd_s=cpu_setup,
                             cores=os.cpu_count())
    mse = np.array([e["mse"] for e in track.errors])
    out["max_rel_dev_vs_oracle"] = float(max(
        np.max(np.abs(d["x"]["r"] - ref["r_x"])) / np.max(np.abs(ref["r_x"])),
        abs(d["x"]["v"] - ref["v_x"]) / ref["v_x"],
        np.max(np.abs(mse - np.array(ref["traj"]["mse_x"])) / np.array(ref["traj"]["mse_x"]))))
    res[key] = out
    print(key, json.dumps(out), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r01_small_configs.json", "w"), indent=1)
