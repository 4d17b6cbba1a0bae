"""Where does the single-cluster variant of the persistent sweep beat the
cooperative-grid one?  Marginal microseconds per EP iteration of ONE sparse-GLM
instance (GaussBernoulli rho=0.1 / Gaussian var=1e-2, alpha=0.5) for a range of
N, for modes 3 (cluster), 2 (grid) and 0 (launch-per-stage + CUDA graph).
-> gpurun_out/r01f_persistent_sizes.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tramp_b200 import _lib
from tramp_b200.priors import GaussBernoulliPrior
from tramp_b200.likelihoods import GaussianLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation, PassCallback

lib = _lib.load()
res = {}
for N in (64, 128, 256, 512, 768, 1000, 1500):
    M = N // 2
    rng = np.random.RandomState(N)
    W = rng.randn(M, N) / np.sqrt(N)
    x = rng.randn(N) * (rng.rand(N) < 0.1)
    y = W @ x + 0.1 * rng.randn(M)
    lin = LinearChannel(W)
    lin._setup()
    model = (GaussBernoulliPrior(size=N, rho=0.1) @ V("x") @ lin @ V("z")
             @ GaussianLikelihood(y=y, var=1e-2)).to_model()
    row = dict(operator_bytes_per_iteration=16 * M * (N + M))
    for label, mode in (("cluster", 3), ("grid", 2), ("graph", 0)):
        lib.trb_set_persistent_sweep(mode)
        ep = ExpectationPropagation(model)
        ep.schedule = "general"
        t = {}
        for n_iter in (100, 500):
            best = 1e30
            for rep in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ep.iterate(max_iter=n_iter, callback=PassCallback())
                ep.get_variables_data(["x"])
                best = min(best, time.perf_counter() - t0)
            t[n_iter] = best
        row[label] = (t[500] - t[100]) / 400 * 1e6
    lib.trb_set_persistent_sweep(-1)
    res[N] = row
    print(N, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r01f_persistent_sizes.json", "w"), indent=1)
