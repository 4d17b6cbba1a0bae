"""The north-star workload through the public API on all GPUs of a box: B independent
sparse-GLM teacher-student instances (GaussBernoulliPrior @ LinearChannel(Gaussian W) @
GaussianLikelihood), sharded over the ranks by contiguous blocks -- no communication
during the sweep, one gather of the results at the end.

    torchrun --standalone --nproc-per-node 8 examples/sharded_instances.py --instances 4096
    python examples/sharded_instances.py --instances 64          # one GPU

Each rank draws, factorises and keeps only its own block (`build_model(start, stop)`);
every rank gets the gathered posterior means, variances, iteration counts and mse
trajectories.  bench.py times the same sweep with operators generated in factored form."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tramp_b200.channels import LinearChannel  # noqa: E402
from tramp_b200.experiments import run_ep_sharded  # noqa: E402
from tramp_b200.likelihoods import GaussianLikelihood  # noqa: E402
from tramp_b200.priors import GaussBernoulliPrior  # noqa: E402
from tramp_b200.variables import SISOVariable as V  # noqa: E402


def instance(index, N, M, rho, var_noise, seed):
    """Instance `index` of the job, drawn as the reference does (gaussian_ensemble.py:19-20,
    gauss_bernoulli_prior.py:38-42, gaussian_channel.py:12-15) from its own seed, so that a
    block does not depend on how the job is sharded."""
    rng = np.random.RandomState(seed + index)
    W = rng.randn(M, N) / np.sqrt(N)
    x = rng.standard_normal(N) * rng.binomial(n=1, size=N, p=rho)
    y = W @ x + np.sqrt(var_noise) * rng.standard_normal(M)
    return W, x, y


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=64)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--rho", type=float, default=0.1)
    ap.add_argument("--var-noise", type=float, default=1e-2)
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    N, M = args.n, int(args.alpha * args.n)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    blocks = {}

    def block(start, stop):
        if (start, stop) not in blocks:
            blocks[(start, stop)] = [np.stack(part) for part in zip(*(
                instance(i, N, M, args.rho, args.var_noise, args.seed) for i in range(start, stop)))]
        return blocks[(start, stop)]

    def build_model(start, stop):
        W, _, y = block(start, stop)
        return (GaussBernoulliPrior(size=N, rho=args.rho, batch=stop - start) @ V("x") @ LinearChannel(W)
                @ V("z") @ GaussianLikelihood(y=y, var=args.var_noise)).to_model()

    res = run_ep_sharded(build_model, args.instances, x_true=lambda a, b: {"x": block(a, b)[1]},
                         max_iter=args.max_iter)
    if int(os.environ.get("RANK", "0")) == 0:
        final = np.array([res["mse"][n - 1, b] for b, n in enumerate(res["n_iter"])])
        print(f"{args.instances} instances on {world} rank(s): N={N} M={M}, "
              f"iterations {res['n_iter'].min()}..{res['n_iter'].max()}, "
              f"mse {final.mean():.4g} (EP variance {res['v']['x'].mean():.4g})")
    if world > 1:
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    main()
