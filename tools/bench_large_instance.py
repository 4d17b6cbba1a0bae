"""BASELINE config 5 at full size: ONE instance, GaussBernoulliPrior(N=65536) @
LinearChannel(alpha=0.6) @ GaussianLikelihood, thin-SVD operators row-sharded
over the ranks, one NCCL all-reduce per half sweep.

  python -m torch.distributed.run --nproc-per-node G tools/bench_large_instance.py [N] [iters]

Synthetic operator: a dense 39321 x 65536 FP64 SVD is hours of setup, so the
factors are drawn directly.  Rank g's singular vectors are Haar-distributed on a
column block of their own (disjoint supports make the shards mutually
orthogonal without a cross-rank Gram-Schmidt); rows are stored at FULL length,
so the streamed bytes are those of a dense operator.  Singular values follow the
Gaussian ensemble (bidiagonal model).  Not a Gaussian W -- stated in the output."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tramp_b200 import synthetic, ops
from tramp_b200.priors import GaussBernoulliPrior
from tramp_b200.likelihoods import GaussianLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation, TrackErrors
from tramp_b200.distributed import instance_shard

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
alpha, rho, var = 0.6, 0.1, 1e-2
M = int(alpha * N); R = M
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
# rank g owns the singular triplets g, g + world, g + 2 world, ...: every shard then sees
# the same spectrum distribution (a contiguous block would give rank 0 all the large ones,
# and the isotropic EP variances would be inconsistent across the blocks)
idx = np.arange(rank, R, world); Rg = idx.size
gen = torch.Generator(device="cuda"); gen.manual_seed(100 + rank)
ldn, ldm = ops.pad_ld(N), ops.pad_ld(M)


def block_rows(Rg, n, ld, rank, world):
    """[1, Rg, ld] with orthonormal rows supported on this rank's column block."""
    c0, c1 = instance_shard(n, rank, world)
    assert Rg <= c1 - c0, "shard has more rows than its column block"
    out = torch.zeros((1, Rg, ld), dtype=torch.float64, device="cuda")
    out[0, :, c0:c1] = synthetic.haar_rows(1, Rg, c1 - c0, gen, chunk=1, ld=c1 - c0)[0]
    return out


Vt = block_rows(Rg, N, ldn, rank, world)
Ut = block_rows(Rg, M, ldm, rank, world)
s_full = synthetic.gaussian_singular_values(1, M, N, seed=5, workers=1)[0]     # same on every rank
s_loc = torch.as_tensor(s_full[idx].copy(), device="cuda")[None]
g2 = torch.Generator(device="cuda"); g2.manual_seed(7)                         # same on every rank
x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g2)
x = x * (torch.rand(N, dtype=torch.float64, device="cuda", generator=g2) < rho)
z = (Ut[0, :, :M].T @ (s_loc[0] * (Vt[0, :, :N] @ x)))
dist.all_reduce(z)
y = z + np.sqrt(var) * torch.randn(M, dtype=torch.float64, device="cuda", generator=g2)
torch.cuda.synchronize(); setup_s = time.time() - t0
lin = LinearChannel.from_sharded_factors(Ut, s_loc, Vt, s_full, Nx=M, Nz=N, group=dist.group.WORLD)
model = (GaussBernoulliPrior(size=N, rho=rho) @ V("x") @ lin @ V("z") @ GaussianLikelihood(y=y, var=var)).to_model()
bytes_it = 16.0 * R * (N + M)
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6550.1
res = dict(config="single large instance (BASELINE configs[4]), thin-SVD operators row-sharded over the ranks",
           N=N, M=M, R=R, n_gpus=world, iters=iters, algorithmic_GB_per_iter=bytes_it / 1e9, setup_s=setup_s,
           schedule="general 4-pass", hbm_peak_gbs=peak,
           operator="synthetic block-orthogonal singular vectors, Gaussian-ensemble spectrum (not a Gaussian W)")
for backend in ("sharded", "sharded_nccl"):
    ep = ExpectationPropagation(model)
    ep.linear_backend = backend
    ep.schedule = "general"
    track = TrackErrors({"x": x})
    ep.iterate(max_iter=3, callback=track)        # warm-up
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ep.iterate(max_iter=iters, callback=track)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    mse = [float(e["mse"]) for e in track.errors]
    res[backend] = dict(
        exchange=("peer memory inside the update kernels (trb_comm.cu)" if backend == "sharded"
                  else "NCCL all-reduce between kernels, staged from Python"),
        ms_per_iter=ms / iters, iterations_per_s=iters / (ms / 1e3),
        GBps_per_gpu=bytes_it / world / (ms / iters / 1e3) / 1e9,
        frac_of_hbm_peak=bytes_it / world / (ms / iters / 1e3) / 1e9 / peak,
        mse_first=mse[0], mse_last=mse[-1])
res["mse_signal"] = float((x**2).mean().item())
if rank == 0:
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/r01_config5_large_instance_{world}gpu.json", "w"), indent=1)
dist.destroy_process_group()
