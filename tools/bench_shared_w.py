"""BASELINE config 4 at full size: phase retrieval with ONE W shared by a batch.
BinaryPrior(p_pos=0.6) @ LinearChannel(N=16384, alpha=1) @ AbsLikelihood, 1024
instances, damping 0.3.  The four operator passes are dense FP64 GEMMs
([B, n] x [n, R]): "gemm" = the DMMA kernels of trb_gemm.cu inside trb_sweep_run,
"cublas" = torch.matmul (library baseline), "gemv" = the HBM-bound batched GEMVs on
the shared operator.  Reports the sweep rate next to a measured cuBLAS DGEMM peak.  Usage: python tools/bench_shared_w.py [N] [B] [iters]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tramp_b200 import synthetic
from tramp_b200.priors import BinaryPrior
from tramp_b200.likelihoods import AbsLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation, TrackErrors, TrackEvolution, JoinCallback

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
M = N
torch.cuda.set_device(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(4)
t0 = time.time()
Vt = synthetic.haar_rows(1, N, N, gen, chunk=1)
Ut = synthetic.haar_rows(1, M, M, gen, chunk=1)
s = torch.as_tensor(synthetic.gaussian_singular_values(1, M, N, 4, workers=1), device="cuda")
x = torch.where(torch.rand((B, N), device="cuda", generator=gen, dtype=torch.float64) < 0.6, 1.0, -1.0).to(torch.float64)
z = ((x @ Vt[0, :, :N].T) * s) @ Ut[0, :, :M]
y = z.abs()
torch.cuda.synchronize(); setup_s = time.time() - t0
lin = LinearChannel.from_factors(Ut, s, Vt, Nx=M, Nz=N, rank=N)
model = (BinaryPrior(size=N, p_pos=0.6, batch=B) @ V("x") @ lin @ V("z") @ AbsLikelihood(y=y)).to_model()
res = {}
for backend in ("gemm", "cublas", "gemv"):
    ep = ExpectationPropagation(model)
    ep.linear_backend = backend
    n_it = iters if backend != "gemv" else 2
    track = TrackErrors({"x": x}, metrics=["sign_mse"])
    ep.iterate(max_iter=2, callback=track, damping=0.3)          # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ep.iterate(max_iter=n_it, callback=JoinCallback([track, TrackEvolution()]), damping=0.3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flops = 4.0 * N * (N + M) * B * n_it
    smse = np.array([e["sign_mse"] for e in track.errors])
    res[backend] = dict(iters=n_it, ms_per_iter=ms / n_it, inst_it_per_s=B * n_it / (ms / 1e3),
                        tflops=flops / (ms / 1e3) / 1e12,
                        sign_mse_first=float(smse[0].mean()), sign_mse_last=float(smse[-1].mean()))
    print(backend, res[backend], flush=True)
# cuBLAS DGEMM peak (8192^3), the denominator for the GEMM path
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn_like(a)
for _ in range(2): a @ b
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["dgemm_peak_tflops"] = 2 * 8192**3 / (best / 1e3) / 1e12
res["gemm_frac_of_dgemm_peak"] = res["gemm"]["tflops"] / res["dgemm_peak_tflops"]
res["cublas_frac_of_dgemm_peak"] = res["cublas"]["tflops"] / res["dgemm_peak_tflops"]
res.update(N=N, M=M, B=B, setup_s=setup_s)
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r01_config4_shared_w.json", "w"), indent=1)
