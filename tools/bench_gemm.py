"""Microbenchmark of the shared-operator FP64 tensor-core GEMMs (trb_gemm.cu)
against cuBLAS DGEMM (torch.matmul) on the BASELINE config-4 shapes:
project T[B,R] = X[B,n] A[R,n]^T and expand O[B,n] = C[B,R] A[R,n].
Usage: python tools/bench_gemm.py [N] [B] [reps]   -> gpurun_out/r01_gemm_microbench.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tramp_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
R = N
torch.cuda.set_device(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
A = torch.randn((1, R, N), dtype=torch.float64, device="cuda", generator=g) / N ** 0.5
X = torch.randn((B, N), dtype=torch.float64, device="cuda", generator=g)
Cf = torch.randn((B, R), dtype=torch.float64, device="cuda", generator=g)
T = torch.zeros((B, R), dtype=torch.float64, device="cuda")
O = torch.zeros((B, N), dtype=torch.float64, device="cuda")
flops = 2.0 * B * R * N


def timed(fn):
    for _ in range(2):
        fn()
    best, tot = 1e30, 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    return dict(ms_best=best, ms_avg=tot / reps, tflops_best=flops / best / 1e9, tflops_avg=flops / (tot / reps) / 1e9)


from tramp_b200 import _lib
res = dict(N=N, R=R, B=B, reps=reps)
_lib.load().trb_gemm_set_variant(1)
res["project_dmma_cpasync"] = timed(lambda: ops.lin_project_gemm(A, R, N, X, B, out=T))
res["expand_dmma_cpasync"] = timed(lambda: ops.lin_expand_gemm(A, R, N, Cf, B, out=O))
probes = () if os.environ.get("BENCH_GEMM_NO_PROBES") else ((2, "diag_noload"), (4, "loadonly"))
for v, nm in probes:
    _lib.load().trb_gemm_set_variant(v)
    res["project_dmma_" + nm] = timed(lambda: ops.lin_project_gemm(A, R, N, X, B, out=T))
    res["expand_dmma_" + nm] = timed(lambda: ops.lin_expand_gemm(A, R, N, Cf, B, out=O))
_lib.load().trb_gemm_set_variant(0)
T.zero_(); O.zero_()
res["project_dmma"] = timed(lambda: ops.lin_project_gemm(A, R, N, X, B, out=T))
t_ref = X @ A[0].T
res["project_max_abs_err_vs_cublas"] = float((T - t_ref).abs().max())
res["project_cublas"] = timed(lambda: torch.matmul(X, A[0].T, out=T))
res["expand_dmma"] = timed(lambda: ops.lin_expand_gemm(A, R, N, Cf, B, out=O))
o_ref = Cf @ A[0]
res["expand_max_abs_err_vs_cublas"] = float((O - o_ref).abs().max())
res["expand_cublas"] = timed(lambda: torch.matmul(Cf, A[0], out=O))
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn_like(a); c = torch.empty_like(a)
fl = flops
flops = 2.0 * 8192 ** 3
res["cublas_dgemm_8192"] = timed(lambda: torch.matmul(a, b, out=c))
flops = fl
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/r01_gemm_microbench_N{N}_B{B}.json", "w"), indent=1)
