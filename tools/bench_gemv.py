"""Micro-benchmark of the batched GEMV kernels at the north-star shape.
Usage: python tools/bench_gemv.py [B] ; prints GB/s per kernel/impl (CUDA events)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tramp_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, M = 4096, 2048
R = M
peak = 6550.1
torch.cuda.set_device(0)
lib = _lib.load()
res = {}
for name, n in (("V", N), ("U", M)):
    A = torch.randn(B, R, n, dtype=torch.float64, device="cuda")
    x = torch.randn(B, n, dtype=torch.float64, device="cuda")
    c = torch.randn(B, R, dtype=torch.float64, device="cuda")
    ns = ops.lin_expand_slots(B, R)
    part = torch.empty(B, ns, n, dtype=torch.float64, device="cuda")
    t = torch.empty(B, R, dtype=torch.float64, device="cuda")
    byts = B * R * n * 8
    st = torch.cuda.current_stream().cuda_stream
    for impl in (1, 2):
        for kind in ("project", "expand"):
            def run():
                if kind == "project":
                    _lib.check(lib.trb_lin_project(A.data_ptr(), A.stride(0), R, n, n, B, x.data_ptr(), n, t.data_ptr(), None, impl, st))
                else:
                    _lib.check(lib.trb_lin_expand(A.data_ptr(), A.stride(0), R, n, n, B, c.data_ptr(), part.data_ptr(), None, impl, st))
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = byts / ms / 1e6
            res[f"{name}_{kind}_impl{impl}"] = dict(ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 3))
            print(name, kind, "impl", impl, f"{ms:.3f} ms  {gbs:.0f} GB/s  {gbs/peak:.3f} of measured peak", flush=True)
    # correctness spot check between impls
    if True:
        t1 = torch.empty_like(t); t2 = torch.empty_like(t)
        lib.trb_lin_project(A.data_ptr(), A.stride(0), R, n, n, B, x.data_ptr(), n, t1.data_ptr(), None, 1, st)
        lib.trb_lin_project(A.data_ptr(), A.stride(0), R, n, n, B, x.data_ptr(), n, t2.data_ptr(), None, 2, st)
        ref = torch.einsum("brn,bn->br", A[:2], x[:2])
        print("project maxdiff impl1/impl2/ref:", (t1 - t2).abs().max().item(), (t1[:2] - ref).abs().max().item())
    del A, x, c, part, t
    torch.cuda.empty_cache()
# copy bandwidth reference (same method as MEASURED_PEAKS)
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("torch copy GB/s (r+w):", 2 * a.numel() * 2 / best / 1e6)
res["copy_gbs"] = 2 * a.numel() * 2 / best / 1e6
del a, b
# setup cost probe: batched eigh of 2048x2048 fp64
W = torch.randn(8, M, N, dtype=torch.float64, device="cuda") / N**0.5
torch.cuda.synchronize(); t0 = time.time()
G = W @ W.transpose(1, 2)
ev, U = torch.linalg.eigh(G)
torch.cuda.synchronize(); dt = time.time() - t0
print(f"eigh(WW^T) 8 x {M}^2 fp64: {dt:.2f} s  ({dt/8:.3f} s/instance)")
res["eigh_s_per_instance"] = dt / 8
t0 = time.time()
U2, s2, Vh2 = torch.linalg.svd(W[:2], full_matrices=False)
torch.cuda.synchronize(); dt = time.time() - t0
print(f"svd 2 x {M}x{N} fp64: {dt:.2f} s ({dt/2:.3f} s/instance)")
res["svd_s_per_instance"] = dt / 2
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_gemv.json", "w"), indent=1)
