"""Set-up of LinearChannel (the factorisation the reference does in its constructor,
channels/linear/linear_channel.py:8-15, 36-46, and counts in EP's total time,
examples/figures/benchmark.py:22): how fast are B matrices of the north-star shape brought
into thin-SVD form on one B200?

Variants timed on the same real Gaussian W [B, M, N] = randn / sqrt(N):

  jacobi         the hand-written set-up (thin_svd_device "jacobi": DMMA Gram, block-Jacobi
                 sweeps of trb_setup.cu, DMMA back-multiplication), with the time of its
                 parts, the number of sweeps and the FP64 rate of the Jacobi kernels
  jacobi_direct  the same kernels on the rows of W itself (no Gram; what "auto" uses for
                 nearly square W)
  gram           round-1 baseline: W W^T, torch.linalg.eigh (cuSOLVER syevd), V = W^T U / s
  svd            round-1 baseline: torch.linalg.svd (cuSOLVER gesvd), on <= 2 instances

Every variant is checked against LAPACK-quality references: singular values against the
library's, orthonormality of both factors, ||Ut W - diag(s) Vt|| / ||W||.  Prints one JSON
line; nothing here is on the EP sweep.

    python tools/bench_setup.py --batch 16 --n 4096 --alpha 0.5
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tramp_b200 import ops  # noqa: E402
from tramp_b200.channels import linear_channel as lc  # noqa: E402


def sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def timed(fn, repeat):
    fn()                                     # warm-up: handles, workspaces
    sync()
    best = float("inf")
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = fn()
        sync()
        best = min(best, time.perf_counter() - t0)
    return best, out


def quality(W, Ut, s, Vt, s_ref=None):
    R = s.shape[-1]
    eye = torch.eye(R, dtype=W.dtype, device=W.device)
    q = {"residual": float(((Ut @ W - s[:, :, None] * Vt).norm() / W.norm()).item()),
         "orth_U": float((Ut @ Ut.transpose(1, 2) - eye).abs().max().item()),
         "orth_V": float((Vt @ Vt.transpose(1, 2) - eye).abs().max().item())}
    if s_ref is not None:
        q["s_rel_dev"] = float(((s - s_ref).abs() / s_ref).max().item())
    return q


def jacobi_parts(W, route):
    """The stages of jacobi_thin_svd timed one by one (CUDA events around each), plus the
    sweep history."""
    B, M, N = W.shape
    L = M if route == "gram" else N
    n_rows, ld = lc._ceil_to(M, ops.JACOBI_ROWS), lc._ceil_to(L, ops.JACOBI_COLS)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    A = torch.zeros((B, n_rows, ld), dtype=torch.float64, device=W.device)
    e0 = ev()
    if route == "gram":
        for b in range(B):
            lc._gram_rows(W[b], A[b])
    else:
        A[:, :M, :N] = W
    e1 = ev()
    work = ops.jacobi_workspace(B, n_rows, ld, W.device)
    sweeps = []
    marks = [e1]
    for _ in range(40):
        off = float(ops.jacobi_sweep(A, work, skip_tol=5e-15, max_inner=lc.JACOBI_INNER_SWEEPS).max().item())
        marks.append(ev())
        sweeps.append(off)
        if off < 1e-13 or (len(sweeps) > 1 and off < 1e-10 and off > 0.5 * sweeps[-2]):
            break
    sync()
    sweep_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
    pairs = (n_rows // 16) * (n_rows // 16 - 1) // 2
    flops_sweep = pairs * 2.0 * 32 * 32 * ld * (10.0 / 16 + 1.0) * B    # Gram (upper tiles) + rotation
    bytes_sweep = (n_rows // 16 - 1) * 3.0 * n_rows * ld * 8 * B        # rows: 2 reads + 1 write per round
    full = [ms for ms in sweep_ms[:-2]] or sweep_ms
    return {"gram_ms_per_instance": e0.elapsed_time(e1) / B,
            "sweeps": len(sweeps), "off": sweeps, "sweep_ms": sweep_ms,
            "jacobi_ms_per_instance": sum(sweep_ms) / B,
            "full_sweep_tflops": flops_sweep / (sum(full) / len(full) * 1e-3) / 1e12,
            "full_sweep_hbm_gbs": bytes_sweep / (sum(full) / len(full) * 1e-3) / 1e9,
            "zsplit": int(work["S"].shape[2])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--skip-svd", action="store_true")
    ap.add_argument("--skip-gram", action="store_true")
    ap.add_argument("--direct", action="store_true", help="also time the route on W itself")
    ap.add_argument("--waves", type=int, default=0, help="trb_jacobi_set_waves (0: library default)")
    ap.add_argument("--inner", type=int, default=0, help="sweeps of the inner 32 x 32 Jacobi (0: default)")
    ap.add_argument("--fused-mask", type=int, default=-1, help="trb_jacobi_set_fused (-1: library default 1)")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device"
    from tramp_b200 import _lib
    if args.waves:
        _lib.load().trb_jacobi_set_waves(args.waves)
    if args.inner:
        lc.JACOBI_INNER_SWEEPS = args.inner
    if args.fused_mask >= 0:
        _lib.load().trb_jacobi_set_fused(args.fused_mask)
    B, N = args.batch, args.n
    M = int(args.alpha * N)
    gen = torch.Generator(device="cuda").manual_seed(0)
    W = torch.randn((B, M, N), dtype=torch.float64, device="cuda", generator=gen) / N**0.5
    line = {"tool": "bench_setup", "B": B, "M": M, "N": N, "waves": args.waves, "inner": lc.JACOBI_INNER_SWEEPS,
            "fused_mask": args.fused_mask,
            "variants": {}}
    wide = W if M <= N else W.transpose(1, 2).contiguous()

    s_ref = None
    if not args.skip_gram:
        t_gram, ref = timed(lambda: lc.thin_svd_device(W, "gram"), args.repeat)
        s_ref = ref[1]
        line["variants"]["gram_cusolver_eigh"] = {"ms_per_instance": 1e3 * t_gram / B, **quality(W, *ref)}
        del ref
    t_j, out = timed(lambda: lc.thin_svd_device(W, "jacobi"), args.repeat)
    line["variants"]["jacobi"] = {"ms_per_instance": 1e3 * t_j / B, **quality(W, *out, s_ref=s_ref),
                                  "parts": jacobi_parts(wide, "gram")}
    if s_ref is None:
        s_ref = out[1]
    del out
    if args.direct:
        t_d, out = timed(lambda: lc.thin_svd_device(W, "jacobi_direct"), args.repeat)
        line["variants"]["jacobi_direct"] = {"ms_per_instance": 1e3 * t_d / B, **quality(W, *out, s_ref=s_ref),
                                             "parts": jacobi_parts(wide, "direct")}
        del out
    if not args.skip_svd:
        k = min(B, 2)
        t_svd, out = timed(lambda: lc.thin_svd_device(W[:k], "svd"), 1)
        line["variants"]["svd_cusolver"] = {"ms_per_instance": 1e3 * t_svd / k, **quality(W[:k], *out, s_ref=s_ref[:k])}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
