"""Set-up of LinearChannel (the factorisation the reference does in its constructor,
channels/linear/linear_channel.py:8-15, 36-46), which DESIGN 11 names as the
end-to-end cost once the sweep runs at the HBM roofline: how fast can B matrices of
the north-star shape be brought into thin-SVD form on one B200?

Variants timed on the same Gaussian W [B, M, N] (all library calls; the point is to
find out what a hand-written batched eigensolver has to beat):

  svd            torch.linalg.svd on the batch                       (cuSOLVER gesvd/gesvdj)
  gram           W W^T, torch.linalg.eigh on the batch, V = W^T U / s (thin_svd_device "gram")
  gram_streams   the same, one matrix per task on S CUDA streams driven by S host threads
                 (syevd is latency-bound on one 2048 x 2048 matrix; independent matrices
                 can overlap)
  gram_parts     the three steps of "gram" timed separately (Gram DGEMM, eigh, back-multiply)
  jacobi_b       block_jacobi_svd with blocks of b rows: every step a batched GEMM or a batched
                 2b x 2b eigen-problem over all instances and block pairs (thin_svd_device "jacobi")

Every variant is checked against the first: singular values to 1e-10 relative and
||Ut W - diag(s) Vt|| / ||W||.  Prints one JSON line; nothing here is on the EP path.

    python tools/bench_setup.py --batch 16 --n 4096 --alpha 0.5 --streams 1 2 4 8 16
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tramp_b200.channels.linear_channel import thin_svd_device, block_jacobi_svd  # noqa: E402


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


def gram_one(W):
    """thin_svd_device(..., "gram") for one [M, N] matrix with M <= N."""
    G = W @ W.T
    ev, U = torch.linalg.eigh(G)
    ev, U = ev.flip(-1), U.flip(-1)
    s = ev.clamp_min(0).sqrt()
    Ut = U.T.contiguous()
    return Ut, s, (Ut @ W) / s[:, None]


def gram_streams(W, n_streams):
    B = W.shape[0]
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    out = [None] * B
    ready = torch.cuda.Event()
    ready.record()

    def work(k):
        with torch.cuda.stream(streams[k]):
            streams[k].wait_event(ready)
            for b in range(k, B, n_streams):
                out[b] = gram_one(W[b])
    with ThreadPoolExecutor(max_workers=n_streams) as pool:
        list(pool.map(work, range(n_streams)))
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    return tuple(torch.stack([o[i] for o in out]) for i in range(3))


def residual(W, Ut, s, Vt):
    return float(((Ut @ W - s[:, :, None] * Vt).norm() / W.norm()).item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--streams", type=int, nargs="*", default=[2, 4, 8, 16])
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--jacobi-blocks", type=int, nargs="*", default=[16, 32])
    ap.add_argument("--skip-svd", action="store_true")
    ap.add_argument("--device", default="cuda", help='"cpu" only smoke-tests the script (no stream variant)')
    args = ap.parse_args()
    dev = args.device
    assert dev == "cpu" or torch.cuda.is_available(), "needs a CUDA device"
    B, N = args.batch, args.n
    M = int(args.alpha * N)
    assert M <= N, "the stream variant is written for M <= N"
    gen = torch.Generator(device=dev).manual_seed(0)
    W = torch.randn((B, M, N), dtype=torch.float64, device=dev, generator=gen) / N**0.5
    line = {"tool": "bench_setup", "B": B, "M": M, "N": N, "variants": {}}

    t_gram, ref = timed(lambda: thin_svd_device(W, "gram"), args.repeat)
    line["variants"]["gram"] = {"s_per_instance": t_gram / B, "residual": residual(W, *ref)}

    # the three steps of "gram"
    t_g, G = timed(lambda: W @ W.transpose(1, 2), args.repeat)
    t_e, (ev, U) = timed(lambda: torch.linalg.eigh(G), args.repeat)
    Ut = U.flip(-1).transpose(1, 2).contiguous()
    s = ev.flip(-1).clamp_min(0).sqrt()
    t_b, _ = timed(lambda: (Ut @ W) / s[:, :, None], args.repeat)
    line["variants"]["gram_parts"] = {"gram_dgemm_s": t_g / B, "eigh_s": t_e / B, "back_multiply_s": t_b / B,
                                      "dgemm_tflops": 2.0 * M * M * N * B / t_g / 1e12}

    for S in (args.streams if dev != "cpu" else []):
        t_s, out = timed(lambda: gram_streams(W, S), args.repeat)
        line["variants"][f"gram_streams_{S}"] = {
            "s_per_instance": t_s / B, "residual": residual(W, *out),
            "s_rel_dev": float(((out[1] - ref[1]).abs() / ref[1]).max().item())}

    for b in args.jacobi_blocks:
        t_j, out = timed(lambda: block_jacobi_svd(W, block=b), 1)
        line["variants"][f"jacobi_{b}"] = {
            "s_per_instance": t_j / B, "residual": residual(W, *out),
            "s_rel_dev": float(((out[1] - ref[1]).abs() / ref[1]).max().item())}

    if not args.skip_svd:
        t_svd, out = timed(lambda: thin_svd_device(W[:min(B, 4)], "svd"), 1)
        line["variants"]["svd"] = {
            "s_per_instance": t_svd / min(B, 4), "residual": residual(W[:min(B, 4)], *out),
            "s_rel_dev": float(((out[1] - ref[1][:min(B, 4)]).abs() / ref[1][:min(B, 4)]).max().item())}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
