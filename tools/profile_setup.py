"""Workload for ncu: a few block-Jacobi rounds of the LinearChannel set-up at the north-star shape
(B instances of a 2048 x 2048 Gram matrix).  Run under
    ncu --set full --clock-control none --import-source on -k regex:k_jacobi -s 30 -c 6 ...
(the first rounds of a sweep are identical in cost to the later ones)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tramp_b200 import ops  # noqa: E402
from tramp_b200.channels import linear_channel as lc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--m", type=int, default=2048)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--sweeps", type=int, default=1)
args = ap.parse_args()
B, M, N = args.batch, args.m, args.n
gen = torch.Generator(device="cuda").manual_seed(0)
W = torch.randn((B, M, N), dtype=torch.float64, device="cuda", generator=gen) / N**0.5
n_rows, ld = lc._ceil_to(M, ops.JACOBI_ROWS), lc._ceil_to(M, ops.JACOBI_COLS)
A = torch.zeros((B, n_rows, ld), dtype=torch.float64, device="cuda")
for b in range(B):
    lc._gram_rows(W[b], A[b])
work = ops.jacobi_workspace(B, n_rows, ld, A.device)
for _ in range(args.sweeps):
    off = ops.jacobi_sweep(A, work, skip_tol=5e-15, max_inner=lc.JACOBI_INNER_SWEEPS)
torch.cuda.synchronize()
print("off", float(off.max()))
