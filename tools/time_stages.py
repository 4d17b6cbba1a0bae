"""In-pipeline duration of every launch of a north-star iteration (CUDA events around each launch,
trb_profile_timeline), for the kernel choices of trb_sweep_run -- the ncu launch list times the
kernels one by one on a cold, idle GPU; this is the same list measured inside the running sweep.

    python tools/time_stages.py [--instances 512] [--n 4096] [--iters 25] [--steps 3]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=512)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--iters", type=int, default=25)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "time_stages.json"))
    args = ap.parse_args()
    import torch
    from tramp_b200 import _lib, synthetic
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, ConstantInit
    lib = _lib.load()
    N, B, iters = args.n, args.instances, args.iters
    M = int(args.alpha * N)
    data = synthetic.gaussian_glm_batch(B, N, M, 0.1, 1e-2, seed=1000, workers=min(16, os.cpu_count() or 8))
    linear = LinearChannel.from_factors(data["Ut"], data["s"], data["Vt"], Nx=M, Nz=N, rank=min(M, N))
    model = (GaussBernoulliPrior(size=N, rho=0.1, batch=B) @ V("x") @ linear @ V("z")
             @ GaussianLikelihood(y=data["y"], var=1e-2)).to_model()
    ep = ExpectationPropagation(model)
    ep.schedule = "general"
    st = ep._ensure_state()
    st["x_true"] = ep._vec_to_dev(data["x"], "x")
    rec = {k: torch.zeros((iters, B), dtype=torch.float64, device="cuda") for k in ("mse", "vx", "vz")}
    ep.configure_damping(None)
    sw = ep._descriptor(rec, iters, None)
    init = ConstantInit(a=0, b=0)

    def step():
        ep.init_message_dag(init)
        st["active"].fill_(1)
        ep._run(sw, 0, iters, True)

    report = {}
    cap = 64 * iters
    ms, kinds = (C.c_double * cap)(), (C.c_int * cap)()
    configs = [("rescale inside the projections + chunked x / z (default)", 1, 3),
               ("nine launches, one CTA per instance", 0, 0),
               ("rescale inside the projections only", 1, 0),
               ("chunked x / z only", 0, 3)]
    for rep in range(2):
        for name, fused, mask in configs:
            lib.trb_set_fused_rescale(fused)
            lib.trb_set_update_kernels(mask)
            step()
            torch.cuda.synchronize()
            rows, total = [], []
            for _ in range(args.steps):
                lib.trb_profile_reset(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                torch.cuda.synchronize()
                n = lib.trb_profile_timeline(ms, kinds, cap)
                lib.trb_profile_reset(0)
                t = np.array(ms[:n]) * 1e3
                k = np.array(kinds[:n])
                per = n // iters                           # the run ends with the restore launch(es)
                assert 0 < n - per * iters < per, (n, iters)
                rows.append(t[:per * iters].reshape(iters, per)[1:])
                kind_row = k[:per]
                total.append(e0.elapsed_time(e1) * 1e3 / iters)
            avg = np.concatenate(rows).mean(axis=0)
            entry = dict(launch_us=[round(float(v), 1) for v in avg], kinds=[int(v) for v in kind_row],
                         sum_operator_passes_us=round(float(avg[kind_row == 1].sum()), 1),
                         sum_update_kernels_us=round(float(avg[kind_row == 0].sum()), 1),
                         iteration_us=round(float(np.mean(total)), 1))
            entry["gaps_us"] = round(entry["iteration_us"] - entry["sum_operator_passes_us"]
                                     - entry["sum_update_kernels_us"], 1)
            report.setdefault(name, []).append(entry)
            print(name, json.dumps(entry), flush=True)
    lib.trb_set_fused_rescale(1)
    lib.trb_set_update_kernels(-1)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(workload=dict(instances=B, N=N, M=M, iters=iters, steps=args.steps), configs=report),
              open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
