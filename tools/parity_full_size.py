"""Full-size parity evidence for the BASELINE configs the test-suite only covers in miniature
(run on a GPU box; results -> gpurun_out/r02_parity_full_size.json, kept under profiles/):

  config3   phase retrieval with ONE W shared by the batch, N = 16384, alpha = 1: REAL Gaussian
            W = randn(N, N) / sqrt(N) factorised by the hand-written set-up (route "direct": square),
            BinaryPrior(p_pos=0.6) @ LinearChannel(W) @ AbsLikelihood, damping 0.3.  Instance 0 of the
            batch against the CPU oracle on the same W, y (full LAPACK SVD on the host), and the
            DMMA GEMM passes against the GEMV passes on the same operator for the whole batch.
  config4   single large instance, REAL Gaussian W at N = 8192, alpha = 0.6, GaussBernoulli / Gaussian
            likelihood: thin-SVD triplets dealt to 2 and 4 ranks (processes sharing the visible GPUs;
            peer-memory exchange of trb_comm.cu), every rank against the CPU oracle and against the
            unsharded sweep.

    python tools/parity_full_size.py [config3] [config4]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def rel_max(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))


def rel_each(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ----------------------------------------------------------------------------- config 3
def config3(N=16384, B=16, iters=20, damping=0.3):
    import torch
    from tramp_b200.priors import BinaryPrior
    from tramp_b200.likelihoods import AbsLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.channels import linear_channel as lc
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    from oracle import tramp_oracle as orc
    np.random.seed(31)
    W = np.random.randn(N, N) / np.sqrt(N)
    x = np.where(np.random.rand(B, N) < 0.6, 1.0, -1.0)
    y = np.abs(x @ W.T)
    t0 = time.perf_counter()
    lin = LinearChannel(W)
    lin._setup()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    res = dict(N=N, M=N, B=B, iters=iters, damping=damping, setup_s=setup_s, setup=dict(lc.LAST_SETUP_STATS))
    out = {}
    for backend in ("gemm", "gemv"):
        model = (BinaryPrior(size=N, p_pos=0.6, batch=B) @ V("x") @ lin @ V("z") @ AbsLikelihood(y=y)).to_model()
        ep = ExpectationPropagation(model)
        ep.linear_backend = backend
        track = TrackErrors({"x": x}, metrics=["mse", "sign_mse"])
        t0 = time.perf_counter()
        ep.iterate(max_iter=iters, callback=track, damping=damping)
        d = ep.get_variables_data()
        out[backend] = dict(rx=np.array(d["x"]["r"]), vx=np.array(d["x"]["v"]), rz=np.array(d["z"]["r"]),
                            vz=np.array(d["z"]["v"]), mse=np.array([e["mse"] for e in track.errors]),
                            s=time.perf_counter() - t0)
    a, b = out["gemm"], out["gemv"]
    res["dmma_vs_gemv"] = dict(instances=B, rx=rel_max(a["rx"], b["rx"]), rz=rel_max(a["rz"], b["rz"]),
                               vx=rel_each(a["vx"], b["vx"]), vz=rel_each(a["vz"], b["vz"]),
                               mse_trajectory=rel_each(a["mse"], b["mse"]))
    print("config3 device done", json.dumps(res["dmma_vs_gemv"]), flush=True)
    t0 = time.perf_counter()
    ref = orc.ep_glm(dict(kind="binary", p_pos=0.6), W, dict(kind="abs", y=y[0]), iters, damping=damping, x_true=x[0])
    res["oracle_s"] = time.perf_counter() - t0
    s_ref = np.linalg.svd(W, compute_uv=False)
    res["singular_values_max_rel_dev_vs_lapack"] = rel_each(lin.s[0].cpu().numpy(), s_ref)
    res["cond_W"] = float(s_ref[0] / s_ref[-1])
    res["vs_oracle_instance0"] = {
        k: dict(rx=rel_max(o["rx"][0], ref["r_x"]), rz=rel_max(o["rz"][0], ref["r_z"]),
                vx=rel_each(o["vx"][0], ref["v_x"]), vz=rel_each(o["vz"][0], ref["v_z"]),
                mse_trajectory=rel_each(o["mse"][:, 0], np.array(ref["traj"]["mse_x"])),
                mse_trajectory_on_signal_scale=float(np.max(np.abs(o["mse"][:, 0] - np.array(ref["traj"]["mse_x"])))))
        for k, o in out.items()}
    res["mse_first_last"] = [float(a["mse"][0, 0]), float(a["mse"][-1, 0])]
    return res


# ----------------------------------------------------------------------------- config 4
def _c4_problem(N, alpha, seed=41):
    rng = np.random.RandomState(seed)
    M = int(alpha * N)
    W = rng.randn(M, N) / np.sqrt(N)
    x = rng.standard_normal(N) * rng.binomial(n=1, size=N, p=0.1)
    y = W @ x + 0.1 * rng.standard_normal(M)
    return W, x, y


def _c4_worker(rank, world, port, N, alpha, iters, fac_path, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ngpu = torch.cuda.device_count()
    if ngpu >= world:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        torch.cuda.set_device(rank % ngpu)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from tramp_b200 import ops
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    W, x, y = _c4_problem(N, alpha)
    M = W.shape[0]
    fac = np.load(fac_path)
    idx = np.arange(rank, fac["s"].size, world)              # triplets dealt round-robin
    lin = LinearChannel.from_sharded_factors(
        ops.padded(fac["Ut"][idx])[None].contiguous(), ops.to_dev(fac["s"][None, idx]),
        ops.padded(fac["Vt"][idx])[None].contiguous(), fac["s"], Nx=M, Nz=N, group=dist.group.WORLD)
    model = (GaussBernoulliPrior(size=N, rho=0.1) @ V("x") @ lin @ V("z") @ GaussianLikelihood(y=y, var=1e-2)).to_model()
    ep = ExpectationPropagation(model)
    ep.schedule = "general"
    track = TrackErrors({"x": x})
    ep.iterate(max_iter=iters, callback=track)
    d = ep.get_variables_data()
    np.savez(out_path % rank, rx=d["x"]["r"], rz=d["z"]["r"], vx=d["x"]["v"], vz=d["z"]["v"],
             mse=np.array([e["mse"] for e in track.errors]), timeout=int(lin.exchange.timeout.item()))
    dist.destroy_process_group()


def config4(N=8192, alpha=0.6, iters=30, worlds=(2, 4)):
    import tempfile
    import torch
    import torch.multiprocessing as mp
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    from oracle import tramp_oracle as orc
    W, x, y = _c4_problem(N, alpha)
    M = W.shape[0]
    t0 = time.perf_counter()
    lin = LinearChannel(W)                                    # hand-written set-up
    lin._setup()
    torch.cuda.synchronize()
    res = dict(N=N, M=M, iters=iters, setup_s=time.perf_counter() - t0, n_gpus_visible=torch.cuda.device_count())
    model = (GaussBernoulliPrior(size=N, rho=0.1) @ V("x") @ lin @ V("z") @ GaussianLikelihood(y=y, var=1e-2)).to_model()
    ep = ExpectationPropagation(model)
    ep.schedule = "general"
    track = TrackErrors({"x": x})
    ep.iterate(max_iter=iters, callback=track)
    d = ep.get_variables_data()
    one = dict(rx=np.array(d["x"]["r"]), rz=np.array(d["z"]["r"]), vx=float(d["x"]["v"]), vz=float(d["z"]["v"]),
               mse=np.array([float(e["mse"]) for e in track.errors]))
    tmp = tempfile.mkdtemp(prefix="trb_c4_")
    fac_path = os.path.join(tmp, "factors.npz")
    np.savez(fac_path, Ut=lin.Ut[0, :, :M].cpu().numpy(), s=lin.s[0].cpu().numpy(), Vt=lin.Vt[0, :, :N].cpu().numpy())
    del ep, model, lin
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    ref = orc.ep_glm(dict(kind="gauss_bernoulli", rho=0.1), W, dict(kind="gaussian", var=1e-2, y=y), iters, x_true=x)
    res["oracle_s"] = time.perf_counter() - t0
    ref_mse = np.array(ref["traj"]["mse_x"])

    def versus(o, r_x, r_z, v_x, v_z, mse):
        return dict(rx=rel_max(o["rx"], r_x), rz=rel_max(o["rz"], r_z), vx=rel_each(o["vx"], v_x),
                    vz=rel_each(o["vz"], v_z), mse_trajectory=rel_each(o["mse"], mse))
    res["unsharded_vs_oracle"] = versus(one, ref["r_x"], ref["r_z"], ref["v_x"], ref["v_z"], ref_mse)
    for world in worlds:
        out = os.path.join(tmp, f"w{world}_rank%d.npz")
        mp.spawn(_c4_worker, args=(world, 29700 + world, N, alpha, iters, fac_path, out), nprocs=world, join=True)
        ranks = [dict(np.load(out % r)) for r in range(world)]
        res[f"sharded_{world}_ways"] = dict(
            vs_oracle=versus(ranks[0], ref["r_x"], ref["r_z"], ref["v_x"], ref["v_z"], ref_mse),
            vs_unsharded=versus(ranks[0], one["rx"], one["rz"], one["vx"], one["vz"], one["mse"]),
            ranks_bit_identical=bool(all(np.array_equal(ranks[0]["rx"], r["rx"]) and np.array_equal(ranks[0]["mse"], r["mse"])
                                         for r in ranks[1:])),
            exchange_timeouts=int(sum(int(r["timeout"]) for r in ranks)))
        print(f"config4 {world} ways", json.dumps(res[f"sharded_{world}_ways"]), flush=True)
    return res


if __name__ == "__main__":
    which = sys.argv[1:] or ["config3", "config4"]
    report = {}
    path = os.path.join(ROOT, "gpurun_out", "r02_parity_full_size.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    for name in which:
        t0 = time.perf_counter()
        report[name] = dict(config3=config3, config4=config4)[name]()
        report[name]["wall_s"] = time.perf_counter() - t0
        json.dump(report, open(path, "w"), indent=1)
        print(name, json.dumps(report[name])[:3000], flush=True)
