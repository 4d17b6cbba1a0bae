"""Measurement blocks that bench.py appends to its JSON line (and that the tools/ scripts run
on their own):

  setup_block        SURVEY 8d inputs: real W = randn(M, N) / sqrt(N) per instance, factorised
                     through the public constructor `LinearChannel(W)` (the hand-written block
                     Jacobi set-up of trb_setup.cu), then swept: set-up time per instance, the
                     end-to-end rate WITH the set-up counted, and parity of the EP result with the
                     oracle on the same W, y (host processes).
  row_sharded_block  BASELINE configs[4]: one instance N = 65536, alpha = 0.6, the thin-SVD operators
                     row-sharded over the ranks, against the same sweep unsharded.
  shared_w_block     BASELINE configs[3]: 1024 instances sharing one W, N = 16384 (FP64 tensor-core
                     GEMMs), the right-hand sides sharded over the ranks.

Nothing here is imported by the package; oracle/ is used as the checker only.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RHO, NOISE_VAR = 0.1, 1e-2


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def reference_instance(np, N, M, seed):
    """One teacher-student instance drawn as the reference does (SURVEY 8d:
    gaussian_ensemble.py:19-20, gauss_bernoulli_prior.py:38-42, gaussian_channel.py:12-15)."""
    rng = np.random.RandomState(seed)
    W = rng.randn(M, N) / np.sqrt(N)
    x = rng.standard_normal(N) * rng.binomial(n=1, size=N, p=RHO)
    y = W @ x + np.sqrt(NOISE_VAR) * rng.standard_normal(M)
    return W, x, y


# ------------------------------------------------------------------ oracle in host processes
def _oracle_worker(job):
    import numpy as np
    from oracle import tramp_oracle as orc
    N, M, seed, iters = job
    W, x, y = reference_instance(np, N, M, seed)
    t0 = time.perf_counter()
    op = orc.LinearOp(W)                        # matrix_rank + full SVD (linear_channel.py:36-46)
    t1 = time.perf_counter()
    out = orc.ep_glm(dict(kind="gauss_bernoulli", rho=RHO), W, dict(kind="gaussian", var=NOISE_VAR, y=y), iters,
                     x_true=x, op=op)
    t2 = time.perf_counter()
    return dict(r_x=out["r_x"], v_x=out["v_x"], mse=np.array(out["traj"]["mse_x"]), setup_s=t1 - t0, sweep_s=t2 - t1)


def oracle_runs(N, M, seeds, iters, threads_per_proc=2):
    """The CPU oracle on independent instances, one process each (all at once)."""
    import multiprocessing as mp
    saved = {k: os.environ.get(k) for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = str(threads_per_proc)   # inherited by the spawned interpreters
    try:
        with mp.get_context("spawn").Pool(len(seeds)) as pool:
            return pool.map(_oracle_worker, [(N, M, s, iters) for s in seeds])
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ------------------------------------------------------------------ set-up with real W
def setup_block(N, M, K, iters, seed0, parity_instances=8, rank=0, world=1, svd_method="auto"):
    """K real Gaussian matrices through LinearChannel(W) + EP(iters); see the module docstring."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.channels import linear_channel as lc
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors, TrackEvolution, JoinCallback
    from tramp_b200 import _lib
    seeds = [seed0 + rank * K + i for i in range(K)]
    inst = [reference_instance(np, N, M, s) for s in seeds]
    W_host = torch.from_numpy(np.stack([w for w, _, _ in inst])).pin_memory()
    x = np.stack([xx for _, xx, _ in inst])
    y_host = torch.from_numpy(np.stack([yy for _, _, yy in inst])).pin_memory()
    x_host = torch.from_numpy(x).pin_memory()

    def run(W_pinned, y_pinned, x_pinned):
        B = W_pinned.shape[0]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        W_dev = W_pinned.to("cuda", non_blocking=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        lin = LinearChannel(W_dev, svd_method=svd_method, keep_W=False)
        lin._setup()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        yv, xv = y_pinned.to("cuda", non_blocking=True), x_pinned.to("cuda", non_blocking=True)
        model = (GaussBernoulliPrior(size=N, rho=RHO, batch=B) @ V("x") @ lin @ V("z")
                 @ GaussianLikelihood(y=yv, var=NOISE_VAR)).to_model()
        ep = ExpectationPropagation(model)
        ep.schedule = "general"
        track = TrackErrors({"x": xv})
        ep.iterate(max_iter=iters, callback=JoinCallback([track, TrackEvolution(ids=["x"])]))
        got = ep.get_variables_data(["x"])
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return dict(h2d=t1 - t0, factor=t2 - t1, sweep=t3 - t2, total=t3 - t0), got, track, lin

    run(W_host[:2], y_host[:2], x_host[:2])                   # warm-up: attributes, allocator
    if world > 1:
        dist.barrier()
    launches0 = int(_lib.load().trb_profile_launches(2))
    tm, got, track, lin = run(W_host, y_host, x_host)
    launches = int(_lib.load().trb_profile_launches(2)) - launches0
    times = torch.tensor([tm["h2d"], tm["factor"], tm["sweep"], tm["total"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    h2d, factor, sweep, total = times.tolist()
    stats = dict(lc.LAST_SETUP_STATS)
    # FP64 work of the Jacobi sweeps: every pair of 16-row blocks once per sweep, Gram (10 of 16
    # tiles) + rotation of 32 rows of length ld; plus the Gram and back-multiplication GEMMs
    R = min(M, N)
    n_rows, ld = -(-R // 32) * 32, -(-R // 64) * 64
    pairs = (n_rows // 16) * (n_rows // 16 - 1) // 2
    flops = stats.get("sweeps", 0) * pairs * 2.0 * 32 * 32 * ld * (10.0 / 16 + 1.0) + 4.0 * R * R * max(M, N)
    block = {
        "what": (f"{K} instances per GPU with REAL W = randn(M, N) / sqrt(N) (SURVEY 8d seeds {seeds[0]}..), "
                 "pinned host W -> LinearChannel(W) (hand-written block-Jacobi set-up, trb_setup.cu) -> "
                 f"EP {iters} iterations through the public API; max over ranks"),
        "instances_per_gpu": K, "svd_method": svd_method, "route": stats.get("route"),
        "jacobi_sweeps": stats.get("sweeps"), "setup_kernel_launches": launches,
        "ms_per_instance": 1e3 * (h2d + factor) / K, "h2d_ms_per_instance": 1e3 * h2d / K,
        "factorisation_ms_per_instance": 1e3 * factor / K, "sweep_ms_per_instance": 1e3 * sweep / K,
        "fp64_tflops": flops * K / factor / 1e12,
        "fp64_frac_of_dmma_peak": flops * K / factor / 1e12 / 36.9,
        "dmma_peak_tflops": 36.9,
        "e2e_incl_setup": {"value": world * K * iters / total, "unit": "instance-iterations/s",
                           "h2d_bytes": int(W_host.numel() + y_host.numel() + x_host.numel()) * 8},
    }
    if rank == 0:
        # the library route it replaces, on the same matrices (round-1 baseline: W W^T, cuSOLVER syevd
        # through torch.linalg.eigh, V = W^T U / s), warm
        k_lib = min(4, K)
        W_lib = W_host[:k_lib].to("cuda")
        lc.thin_svd_device(W_lib[:1], "gram")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lc.thin_svd_device(W_lib, "gram")
        torch.cuda.synchronize()
        block["library_gram_eigh_ms_per_instance"] = 1e3 * (time.perf_counter() - t0) / k_lib
        del W_lib
    if rank == 0 and parity_instances > 0:
        P = min(parity_instances, K)
        t0 = time.perf_counter()
        refs = oracle_runs(N, M, seeds[:P], iters)
        wall = time.perf_counter() - t0
        worst = 0.0
        mse_dev = np.array([[e["mse"][b] for b in range(P)] for e in track.errors])
        s_dev = 0.0
        for b, ref in enumerate(refs):
            worst = max(worst, float(np.max(np.abs(got["x"]["r"][b] - ref["r_x"])) / np.max(np.abs(ref["r_x"]))),
                        float(abs(got["x"]["v"][b] - ref["v_x"]) / ref["v_x"]),
                        float(np.max(np.abs(mse_dev[:, b] - ref["mse"]) / ref["mse"])))
        s_ref = np.linalg.svd(inst[0][0], compute_uv=False)
        s_dev = float(np.max(np.abs(lin.s[0].cpu().numpy() - s_ref) / s_ref))
        block["parity_vs_oracle"] = {"instances": P, "iterations": iters, "max_rel_dev_r_v_mse": worst, "tol": 1e-9,
                                     "singular_values_max_rel_dev_vs_lapack": s_dev, "oracle_wall_s": wall}
        cpu_setup = float(np.mean([r["setup_s"] for r in refs]))
        cpu_sweep = float(np.mean([r["sweep_s"] for r in refs]))
        block["cpu_port_incl_setup"] = {
            "value": P * iters / max(r["setup_s"] + r["sweep_s"] for r in refs), "unit": "instance-iterations/s",
            "sample": (f"{P} of these instances at once, one process and 2 BLAS threads each: matrix_rank + full SVD "
                       f"{cpu_setup:.1f} s + {iters} iterations {cpu_sweep:.1f} s per instance"),
            "cores": 2 * P}
    return block


# ------------------------------------------------------------------ config 5: row-sharded instance
def _block_rows(torch, synthetic, Rg, n, ld, shard, world, seed):
    """[1, Rg, ld] with orthonormal Haar rows supported on shard `shard`'s column block."""
    from tramp_b200.distributed import instance_shard
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    c0, c1 = instance_shard(n, shard, world)
    assert Rg <= c1 - c0, "shard has more rows than its column block"
    out = torch.zeros((1, Rg, ld), dtype=torch.float64, device="cuda")
    out[0, :, c0:c1] = synthetic.haar_rows(1, Rg, c1 - c0, gen, chunk=1, ld=c1 - c0)[0]
    return out


def row_sharded_block(N=65536, alpha=0.6, iters=20, rank=0, world=1, compare_unsharded=True, backends=("sharded",)):
    """One instance whose singular triplets are dealt round-robin to the ranks; the two expansions
    per iteration are summed over the ranks through peer memory inside the update kernels.
    Synthetic operator (a dense 39321 x 65536 FP64 SVD is not a benchmark set-up): rank g's singular
    vectors are Haar-distributed on a column block of their own, stored at FULL length, Gaussian-
    ensemble spectrum -- not a Gaussian W.  Rank 0 then rebuilds every shard (same seeds) and runs
    the same sweep unsharded."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tramp_b200 import synthetic, ops
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    M = int(alpha * N)
    R = M
    ldn, ldm = ops.pad_ld(N), ops.pad_ld(M)
    t0 = time.time()
    idx = np.arange(rank, R, world)

    def shard(g):
        Rg = np.arange(g, R, world).size
        return (_block_rows(torch, synthetic, Rg, M, ldm, g, world, 300 + g),
                _block_rows(torch, synthetic, Rg, N, ldn, g, world, 100 + g))
    Ut, Vt = shard(rank)
    s_full = synthetic.gaussian_singular_values(1, M, N, seed=5, workers=1)[0]     # same on every rank
    s_loc = torch.as_tensor(s_full[idx].copy(), device="cuda")[None]
    g2 = torch.Generator(device="cuda")
    g2.manual_seed(7)                                                               # same on every rank
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g2)
    x = x * (torch.rand(N, dtype=torch.float64, device="cuda", generator=g2) < RHO)
    z = Ut[0, :, :M].T @ (s_loc[0] * (Vt[0, :, :N] @ x))
    if world > 1:
        dist.all_reduce(z)
    y = z + np.sqrt(NOISE_VAR) * torch.randn(M, dtype=torch.float64, device="cuda", generator=g2)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    group = dist.group.WORLD if world > 1 else None
    if world > 1:
        lin = LinearChannel.from_sharded_factors(Ut, s_loc, Vt, s_full, Nx=M, Nz=N, group=group)
    else:
        lin = LinearChannel.from_factors(Ut, s_loc, Vt, Nx=M, Nz=N, rank=R)

    def model_of(linear):
        return (GaussBernoulliPrior(size=N, rho=RHO) @ V("x") @ linear @ V("z")
                @ GaussianLikelihood(y=y, var=NOISE_VAR)).to_model()
    bytes_it = 16.0 * R * (N + M)
    peak, peak_src = hbm_peak()
    res = dict(config="single large instance (BASELINE configs[4]), thin-SVD operators row-sharded over the ranks",
               N=N, M=M, R=R, n_gpus=world, iters=iters, algorithmic_GB_per_iter=bytes_it / 1e9, setup_s=setup_s,
               schedule="general 4-pass", hbm_peak_gbs=peak, peak_source=peak_src,
               operator="synthetic block-orthogonal singular vectors, Gaussian-ensemble spectrum (not a Gaussian W)")
    final = None
    for backend in (backends if world > 1 else ("gemv",)):
        ep = ExpectationPropagation(model_of(lin))
        if world > 1:
            ep.linear_backend = backend
        ep.schedule = "general"
        track = TrackErrors({"x": x})
        ep.iterate(max_iter=3, callback=track)        # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ep.iterate(max_iter=iters, callback=track)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        mse = [float(e["mse"]) for e in track.errors]
        res[backend] = dict(
            exchange={"sharded": "peer memory inside the update kernels (trb_comm.cu)",
                      "sharded_nccl": "NCCL all-reduce between kernels, staged from Python",
                      "gemv": "none (one GPU)"}[backend],
            ms_per_iter=ms / iters, iterations_per_s=iters / (ms / 1e3),
            GBps_per_gpu=bytes_it / world / (ms / iters / 1e3) / 1e9,
            frac_of_hbm_peak=bytes_it / world / (ms / iters / 1e3) / 1e9 / peak,
            mse_first=mse[0], mse_last=mse[-1])
        if final is None:
            d = ep.get_variables_data()
            final = dict(rx=np.array(d["x"]["r"]), vx=float(d["x"]["v"]), rz=np.array(d["z"]["r"]), mse=np.array(mse))
        del ep
    res["mse_signal"] = float((x**2).mean().item())
    if world > 1 and compare_unsharded:
        # rank 0: the same operator in one piece, the same sweep (the ranks deal the triplets
        # round-robin: triplet j of shard g is global triplet g + j * world)
        del lin, Ut, Vt
        torch.cuda.empty_cache()
        dev = None
        if rank == 0:
            Ut_f = torch.zeros((1, R, ldm), dtype=torch.float64, device="cuda")
            Vt_f = torch.zeros((1, R, ldn), dtype=torch.float64, device="cuda")
            for g in range(world):
                u, v = shard(g)
                Ut_f[0, g::world] = u[0]
                Vt_f[0, g::world] = v[0]
                del u, v
            lin1 = LinearChannel.from_factors(Ut_f, torch.as_tensor(s_full, device="cuda")[None], Vt_f, Nx=M, Nz=N, rank=R)
            ep = ExpectationPropagation(model_of(lin1))
            ep.schedule = "general"
            track = TrackErrors({"x": x})
            ep.iterate(max_iter=3, callback=track)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ep.iterate(max_iter=iters, callback=track)
            e1.record()
            torch.cuda.synchronize()
            d = ep.get_variables_data()
            mse1 = np.array([float(e["mse"]) for e in track.errors])
            dev = max(float(np.max(np.abs(final["rx"] - d["x"]["r"])) / np.max(np.abs(d["x"]["r"]))),
                      float(np.max(np.abs(final["rz"] - d["z"]["r"])) / np.max(np.abs(d["z"]["r"]))),
                      abs(final["vx"] - float(d["x"]["v"])) / float(d["x"]["v"]),
                      float(np.max(np.abs(final["mse"] - mse1) / mse1)))
            res["unsharded_on_rank0"] = dict(ms_per_iter=e0.elapsed_time(e1) / iters,
                                             frac_of_hbm_peak=bytes_it / (e0.elapsed_time(e1) / iters / 1e3) / 1e9 / peak)
            res["max_rel_dev_sharded_vs_unsharded"] = dev
            res["speedup_vs_one_gpu"] = res["unsharded_on_rank0"]["ms_per_iter"] / res[backends[0]]["ms_per_iter"]
            del ep, lin1, Ut_f, Vt_f
        dist.barrier()
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ config 3: shared W
def shared_w_block(N=16384, B_total=1024, iters=10, rank=0, world=1, check_instances=4):
    """Phase retrieval with ONE W for the whole batch: BinaryPrior(p_pos=0.6) @ LinearChannel(N, alpha=1)
    @ AbsLikelihood, damping 0.3.  W's factors are replicated on every rank (drawn from the same seed),
    the B_total right-hand sides are sharded: strong scaling of a fixed batch.  The operator
    passes are FP64 tensor-core GEMMs (k_dgemm_dmma_tma); a few instances are re-run through the
    HBM-bound GEMV kernels on the same operator and compared."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tramp_b200 import synthetic
    from tramp_b200.priors import BinaryPrior
    from tramp_b200.likelihoods import AbsLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    from tramp_b200.distributed import instance_shard
    M = N
    b0, b1 = instance_shard(B_total, rank, world)
    B = b1 - b0
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    t0 = time.time()
    Vt = synthetic.haar_rows(1, N, N, gen, chunk=1)
    Ut = synthetic.haar_rows(1, M, M, gen, chunk=1)
    s = torch.as_tensor(synthetic.gaussian_singular_values(1, M, N, 4, workers=1), device="cuda")
    x_all = torch.where(torch.rand((B_total, N), device="cuda", generator=gen, dtype=torch.float64) < 0.6, 1.0, -1.0)
    x = x_all[b0:b1].to(torch.float64).contiguous()
    del x_all
    y = (((x @ Vt[0, :, :N].T) * s) @ Ut[0, :, :M]).abs()
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    lin = LinearChannel.from_factors(Ut, s, Vt, Nx=M, Nz=N, rank=N)

    def run(backend, xs, ys, n_it, timed=True):
        model = (BinaryPrior(size=N, p_pos=0.6, batch=xs.shape[0]) @ V("x") @ lin @ V("z") @ AbsLikelihood(y=ys)).to_model()
        ep = ExpectationPropagation(model)
        ep.linear_backend = backend
        track = TrackErrors({"x": xs}, metrics=["sign_mse"])
        ep.iterate(max_iter=2, callback=track, damping=0.3)          # warm-up
        if world > 1 and timed:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ep.iterate(max_iter=n_it, callback=track, damping=0.3)
        e1.record()
        torch.cuda.synchronize()
        d = ep.get_variables_data(["x"])
        smse = np.array([e["sign_mse"] for e in track.errors])
        return e0.elapsed_time(e1), d, smse

    ms, d_gemm, smse = run("gemm", x, y, iters)
    mm = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mm, op=dist.ReduceOp.MAX)
    ms = float(mm.item())
    flops = 4.0 * N * (N + M) * B_total * iters
    res = dict(config="phase retrieval, one W shared by the batch (BASELINE configs[3]); alpha = 1, BinaryPrior(p_pos=0.6), "
                      "AbsLikelihood, damping 0.3; W replicated, right-hand sides sharded over the ranks",
               N=N, M=M, instances=B_total, instances_per_gpu=B, n_gpus=world, iters=iters, setup_s=setup_s,
               kernel="k_dgemm_dmma_tma (FP64 tensor cores, DMMA m8n8k4; 2-D TMA ring)",
               ms_per_iter=ms / iters, value=B_total * iters / (ms / 1e3), unit="instance-iterations/s",
               tflops_total=flops / (ms / 1e3) / 1e12, tflops_per_gpu=flops / world / (ms / 1e3) / 1e12,
               dmma_peak_tflops=36.9, frac_of_dmma_peak_per_gpu=flops / world / (ms / 1e3) / 1e12 / 36.9,
               sign_mse_first=float(smse[0].mean()), sign_mse_last=float(smse[-1].mean()))
    if check_instances and B >= check_instances:
        k = check_instances
        _, d_a, sm_a = run("gemm", x[:k].contiguous(), y[:k].contiguous(), iters, timed=False)
        _, d_b, sm_b = run("gemv", x[:k].contiguous(), y[:k].contiguous(), iters, timed=False)
        ra, rb = np.asarray(d_a["x"]["r"]), np.asarray(d_b["x"]["r"])
        dev = max(float(np.max(np.abs(ra - rb)) / np.max(np.abs(rb))),
                  float(np.max(np.abs(np.asarray(d_a["x"]["v"]) - np.asarray(d_b["x"]["v"])) / np.asarray(d_b["x"]["v"]))),
                  float(np.max(np.abs(sm_a - sm_b) / np.maximum(sm_b, 1e-300))))
        full = np.asarray(d_gemm["x"]["r"])[:k]
        res["max_rel_dev_dmma_vs_gemv"] = dev
        res["max_rel_dev_batch_vs_subbatch"] = float(np.max(np.abs(full - ra)) / np.max(np.abs(ra)))
        res["checked_instances"] = k
    del lin, Ut, Vt
    torch.cuda.empty_cache()
    return res
