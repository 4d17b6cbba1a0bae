"""world_size-2 gloo test of the instance-sharding plumbing (runs on CPU)."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_instance_shard_covers_everything():
    from tramp_b200.distributed import instance_shard
    for B, G in ((4096, 8), (10, 3), (7, 8), (512, 1)):
        blocks = [instance_shard(B, r, G) for r in range(G)]
        assert blocks[0][0] == 0 and blocks[-1][1] == B
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(G - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tramp_b200.distributed import instance_shard, gather_records, max_over_ranks
    B, iters = 7, 3
    a, b = instance_shard(B, rank, world)
    # each rank "computes" the records of its own instances
    local = torch.arange(a, b, dtype=torch.float64)[None, :] + 100.0 * torch.arange(iters, dtype=torch.float64)[:, None]
    full = gather_records(local)
    slow = max_over_ranks(1.0 + rank, torch.device("cpu"))
    if rank == 0:
        np.save(out, np.concatenate([full.numpy().ravel(), [slow]]))
    dist.destroy_process_group()


def test_gather_records_two_ranks(tmp_path):
    out = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    expect = (np.arange(7)[None, :] + 100.0 * np.arange(3)[:, None]).ravel()
    assert np.array_equal(got[:-1], expect)
    assert got[-1] == 2.0


def _se_grid_worker(rank, world, port, out):
    """State-Evolution grid sharded over two ranks (gloo); the two SE kernels are
    emulated by the oracle (tests/_emulated_device.py), the sharding and the
    gather are the code under test."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import _emulated_device
    fake = _emulated_device.install(setattr)
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.experiments import run_state_evolution_grid
    kw = dict(prior_type="gauss_bernoulli", output_type="gaussian", prior_rho=0.3, output_var=1e-3)
    alphas = [0.2, 0.4, 0.6, 0.8, 1.0]
    grid = run_state_evolution_grid(["x"], [glm_state_evolution(alpha=a, **kw) for a in alphas],
                                    max_iter=100)
    assert fake.calls["trb_se_run"] == 1                 # one launch per rank for its block
    np.save(out % rank, np.array([[r[0]["v"], r[0]["n_iter"]] for r in grid]))
    dist.destroy_process_group()


def test_state_evolution_grid_sharded_over_two_ranks(tmp_path, monkeypatch):
    from tests import _emulated_device
    out = str(tmp_path / "grid_rank%d.npy")
    port = 29500 + ((os.getpid() + 7) % 500)
    mp.spawn(_se_grid_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    assert np.array_equal(r0, r1) and r0.shape == (5, 2)          # every rank holds the whole grid
    # single process, same emulated kernels
    _emulated_device.install(monkeypatch.setattr)
    from tramp_b200.models import glm_state_evolution
    from tramp_b200.experiments import run_state_evolution_grid
    kw = dict(prior_type="gauss_bernoulli", output_type="gaussian", prior_rho=0.3, output_var=1e-3)
    single = run_state_evolution_grid(["x"], [glm_state_evolution(alpha=a, **kw)
                                              for a in [0.2, 0.4, 0.6, 0.8, 1.0]], max_iter=100)
    assert np.array_equal(r0, np.array([[r[0]["v"], r[0]["n_iter"]] for r in single]))
    assert np.all(np.diff(r0[:, 0]) < 0)


def _ep_problem(B=5, N=40, M=24, seed=3):
    rng = np.random.RandomState(seed)
    W = rng.randn(B, M, N) / np.sqrt(N)
    x = rng.randn(B, N) * (rng.rand(B, N) < 0.2)
    y = np.einsum("bmn,bn->bm", W, x) + 0.1 * rng.randn(B, M)
    return W, x, y


def _ep_builders():
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.channels import LinearChannel
    from tramp_b200.variables import SISOVariable as V
    W, x, y = _ep_problem()

    def build_model(a, b):
        return (GaussBernoulliPrior(size=W.shape[2], rho=0.2, batch=b - a) @ V("x") @ LinearChannel(W[a:b])
                @ V("z") @ GaussianLikelihood(y=y[a:b], var=1e-2)).to_model()
    return build_model, (lambda a, b: {"x": x[a:b]}), W.shape[0]


def _ep_sharded_worker(rank, world, port, out):
    """run_ep_sharded over two ranks (gloo): each rank sweeps its own block of instances
    (emulated device), the results are all-gathered."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import _emulated_device
    fake = _emulated_device.install(setattr)
    from tramp_b200.experiments import run_ep_sharded
    build_model, x_true, B = _ep_builders()
    res = run_ep_sharded(build_model, B, x_true=x_true, max_iter=40)
    assert fake.calls["trb_sweep_run"] >= 1
    np.savez(out % rank, rx=res["r"]["x"], vx=res["v"]["x"], vz=res["v"]["z"], n_iter=res["n_iter"], mse=res["mse"])
    dist.destroy_process_group()


def test_ep_instances_sharded_over_two_ranks(tmp_path, monkeypatch):
    from tests import _emulated_device
    out = str(tmp_path / "ep_rank%d.npz")
    port = 29500 + ((os.getpid() + 13) % 500)
    mp.spawn(_ep_sharded_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    for k in r0.files:                                   # every rank holds every instance
        assert np.array_equal(r0[k], r1[k], equal_nan=True)
    assert r0["rx"].shape == (5, 40) and r0["vx"].shape == (5,) and r0["mse"].shape == (40, 5)
    # the same five instances in one process: a block of 5 instead of blocks of 3 and 2
    _emulated_device.install(monkeypatch.setattr)
    from tramp_b200.experiments import run_ep_sharded
    build_model, x_true, B = _ep_builders()
    single = run_ep_sharded(build_model, B, x_true=x_true, max_iter=40)
    assert np.array_equal(single["n_iter"], r0["n_iter"]) and len(set(r0["n_iter"].tolist())) > 1
    np.testing.assert_allclose(single["r"]["x"], r0["rx"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(single["v"]["z"], r0["vz"], rtol=1e-12)
    np.testing.assert_allclose(single["mse"], r0["mse"], rtol=1e-12, equal_nan=True)
    W, x, y = _ep_problem()
    # the returned estimate is the last recorded iterate -- or the one before it when
    # EarlyStoppingEP saw a divergence and restored the previous messages (callbacks.py:277-283)
    achieved = np.mean((r0["rx"] - x)**2, axis=1)
    for b, n in enumerate(r0["n_iter"]):
        assert np.isclose(r0["mse"][n - 1, b], achieved[b], rtol=1e-9) or \
            np.isclose(r0["mse"][n - 2, b], achieved[b], rtol=1e-9)
        assert np.all(np.isnan(r0["mse"][n:, b]))
