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
