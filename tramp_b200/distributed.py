"""Multi-GPU plumbing: the EP path shards by INSTANCE -- every GPU
owns a contiguous block of independent teacher-student instances, its own
operators and messages, and runs the device sweep with no data-path
collective.  torch.distributed (NCCL over NVLink on the box, gloo in the CPU
tests) is used only to gather the per-iteration records at the end."""


def instance_shard(n_instances, rank, world):
    """Contiguous block [start, stop) of the rank: instance i lives on GPU
    floor(i * world / n_instances)."""
    start = (rank * n_instances + world - 1) // world
    stop = ((rank + 1) * n_instances + world - 1) // world
    return start, stop


def gather_records(local, group=None):
    """All-gather per-instance records along the instance axis.

    local: tensor [..., B_local] (instances last); shards may differ in size by
    one.  Returns the concatenated tensor [..., B_total] on every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[-1]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    pad = torch.zeros(local.shape[:-1] + (width,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[..., :n] for o, n in zip(out, sizes)], dim=-1)


def max_over_ranks(value, device, group=None):
    """Max of a host scalar over ranks (timing: a multi-GPU step takes as long as
    its slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
