"""Multi-GPU plumbing: the EP path shards by INSTANCE -- every GPU
owns a contiguous block of independent teacher-student instances, its own
operators and messages, and runs the device sweep with no data-path
collective.  torch.distributed (NCCL over NVLink on the box, gloo in the CPU
tests) is used only to gather the per-iteration records at the end."""


def collective_device(group=None, fallback="cuda"):
    """Device the tensors of a collective must live on: NCCL moves device memory, gloo (CPU
    tests, or several ranks sharing ONE GPU, which NCCL refuses) wants host tensors for
    all_gather."""
    import torch.distributed as dist
    return "cpu" if dist.get_backend(group) == "gloo" else fallback


class PeerExchange:
    """The peer-memory exchange of a row-sharded operator (include/tramp_b200.h
    trb_comm_*): every rank's buffer is mapped into every other rank over NVLink
    (CUDA IPC); torch.distributed only carries the 64-byte handles, once."""

    def __init__(self, vec_doubles, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_RANKS:
            raise ValueError(f"a peer exchange spans at most {_lib.MAX_RANKS} GPUs of one node")
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        _lib.check(lib.trb_comm_create(self.rank, self.world, int(vec_doubles), C.byref(ptr), handle))
        self.ptr = ptr
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=collective_device(group))
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        blob = b"".join(bytes(t.cpu().tolist()) for t in every)
        _lib.check(lib.trb_comm_connect(self.ptr, blob))
        dist.barrier(group=group)           # nobody publishes before everyone is mapped
        self.timeout = torch.zeros(1, dtype=torch.int32, device="cuda")

    def all_reduce(self, tensor):
        """In-place sum over the ranks (same protocol as inside the sweep)."""
        from . import _lib
        assert tensor.is_contiguous() and tensor.dtype.itemsize == 8
        _lib.check(_lib.load().trb_comm_all_reduce(self.ptr, tensor.data_ptr(), tensor.numel(),
                                                   self.timeout.data_ptr(), _lib.current_stream()))
        return tensor

    def close(self):
        from . import _lib
        if self.ptr:
            _lib.load().trb_comm_destroy(self.ptr)
            self.ptr = None


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
    home = local.device
    local = local.to(collective_device(group, fallback=home))
    n_local = torch.tensor([local.shape[-1]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    pad = torch.zeros(local.shape[:-1] + (width,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[..., :n] for o, n in zip(out, sizes)], dim=-1).to(home)


def max_over_ranks(value, device, group=None):
    """Max of a host scalar over ranks (timing: a multi-GPU step takes as long as
    its slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = t.to(collective_device(group, fallback=device))
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
