"""Multi-GPU plumbing for the rollout engine: one process per GPU (torchrun), environments sharded contiguously by GLOBAL id, no collective
on the rollout path.  The only collective is the optional gather of trajectory slabs for a central learner (NCCL over NVLink on GPUs;
gloo on CPU tensors in the host-logic tests).

The reference has no distributed backend at all (SURVEY.md finding 8: parallelism = independent OS processes); this module is therefore
new surface, kept to the two things a learner needs: who owns which environments, and how to see everybody's trajectories."""
import os

import torch
import torch.distributed as dist


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(rank, world, n_global):
    """contiguous global env-id range [first, first + count) owned by `rank`; counts differ by at most one"""
    base, rem = divmod(n_global, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def init_process_group(backend=None):
    rank, local, world = rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def allgather_trajectories(slab, world=None, n_global=None):
    """slab: this rank's [T, n_local, D] trajectory tensor (device tensor written by rollout/collect) -> [T, n_global, D] on every rank,
    environments in global-id order (one all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed).
    n_global: total number of environments when the shards follow shard_range() -- the shard sizes are then known locally and the call is a single
    collective without a host synchronisation; otherwise the sizes are exchanged first."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return slab
    world = dist.get_world_size() if world is None else world
    T, n_local, D = slab.shape
    if n_global is not None:
        counts = [shard_range(r, world, n_global)[1] for r in range(world)]
        if counts[dist.get_rank()] != n_local:
            raise ValueError("allgather_trajectories: this rank holds %d environments, shard_range gives %d" % (n_local, counts[dist.get_rank()]))
    else:
        counts = [torch.zeros(1, dtype=torch.int64, device=slab.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=slab.device))
        counts = [int(c.item()) for c in counts]
    if len(set(counts)) == 1:
        out = torch.empty((world * T, n_local, D), dtype=slab.dtype, device=slab.device)   # rank-major concatenation along dim 0
        dist.all_gather_into_tensor(out, slab.contiguous())
        return out.view(world, T, n_local, D).permute(1, 0, 2, 3).reshape(T, world * n_local, D)
    cmax = max(counts)                              # ragged shards: pad to the largest shard, gather, trim
    padded = torch.zeros((T, cmax, D), dtype=slab.dtype, device=slab.device)
    padded[:, :n_local] = slab
    out = torch.empty((world * T, cmax, D), dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(out, padded)
    out = out.view(world, T, cmax, D)
    return torch.cat([out[r, :, :counts[r]] for r in range(world)], dim=1)
