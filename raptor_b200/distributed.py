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


def unique_id_bytes(uid):
    """all 128 bytes of an ncclUniqueId structure (reading its c_char array as a VALUE stops at the first NUL byte and the other ranks would join a different id)"""
    import ctypes
    return ctypes.string_at(ctypes.byref(uid), ctypes.sizeof(uid))


# ---- the same gather through the engine's C ABI (b200l2f_allgather_trajectories): NCCL bound by the engine at run time, communicator owned by the caller ----------
class NcclCommunicator:
    """an ncclComm_t for this process (one rank per GPU), created with the libnccl the process already carries (torch's): rank 0 draws the unique id, the
    ranks exchange it through the initialised torch.distributed group, every rank calls ncclCommInitRank.  Use: comm = NcclCommunicator(); comm.handle"""

    def __init__(self):
        import ctypes
        if not dist.is_initialized():
            raise RuntimeError("NcclCommunicator needs an initialised torch.distributed process group (for the unique-id exchange)")
        self._ctypes = ctypes
        self.lib = self._load()
        rank, world = dist.get_rank(), dist.get_world_size()

        class UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_char * 128)]
        uid = UniqueId()
        if rank == 0:
            self._check(self.lib.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
        box = [unique_id_bytes(uid) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctypes.memmove(ctypes.byref(uid), box[0].ljust(128, b"\0"), 128)
        self.handle = ctypes.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        self._check(self.lib.ncclCommInitRank(ctypes.byref(self.handle), world, uid, rank), "ncclCommInitRank")
        self.rank, self.world = rank, world

    def _load(self):
        ctypes = self._ctypes
        for name in (os.environ.get("B200L2F_NCCL_LIB"), "libnccl.so.2"):
            if not name:
                continue
            try:
                return ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
            except OSError:
                pass
        import glob
        import sys
        for base in sys.path:                       # torch's wheel: site-packages/nvidia/nccl/lib/libnccl.so.2
            hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2"))
            if hits:
                return ctypes.CDLL(hits[0], mode=ctypes.RTLD_GLOBAL)
        raise OSError("libnccl.so.2 not found")

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed with ncclResult_t %d" % (what, rc))

    def destroy(self):
        if self.handle:
            self.lib.ncclCommDestroy.argtypes = [self._ctypes.c_void_p]
            self.lib.ncclCommDestroy(self.handle)
            self.handle = None


def allgather_trajectories_native(env, comm, slab):
    """slab: this rank's [T, n_local, D] float32 CUDA tensor (every rank the same shape) -> [T, world * n_local, D], through b200l2f_allgather_trajectories on the
    engine's stream (ordered behind the kernel that wrote the slab)."""
    import ctypes
    T, n_local, D = slab.shape
    slab = slab.contiguous()
    out = torch.empty((comm.world * T, n_local, D), dtype=torch.float32, device=slab.device)
    ranks = ctypes.c_int32(0)
    env._check(env._lib.b200l2f_allgather_trajectories(env._h, comm.handle, slab.data_ptr(), out.data_ptr(), slab.numel(), ctypes.byref(ranks)))
    assert ranks.value == comm.world
    env.synchronize()   # the gather runs on the ENGINE's stream; the re-layout below is a torch kernel on torch's stream
    return out.view(comm.world, T, n_local, D).permute(1, 0, 2, 3).reshape(T, comm.world * n_local, D)
