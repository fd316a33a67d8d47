#!/usr/bin/env python
"""tools/check_sharded_allgather.py -- run under torchrun on N GPUs (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_sharded_allgather.py
Every rank rolls out ITS shard of a global set of environments (RNG streams and initial conditions keyed by global id, no collective on the rollout
path), the trajectory slabs are all-gathered over NCCL (raptor_b200.distributed.allgather_trajectories: the optional learner feed), and rank 0
compares the result with ONE handle that owns all environments: identical bits for any number of GPUs.  Also times the all-gather."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
from raptor_b200.distributed import NcclCommunicator, allgather_trajectories, allgather_trajectories_native, init_process_group, shard_range  # noqa: E402

DR = [1.5, 5.0, 40, 1200, 0.02, 5.0, 0.1, 0.03, 0.10, 0.03, 0.30, 0.005, 0.05, 0.0, 0.3]


def make_env(n, first, local):
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, device=local, first_env_id=first)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(seed=77, warmup=16)
    env.sample_initial_parameters(); env.sample_initial_state(); env.load_policy()
    return env


def main():
    rank, local, world = init_process_group("nccl")
    dev = torch.device("cuda", local)
    for n_global, T in ((4096 + 37, 50), (65536 * world, 64)):          # a ragged split and a large one
        first, cnt = shard_range(rank, world, n_global)
        env = make_env(cnt, first, local)
        slab = torch.zeros((T, cnt, 4), dtype=torch.float32, device=dev)
        rew = torch.zeros((T, cnt), dtype=torch.float32, device=dev)
        env.rollout(T, out={"actions": slab, "rewards": rew})
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        full = allgather_trajectories(slab)                              # warm-up (NCCL connection setup)
        dist.barrier(); torch.cuda.synchronize(dev)
        e0.record(); full = allgather_trajectories(slab, n_global=n_global); e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        full_r = allgather_trajectories(rew[:, :, None])[:, :, 0]
        assert full.shape == (T, n_global, 4)
        if rank == 0:
            one = make_env(n_global, 0, local)
            ref_a = torch.zeros((T, n_global, 4), dtype=torch.float32, device=dev)
            ref_r = torch.zeros((T, n_global), dtype=torch.float32, device=dev)
            one.rollout(T, out={"actions": ref_a, "rewards": ref_r})
            torch.cuda.synchronize(dev)
            same = bool(torch.equal(full, ref_a)) and bool(torch.equal(full_r, ref_r))
            gb = full.numel() * 4 / 1e9
            print("n_global=%d T=%d world=%d: all-gathered trajectories %s the single-handle rollout; all_gather of %.3f GB in %.3f ms (%.1f GB/s received per GPU)"
                  % (n_global, T, world, "EQUAL (bit-exact)" if same else "DIFFER from", gb, ms, gb * (world - 1) / world / ms * 1e3), flush=True)
            assert same
        dist.barrier()
    # ---- the same gather through the C ABI (b200l2f_allgather_trajectories: NCCL bound by the engine, enqueued on the engine's stream), timed on a
    # ---- config-4-sized slab: 262 144 envs x 32 steps x 37 floats = 1.24 GB per GPU
    comm = NcclCommunicator()
    n, T, D = 262144, 32, 37
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, device=local, first_env_id=rank * n)
    slab = torch.full((T, n, D), float(rank + 1), dtype=torch.float32, device=dev)
    for it in range(3):
        torch.cuda.synchronize(dev); dist.barrier()
        t0 = time.perf_counter()
        full = allgather_trajectories_native(env, comm, slab)
        env.synchronize()
        dt = time.perf_counter() - t0
    ok = all(bool((full[:, r * n:(r + 1) * n] == float(r + 1)).all()) for r in range(world))
    ref = allgather_trajectories(slab, n_global=n * world)
    ok = ok and bool(torch.equal(ref, full))
    t = torch.tensor([dt], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = slab.numel() * 4 / 1e9
        print("b200l2f_allgather_trajectories: %d ranks x %.2f GB slabs in %.2f ms (max over ranks) = %.1f GB/s received per GPU; contents %s"
              % (world, gb, float(t.item()) * 1e3, gb * (world - 1) / float(t.item()), "correct, equal to the torch.distributed gather" if ok else "WRONG"), flush=True)
    assert ok
    comm.destroy()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
