"""Host-side multi-process logic on CPU (gloo, world_size 2): shard ownership by global environment id and the trajectory gather.
The rollout itself has no collective; the engine's shard invariance on real GPUs is asserted in test_gpu_parity.py::test_full_size_properties."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_and_order():
    sys.path.insert(0, ROOT)
    from raptor_b200.distributed import shard_range
    for n in [1, 7, 64, 65536, 8388608 + 3]:
        for world in [1, 2, 3, 8]:
            nxt = 0
            for r in range(world):
                first, cnt = shard_range(r, world, n)
                assert first == nxt and cnt >= 0
                nxt += cnt
            assert nxt == n


SCRIPT = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    from raptor_b200.distributed import init_process_group, shard_range, allgather_trajectories
    rank, local, world = init_process_group("gloo")
    assert world == 2
    for n_global in (10, 11):                      # equal and ragged shards
        first, cnt = shard_range(rank, world, n_global)
        T, D = 3, 5
        ids = torch.arange(first, first + cnt, dtype=torch.float32)
        slab = ids[None, :, None].expand(T, cnt, D).contiguous() + torch.arange(T, dtype=torch.float32)[:, None, None] * 1000
        full = allgather_trajectories(slab)
        assert full.shape == (T, n_global, D), full.shape
        assert torch.equal(full, allgather_trajectories(slab, n_global=n_global))
        want = torch.arange(n_global, dtype=torch.float32)[None, :, None].expand(T, n_global, D) + torch.arange(T, dtype=torch.float32)[:, None, None] * 1000
        assert torch.equal(full, want)
    sys.stdout.write("rank" + str(rank) + "-ok" + chr(10)); sys.stdout.flush()
""") % ROOT


def test_gloo_world_size_2_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_reference_arm_under_torchrun_only_rank0_works(tmp_path):
    """bench.py --impl reference: rank 0 prints the line, the other ranks exit 0 without work"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29618",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-envs-per-thread", "16", "--rollout-steps", "50"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and '"impl": "reference"' in lines[0] and '"n_gpus": 2' in lines[0]


def test_nccl_unique_id_marshalling_keeps_embedded_nul_bytes():
    """regression: the 128-byte ncclUniqueId travels between ranks as bytes; a c_char array read as a value is cut at its first NUL (found on the first 2-GPU run
    of b200l2f_allgather_trajectories: ncclCommInitRank failed with a remote error on every rank but 0)"""
    import ctypes
    from raptor_b200.distributed import unique_id_bytes

    class UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]
    uid = UniqueId()
    raw = bytes([7, 0, 9, 0, 0, 200] + list(range(122)))
    ctypes.memmove(ctypes.byref(uid), raw, 128)
    assert bytes(uid.internal) != raw              # the trap
    assert unique_id_bytes(uid) == raw
