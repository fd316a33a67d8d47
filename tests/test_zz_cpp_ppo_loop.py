"""The PPO loop step's data path written in C++ against the header-only shim (tests/cpp/ppo_loop.cpp), on the GPU.  Kept in a file that sorts last:
the driver runs the GPU suite with -x, and this test depends on a host compiler and a second process besides the engine."""
import os
import subprocess

import numpy as np
import pytest

from test_cpp_shim import PPO_EXE, build_exe


@pytest.mark.gpu
@pytest.mark.parametrize("spec_name", ["raptor", "default"])
def test_cpp_ppo_loop_equals_the_python_mirror(tmp_path, spec_name):
    """collect -> critic values -> GAE -> normalizer, two iterations, written in C++ against the shim: the dataset and the normalizer it ends with are
    bit-identical to the same calls made through the Python mirror (one C ABI underneath; the parity of that ABI with the oracle is
    tests/test_gpu_parity.py::test_ppo_collect_vs_oracle / test_learner_feed_vs_oracle)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200 as rb
    from conftest import random_mlp_blob
    if not os.path.exists(PPO_EXE):
        build_exe("ppo_loop")
    spec, obs = (rb.SPEC_RAPTOR_DR, 22) if spec_name == "raptor" else (rb.SPEC_DEFAULT_DR, 82)
    n, T, limit = 200, 24, 9
    rs = np.random.RandomState(9)
    actor, critic = random_mlp_blob(rs, obs, 4, True, True), random_mlp_blob(rs, obs, 1, True, False)
    blobs, out = str(tmp_path / "blobs.f32"), str(tmp_path / "out.f32")
    np.concatenate([actor, critic]).astype(np.float32).tofile(blobs)
    r = subprocess.run([PPO_EXE, spec_name, blobs, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, np.float32)
    D = obs + 15
    got, got_mean, got_std = raw[:(T + 1) * n * D].reshape(-1, D), raw[(T + 1) * n * D:][:obs], raw[(T + 1) * n * D + obs:]
    env = rb.VectorEnvironment(n, spec)
    row = env.get_environment_parameters()
    row[124:139] = np.array([1.5, 5.0, 40, 1200, 0.02, 5.0, 0.1, 0.03, 0.10, 0.03, 0.30, 0.005, 0.05, 0.0, 0.3], np.float32)
    env.set_environment_parameters(row)
    env.load_policy(actor, arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN)
    env.load_critic(critic, standardize=1)
    env.initialize_rng(77, warmup=0)
    env.initial_parameters(); env.initial_state(); env.collect_reset()
    mean, std, age = np.zeros(obs, np.float32), np.ones(obs, np.float32), 0
    for _ in range(2):
        data = env.collect(T, limit)
        env.evaluate_values(data, T)
        env.estimate_generalized_advantages(data, T, 0.99, 0.95, False)
        age = env.normalizer_update(data, T, mean, std, age)
    assert data[:T * n, obs + 11].sum() > 0 and np.abs(data[:, obs + 13]).max() > 0
    assert np.array_equal(got.view(np.uint32), data.view(np.uint32))
    assert np.array_equal(got_mean, mean) and np.array_equal(got_std, std) and age == 2


@pytest.mark.gpu
def test_cpp_sac_data_path_equals_the_python_mirror(tmp_path):
    """off-policy runner steps + gather_batch (SEQUENCE_LENGTH 1 and 6) written in C++ against the shim: rings and batches bit-identical to the same calls through the
    Python mirror (parity of the ABI with the oracle: test_off_policy_steps_vs_oracle / test_gather_batch_sequential_vs_oracle)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200 as rb
    from conftest import random_mlp_blob
    build_exe("ppo_loop")
    n, cap, obs, Bsz, Lq = 64, 64, 26, 32, 6
    actor = random_mlp_blob(np.random.RandomState(12), obs, 8, False, False)
    blob, out = str(tmp_path / "actor.f32"), str(tmp_path / "out.f32")
    actor.astype(np.float32).tofile(blob)
    r = subprocess.run([PPO_EXE, "sac", blob, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, np.float32)
    env = rb.VectorEnvironment(n, rb.SPEC_TEACHER_DR)
    row = env.get_environment_parameters()
    row[124:139] = np.array([1.5, 5.0, 40, 1200, 0.02, 5.0, 0.1, 0.03, 0.10, 0.03, 0.30, 0.005, 0.05, 0.0, 0.3], np.float32)
    env.set_environment_parameters(row)
    env.load_policy(actor, arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL)
    env.initialize_rng(91, warmup=0)
    env.initial_parameters(); env.initial_state(); env.collect_reset()
    replay = env.new_replay_buffers(cap)
    env.off_policy_steps(100, 30, replay)
    r1 = np.arange(Bsz, dtype=np.uint64) + np.uint64(0xAAAAAAAA + 5000)
    r6 = np.arange(Bsz, dtype=np.uint64) + np.uint64(0xAAAAAAAA + 6000)
    b1 = env.gather_batch(replay, r1, 500)
    b6 = env.gather_batch(replay, r6, 500, env_begin=16, env_count=24, sequence_length=Lq, include_first_step_in_targets=True, always_sample_from_initial_state=False,
                          random_seq_length=True, nominal_sequence_length_probability=0.25)
    want = [replay["data"], replay["position"], replay["full"], b1["observations_actions"], b1["rewards"], b1["terminated"]] + \
           [b6[k] for k in ("observations_actions", "rewards", "terminated", "reset", "next_reset", "final_step_mask", "next_final_step_mask")]
    want = np.concatenate([np.asarray(w, np.float32).ravel() for w in want])
    assert raw.shape == want.shape and np.array_equal(raw.view(np.uint32), want.view(np.uint32))
    assert replay["full"].all() and b6["final_step_mask"].sum() >= Bsz and b6["final_step_mask"][:Lq - 1].sum() > 0
