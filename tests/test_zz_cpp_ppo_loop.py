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
