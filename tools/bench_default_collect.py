#!/usr/bin/env python
"""tools/bench_default_collect.py -- PPO collection + learner feed on the DEFAULT spec (the PPO zoo's environment, rl/zoo/l2f/ppo.h: H = 16 action
history, 82-wide observation, 97-float dataset rows): k_collect_ts<DEFAULT> (tcgen05, K = 88 first layer) and its CUDA-core twin k_collect; k_values<82> / k_gae.  One JSON line each.
Timing: CUDA events on the engine's stream, 3 warm-ups, L2 flushed between timed launches."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
from bench_configs import DR, mlp_blob, timed  # noqa: E402


def main(gemm):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rs = np.random.RandomState(0)
    n, T, obs = 65536, 128, 82
    D = obs + 15
    env = rb.VectorEnvironment(n, rb.SPEC_DEFAULT_DR, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(4, warmup=16); env.initial_parameters(); env.initial_state()
    env.load_policy(mlp_blob(rs, obs, 4, True, True), arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=gemm)
    data = torch.zeros(((T + 1) * n, D), dtype=torch.float32, device=dev)
    ms = timed(lambda: env.collect(T, 500, data), lambda: env.collect_reset(), stream=stream, flush=flush)
    written = n * T * (obs + 12) * 4 + n * obs * 4
    kernel_name = env.last_kernel()
    env.load_critic(mlp_blob(rs, obs, 1, True, False), standardize=1)
    nop = lambda: None
    ms_f = timed(lambda: env.values_and_advantages(data, T), nop, stream=stream, flush=flush)
    mean, std = np.zeros(obs, np.float32), np.ones(obs, np.float32)
    ms_n = timed(lambda: env.normalizer_update(data, T, mean, std, 0), nop, stream=stream, flush=flush)
    rows = (T + 1) * n
    print(json.dumps({"config": "PPO zoo environment (DEFAULT spec, DR): %d envs x %d-step collection, PPO MLP 82-64-64-4, dataset rows [(T+1)N, 97] in HBM (%.1f GB)" % (n, T, rows * D * 4 / 1e9),
                      "kernel": env.last_kernel() if False else kernel_name, "env_steps_per_s": n * T / ms * 1e3, "ms_per_launch": ms, "dataset_bytes_written": written,
                      "hbm_write_gbs": written / ms / 1e6, "mean_reward": float(data[: T * n, obs + 9].mean().item()), "truncated_fraction": float(data[: T * n, obs + 11].mean().item()),
                      "values_gae_ms": ms_f, "values_gae_rows_per_s": rows / ms_f * 1e3, "normalizer_update_ms": ms_n, "normalizer_hbm_gbs": 2 * T * n * D * 4 / ms_n / 1e6}))


if __name__ == "__main__":
    main(rb.GEMM_TCGEN05_3XTF32)
    main(rb.GEMM_FP32_CUDA_CORES)
