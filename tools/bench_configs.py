#!/usr/bin/env python
"""tools/bench_configs.py -- throughput of the other BASELINE configs (3: 1M envs + SAC-teacher MLP, 4: PPO collection with write-back).
Prints one JSON object per config.  Timing: CUDA events on the engine's stream, 3 warm-ups, L2 flushed between timed launches."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402

DR = [1.5, 5.0, 40, 1200, 0.02, 5.0, 0.1, 0.03, 0.10, 0.03, 0.30, 0.005, 0.05, 0.0, 0.3]


def mlp_blob(rs, i, o, std, ls):
    parts = []
    if std:
        parts += [np.zeros(i), np.ones(i)]
    for (oo, ii) in [(64, i), (64, 64), (o, 64)]:
        b = np.sqrt(6.0 / ii)
        parts += [rs.uniform(-b, b, (oo, ii)).ravel() * (0.3 if oo == o else 1.0), np.zeros(oo)]
    if ls:
        parts.append(np.log(np.full(4, 0.5)))
    return np.concatenate(parts).astype(np.float32)


def timed(fn, reset, steps=3, warmup=3, stream=None, flush=None):
    for _ in range(warmup):
        reset(); fn()
    ms = []
    for _ in range(steps):
        reset(); flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms)


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rs = np.random.RandomState(0)
    out = []
    gemm = rb.GEMM_FP32_CUDA_CORES if "--fp32-gemm" in sys.argv else rb.GEMM_TCGEN05_3XTF32
    gemm_name = "fp32 CUDA cores" if "--fp32-gemm" in sys.argv else "tcgen05 3xTF32"
    # ---- config 3
    n, T = 1048576, 100
    env = rb.VectorEnvironment(n, rb.SPEC_TEACHER_DR, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(3, warmup=16); env.sample_initial_parameters(); env.sample_initial_state()
    env.load_policy(mlp_blob(rs, 26, 8, False, False), arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=gemm)
    s0 = torch.from_numpy(env.get_state()).to(dev); ret = torch.zeros(n, device=dev)
    ms = timed(lambda: env.rollout(T, out={"returns": ret}), lambda: env.set_state(s0), stream=stream, flush=flush)
    out.append({"config": "3: 1 048 576 envs, per-env DR, SAC-teacher MLP 26-64-64-8 + squash (eval), %d-step rollout" % T, "env_steps_per_s": n * T / ms * 1e3, "ms_per_launch": ms,
                "algorithmic_flop_per_env_step": 1100 + 2 * 6272, "fp32_tflops": n * T / ms * 1e3 * (1100 + 2 * 6272) / 1e12})
    del env, s0, ret
    # ---- config 4
    n, T = 262144, 256
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(4, warmup=16); env.initial_parameters(); env.initial_state()
    env.load_policy(mlp_blob(rs, 22, 4, True, True), arch=rb.POLICY_MLP, input_dim=22, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=gemm)
    data = torch.zeros(((T + 1) * n, 37), dtype=torch.float32, device=dev)
    ms = timed(lambda: env.collect(T, 500, data), lambda: env.collect_reset(), stream=stream, flush=flush)
    written = n * T * 34 * 4 + n * 22 * 4
    out.append({"config": "4: 262 144 envs x 256-step PPO collection, PPO MLP 22-64-64-4 (standardize, learned log_std), DR resets, dataset rows [(T+1)N, 37] in HBM",
                "env_steps_per_s": n * T / ms * 1e3, "ms_per_launch": ms, "dataset_bytes_written": written, "hbm_write_gbs": written / ms / 1e6,
                "mean_reward": float(data[: T * n, 31].mean().item()), "truncated_fraction": float(data[: T * n, 33].mean().item())})
    # ---- learner feed on the config-4 dataset (stays in HBM): critic values + GAE in one pass, stand-alone GAE, normalizer update
    env.load_critic(mlp_blob(rs, 22, 1, True, False), standardize=1, gemm=gemm)
    rows = (T + 1) * n
    nop = lambda: None
    ms_f = timed(lambda: env.values_and_advantages(data, T), nop, stream=stream, flush=flush)
    ms_v = timed(lambda: env.evaluate_values(data, T), nop, stream=stream, flush=flush)
    ms_g = timed(lambda: env.estimate_generalized_advantages(data, T), nop, stream=stream, flush=flush)
    mean, std = np.zeros(22, np.float32), np.ones(22, np.float32)
    ms_n = timed(lambda: env.normalizer_update(data, T, mean, std, 0), nop, stream=stream, flush=flush)
    row_bytes = 37 * 4
    out.append({"config": "4 learner feed: critic 22-64-64-1 values + GAE on the [(T+1)N, 37] dataset in HBM (%.1f GB)" % (rows * row_bytes / 1e9),
                "fused_values_gae_ms": ms_f, "fused_rows_per_s": rows / ms_f * 1e3,
                "fused_hbm_gbs_algorithmic": (rows * (22 + 3) * 4 + rows * 3 * 4) / ms_f / 1e6, "fused_hbm_gbs_rows_streamed": (rows * row_bytes + rows * 12) / ms_f / 1e6,
                "values_only_ms": ms_v, "gae_only_ms": ms_g, "gae_hbm_gbs_algorithmic": rows * 6 * 4 / ms_g / 1e6,
                "normalizer_update_ms": ms_n, "normalizer_hbm_gbs": 2 * T * n * row_bytes / ms_n / 1e6,
                "collect_plus_feed_env_steps_per_s": n * T / (ms + ms_f) * 1e3})
    del env, data
    # ---- foundation-policy DAgger epoch: gather_epoch for NUM_TEACHERS = 1000 x NUM_EPISODES = 10 x EPISODE_STEP_LIMIT = 500 (post_training/config.h)
    nt, E, T = 1000, 10, 500
    n = nt * E
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(5, warmup=16); env.sample_initial_parameters()
    p = env.get_parameters(); env.set_parameters(np.ascontiguousarray(np.repeat(p[::E], E, axis=0)))
    env.sample_initial_state(); env.load_policy(gemm=gemm)
    env.load_teachers(np.stack([mlp_blob(rs, 26, 8, False, False) for _ in range(nt)]), None, episodes_per_teacher=E, gemm=gemm)
    s0 = torch.from_numpy(env.get_state()).to(dev)
    cap = n * T
    ds = dict(input_student=torch.zeros((cap, 22), device=dev), output_target=torch.zeros((cap, 4), device=dev), truncated=torch.zeros(cap, dtype=torch.uint8, device=dev),
              reset=torch.zeros(cap, dtype=torch.uint8, device=dev), episode_start=torch.zeros(n, dtype=torch.int32, device=dev))
    res = {}
    def gather():
        res["rows"] = env.dagger_gather(T, out=ds)["rows"]
    def reset():
        env.set_state(s0); env.policy_reset()
    ms_d = timed(gather, reset, stream=stream, flush=flush)
    ms_r = timed(lambda: env.rollout(T), reset, stream=stream, flush=flush)
    out.append({"config": "DAgger epoch data path: gather_epoch for 1000 teachers x 10 episodes x 500 steps (student rollout + add_to_dataset, device-resident dataset)",
                "ms_per_epoch": ms_d, "env_steps_per_s": n * T / ms_d * 1e3, "dataset_rows": res["rows"], "rows_per_s": res["rows"] / ms_d * 1e3,
                "student_rollout_only_ms": ms_r})
    del env, s0, ds
    # ---- SAC-teacher collection (off-policy runner steps): 262 144 environments x 256 steps into per-environment replay rings of 256 rows
    n, T, cap = 262144, 256, 256
    env = rb.VectorEnvironment(n, rb.SPEC_TEACHER_DR, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(6, warmup=16); env.initial_parameters(); env.initial_state()
    env.load_policy(mlp_blob(rs, 26, 8, False, False), arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=gemm)
    replay = env.new_replay_buffers(cap, device=True)
    def reset_runner():
        env.collect_reset()
        replay["position"].zero_(); replay["full"].zero_(); replay["current_episode_start"].zero_()
    ms_o = timed(lambda: env.off_policy_steps(T, 500, replay), reset_runner, stream=stream, flush=flush)
    written = n * T * (59 + 1) * 4
    out.append({"config": "off-policy runner: 262 144 envs x 256 steps, SAC actor 26-64-64-8 in Rollout mode, DR resets, replay rings [N][256][59] in HBM (%.1f GB)" % (n * cap * 59 * 4 / 1e9),
                "env_steps_per_s": n * T / ms_o * 1e3, "ms_per_launch": ms_o, "replay_bytes_written": written, "hbm_write_gbs": written / ms_o / 1e6,
                "mean_reward": float(replay["data"][:, :, 30].mean().item()), "truncated_fraction": float(replay["data"][:, :, 58].mean().item())})
    for o in out:
        o["gemm"] = gemm_name
        print(json.dumps(o))


if __name__ == "__main__":
    main()
