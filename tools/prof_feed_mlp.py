#!/usr/bin/env python
"""tools/prof_feed_mlp.py -- the kernels of configs 3 / 4 that had no ncu capture yet, at sizes ncu can replay: k_rollout_mlp_ts (SAC-teacher MLP
rollout), k_values_ts (critic values + GAE over a collected dataset) and k_collect<SpecDefault> (PPO zoo environment, CUDA cores).  Each is
launched twice (cold, warm); capture with  ncu --set full -k regex:'k_rollout_mlp_ts|k_values_ts|k_collect' ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
from tools.bench_configs import DR, mlp_blob  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
rs = np.random.RandomState(0)


def timed(what, units, fn):
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        print("%s: %.3f ms  %.3e /s" % (what, a.elapsed_time(b), units / a.elapsed_time(b) * 1e3), flush=True)


def make(n, spec):
    env = rb.VectorEnvironment(n, spec, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
    return env


# config 3 shape: SAC-teacher MLP rollout
n, T = 262144, 50
env = make(n, rb.SPEC_TEACHER_DR)
env.initialize_rng(3, warmup=16); env.sample_initial_parameters(); env.sample_initial_state()
env.load_policy(mlp_blob(rs, 26, 8, False, False), arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL)
ret = torch.zeros(n, device=dev)
timed("k_rollout_mlp_ts %d x %d env-steps" % (n, T), n * T, lambda: env.rollout(T, out={"returns": ret}))
del env
# config 4 learner feed: collect a dataset, then critic values + GAE in one pass
n, T = 65536, 64
env = make(n, rb.SPEC_RAPTOR_DR)
env.initialize_rng(4, warmup=16); env.initial_parameters(); env.initial_state()
env.load_policy(mlp_blob(rs, 22, 4, True, True), arch=rb.POLICY_MLP, input_dim=22, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN)
data = torch.zeros(((T + 1) * n, 37), dtype=torch.float32, device=dev)
env.collect_reset(); env.collect(T, 500, data)
env.load_critic(mlp_blob(rs, 22, 1, True, False), standardize=1)
timed("k_values_ts %d rows" % ((T + 1) * n), (T + 1) * n, lambda: env.values_and_advantages(data, T))
del env, data
# PPO zoo environment (DEFAULT spec): CUDA-core collection
n, T = 32768, 32
env = make(n, rb.SPEC_DEFAULT_DR)
env.initialize_rng(5, warmup=16); env.initial_parameters(); env.initial_state()
env.load_policy(mlp_blob(rs, 82, 4, True, True), arch=rb.POLICY_MLP, input_dim=82, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN)
data = torch.zeros(((T + 1) * n, 97), dtype=torch.float32, device=dev)
env.collect_reset()
timed("k_collect<SpecDefault> %d x %d env-steps" % (n, T), n * T, lambda: env.collect(T, 500, data))
