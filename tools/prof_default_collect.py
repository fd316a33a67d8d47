#!/usr/bin/env python
"""tools/prof_default_collect.py [n] [T] -- one DEFAULT-spec PPO collection (k_collect_ts<DEFAULT>) at a size ncu can replay."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
from tools.bench_configs import DR, mlp_blob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
obs = 82
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
env = rb.VectorEnvironment(n, rb.SPEC_DEFAULT_DR, stream=stream.cuda_stream)
row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
env.initialize_rng(4, warmup=16); env.initial_parameters(); env.initial_state()
env.load_policy(mlp_blob(np.random.RandomState(0), obs, 4, True, True), arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN)
data = torch.zeros(((T + 1) * n, obs + 15), dtype=torch.float32, device=dev)
env.collect_reset()
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); env.collect(T, 500, data); b.record(stream)
    torch.cuda.synchronize()
    print("%s %d x %d: %.3f ms  %.3e env-steps/s" % (env.last_kernel(), n, T, a.elapsed_time(b), n * T / a.elapsed_time(b) * 1e3))
