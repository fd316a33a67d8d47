#!/usr/bin/env python
"""tools/prof_off_policy.py -- off-policy runner launches at a size ncu can replay (usage: prof_off_policy.py [n] [T] [capacity] [launches])."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
from tools.bench_configs import DR, mlp_blob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 64
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
env = rb.VectorEnvironment(n, rb.SPEC_TEACHER_DR, stream=stream.cuda_stream)
row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32); env.set_environment_parameters(row)
env.initialize_rng(6, warmup=16); env.initial_parameters(); env.initial_state()
env.load_policy(mlp_blob(np.random.RandomState(0), 26, 8, False, False), arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL)
replay = env.new_replay_buffers(cap, device=True)
env.collect_reset()
for _ in range(launches):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); env.off_policy_steps(T, 500, replay); b.record(stream)
    torch.cuda.synchronize()
    print("off_policy_steps %d x %d: %.3f ms  %.3e env-steps/s" % (n, T, a.elapsed_time(b), n * T / a.elapsed_time(b) * 1e3))
