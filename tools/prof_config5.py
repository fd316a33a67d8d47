#!/usr/bin/env python
"""tools/prof_config5.py [n] [T] -- one config-5 shard launch (1 048 576 envs, Raptor GRU policy, DR) at a size ncu can replay; prints the kernel name and rate."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, stream=stream.cuda_stream)
row = env.get_environment_parameters(); row[124:139] = np.array(bench.DR_RANGES, np.float32); env.set_environment_parameters(row)
env.initialize_rng(20250925, warmup=16); env.sample_initial_parameters(); env.sample_initial_state(); env.load_policy(gemm=rb.GEMM_TCGEN05_3XTF32)
ret = torch.zeros(n, device=dev)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); env.rollout(T, out={"returns": ret}); b.record(stream)
    torch.cuda.synchronize()
    print("%s: %d x %d in %.3f ms = %.3e env-steps/s" % (env.last_kernel(), n, T, a.elapsed_time(b), n * T / a.elapsed_time(b) * 1e3))
