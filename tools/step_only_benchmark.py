#!/usr/bin/env python
"""tools/step_only_benchmark.py -- the like-for-like line against the reference's own GPU kernel (tools/ref_gpu_benchmark.sh runs that one):
4096 x 512 = 2 097 152 environments of the DEFAULT spec (the reference benchmark's state / history layout), nominal parameters, action 0, step() only,
N iterations with the state in registers (b200l2f_step_repeated).  Prints Msteps/s like the reference binary."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402

n, T = 4096 * 512, 2000
for spec, name in ((rb.SPEC_DEFAULT, "DEFAULT spec (H = 16: the reference benchmark's state)"), (rb.SPEC_RAPTOR, "RAPTOR spec (H = 1, Langevin target)")):
    env = rb.VectorEnvironment(n, spec)
    env.initialize_rng(0, warmup=0); env.initial_parameters(); env.initial_state()
    a = np.zeros(4, np.float32)
    env.step_repeated(a, 10); env.synchronize()
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        env.step_repeated(a, T); env.synchronize()
        best = min(best, time.perf_counter() - t0)
    s = env.get_state()
    print("b200l2f_step_repeated, %s: %d envs x %d steps in %.2f ms = %.1f Msteps/s (finite states: %s, mean altitude %.4f)"
          % (name, n, T, best * 1e3, n * T / best / 1e6, bool(np.isfinite(s).all()), float(s[:, 2].mean())), flush=True)
    del env
