#!/usr/bin/env python
"""tools/e2e_timeline.py -- where the end-to-end step time of bench.py's pipelined loop goes: CUDA events around every rollout call of the asynchronous
pipeline (kernel time inside the pipeline, gap between consecutive rollouts on the main stream), host time per call."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raptor_b200 as rb  # noqa: E402
import bench  # noqa: E402

n, T, steps = 65536, 1000, 12
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, stream=stream.cuda_stream)
row = env.get_environment_parameters(); row[124:139] = np.array(bench.DR_RANGES, np.float32); env.set_environment_parameters(row)
env.initialize_rng(seed=1, warmup=16); env.sample_initial_parameters(); env.sample_initial_state(); env.load_policy()
pin = lambda a: torch.from_numpy(a).pin_memory()
p0, s0 = pin(env.get_parameters()), pin(env.get_state())
hs = [pin(np.zeros((n, 48), np.float32)) for _ in range(2)]; hr = [pin(np.zeros(n, np.float32)) for _ in range(2)]
rd = [torch.zeros(n, device=dev) for _ in range(2)]
for variant in ("full", "full-no-returns", "full-state-last", "no-download", "no-upload", "kernel-only"):
    for rep in range(2):
        ev = []; host = []
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        if variant in ("full", "full-no-returns", "full-state-last", "no-download"):
            env.set_parameters_async(p0.numpy()); env.set_state_async(s0.numpy())
        for k in range(steps):
            b = k & 1
            h0 = time.perf_counter()
            env.policy_reset()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            h1 = time.perf_counter()
            env.rollout(T, out={"returns": rd[b]})
            h2 = time.perf_counter()
            c.record(stream)
            if variant in ("full", "no-upload", "full-no-returns", "full-state-last"):
                if k >= 1:
                    env.transfers_synchronize(uploads=False, downloads=True)
                h3 = time.perf_counter()
                if variant == "full-state-last":
                    env.copy_to_host_async(hr[b].numpy(), rd[b]); env.get_state_async(hs[b].numpy())
                else:
                    env.get_state_async(hs[b].numpy())
                    if variant != "full-no-returns":
                        env.copy_to_host_async(hr[b].numpy(), rd[b])
            else:
                h3 = time.perf_counter()
            h4 = time.perf_counter()
            if variant in ("full", "full-no-returns", "full-state-last", "no-download") and k + 1 < steps:
                env.set_parameters_async(p0.numpy()); env.set_state_async(s0.numpy())
            h5 = time.perf_counter()
            ev.append((a, c)); host.append((h1 - h0, h2 - h1, h3 - h2, h4 - h3, h5 - h4))
        env.transfers_synchronize(); env.synchronize(); torch.cuda.synchronize()
        total = time.perf_counter() - t_all
    kern = [a.elapsed_time(c) for a, c in ev]
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(steps - 1)]
    hm = np.array(host[2:]).mean(0) * 1e3
    print("%-12s total %.2f ms/step | rollout call on GPU %.3f ms | gap between rollouts %.3f ms | host ms: reset %.3f rollout() %.3f dl-wait %.3f dl-enqueue %.3f upload-enqueue %.3f"
          % (variant, total / steps * 1e3, np.mean(kern[2:]), np.mean(gaps[2:]), *hm), flush=True)
