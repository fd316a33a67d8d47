#!/usr/bin/env python
"""tools/bench_collect.py [--n N] [--T T] -- BASELINE config 4 alone (k_collect_ts): 262 144 envs x 256-step PPO collection, CUDA events, L2 flushed.
B200L2F_LIB selects a variant library, B200L2F_NO_BULK_ROWS=1 the element-loop write-back."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raptor_b200 as rb  # noqa: E402
from tools.bench_configs import DR, mlp_blob, timed  # noqa: E402


def main():
    n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 262144
    T = int(sys.argv[sys.argv.index("--T") + 1]) if "--T" in sys.argv else 256
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rs = np.random.RandomState(0)
    mlp_blob(rs, 26, 8, False, False)   # same random stream position as bench.py / bench_configs.py
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR, stream=stream.cuda_stream)
    row = env.get_environment_parameters(); row[124:139] = np.array(DR, np.float32)
    if "--no-term" in sys.argv: row[114] = 0.0   # termination off: only the forced reset at t = 0 (upper bound with the reset path out of the way)
    env.set_environment_parameters(row)
    env.initialize_rng(4, warmup=16); env.initial_parameters(); env.initial_state()
    env.load_policy(mlp_blob(rs, 22, 4, True, True), arch=rb.POLICY_MLP, input_dim=22, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=rb.GEMM_TCGEN05_3XTF32)
    data = torch.zeros(((T + 1) * n, 37), dtype=torch.float32, device=dev)
    ms = timed(lambda: env.collect(T, 500, data), lambda: env.collect_reset(), stream=stream, flush=flush, steps=5)
    written = n * T * 34 * 4 + n * 22 * 4
    print(json.dumps({"tag": os.environ.get("TAG", ""), "env_steps_per_s": n * T / ms * 1e3, "ms_per_launch": ms, "hbm_write_gbs": written / ms / 1e6, "kernel": env.last_kernel(),
                      "mean_reward": float(data[: T * n, 31].mean().item()), "truncated_fraction": float(data[: T * n, 33].mean().item()),
                      "checksum": float(data[:, :34].double().abs().sum().item())}))


if __name__ == "__main__":
    main()
