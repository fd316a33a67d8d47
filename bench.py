#!/usr/bin/env python
"""bench.py -- the headline benchmark: quadrotor env-steps/sec of the fused rollout hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs-per-gpu E] [--rollout-steps T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One bench "step" = one pass of the hot path over one batch: ONE fused persistent-kernel launch that advances `envs-per-gpu`
environments by `rollout-steps` closed-loop control steps (observe -> Raptor GRU actor -> RK4 step -> reward -> terminated) on each GPU.
Workload at N=1 = BASELINE.json configs[1] (65 536 envs, Raptor GRU policy, 1000-step rollout on 1xB200); for N>1 every GPU runs the
same shard size (weak scaling, environments keyed by global id, no collective on the rollout path).

value   : whole-job env-steps/s, inputs resident in HBM, CUDA-event time of the K launches (max over ranks), L2 flushed between launches.
e2e     : same metric through the public C-ABI/`VectorEnvironment` calls with HOST buffers: per step set_parameters + set_state (H2D),
          policy_reset, rollout, get_state + returns (D2H); wall clock between device synchronisations, max over ranks.
roofline: see DESIGN.md "Rooflines" -- algorithmic FLOPs of the dominant kernel against the measured peaks in MEASURED_PEAKS.json.
cpu_baseline / --impl reference: the reference's own CPU implementation (oracle/_ref, compiled from /root/reference) -- or the plain-C
          port when that library is absent -- timed on this host's cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# algorithmic cost per env-step (restated in DESIGN.md "Rooflines")
FLOP_ENV = 1100.0          # reference algorithm: RK4 (4 dynamics evaluations) + observe + reward, fp32
FLOP_POLICY_GEMM = 3904.0  # 2 * (22*16 + 48*16 + 48*16 + 16*4) multiply-accumulates, tensor-core eligible
FLOP_POLICY_GATES = 150.0
# what k_rollout_raptor_ts EXECUTES on the CUDA cores per environment step (ncu, profiles/r01_ncu_k_rollout_raptor_ts_v9_default_bench.txt):
# fadd + fmul + 2 ffma thread instructions, and all thread instructions.  The actor's GEMMs run on tcgen05 and are not in these counts.
TS_CUDA_CORE_FLOP = 1283.0
TS_THREAD_INSTRUCTIONS = 1302.0
TS_WARP_INSTRUCTIONS = 1416.0       # smsp__inst_executed.sum / (65536 / 32 warps x 1000 steps), same capture
BYTES_PER_ENV_LAUNCH = 4.0 * (2 * (48 + 16 + 2) + 145)   # read+write state, hidden, rng; read parameters (once per launch)

DR_RANGES = [1.5, 5.0, 40, 1200, 0.02, 5.0, 0.1, 0.03, 0.10, 0.03, 0.30, 0.005, 0.05, 0.0, 0.3]  # sample_dynamics_parameters.cpp:48-64


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


class ClockSampler:
    """samples SM clock / throttle reasons of one GPU (NVML, every 5 ms) while the timed region runs; nvidia-smi as a fallback"""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.nvml = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        if self.nvml is None:
            return
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                util = n.nvmlDeviceGetUtilizationRates(self.handle).gpu
                self.rows.append((sm, reasons, util))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1)
        if self.nvml is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        n = self.nvml
        names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = [r[0] for r in self.rows]
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.max_sm), "reasons": reasons, "samples": len(sm)}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


def cpu_reference_rate(envs_per_thread, T, spec, seed=1):
    """times the reference's CPU implementation (all host threads) on a bounded sample; returns (env-steps/s, dict)"""
    from oracle import binding as B
    from conftest import foundation_dr_env_params
    kind = "reference" if B.Ref.available(fast=True) else "port"
    lib = B.Ref(fast=True) if kind == "reference" else B.Port(fast=True)
    helper = B.Port()   # samplers for the inputs (not timed)
    threads = lib.hardware_threads()
    n = envs_per_thread * threads
    rng = helper.rng_states(seed, 64, warmup=16)
    env_p = foundation_dr_env_params(helper, spec)
    p64 = helper.sample_initial_parameters_n(spec, env_p, rng)
    s64 = helper.sample_initial_state_n(spec, p64, rng)
    params = np.ascontiguousarray(np.tile(p64, (n // 64 + 1, 1))[:n])
    states = np.ascontiguousarray(np.tile(s64, (n // 64 + 1, 1))[:n])
    rngs = np.ascontiguousarray(np.tile(rng, n // 64 + 1)[:n])
    hidden = np.tile(np.load(os.path.join(ROOT, "tests", "golden", "raptor_kat.npz"))["h0"], (n, 1)).astype(np.float32)
    gstep = np.zeros(n, np.int32)
    if kind == "reference":
        run = lambda: lib.rollout(spec, params, states, rngs, T, hidden=hidden, gru_step=gstep, threads=threads, record=False)
    else:
        from raptor_b200 import raptor_policy_blob
        pol = lib.make_policy(raptor_policy_blob())
        run = lambda: lib.rollout(spec, pol, params, states, rngs, T, hidden=hidden, gru_step=gstep, threads=threads, record=False)
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    rate = n * T / dt
    info = {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": kind,
            "sample": "%d envs (%d per thread x %d threads) x %d steps, same spec/policy/DR inputs, %s flags" % (n, envs_per_thread, threads, T, "-Ofast -march=x86-64-v3" if kind == "reference" else "-O3 -march=x86-64-v3"),
            "seconds": dt}
    return rate, info


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    import raptor_b200 as rb  # constants only
    spec = rb.SPEC_RAPTOR_DR
    T = args.rollout_steps
    rates = []
    info = None
    for i in range(args.warmup + args.steps):
        r, info = cpu_reference_rate(args.cpu_envs_per_thread, T, spec, seed=1 + i)
        if i >= args.warmup:
            rates.append((r, info["seconds"]))
    total_steps = sum(r * s for r, s in rates)
    total_time = sum(s for _, s in rates)
    value = total_steps / total_time
    info["value"] = value
    line = {"impl": "reference", "metric": "quadrotor env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_time / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "cpu_baseline": info,
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    return {"workload": "BASELINE configs[1]: %d envs/GPU x %d-step closed-loop rollout, Raptor GRU policy (Dense22-16/GRU16/Dense16-4), foundation-policy env spec (H=1, OBS 22, Langevin targets), per-env domain-randomised dynamics" % (args.envs_per_gpu, args.rollout_steps),
            "envs_per_gpu": args.envs_per_gpu, "rollout_steps": args.rollout_steps, "global_envs": args.envs_per_gpu * world,
            "env_steps_per_bench_step": args.envs_per_gpu * world * args.rollout_steps, "parallelism": "env-shards x%d (no collective on the rollout path)" % world,
            "l2": "flushed (256 MiB write) before every timed launch", "gemm": "fp32 CUDA cores" if not args.tcgen05 else "tcgen05 kind::tf32, 3xTF32 split, A operand and accumulators in TMEM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--rollout-steps", type=int, default=1000)
    ap.add_argument("--cpu-envs-per-thread", type=int, default=8192)   # ~9 s of CPU work per sample on the 16 host threads of the GPU box
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32-gemm", action="store_true", help="actor GEMMs on the fp32 CUDA cores instead of tcgen05 (3xTF32)")
    ap.add_argument("--accurate-math", action="store_true")
    args = ap.parse_args()
    args.tcgen05 = not args.fp32_gemm
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local, world = dist_env()
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import raptor_b200 as rb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, T = args.envs_per_gpu, args.rollout_steps
    spec = rb.SPEC_RAPTOR_DR
    stream = torch.cuda.Stream(dev)        # a real (non-NULL) stream shared by torch (events, L2 flush) and the engine's launches
    torch.cuda.set_stream(stream)
    env = rb.VectorEnvironment(n, spec, device=local, first_env_id=rank * n, flags=rb.FLAG_ACCURATE_MATH if args.accurate_math else 0, stream=stream.cuda_stream)
    row = env.get_environment_parameters()
    row[124:139] = np.array(DR_RANGES, np.float32)
    env.set_environment_parameters(row)
    env.initialize_rng(seed=20250925, warmup=16)
    env.sample_initial_parameters()
    env.sample_initial_state()
    env.load_policy(gemm=rb.GEMM_TCGEN05_3XTF32 if args.tcgen05 else rb.GEMM_FP32_CUDA_CORES)
    params0 = env.get_parameters()
    state0 = env.get_state()
    state0_dev = torch.from_numpy(state0).to(dev)
    returns_dev = torch.zeros(n, dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reset_inputs():
        env.set_state(state0_dev)      # device -> device, untimed
        env.policy_reset()

    # ---- kernel-resident throughput ("value")
    for _ in range(args.warmup):
        reset_inputs()
        env.rollout(T, out={"returns": returns_dev})
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = env.kernel_launches
    rollout_launches = 0
    events = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        reset_inputs()
        flush.fill_(1)                 # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        env.rollout(T, out={"returns": returns_dev})
        e1.record(stream)
        rollout_launches += 1
        events.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    ms = [a.elapsed_time(b) for a, b in events]
    t_local = sum(ms) / 1e3
    launches = env.kernel_launches - launches0
    t_max = torch.tensor([t_local], dtype=torch.float64, device=dev)
    l_sum = torch.tensor([rollout_launches], dtype=torch.int64, device=dev)
    if use_dist:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(l_sum, op=dist.ReduceOp.SUM)
    t_total = float(t_max.item())
    total_env_steps = float(n) * world * T * args.steps
    value = total_env_steps / t_total
    mean_return = float(returns_dev.mean().item())

    # ---- end-to-end through the public API with host buffers ("e2e"): inputs and results live in PINNED host memory (numpy views of
    # ---- torch pinned tensors), every step copies them H2D / D2H inside the timed region
    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t.numpy(), t
    params0, _keep_p = pinned(params0)
    state0, _keep_s = pinned(state0)
    host_state, _keep_hs = pinned(state0.copy())
    host_ret, _keep_hr = pinned(np.zeros(n, np.float32))
    for _ in range(2):
        env.set_parameters(params0); env.set_state(state0); env.policy_reset(); env.rollout(T, out={"returns": host_ret}); env.get_state(out=host_state)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        env.set_parameters(params0)                       # H2D  n*145*4
        env.set_state(state0)                             # H2D  n*48*4
        env.policy_reset()
        env.rollout(T, out={"returns": host_ret})         # D2H  n*4 (episode returns)
        env.get_state(out=host_state)                     # D2H  n*48*4
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = float(n) * world * T * e2e_steps / float(t_e2e.item())
    h2d = n * (145 + env.STATE_DIM) * 4
    d2h = n * (env.STATE_DIM + 1) * 4

    if rank == 0:
        peaks = load_peaks()
        per_launch_s = t_total / args.steps
        steps_per_s_gpu = n * T / per_launch_s
        tensor_ach = steps_per_s_gpu * FLOP_POLICY_GEMM / 1e12
        hbm_ach = n * BYTES_PER_ENV_LAUNCH / per_launch_s / 1e9
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        issue_peak = 148 * 128 * peaks["sm_max_mhz"] * 1e6 / 1e12          # thread instructions / s: 4 schedulers x 32 lanes per SM and clock
        if args.tcgen05:
            fp32_ach = steps_per_s_gpu * TS_CUDA_CORE_FLOP / 1e12
            fp32_note = "executed CUDA-core fp32 FLOPs of k_rollout_raptor_ts (ncu op counts per env-step x measured rate); the actor GEMMs are on the tensor pipe"
            warp_issue_peak = 148 * 4 * peaks["sm_max_mhz"] * 1e6 / 1e12    # warp instructions / s: one per scheduler and clock, 4 schedulers per SM
            warp_issue_ach = steps_per_s_gpu / 32.0 * TS_WARP_INSTRUCTIONS / 1e12
            issue = {"achieved": warp_issue_ach, "peak": warp_issue_peak, "unit": "T warp-instructions/s", "frac": warp_issue_ach / warp_issue_peak,
                     "warp_instructions_per_warp_step": TS_WARP_INSTRUCTIONS, "thread_instructions_per_env_step": TS_THREAD_INSTRUCTIONS,
                     "note": "the ceiling that binds this kernel: instruction issue slots (148 SMs x 4 schedulers x max SM clock); executed warp instructions per "
                             "warp-step (32 environments) from the committed ncu capture, rate measured live; ncu's own sm__inst_issued is 61.1 % of active cycles"}
        else:
            fp32_ach = steps_per_s_gpu * (FLOP_ENV + FLOP_POLICY_GEMM + FLOP_POLICY_GATES) / 1e12
            fp32_note = "algorithmic fp32 FLOPs (SURVEY 8d) of the CUDA-core kernel, actor GEMMs included"
            issue = None
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed `ncu --set full` capture of this exact configuration
        # (39.42 MB read + 2.02 MB written: parameters / state in, state out; the per-chunk hand-over lives in the 126 MB L2)
        traffic, traffic_src = None, None
        if args.tcgen05 and n == 65536 and T == 1000 and not args.accurate_math:
            traffic, traffic_src = 41.44e6, "profiles/r01_ncu_k_rollout_raptor_ts_v9_default_bench.txt"
        roofline = {"bound": "tensor", "achieved": tensor_ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tensor_ach / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long launch)",
                    "kernel": "k_rollout_raptor_ts" if args.tcgen05 else "k_rollout_raptor", "kernel_ms": 1e3 * per_launch_s,
                    "hbm": {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": n * BYTES_PER_ENV_LAUNCH},
                    "fp32_issue": {"achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak, "note": fp32_note},
                    "issue": issue,
                    "algorithmic_flop_per_env_step": FLOP_ENV + FLOP_POLICY_GEMM + FLOP_POLICY_GATES}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cpu = cpu_reference_rate(args.cpu_envs_per_thread, T, spec)
            except Exception as ex:  # the checker is optional for the number, never for the tests
                cpu = {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        line = {"metric": "quadrotor env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
                "gpu_launches": int(l_sum.item()), "gpu_launches_incl_input_reset": int(launches) * world,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "wall_s_timed_region": wall, "mean_episode_return": mean_return}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
