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
FLOP_TEACHER_GEMM = 2.0 * (26 * 64 + 64 * 64 + 64 * 8)   # config 3 actor (SAC teacher 26-64-64-8)
FLOP_PPO_GEMM = 2.0 * (22 * 64 + 64 * 64 + 64 * 4)       # config 4 actor (PPO 22-64-64-4)
BYTES_PER_ENV_LAUNCH = 4.0 * (2 * (48 + 16 + 2) + 145)   # read+write state, hidden, rng; read parameters (once per launch)
# What the kernels EXECUTE per environment step (warp instructions, CUDA-core fp32 FLOPs, MUFU operations, DRAM bytes per launch) is not written
# here: tools/ncu_counters.py extracts it from the committed `ncu --set full` captures into profiles/kernel_counters.json, keyed by the kernel
# name the engine reports (b200l2f_last_kernel) and the launch shape.  A launch without a capture reports those fractions as null.
COUNTERS_PATH = os.path.join(ROOT, "profiles", "kernel_counters.json")

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


def load_counters():
    try:
        return json.load(open(COUNTERS_PATH))
    except Exception:
        return {}


def roofline_object(kernel, n, T, per_launch_s, peaks, counters, gemm_flop, tcgen05=True):
    """the contract's roofline object for one fused kernel launch: the ceiling that binds first (instruction issue slots), with the other
    ceilings -- MUFU, executed and algorithmic fp32, tensor pipe, HBM -- as sub-objects"""
    rate = n * T / per_launch_s                                           # env-steps/s of this GPU
    clk = peaks["sm_max_mhz"] * 1e6
    c = counters.get("%s|%d|%d" % (kernel, n, T))
    if c is None:                                                         # same kernel captured at another launch shape: per-step counts carry over, traffic does not
        same = [v for k, v in counters.items() if k.split("|")[0] == kernel]
        c = dict(same[0], dram_bytes_per_launch=None, source=same[0]["source"] + " (per-step counts of another launch shape)") if same else None
    issue_peak = 148 * 4 * clk / 1e12                                     # warp instructions / s: one per scheduler and clock
    fp32_peak = 148 * 128 * 2 * clk / 1e12
    mufu_peak = 148 * 16 * clk / 1e12
    algo = FLOP_ENV + FLOP_POLICY_GATES
    r = {"bound": "issue", "achieved": None, "peak": issue_peak, "unit": "T warp-instructions/s", "frac": None, "traffic": None,
         "kernel": kernel, "kernel_ms": 1e3 * per_launch_s,
         "what": "instruction issue slots (148 SMs x 4 schedulers x max SM clock): the first ceiling this latency-bound CUDA-core kernel meets; achieved = executed warp "
                 "instructions per 32 environment steps (ncu capture) x the live rate.  Fewer instructions per step lower this fraction at equal speed -- read it together with "
                 "fp32_algorithmic, the fixed-work fraction",
         "peak_source": "148 x 4 x %.0f MHz (%s)" % (peaks["sm_max_mhz"], peaks["source"])}
    if c:
        ach = rate / 32.0 * c["warp_instructions_per_warp_step"] / 1e12
        r.update({"achieved": ach, "frac": ach / issue_peak, "traffic": c["dram_bytes_per_launch"], "counters_source": c["source"],
                  "warp_instructions_per_warp_step": c["warp_instructions_per_warp_step"], "ncu_issue_slots_busy_pct": c.get("issue_slots_busy_pct"),
                  "ncu_pipe_pct": c.get("pipe_pct"), "registers": c.get("registers")})
        r["mufu"] = {"achieved": rate * c["mufu_per_env_step"] / 1e12, "peak": mufu_peak, "unit": "T MUFU ops/s", "frac": rate * c["mufu_per_env_step"] / 1e12 / mufu_peak,
                     "per_env_step": c["mufu_per_env_step"]}
        r["fp32_executed"] = {"achieved": rate * c["cuda_core_fp32_flop_per_env_step"] / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                              "frac": rate * c["cuda_core_fp32_flop_per_env_step"] / 1e12 / fp32_peak, "per_env_step": c["cuda_core_fp32_flop_per_env_step"],
                              "note": "CUDA-core fp32 FLOPs actually executed (SASS op counts of the capture); the actor GEMMs run on the tensor pipe"}
    r["fp32_algorithmic"] = {"achieved": rate * algo / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": rate * algo / 1e12 / fp32_peak, "per_env_step": algo,
                             "note": "SURVEY 8(d): environment + gate FLOPs of the reference algorithm (fixed work) against 148 SM x 128 lanes x 2 x max clock"}
    r["tensor"] = {"achieved": rate * gemm_flop / 1e12, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": rate * gemm_flop / 1e12 / peaks["bf16_tflops_sustained"],
                   "note": "algorithmic GEMM FLOPs of the actor against the measured sustained bf16 peak (TF32 runs at half of it, the 3xTF32 split issues 3 products)" if tcgen05 else "actor on CUDA cores"}
    hbm = n * BYTES_PER_ENV_LAUNCH / per_launch_s / 1e9
    r["hbm"] = {"achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": n * BYTES_PER_ENV_LAUNCH}
    r["algorithmic_flop_per_env_step"] = FLOP_ENV + gemm_flop + FLOP_POLICY_GATES
    return r


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


def run_other_configs(rb, torch, dev, stream, flush, rank, world, use_dist):
    """BASELINE.json configs 3, 4 and 5 (this GPU's shard), each one fused launch timed with CUDA events after 3 warm-ups, L2 flushed in between; values are whole-job
    (sum over ranks / max time).  Returns the `configs` object of the JSON line."""
    if use_dist:
        import torch.distributed as dist
    peaks, counters = load_peaks(), load_counters()
    rs = np.random.RandomState(0)

    def timed(fn, reset, steps=3, warmup=3):
        for _ in range(warmup):
            reset(); fn()
        ms = []
        for _ in range(steps):
            reset(); flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize(dev)
            ms.append(a.elapsed_time(b))
        t = torch.tensor([sum(ms) / len(ms)], dtype=torch.float64, device=dev)
        if use_dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def dr_env(n, spec, seed):
        env = rb.VectorEnvironment(n, spec, device=dev.index, first_env_id=rank * n, stream=stream.cuda_stream)
        row = env.get_environment_parameters(); row[124:139] = np.array(DR_RANGES, np.float32); env.set_environment_parameters(row)
        env.initialize_rng(seed, warmup=16)
        return env
    out = {}
    # ---- config 3: 1 048 576 envs, per-env DR (sample_initial_parameters), random-init MLP actor (SAC-teacher shape 26-64-64-8), T = 500
    n, T = 1048576, 500
    env = dr_env(n, rb.SPEC_TEACHER_DR, 3)
    env.sample_initial_parameters(); env.sample_initial_state()
    env.load_policy(mlp_blob(rs, 26, 8, False, False), arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=rb.GEMM_TCGEN05_3XTF32)
    s0 = torch.from_numpy(env.get_state()).to(dev); ret = torch.zeros(n, device=dev)
    ms = timed(lambda: env.rollout(T, out={"returns": ret}), lambda: env.set_state(s0))
    k = env.last_kernel()
    out["config3"] = {"workload": "BASELINE configs[2]: %d envs/GPU x %d steps, per-env domain-randomised params, random-init MLP actor 26-64-64-8 (+ squash, evaluation mode)" % (n, T),
                      "value": n * world * T / ms * 1e3, "unit": "env-steps/s", "ms_per_launch": ms, "kernel": k,
                      "roofline": roofline_object(k, n, T, ms / 1e3, peaks, counters, FLOP_TEACHER_GEMM)}
    del env, s0, ret
    # ---- config 4: 262 144 envs x 256-step PPO collection with obs / action / reward / done write-back into the HBM dataset
    n, T = 262144, 256
    env = dr_env(n, rb.SPEC_RAPTOR_DR, 4)
    env.initial_parameters(); env.initial_state()
    env.load_policy(mlp_blob(rs, 22, 4, True, True), arch=rb.POLICY_MLP, input_dim=22, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=rb.GEMM_TCGEN05_3XTF32)
    data = torch.zeros(((T + 1) * n, 37), dtype=torch.float32, device=dev)
    ms = timed(lambda: env.collect(T, 500, data), lambda: env.collect_reset())
    written = n * T * 34 * 4 + n * 22 * 4
    k = env.last_kernel()
    rl = roofline_object(k, n, T, ms / 1e3, peaks, counters, FLOP_PPO_GEMM)
    rl["hbm"] = {"achieved": written / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": written / ms / 1e6 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": written,
                 "note": "algorithmic bytes: the 34 of 37 floats per row that collect produces (on_policy_runner.h:42-64) + the final observations; the kernel writes whole 148-byte rows (learner columns zero-filled) as one TMA bulk copy per warp-step"}
    out["config4"] = {"workload": "BASELINE configs[3]: %d envs/GPU x %d-step PPO rollout collection, PPO actor 22-64-64-4 (standardize, learned log_std), DR resets, dataset [(T+1)N, 37] in HBM" % (n, T),
                      "value": n * world * T / ms * 1e3, "unit": "env-steps/s", "ms_per_launch": ms, "kernel": k, "dataset_bytes_written": written, "roofline": rl,
                      "mean_reward": float(data[: T * n, 31].mean().item()), "truncated_fraction": float(data[: T * n, 33].mean().item())}
    if use_dist and world > 1:
        # SURVEY 8(e)'s one optional exchange: config 4 sharded over the job (262 144 / world environments per rank), every rank ends with the whole dataset.
        # b200l2f_allgather_trajectories (NCCL over NVLink, enqueued on the engine's stream behind the kernel that wrote the slab); never on the rollout path.
        try:
            import ctypes
            from raptor_b200.distributed import NcclCommunicator
            comm = NcclCommunicator()
            nl = n // world
            slab = data.view(T + 1, n, 37)[:, :nl].contiguous()
            gathered = torch.empty((world, T + 1, nl, 37), dtype=torch.float32, device=dev)
            ranks = ctypes.c_int32(0)

            def gather():
                env._check(env._lib.b200l2f_allgather_trajectories(env._h, comm.handle, slab.data_ptr(), gathered.data_ptr(), slab.numel(), ctypes.byref(ranks)))
            ms_g = timed(gather, lambda: None)
            ok = bool(torch.equal(gathered[rank], slab)) and ranks.value == world
            out["config4"]["allgather"] = {"what": "all-gather of the config-4 dataset sharded over the job: %d envs/rank x %d rows x 37 floats -> the whole dataset on every rank" % (nl, T + 1),
                                           "ms": ms_g, "bytes_in_per_rank": slab.numel() * 4, "bytes_out_per_rank": gathered.numel() * 4,
                                           "algbw_gbs": gathered.numel() * 4 / ms_g / 1e6, "busbw_gbs": gathered.numel() * 4 * (world - 1) / world / ms_g / 1e6,
                                           "own_shard_intact": ok, "entry": "b200l2f_allgather_trajectories (NCCL, %d ranks)" % ranks.value}
            comm.destroy()
            del slab, gathered
        except Exception as e:   # the exchange is optional: never lose the bench line over it
            out["config4"]["allgather"] = {"error": repr(e)[:200]}
    del env, data
    # ---- config 5: 1 048 576 envs per GPU, Raptor checkpoint (weak scaling across the ranks of this job)
    n, T = 1048576, 1000
    env = dr_env(n, rb.SPEC_RAPTOR_DR, 20250925)
    env.sample_initial_parameters(); env.sample_initial_state(); env.load_policy(gemm=rb.GEMM_TCGEN05_3XTF32)
    s0 = torch.from_numpy(env.get_state()).to(dev); ret = torch.zeros(n, device=dev)

    def reset5():
        env.set_state(s0); env.policy_reset()
    ms = timed(lambda: env.rollout(T, out={"returns": ret}), reset5)
    k = env.last_kernel()
    out["config5"] = {"workload": "BASELINE configs[4]: %d envs sharded across %d GPU(s) (%d per GPU), Raptor GRU policy, %d-step rollout, weak scaling" % (n * world, world, n, T),
                      "value": n * world * T / ms * 1e3, "unit": "env-steps/s", "ms_per_launch": ms, "kernel": k, "n_gpus": world,
                      "roofline": roofline_object(k, n, T, ms / 1e3, peaks, counters, FLOP_POLICY_GEMM)}
    del env, s0, ret
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 3 / 4 / 5 block of the JSON line")
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
        k0 = env.kernel_launches
        e0.record(stream)
        env.rollout(T, out={"returns": returns_dev})
        e1.record(stream)
        rollout_launches += env.kernel_launches - k0      # the fused rollout kernel + the two small status-reduction kernels behind it
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

    kernel_name = env.last_kernel()

    # ---- end-to-end through the public API with host buffers ("e2e"): inputs and results live in PINNED host memory (numpy views of torch pinned
    # ---- tensors); every step uploads its parameters + initial states (H2D) and downloads the final states + episode returns (D2H) inside the timed
    # ---- region.  The transfers go through the engine's asynchronous twins (b200l2f_set_*_async / get_state_async / copy_to_host_async): the upload
    # ---- for rollout k+1 and the download of rollout k run on the handle's copy streams while the main stream executes, the host waits once per
    # ---- step for the downloads of the PREVIOUS step (double-buffered results) -- the way a training loop that consumes rollouts would drive it.
    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t.numpy(), t
    params0, _keep_p = pinned(params0)
    state0, _keep_s = pinned(state0)
    host_state = [pinned(state0.copy()) for _ in range(2)]
    host_ret = [pinned(np.zeros(n, np.float32)) for _ in range(2)]
    ret_dev = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2)]

    def e2e_run(steps):
        consumed = 0.0
        env.set_parameters_async(params0); env.set_state_async(state0)              # upload for step 0
        for k in range(steps):
            b = k & 1
            env.policy_reset()
            env.rollout(T, out={"returns": ret_dev[b]})                              # main stream, asynchronous
            if k >= 1:                                                               # the ONE host wait of the step: downloads of step k-1 (rollout k is already queued)
                env.transfers_synchronize(uploads=False, downloads=True)
                consumed += float(host_ret[(k - 1) & 1][0][0]) + float(host_state[(k - 1) & 1][0][0, 0])   # the host reads the results of step k-1
            env.get_state_async(host_state[b][0])                                    # D2H  n*48*4, ordered after rollout k
            env.copy_to_host_async(host_ret[b][0], ret_dev[b])                       # D2H  n*4 (episode returns)
            if k + 1 < steps:
                env.set_parameters_async(params0)                                    # H2D  n*145*4 for step k+1: overlaps rollout k
                env.set_state_async(state0)                                          # H2D  n*48*4
        env.transfers_synchronize()
        env.synchronize()
        return consumed
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(args.steps, 20)          # a stream of rollouts: the pipeline's fill (first upload) and drain (last download) amortise over the run
    e2e_run(e2e_steps)
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = float(n) * world * T * e2e_steps / float(t_e2e.item())
    h2d = n * (145 + env.STATE_DIM) * 4
    d2h = n * (env.STATE_DIM + 1) * 4
    # the pipeline computes what the synchronous calls compute: one more step each way from the same RNG streams, bit for bit
    rng_mark = env.get_rng()
    e2e_run(1)
    piped = host_ret[0][0].copy()
    env.set_rng(rng_mark); env.set_parameters(params0); env.set_state(state0); env.policy_reset()
    sync_ret = env.rollout(T, record=("returns",))["returns"]
    e2e_check = bool(np.array_equal(piped, sync_ret))

    # ---- the other BASELINE configs on this GPU (3: 1M envs + teacher MLP, 4: PPO collection with write-back, 5: the 1M-env Raptor shard), CUDA-event timed
    configs = None
    if not args.no_configs:
        del env
        configs = run_other_configs(rb, torch, dev, stream, flush, rank, world, use_dist)

    if rank == 0:
        peaks = load_peaks()
        per_launch_s = t_total / args.steps
        counters = load_counters()
        roofline = roofline_object(kernel_name, n, T, per_launch_s, peaks, counters, FLOP_POLICY_GEMM, tcgen05=args.tcgen05)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cpu = cpu_reference_rate(args.cpu_envs_per_thread, T, spec)
            except Exception as ex:  # the checker is optional for the number, never for the tests
                cpu = {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        line = {"metric": "quadrotor env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "how": "asynchronous C-ABI transfers from / to pinned host buffers, overlapped with the kernels (one host wait per step); pipelined step == synchronous-call step, bit for bit: %s" % e2e_check},
                "gpu_launches": int(l_sum.item()), "gpu_launches_incl_input_reset": int(launches) * world,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "configs": configs,
                "wall_s_timed_region": wall, "mean_episode_return": mean_return}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
