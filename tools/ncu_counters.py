#!/usr/bin/env python
"""tools/ncu_counters.py REPORT.ncu-rep --n ENVS --T STEPS [--out profiles/kernel_counters.json] [--source NAME]

Reads one `ncu --set full --import-source on` capture of a fused rollout / collection kernel and writes the per-environment-step counters that
bench.py turns into roofline fractions (instead of literals in bench.py):
  warp instructions per warp-step, thread instructions, executed CUDA-core fp32 FLOPs (from the SASS page: FADD / FMUL = 1, FFMA = 2,
  FADD2 / FMUL2 = 2, FFMA2 = 4 per predicated-on thread instruction), MUFU operations, DRAM bytes per launch, pipe utilisations.
Entries are keyed "<kernel as b200l2f_last_kernel names it>|<envs>|<steps>".  Runs where ncu is installed; no GPU needed."""
import argparse
import csv
import io
import json
import os
import re
import subprocess

FLOPS = {"FADD": 1, "FMUL": 1, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FFMA2": 4}


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def engine_kernel_name(demangled):
    m = re.search(r"(k_[a-z0-9_]+)", demangled)
    base = m.group(1) if m else demangled
    if base == "k_rollout_raptor_ts":      # k_rollout_raptor_ts<EnvSpec<...>, FAST, UNIFORM, AXIAL, NOISE, CTAS, RECORD>
        tail = re.search(r">((?:,\s*\d+)+)\s*>\s*\(", demangled)
        ints = [int(x) for x in re.findall(r"\d+", tail.group(1))] if tail else []
        return "k_rollout_raptor_ts<CTAS=%d>" % (ints[4] if len(ints) >= 5 else 3)
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--T", type=int, required=True)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_counters.json"))
    ap.add_argument("--source", default=None, help="name of the committed summary this entry comes from")
    ap.add_argument("--key", default=None)
    a = ap.parse_args()
    rows = ncu_csv(a.report, "raw")
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))

    def get(name):
        for k, v in d.items():
            if k == name or k.endswith(name):
                return num(v)
        return None
    kernel = engine_kernel_name(d.get("Kernel Name", ""))
    env_steps = float(a.n) * a.T
    warp_steps = env_steps / 32.0
    src = ncu_csv(a.report, "source", ["--print-source", "sass"])
    h = None
    ops = {}
    thread_instructions = 0.0
    for r in src:
        if r and r[0] == "Address":
            h = r
            i_src, i_thr, i_all = h.index("Source"), h.index("Predicated-On Thread Instructions Executed"), h.index("Thread Instructions Executed")
            continue
        if h is None or len(r) < len(h):
            continue
        parts = r[i_src].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0.0) + (num(r[i_thr]) or 0.0)
        thread_instructions += num(r[i_all]) or 0.0
    flop = sum(FLOPS[o] * c for o, c in ops.items() if o in FLOPS)
    entry = {
        "kernel": kernel, "envs": a.n, "steps": a.T,
        "warp_instructions_per_warp_step": get("smsp__inst_executed.sum") / warp_steps,
        "thread_instructions_per_env_step": thread_instructions / env_steps,
        "cuda_core_fp32_flop_per_env_step": flop / env_steps,
        "mufu_per_env_step": ops.get("MUFU", 0.0) / env_steps,
        "tcgen05_mma_per_tile_step": (get("smsp__sass_inst_executed_op_utcmma.sum") or 0.0) / (env_steps / 128.0),
        "dram_bytes_per_launch": (get("dram__bytes_read.sum") or 0.0) * 1.0 + (get("dram__bytes_write.sum") or 0.0) * 1.0,
        "ncu_ms": None, "issue_slots_busy_pct": get("sm__inst_issued.avg.pct_of_peak_sustained_active"),
        "pipe_pct": {"fma": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), "alu": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     "xu": get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), "lsu": get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                     "tensor": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")},
        "registers": get("launch__registers_per_thread"), "warps_per_scheduler": get("smsp__warps_active.avg.per_cycle_active"),
        "source": a.source or os.path.basename(a.report),
    }
    # units: ncu prints dram bytes in the unit of the second row (Mbyte / Kbyte / byte)
    units = dict(zip(hdr, rows[1]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = [k for k in d if k.endswith("dram__bytes_read.sum")][0]
    wr = [k for k in d if k.endswith("dram__bytes_write.sum")][0]
    entry["dram_bytes_per_launch"] = num(d[rd]) * scale.get(units[rd], 1.0) + num(d[wr]) * scale.get(units[wr], 1.0)
    t = [k for k in d if k.endswith("gpu__time_duration.sum")][0]
    entry["ncu_ms"] = num(d[t]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[t].replace("second", "s").replace("msecond", "ms"), 1.0) if units[t] in ("ms", "us", "ns", "s") else num(d[t]) * {"msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(units[t], 1.0)
    key = a.key or "%s|%d|%d" % (kernel, a.n, a.T)
    try:
        table = json.load(open(a.out))
    except Exception:
        table = {}
    table[key] = entry
    json.dump(table, open(a.out, "w"), indent=1, sort_keys=True)
    print(key, json.dumps(entry))


if __name__ == "__main__":
    main()
