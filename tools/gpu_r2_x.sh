#!/bin/bash
# round 2, call x: what the driver runs at round end -- smoke, GPU suite, both bench arms -- plus the ncu launch list of the bench command
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_final_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_final_launches.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_final_bench.json").read().strip().splitlines()[-1])
print("value %.3fe9 e2e %.3fe9 frac %.3f fp32 %.3f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["roofline"]["frac"], d["roofline"]["fp32_algorithmic"]["frac"]))
for k, v in d["configs"].items(): print(k, "%.3fe9" % (v["value"] / 1e9), v["ms_per_launch"], v["roofline"]["frac"], v["roofline"]["hbm"]["frac"])
r = json.loads(open("gpurun_out/r02_final_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
for K in k_rollout_mlp_ts k_collect_ts k_rollout_raptor_ts; do timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/r02_final_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_ncu_$K.log 2>&1; tail -1 gpurun_out/r02_final_ncu_$K.log | cut -c1-120; done
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warning > gpurun_out/r02_final_configs.jsonl
timeout 300 python tools/bench_default_collect.py 2>&1 | grep -v Warning > gpurun_out/r02_final_default_collect.jsonl
