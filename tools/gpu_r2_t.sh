#!/bin/bash
# round 2, call t: the bench line with the refreshed kernel counters, the reference arm, and the ncu launch list of the bench command
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_t_bench.json 2> gpurun_out/r02_t_bench.err; tail -2 gpurun_out/r02_t_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_t_bench_reference.json 2>/dev/null
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_t_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_t_launches.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_t_bench.json").read().strip().splitlines()[-1])
print("value %.3fe9 e2e %.3fe9 frac %.3f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["roofline"]["frac"]))
for k, v in d["configs"].items(): print(k, "%.3fe9" % (v["value"] / 1e9), v["roofline"]["frac"], v["roofline"]["hbm"]["frac"])
r = json.loads(open("gpurun_out/r02_t_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
