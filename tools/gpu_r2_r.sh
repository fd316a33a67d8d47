#!/bin/bash
# round 2, call r: full GPU suite + default bench line after the collect / sampler / RNG changes
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r02_r_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_r_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_r_bench.json 2> gpurun_out/r02_r_bench.err; tail -3 gpurun_out/r02_r_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_r_bench.json").read().strip().splitlines()[-1])
print("value %.3fe9 e2e %.3fe9 ratio %.3f kernel %s issue frac %s launches %s" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["value"] / d["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches"]))
for k, v in (d.get("configs") or {}).items():
    print(k, "%.3fe9" % (v["value"] / 1e9), v["ms_per_launch"], v["kernel"])
PY
timeout 300 python tools/bench_default_collect.py 2>&1 | grep -v Warning | tee gpurun_out/r02_r_default_collect.jsonl | cut -c1-300
