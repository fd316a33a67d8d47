#!/bin/bash
# round 2, GPU call H: new tests (step_repeated, DEFAULT tcgen05, status, async), like-for-like step-only line next to the reference GPU kernel, initcheck re-run
TAG=${1:-r02_h}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
(tools/ref_gpu_benchmark.sh run; python tools/step_only_benchmark.py) > gpurun_out/${TAG}_step_only_vs_reference_gpu.log 2>&1; cat gpurun_out/${TAG}_step_only_vs_reference_gpu.log
compute-sanitizer --error-exitcode 7 --print-limit 5 --tool initcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ppo_collect or off_policy_steps or runner_edge" > gpurun_out/${TAG}_initcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/${TAG}_initcheck.log | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.3fe9 e2e %.3fe9 ratio %.3f kernel %s issue frac %s launches %s e2e how: %s" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["value"] / d["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["how"][-60:]))
for k, v in (d.get("configs") or {}).items():
    print(k, "%.3fe9" % (v["value"] / 1e9), v["ms_per_launch"], v["kernel"], v["roofline"]["frac"])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
