#!/bin/bash
# round 2, GPU call F: full suite, default bench line, ncu captures of the four fused kernels the bench line reports (for profiles/kernel_counters.json)
TAG=${1:-r02_f}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.3fe9 e2e %.3fe9 ratio %.3f kernel %s issue frac %s launches %s" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["value"] / d["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches"]))
for k, v in (d.get("configs") or {}).items():
    print(k, "%.3fe9" % (v["value"] / 1e9), v["ms_per_launch"], v["kernel"])
PY
B200L2F_TS_RECORD=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('RECORD=1 variant', round(d['value']/1e9,3), round(d['ms_per_step'],4))"
for K in k_rollout_raptor_ts k_rollout_mlp_ts k_collect_ts; do
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${TAG}_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$K.log | cut -c1-160
done
# the 1M-environment Raptor shard (config 5, CTAS = 4): the second k_rollout_raptor_ts kernel of the run is the warm-up of config 5 -> skip the headline's launches
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout_raptor_ts -s 60 -c 1 -o gpurun_out/${TAG}_k_rollout_raptor_ts_config5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_config5.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_config5.log | cut -c1-160
