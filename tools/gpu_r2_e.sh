#!/bin/bash
# round 2, GPU call E: full suite, default bench line (async e2e, configs object), CTAS / chunk experiments, Langevin branch variant, ncu capture
TAG=${1:-r02_e}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/%s_bench.json" % "r02_e").read().strip().splitlines()[-1])
print("value %.3fe9 e2e %.3fe9 ratio %.3f kernel %s issue frac %s" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["value"] / d["value"], d["roofline"]["kernel"], d["roofline"]["frac"]))
for k, v in (d.get("configs") or {}).items():
    print(k, "%.3fe9" % (v["value"] / 1e9), v["ms_per_launch"], v["kernel"])
PY
for cfg in "3:" "4:" "4:16" "4:32" "3:16"; do
  c=${cfg%%:*}; ch=${cfg#*:}
  B200L2F_TS_CTAS=$c B200L2F_CHUNKS=$ch python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ctas=$c chunks=$ch', round(d['value']/1e9,3), round(d['ms_per_step'],4), d['roofline']['kernel'])"
done 2>&1 | tee gpurun_out/${TAG}_ctas_chunks.log
bash tools/run_variants.sh 2>&1 | tee gpurun_out/${TAG}_variants.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/${TAG}_k_rollout python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-200
