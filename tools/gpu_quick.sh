#!/bin/bash
# quick GPU iteration: x2 parity + timing (tools/check_x2.py), optional ncu of the bench launch (NCU=1), tag $1
TAG=${1:-quick}
mkdir -p gpurun_out
(timeout 600 python tools/check_x2.py --time) > gpurun_out/${TAG}_check_x2.log 2>&1; echo "check_x2 rc=$?" >> gpurun_out/${TAG}_check_x2.log
grep -E "PARITY|time|rc=" gpurun_out/${TAG}_check_x2.log
if [ "${NCU:-0}" = "1" ]; then
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/${TAG}_k_rollout python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-200
fi
