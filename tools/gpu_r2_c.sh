#!/bin/bash
# round 2, GPU call C: packed single-environment dynamics (ts v2): full GPU suite, bench, ncu capture, x2 check
TAG=${1:-r02_c}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
(timeout 600 python tools/check_x2.py --time) > gpurun_out/${TAG}_check_x2.log 2>&1; grep -E "PARITY|time" gpurun_out/${TAG}_check_x2.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/${TAG}_k_rollout python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-200
python tools/bench_configs.py > gpurun_out/${TAG}_configs34.jsonl 2> gpurun_out/${TAG}_configs34.err
cut -c1-250 gpurun_out/${TAG}_configs34.jsonl
