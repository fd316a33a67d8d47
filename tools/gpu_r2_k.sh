#!/bin/bash
# round 2, call k: k_collect_lag (resets on a fifth warp) -- parity + timing vs k_collect_ts
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "collect" 2>&1 | tail -15) > gpurun_out/r02_k_pytest_collect.log
cat gpurun_out/r02_k_pytest_collect.log
{
for rep in 1 2; do
TAG=lag timeout 300 python tools/bench_collect.py
TAG=inline B200L2F_COLLECT_LAG=0 timeout 300 python tools/bench_collect.py
done
TAG=lag_no_term timeout 300 python tools/bench_collect.py --no-term
} 2>&1 | grep -v Warning | tee gpurun_out/r02_k_collect_lag.log
