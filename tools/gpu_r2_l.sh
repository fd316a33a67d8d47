#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "collect" 2>&1 | tail -15) > gpurun_out/r02_l_pytest_collect.log
cat gpurun_out/r02_l_pytest_collect.log
{
TAG=lag timeout 300 python tools/bench_collect.py
TAG=lag_no_term timeout 300 python tools/bench_collect.py --no-term
TAG=inline B200L2F_COLLECT_LAG=0 timeout 300 python tools/bench_collect.py
} 2>&1 | grep -v Warning | tee gpurun_out/r02_l_collect_lag.log


