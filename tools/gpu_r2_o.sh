#!/bin/bash
# round 2, call o: FAST in-kernel resets (MUFU twins of the samplers) -- parity + timing vs the accurate resets
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "collect or off_policy or runner or dagger or learner" 2>&1 | tail -15) > gpurun_out/r02_o_pytest.log
tail -4 gpurun_out/r02_o_pytest.log
{
for rep in 1 2; do
TAG=fast_reset timeout 300 python tools/bench_collect.py
TAG=accurate_reset B200L2F_LIB=$PWD/raptor_b200/lib/variants/libb200l2f_slowreset.so timeout 300 python tools/bench_collect.py
done
TAG=fast_reset_lag B200L2F_COLLECT_LAG=1 timeout 300 python tools/bench_collect.py
} 2>&1 | grep -v Warning | tee gpurun_out/r02_o_fast_reset.log
