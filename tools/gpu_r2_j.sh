#!/bin/bash
# round 2, call j: k_collect_ts write-back as one bulk copy per warp-step (parity + timing vs the element loop, unrolled-RK4 variant)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "collect or learner or runner or off_policy" 2>&1 | tail -5) > gpurun_out/r02_j_pytest_collect.log
cat gpurun_out/r02_j_pytest_collect.log
{
for rep in 1 2; do
TAG=bulk timeout 300 python tools/bench_collect.py
TAG=element_loop B200L2F_NO_BULK_ROWS=1 timeout 300 python tools/bench_collect.py
TAG=bulk_unrolled_rk4 B200L2F_LIB=$PWD/raptor_b200/lib/variants/libb200l2f_unrolled.so timeout 300 python tools/bench_collect.py
done
} 2>&1 | grep -v Warning | tee gpurun_out/r02_j_collect_variants.log
