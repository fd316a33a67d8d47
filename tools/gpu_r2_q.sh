#!/bin/bash
# round 2, call q: deferred parameter write-back + by-value RNG state in the out-of-line draw -- parity + timing
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "collect or off_policy or runner or dagger or learner or mlp or noise" 2>&1 | tail -15) > gpurun_out/r02_q_pytest.log
tail -4 gpurun_out/r02_q_pytest.log
{
for rep in 1 2; do
TAG=deferred_flush_value_rng timeout 300 python tools/bench_collect.py
done
TAG=no_term timeout 300 python tools/bench_collect.py --no-term
} 2>&1 | grep -v Warning | tee gpurun_out/r02_q_collect.log
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warning | tee gpurun_out/r02_q_configs.jsonl | cut -c1-260
