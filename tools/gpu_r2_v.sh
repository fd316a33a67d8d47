#!/bin/bash
# round 2, call v: packed ReLU (v + |v| with the 0.5 folded into the next layer's weights) + packed lo-part subtraction in the MLP tensor-core forward
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_v_pytest.log
tail -4 gpurun_out/r02_v_pytest.log
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warning | tee gpurun_out/r02_v_configs.jsonl | cut -c1-230
TAG=packed_relu timeout 300 python tools/bench_collect.py 2>&1 | grep -v Warning | tee gpurun_out/r02_v_collect.log
timeout 300 python tools/bench_default_collect.py 2>&1 | grep -v Warning | tee gpurun_out/r02_v_default_collect.jsonl | cut -c1-260
