#!/bin/bash
# round 2, call zd: last MLP layer (<= 4 outputs) on the CUDA cores instead of a third tensor-core round trip -- parity + timing
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/r02_zd_pytest.log
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warning | tee gpurun_out/r02_zd_configs.jsonl | cut -c1-330
TAG=l3_cuda timeout 300 python tools/bench_collect.py 2>&1 | grep -v Warning | tee gpurun_out/r02_zd_collect.log
timeout 300 python tools/bench_default_collect.py 2>&1 | grep -v Warning | head -1 | cut -c1-420
