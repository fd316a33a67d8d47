mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) | tee gpurun_out/pytest_gpu9.log
timeout -s KILL 600 python tools/bench_configs.py 2>&1 | tail -4 | tee gpurun_out/bench_configs_r01.json
