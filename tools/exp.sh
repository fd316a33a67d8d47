mkdir -p gpurun_out
(timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) | tee gpurun_out/pytest_gpu16.log
(timeout -s KILL 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200) | tee gpurun_out/bench_quick.log
(B200L2F_DYNAMICS=general timeout -s KILL 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200) | tee gpurun_out/bench_quick_general.log
(timeout -s KILL 200 python tools/bench_configs.py 2>&1 | tail -2 | cut -c1-260) | tee gpurun_out/configs34_ts.log
