mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "mlp or collect or samplers" 2>&1 | tail -8) | tee gpurun_out/pytest_mlp_ts.log
(timeout -s KILL 200 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-400) | tee gpurun_out/configs34_ts.log
(B200L2F_RK4=rolled timeout -s KILL 200 python tools/bench_configs.py 2>&1 | tail -1 | cut -c1-400) | tee gpurun_out/configs34_ts_rolled.log
