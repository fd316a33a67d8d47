mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "mlp or collect" 2>&1 | tail -8) | tee gpurun_out/pytest_mlp_ts.log
(timeout -s KILL 200 python tools/bench_configs.py 2>&1 | tail -4) | tee gpurun_out/configs34_ts.log
(timeout -s KILL 200 python tools/bench_configs.py --fp32-gemm 2>&1 | tail -4) | tee gpurun_out/configs34_fp32.log
