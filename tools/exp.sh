mkdir -p gpurun_out
(B200L2F_A=tmem timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05 or per_environment" 2>&1 | tail -15) | tee gpurun_out/pytest_ts.log
