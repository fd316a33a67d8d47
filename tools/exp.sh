mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) | tee gpurun_out/pytest_gpu8.log
