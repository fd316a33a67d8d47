#!/bin/bash
# round 2, GPU call A: x2 kernel parity + timing, regression suite on the TS kernel, reference GPU kernel, bench lines, ncu capture of x2
TAG=r02_a
mkdir -p gpurun_out
(timeout 600 python tools/check_x2.py --time) > gpurun_out/${TAG}_check_x2.log 2>&1; echo "check_x2 rc=$?" >> gpurun_out/${TAG}_check_x2.log
tail -32 gpurun_out/${TAG}_check_x2.log
(time B200L2F_X2=0 timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest_gpu_ts.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu_ts.log
tools/ref_gpu_benchmark.sh run > gpurun_out/${TAG}_ref_gpu_benchmark.log 2>&1; cat gpurun_out/${TAG}_ref_gpu_benchmark.log
B200L2F_X2=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_ts.json 2> gpurun_out/${TAG}_bench_ts.err; cut -c1-300 gpurun_out/${TAG}_bench_ts.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_x2.json 2> gpurun_out/${TAG}_bench_x2.err; cut -c1-300 gpurun_out/${TAG}_bench_x2.json
B200L2F_CHUNKS=16 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_x2_c16.json 2> gpurun_out/${TAG}_bench_x2_c16.err; cut -c1-300 gpurun_out/${TAG}_bench_x2_c16.json
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/${TAG}_k_rollout_raptor_x2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
