mkdir -p gpurun_out
(timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) | tee gpurun_out/pytest_gpu11.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_final_n1.json; cut -c1-1500 gpurun_out/bench_r01_final_n1.json
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/prof_v4_ts python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rollout-steps 200 > gpurun_out/ncu_v4.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1
