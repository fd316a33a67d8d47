mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/prof_v6_ts_T1000 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v6.log 2>&1
tail -2 gpurun_out/ncu_v6.log | cut -c1-200
