mkdir -p gpurun_out
(timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) | tee gpurun_out/pytest_gpu12.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke2.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -o gpurun_out/prof_v4_ts_T1000 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v4b.log 2>&1
