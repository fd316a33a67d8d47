mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k tcgen05 2>&1 | tail -25) | tee gpurun_out/pytest_tc.log
for g in "" "--tcgen05"; do for n in 65536 1048576; do
  T=1000; if [ $n = 1048576 ]; then T=200; fi
  echo "== gemm=$g n=$n" | tee -a gpurun_out/exp4.log
  timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --envs-per-gpu $n --rollout-steps $T $g 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['fp32_issue']['frac'], d['e2e']['value'], d['mean_episode_return'])" | tee -a gpurun_out/exp4.log
done; done
