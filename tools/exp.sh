mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) | tee gpurun_out/pytest_gpu7.log
for g1 in cuda tc; do for n in 65536 1048576; do
  T=1000; if [ $n = 1048576 ]; then T=200; fi
  echo "== g1=$g1 n=$n" | tee -a gpurun_out/exp7.log
  B200L2F_G1=$g1 timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --envs-per-gpu $n --rollout-steps $T 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['fp32_issue']['frac'], d['e2e']['value'], d['mean_episode_return'])" | tee -a gpurun_out/exp7.log
done; done
