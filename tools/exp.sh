mkdir -p gpurun_out
for r in unrolled rolled; do for n in 65536 1048576; do
  T=1000; if [ $n = 1048576 ]; then T=200; fi
  echo "== rk4=$r n=$n" | tee -a gpurun_out/exp11.log
  B200L2F_RK4=$r timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --envs-per-gpu $n --rollout-steps $T 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['fp32_issue']['frac'], d['e2e']['value'], d['mean_episode_return'])" | tee -a gpurun_out/exp11.log
done; done
