# bench every library under raptor_b200/lib/variants (value / ms / e2e), twice, interleaved; then once at 1 048 576 environments x 200 steps
for rep in 1 2; do
for f in raptor_b200/lib/variants/libb200l2f_*.so; do
  v=$(basename $f .so | sed 's/libb200l2f_//')
  B200L2F_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3), d['clocks']['sm_mhz'])"
done; done
for f in raptor_b200/lib/variants/libb200l2f_*.so; do
  v=$(basename $f .so | sed 's/libb200l2f_//')
  B200L2F_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --envs-per-gpu 1048576 --rollout-steps 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v 1M x 200:', round(d['value']/1e9,3), round(d['ms_per_step'],4))"
done
