# last gpurun call of a session: the driver's own checks (pytest -m gpu -x, smoke) + time-chunk sweep of the hot kernel + both bench arms
TAG=${1:-r01_s11}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in default 16 13 6 25; do
  if [ $c = default ]; then unset B200L2F_CHUNKS; else export B200L2F_CHUNKS=$c; fi
  B200L2F_VERBOSE=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_chunks_$c.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c', round(d['value']/1e9,3), 'e9 env-steps/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e9,3), 'sm', d['clocks']['sm_mhz'])"
  grep -m1 "schedule:" gpurun_out/${TAG}_chunks_$c.err
done 2>&1 | tee gpurun_out/${TAG}_chunk_sweep.log
unset B200L2F_CHUNKS
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-330 gpurun_out/${TAG}_bench.json
