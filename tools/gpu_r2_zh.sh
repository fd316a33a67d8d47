#!/bin/bash
# round 2, call zh: time-chunk sweep of the final hot kernel (B200L2F_CHUNKS) at the headline launch
mkdir -p gpurun_out
{
for c in 8 10 12 14 16 18 20 24 28 32; do
B200L2F_CHUNKS=$c python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c', round(d['value']/1e9,3), round(d['ms_per_step'],4))"
done
} | tee gpurun_out/r02_zh_chunk_sweep.log
