#!/bin/bash
# round 2, call ze: CUDA-core last layer for EIGHT outputs (SAC teacher actor, config 3) vs the tensor-core round trip
mkdir -p gpurun_out
{
for rep in 1 2; do
echo base; python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['config3']; print(round(c['value']/1e9,3), round(c['ms_per_launch'],3))"
echo l3cuda8; B200L2F_LIB=$PWD/raptor_b200/lib/variants/libb200l2f_l3cuda8.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['config3']; print(round(c['value']/1e9,3), round(c['ms_per_launch'],3))"
done
} 2>&1 | tee gpurun_out/r02_ze_l3cuda8.log
B200L2F_LIB=$PWD/raptor_b200/lib/variants/libb200l2f_l3cuda8.so timeout 600 python -m pytest tests -m gpu -x -q -k "teacher_mlp or config3 or mlp_tensor_core" 2>&1 | tail -2
