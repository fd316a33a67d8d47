#!/bin/bash
# tools/build_variant_unit.sh UNIT NAME "-DFLAG=0 ..." -- like build_variant.sh for any translation unit (UNIT = collect_ts, rollout_mlp_ts, ...):
# links raptor_b200/lib/variants/libb200l2f_NAME.so from the regular objects with UNIT.cu recompiled with the given macros (B200L2F_LIB=<path> selects it)
set -e
UNIT=$1; NAME=$2; FLAGS=$3
cd "$(dirname "$0")/.."
mkdir -p raptor_b200/lib/variants
O=raptor_b200/lib/variants/${UNIT}_$NAME.o
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v $FLAGS -c -o $O raptor_b200/csrc/$UNIT.cu 2> raptor_b200/lib/variants/ptxas_$NAME.log
OBJS=$(ls raptor_b200/lib/obj/*.o | grep -v "/$UNIT.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o raptor_b200/lib/variants/libb200l2f_$NAME.so $OBJS $O -ldl
grep -o "Used [0-9]* registers" raptor_b200/lib/variants/ptxas_$NAME.log | sort | uniq -c | tr '\n' ' '; echo " <- $NAME"
