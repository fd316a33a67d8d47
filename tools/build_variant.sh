#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=0 ..." -- links raptor_b200/lib/variants/libb200l2f_NAME.so from the regular objects, with rollout_ts.cu (the hot
# kernel's translation unit) recompiled with the given macros; select it at run time with B200L2F_LIB=<path> (tuning experiments only)
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/.."
mkdir -p raptor_b200/lib/variants
O=raptor_b200/lib/variants/rollout_ts_$NAME.o
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v $FLAGS -c -o $O raptor_b200/csrc/rollout_ts.cu 2> raptor_b200/lib/variants/ptxas_$NAME.log
OBJS=$(ls raptor_b200/lib/obj/*.o | grep -v rollout_ts.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o raptor_b200/lib/variants/libb200l2f_$NAME.so $OBJS $O
grep -A2 "k_rollout_raptor_tsINS_7EnvSpecILi1ELb1ELi1ELb0EEELb1ELb1ELb1ELb0ELi3ELb0" raptor_b200/lib/variants/ptxas_$NAME.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $NAME"
