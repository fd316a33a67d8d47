#!/bin/bash
# tools/ref_gpu_benchmark.sh [build|run] -- the reference's OWN GPU kernel on the same box (simulate_parallel,
# /root/reference/rl-tools/src/rl/environments/l2f/cuda/benchmark.cu:98-128, launched :209-211): step() only, action 0, no policy, no observe / reward,
# one shared parameter set per block, 4096 x 512 environments x N_ITERATIONS steps in registers.
#   build  (in the container, needs /root/reference): nvcc on the file where it lies -> oracle/_ref/ref_gpu_benchmark (unmodified: 1e6 iterations) and
#          oracle/_ref/ref_gpu_benchmark_short (the same file with ITS OWN constant N_ITERATIONS set to 20000 by sed on a temporary copy, so that the run
#          takes seconds instead of minutes; nothing else differs).  Outputs only under oracle/_ref/ (git-ignored, travels to the GPU box).
#   run    (on the GPU box): runs the short binary (and the unmodified one under `timeout` when REF_GPU_FULL=1), prints its own report lines.
set -e
cd "$(dirname "$0")/.."
REF=/root/reference/rl-tools
SRC=$REF/src/rl/environments/l2f/cuda/benchmark.cu
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -use_fast_math -std=c++17 -I$REF/include"
case "${1:-run}" in
build)
    mkdir -p oracle/_ref
    nvcc $FLAGS -o oracle/_ref/ref_gpu_benchmark $SRC > /dev/null 2>&1
    TMP=$(mktemp -d)
    sed 's/constexpr size_t N_ITERATIONS = 1000000;/constexpr size_t N_ITERATIONS = 20000;/' $SRC > $TMP/benchmark_short.cu
    grep -q "N_ITERATIONS = 20000" $TMP/benchmark_short.cu
    nvcc $FLAGS -o oracle/_ref/ref_gpu_benchmark_short $TMP/benchmark_short.cu > /dev/null 2>&1
    rm -rf $TMP
    ls -la oracle/_ref/ref_gpu_benchmark oracle/_ref/ref_gpu_benchmark_short
    ;;
run)
    echo "# reference GPU kernel simulate_parallel (step only, zero action, no policy), N_ITERATIONS = 20000 (its own constant, shortened)"
    timeout 300 oracle/_ref/ref_gpu_benchmark_short | grep -E "Name|Number of SMs|Simulation time|Simluation dt"
    if [ "${REF_GPU_FULL:-0}" = "1" ]; then
        echo "# unmodified (N_ITERATIONS = 1000000)"
        timeout 900 oracle/_ref/ref_gpu_benchmark | grep -E "Simulation time"
    fi
    ;;
esac
