#!/bin/bash
# tools/sanitize.sh [TAG] -- compute-sanitizer passes over the GPU suite (run on the GPU box); logs under gpurun_out/, summaries to copy into profiles/.
#   memcheck : the whole `-m gpu` suite (out-of-bounds / misaligned accesses, leaks of device allocations at exit)
#   racecheck: the tests that exercise the cross-CTA hand-over of the time-chunked scheduler and the shared-memory staging of the fused kernels
#              (test_time_chunked_scheduler_is_transparent, test_fused_rollout_vs_golden, test_bench_configuration_vs_oracle at a reduced size via B200L2F_TEST_SMALL=1)
#   initcheck: the runner / collection tests (uninitialised global memory reads)
# compute-sanitizer slows kernels 10-100x: the racecheck / initcheck selections are the small-size tests.
TAG=${1:-sanitize}
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 7 --print-limit 20"
run(){ name=$1; shift; echo "== $name: $*"; ( time timeout 3000 "$@" ) > gpurun_out/${TAG}_$name.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_$name.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/${TAG}_$name.log | tail -4; }
run memcheck  $CS --tool memcheck --leak-check no python -m pytest tests -m gpu -q -x -k "not full_size and not bench_configuration and not readme_script and not config3_and_config4"
run racecheck $CS --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "time_chunked or fused_rollout_vs_golden or asynchronous_transfers or last_status or (ppo_collect and 4-tcgen05)"
run initcheck $CS --tool initcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ppo_collect or off_policy_steps or runner_edge or gather_batch_sequential"
