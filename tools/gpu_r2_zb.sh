#!/bin/bash
# round 2, call zb: time-chunked collection for the DEFAULT spec (one CTA per SM): parity with forced chunks + chunk sweep
mkdir -p gpurun_out
(B200L2F_COLLECT_CHUNKS=5 timeout 900 python -m pytest tests -m gpu -x -q -k "ppo_collect or cpp_ppo_loop or learner_feed" 2>&1 | tail -5) > gpurun_out/r02_zb_pytest_chunks5.log; tail -3 gpurun_out/r02_zb_pytest_chunks5.log
(timeout 900 python -m pytest tests -m gpu -x -q -k "ppo_collect or cpp_ppo_loop" 2>&1 | tail -3)
{
for c in 1 2 4 8 16 32; do echo "chunks $c"; B200L2F_COLLECT_CHUNKS=$c timeout 300 python tools/prof_default_collect.py | tail -1; done
echo "default"; timeout 300 python tools/prof_default_collect.py | tail -1
} 2>&1 | grep -v Warning | tee gpurun_out/r02_zb_default_chunks.log
