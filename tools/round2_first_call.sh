#!/bin/bash
# First GPU call of round 2: everything that was written after the round-1 GPU budget ran out, in one go.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/ (merged back by gpurun).  Every step has its own timeout; a failing step does not stop the rest.
mkdir -p gpurun_out
echo "== pending + validated GPU tests"
HOLD_RUN_VARIANTS=1 timeout 800 python -m pytest tests -q -m gpu -rxX 2>&1 | tail -25 | tee gpurun_out/r2_gpu_suite.log
echo "== SDF-kernel variants (time, error) and whole-step A/B"
timeout 240 python tools/exp_matrix.py 23 2>&1 | grep -v -i "warn" | tee gpurun_out/r2_exp_matrix.log
echo "== cycle accounting: single-CTA and pair kernels"
HOLD_TC_PROF=1 timeout 120 python tools/prof_pair.py 0:0 1:0 1:32 1:160 2>&1 | grep -v -i "warn" | tee gpurun_out/r2_prof.log
echo "== tcgen05 background"
HOLD_RUN_VARIANTS=1 HOLD_BG_TC=1 timeout 120 python -m pytest tests/test_gpu_background.py -q -rxX 2>&1 | tail -5 | tee gpurun_out/r2_bg_tc.log
