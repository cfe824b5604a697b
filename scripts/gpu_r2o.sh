#!/bin/bash
# round 2, visit O: decomposition of the backward stage kernel's time by save level and store switches (timing experiments)
set -u
mkdir -p gpurun_out
for mode in none inputs all; do
  for fl in 0 16 48; do
    AB200_STAGE_TIMING_ONLY=1 AB200_STAGE_FLAGS=$fl timeout 600 python scripts/prof_c3_step.py 250112 2 $mode kineto > gpurun_out/r2o_${mode}_$fl.log 2>&1
    echo "== $mode flags $fl exit $?"; grep -E "stage_bwd_tc|stage_fwd2_tc|wgrad_tc_kernel|kernel time" gpurun_out/r2o_${mode}_$fl.log
  done
done
