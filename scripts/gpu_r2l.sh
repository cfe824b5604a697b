#!/bin/bash
# round 2, visit L: per-kernel durations (CUPTI) of the chunk step, saved_operands inputs vs all
set -u
mkdir -p gpurun_out
for mode in inputs all; do
  timeout 600 python scripts/prof_c3_step.py 250112 3 $mode kineto > gpurun_out/r2l_step_$mode.log 2>&1
  echo "step $mode exit $?"; grep -A18 "^rep 2" gpurun_out/r2l_step_$mode.log
done
