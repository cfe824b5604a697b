#!/bin/bash
# round 2, visit M: where does the time of the saving forward / spilling backward go?  (timing experiments, results invalid)
set -u
mkdir -p gpurun_out
for fl in 64; do
  AB200_STAGE_TIMING_ONLY=1 AB200_STAGE_FLAGS=$fl timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2m_flags$fl.log 2>&1
  echo "flags $fl exit $?"; grep -A5 "^rep 2" gpurun_out/r2m_flags$fl.log
done
