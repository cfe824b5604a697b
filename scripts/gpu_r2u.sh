#!/bin/bash
# round 2, visit U: kernel shares of the chunk step with the final code: CUPTI table (warm) and the ncu launch list (cold, serialised)
set -u
mkdir -p gpurun_out
timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2u_step_all.log 2>&1
echo "step exit $?"; grep -A14 "^rep 2" gpurun_out/r2u_step_all.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2u_launches.csv python scripts/prof_c3_step.py 250112 2 all > gpurun_out/r2u_ncu.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/r2u_ncu.log
