#!/bin/bash
# round 2, visit S: ncu --set full of the three stage launches in their shipped form (saved_operands = all), one ncu process
set -u
mkdir -p gpurun_out
timeout 300 python scripts/prof_c3_step.py 250112 1 all > gpurun_out/r2s_plain.log 2>&1
echo "plain exit $?"; tail -n 2 gpurun_out/r2s_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'stage_fwd2_tc_kernel|stage_bwd_tc_kernel|wgrad_tc_kernel' -s 35 -c 4 -f -o /tmp/prof_stage python scripts/prof_c3_step.py 250112 1 all > gpurun_out/r2s_ncu.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_stage.ncu-rep --page raw --csv > gpurun_out/r2s_raw_stage.csv 2>/dev/null
ncu -i /tmp/prof_stage.ncu-rep --page source --csv --kernel-name regex:stage_bwd_tc_kernel > gpurun_out/r2s_source_bwd.csv 2>/dev/null
ncu -i /tmp/prof_stage.ncu-rep --page source --csv --kernel-name regex:stage_fwd2_tc_kernel > gpurun_out/r2s_source_fwd2.csv 2>/dev/null
ls -la gpurun_out | grep r2s
