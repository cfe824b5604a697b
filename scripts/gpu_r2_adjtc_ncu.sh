#!/bin/bash
# ncu --set full of the kernels of one continuous-rk4 step (configs[4] mode) over 250,112 agents; the plain run of the same program
# exited 0 in scripts/gpu_r2_adjtc2.sh.  Only the raw-page CSV travels back.
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k 'regex:aug_stage|stage_bwd_tc|stage_fwd_tc|wgrad_tc' --launch-skip 100 -c 22 \
  -f -o /tmp/adjtc python scripts/prof_c5_contrk4.py 250112 1 > gpurun_out/adjtc_ncu.log 2>&1
echo "ncu exit: $?"; tail -2 gpurun_out/adjtc_ncu.log
ncu -i /tmp/adjtc.ncu-rep --page raw --csv > gpurun_out/adjtc_raw.csv 2>/dev/null
ls -la gpurun_out/adjtc_raw.csv
python scripts/ncu_misc_summary.py gpurun_out/adjtc_raw.csv "r02: ncu --set full --clock-control none of scripts/prof_c5_contrk4.py 250112 1 (continuous adjoint on the stage kernels, matching launches 100..121 of the step: one augmented rk4 step of the backward pass): one line per kernel = the longest of its n captured launches"
