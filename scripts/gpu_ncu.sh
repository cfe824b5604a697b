#!/bin/bash
# ncu evidence for the bench command (1 GPU): launch list with per-launch device time, then one full capture of the
# dominant kernels.  Each ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --agents 131072 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"stage_bwd_tc_kernel|stage_fwd_tc_kernel|wgrad_tc_kernel" -s 1500 -c 6 -o gpurun_out/prof_c3_stage $CMD > gpurun_out/ncu_full.log 2>&1
echo "full-capture exit: $?"
ls -la gpurun_out | tail -8
