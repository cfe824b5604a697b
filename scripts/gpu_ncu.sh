#!/bin/bash
# ncu evidence for the bench command (1 GPU): launch list with per-launch device time, then one full capture of the
# dominant kernel.  Each ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --workload c2 --precision bf16 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_bf16.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rk4_tc_kernel -s 2 -c 1 -o gpurun_out/prof_rk4_tc_c2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full-capture exit: $?"
ls -la gpurun_out | tail -12
