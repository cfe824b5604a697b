#!/bin/bash
# ncu evidence for the bench command (1 GPU): launch list with per-launch device time (shares of one training step).
# The ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --agents 37888 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 1300 --csv --log-file gpurun_out/launches_c3_dopri5.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
ls -la gpurun_out | tail -5
