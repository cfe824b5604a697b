#!/bin/bash
# round 2, visit X: device-timeline idle analysis of a c3 chunk-step at the 1-GPU chunk size and at the 8-GPU shard size
set -u
mkdir -p gpurun_out
timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2x_idle_250k.txt 2>&1
echo "250k exit $?"; tail -n 32 gpurun_out/r2x_idle_250k.txt
timeout 600 python scripts/prof_c3_step.py 125056 3 all kineto > gpurun_out/r2x_idle_125k.txt 2>&1
echo "125k exit $?"; tail -n 32 gpurun_out/r2x_idle_125k.txt
