#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint_tc.py -q 2>&1 | tail -8 > gpurun_out/pytest_adjtc.log
tail -n 4 gpurun_out/pytest_adjtc.log
timeout 600 python scripts/prof_c5_contrk4.py 1000000 2 > gpurun_out/c5_contrk4_kernel_shares.txt 2>&1
grep -E "^rep|kernel time|%" gpurun_out/c5_contrk4_kernel_shares.txt | head -10
