#!/bin/bash
# GPU visit: continuous-rk4 mode -- tests again, kernel shares (CUPTI), ncu launch list of one 250k-agent step
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint_tc.py -q 2>&1 | tail -6 > gpurun_out/pytest_adjtc.log
tail -n 3 gpurun_out/pytest_adjtc.log
timeout 600 python scripts/prof_c5_contrk4.py 1000000 2 > gpurun_out/c5_contrk4_kernel_shares.txt 2>&1
tail -n 18 gpurun_out/c5_contrk4_kernel_shares.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c5_contrk4_launches.csv \
  python scripts/prof_c5_contrk4.py 250112 1 > gpurun_out/c5_contrk4_ncu.log 2>&1
tail -n 3 gpurun_out/c5_contrk4_ncu.log; wc -l gpurun_out/c5_contrk4_launches.csv
