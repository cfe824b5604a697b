#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint_tc.py -q 2>&1 | tail -6 > gpurun_out/pytest_adjtc.log
tail -n 3 gpurun_out/pytest_adjtc.log
timeout 600 python scripts/prof_c5_contrk4.py 1000000 2 > gpurun_out/c5_contrk4_kernel_shares.txt 2>&1
grep -E "^rep|kernel time|%" gpurun_out/c5_contrk4_kernel_shares.txt | head -10
timeout 900 python bench.py --workload c5 --adjoint-mode continuous-rk4 --steps 2 --warmup 1 \
  > gpurun_out/bench_c5_contrk4_8M.json 2> gpurun_out/bench_c5_contrk4_8M.err; echo "exit $?" >> gpurun_out/bench_c5_contrk4_8M.err
tail -n 2 gpurun_out/bench_c5_contrk4_8M.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_contrk4_8M.json')); print(d['ms_per_step'], d['agent_days_per_s'], d['value'], d['peak_mem_gb'], d['e2e']['agent_days_per_s'])"
