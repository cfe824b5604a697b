#!/bin/bash
# final-state verification of round 2: GPU suite, smoke, default bench, kernel shares (CUPTI) and ncu launch list of one chunk-step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
tail -n 1 gpurun_out/bench_c3_dopri5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_dopri5.json')); print(d['ms_per_step'], d['agent_days_per_s'], d['e2e']['agent_days_per_s'], d['roofline']['frac'], [k['kernel_ms'] for k in d['roofline']['stage_kernels']], d['clocks'])"
timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/c3_kernel_shares.txt 2>&1
grep -E "^rep 2|kernel time|%" gpurun_out/c3_kernel_shares.txt | head -9
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c3_launches.csv \
  python scripts/prof_c3_step.py 250112 1 all > gpurun_out/c3_ncu.log 2>&1
tail -n 2 gpurun_out/c3_ncu.log; wc -l gpurun_out/c3_launches.csv
