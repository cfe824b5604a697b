#!/bin/bash
# final-state verification of round 2 (after the continuous-adjoint restructuring)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
tail -n 1 gpurun_out/bench_c3_dopri5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_dopri5.json')); print(d['ms_per_step'], d['agent_days_per_s'], d['e2e']['agent_days_per_s'], d['roofline']['frac'], d['clocks'])"
for A in 1000000 8000000; do
timeout 900 python bench.py --workload c5 --adjoint-mode continuous-rk4 --agents $A --steps 2 --warmup 1 \
  > gpurun_out/bench_c5_contrk4_$A.json 2> gpurun_out/bench_c5_contrk4_$A.err; echo "exit $?" >> gpurun_out/bench_c5_contrk4_$A.err
tail -n 1 gpurun_out/bench_c5_contrk4_$A.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_contrk4_$A.json')); print($A, d['ms_per_step'], d['agent_days_per_s'], d['value'], d['peak_mem_gb'], d['e2e']['agent_days_per_s'])"
done
timeout 600 python scripts/prof_c5_contrk4.py 1000000 2 > gpurun_out/c5_contrk4_kernel_shares.txt 2>&1
grep -E "^rep 1|kernel time|%" gpurun_out/c5_contrk4_kernel_shares.txt | head -7
