#!/bin/bash
# round 2, visit D: C-side attempt builders + reordered dense rows; source-level ncu of the forward and backward stage kernels
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
timeout 300 python bench.py --agents 125000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_125k.json 2> gpurun_out/bench_c3_125k.err; echo "exit $?" >> gpurun_out/bench_c3_125k.err
timeout 300 python scripts/prof_c3_step.py 333440 2 > gpurun_out/prof_c3_plain.log 2>&1 && \
for K in stage_fwd2_tc_kernel:8 stage_bwd_tc_kernel:3; do
  name=${K%%:*}; skip=${K##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o /tmp/prof_$name python scripts/prof_c3_step.py 333440 1 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv > gpurun_out/source_$name.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/raw_$name.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
tail -n 6 gpurun_out/pytest_gpu.log; for f in gpurun_out/bench*.err; do echo $f; tail -n 2 $f; done; cat gpurun_out/prof_c3_plain.log
for f in bench_c3_dopri5 bench_c3_125k; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'],'ms', 'days/s', d.get('agent_days_per_s'), 'steps', d['config']['solver_steps']['accepted_per_trajectory'], d['config']['solver_steps']['rejected_per_trajectory'])"; done
