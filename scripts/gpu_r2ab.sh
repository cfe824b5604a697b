#!/bin/bash
# round 2, visit AB (re-run of visit V with the final code): the other bench workloads / options after the bench refactor (saved operands, wave-aligned chunks, 3-kernel roofline)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -n 2 gpurun_out/$name.err; python -c "
import json
l=[x for x in open('gpurun_out/$name.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); r=d['roofline']; print('   ', round(d['value']/1e6,2), 'M a-s/s', round(d['ms_per_step'],1), 'ms days/s', round(d.get('agent_days_per_s') or 0), d['config'].get('solver_steps'), 'roof', r['bound'], round(r['frac'],3), r['kernel'][:30], 'mem', round(d['peak_mem_gb'],1))" 2>/dev/null; }
run r2ab_c3_inputs --saved-operands inputs --agents 300000 --steps 2 --warmup 1 --no-cpu-baseline
run r2ab_c3_none --saved-operands none --agents 300000 --steps 2 --warmup 1 --no-cpu-baseline
run r2ab_c3_rk4 --solver rk4 --steps 2 --warmup 1 --no-cpu-baseline
run r2ab_c2 --workload c2 --no-cpu-baseline
run r2ab_c5 --workload c5 --agents 1000000 --steps 1 --warmup 1 --no-cpu-baseline
run r2ab_c3_ce --loss ce --agents 265216 --steps 2 --warmup 1 --no-cpu-baseline
run r2ab_c3_f32 --precision f32 --agents 50000 --steps 1 --warmup 1 --no-cpu-baseline
