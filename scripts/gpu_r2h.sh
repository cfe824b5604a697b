#!/bin/bash
# round 2, visit H: the other bench workloads / options still run after the round-2 changes
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -n 2 gpurun_out/$name.err; python -c "
import json
l=[x for x in open('gpurun_out/$name.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print('   ', d['metric'], round(d['value']/1e6,2), 'M a-s/s', round(d['ms_per_step'],1), 'ms', d.get('agent_days_per_s'), d['config'].get('solver_steps'))" 2>/dev/null; }
run bench_c5_small --workload c5 --agents 700000 --steps 1 --warmup 1 --no-cpu-baseline
run bench_c3_ce --loss ce --agents 333440 --steps 2 --warmup 1 --no-cpu-baseline
run bench_c3_f32_rk4 --precision f32 --solver rk4 --agents 100000 --steps 1 --warmup 1 --no-cpu-baseline
run bench_c2_f32 --workload c2 --precision f32 --no-cpu-baseline
run bench_c3_weak --scaling weak --agents 200000 --steps 1 --warmup 1 --no-cpu-baseline
