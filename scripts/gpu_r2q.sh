#!/bin/bash
# round 2, visit Q: full GPU suite + default bench line with saved_operands=all and wave-aligned chunks
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/r2q_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_c3.json 2> gpurun_out/r2q_bench_c3.err
echo "bench exit $?"; tail -n 2 gpurun_out/r2q_bench_c3.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2q_bench_c3.json') if l.startswith('{')][-1])
print(d['ms_per_step'],'ms', d['value'], d.get('agent_days_per_s'), d['e2e'], d['roofline'], d.get('peak_mem_gb'), d['config'].get('agent_chunks'))"
