#!/bin/bash
# final-state verification: GPU suite, smoke, default bench, reference arms
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
tail -n 1 gpurun_out/bench_c3_dopri5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_dopri5.json')); print(d['ms_per_step'], d['agent_days_per_s'], d['e2e']['agent_days_per_s'], d['roofline']['frac'], [k['kernel_ms'] for k in d['roofline']['stage_kernels']], d['clocks'])"
timeout 400 python bench.py --impl reference > gpurun_out/bench_c3_reference_cpu.json 2> gpurun_out/bench_c3_reference_cpu.err; echo "exit $?" >> gpurun_out/bench_c3_reference_cpu.err
cut -c1-300 gpurun_out/bench_c3_reference_cpu.json
