#!/bin/bash
# round 2, visit W: final evidence pass: full GPU suite, smoke, default bench line, reference arms
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/r2w_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2w_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/r2w_smoke.log
timeout 1200 python bench.py > gpurun_out/r2w_bench_c3.json 2> gpurun_out/r2w_bench_c3.err
echo "bench exit $?"; tail -n 2 gpurun_out/r2w_bench_c3.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_ref_cpu.json 2> gpurun_out/r2w_bench_ref_cpu.err
echo "ref cpu exit $?"; tail -n 1 gpurun_out/r2w_bench_ref_cpu.json | cut -c1-400
timeout 900 python bench.py --impl reference-gpu --steps 3 --warmup 1 > gpurun_out/r2w_bench_ref_gpu.json 2> gpurun_out/r2w_bench_ref_gpu.err
echo "ref gpu exit $?"; tail -n 1 gpurun_out/r2w_bench_ref_gpu.json | cut -c1-400
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2w_bench_c3.json') if l.startswith('{')][-1])
r=d['roofline']
print(round(d['ms_per_step'],1),'ms days/s', round(d['agent_days_per_s']), 'e2e', round(d['e2e']['agent_days_per_s']), 'roof', r['bound'], round(r['frac'],3), 'tw', round(r['stage_kernels_time_weighted_frac_hbm'],3), 'launches', d['gpu_launches'], 'mem', round(d['peak_mem_gb'],1), 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks'])"
