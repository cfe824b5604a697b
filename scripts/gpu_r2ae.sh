#!/bin/bash
# round 2, visit AE: evidence pass with the final code (after the seminorm option and the bench adjoint modes): full GPU suite, smoke, default
# bench line, reference arms, the 125,000-agent line (the per-GPU load at N = 8)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ae_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/r2ae_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ae_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/r2ae_smoke.log
timeout 1200 python bench.py > gpurun_out/r2ae_bench_c3.json 2> gpurun_out/r2ae_bench_c3.err
echo "bench exit $?"; tail -n 2 gpurun_out/r2ae_bench_c3.err
timeout 600 python bench.py --agents 125000 --no-cpu-baseline > gpurun_out/r2ae_bench_c3_125k.json 2> gpurun_out/r2ae_bench_c3_125k.err
echo "bench 125k exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ae_bench_ref_cpu.json 2> gpurun_out/r2ae_bench_ref_cpu.err
echo "ref cpu exit $?"
timeout 900 python bench.py --impl reference-gpu --steps 3 --warmup 1 > gpurun_out/r2ae_bench_ref_gpu.json 2> gpurun_out/r2ae_bench_ref_gpu.err
echo "ref gpu exit $?"
python -c "
import json
for f in ('r2ae_bench_c3','r2ae_bench_c3_125k','r2ae_bench_ref_cpu','r2ae_bench_ref_gpu'):
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    r=d.get('roofline',{})
    print(f, round(d['ms_per_step'],1),'ms days/s', round(d['agent_days_per_s']), 'e2e', round(d['e2e'].get('agent_days_per_s',0)), 'roof', r.get('bound'), round(r.get('frac',0),3), 'tw', round(r.get('stage_kernels_time_weighted_frac_hbm',0),3), 'launches', d.get('gpu_launches'), 'mem', d.get('peak_mem_gb'), d.get('clocks'))"
