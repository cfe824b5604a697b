#!/bin/bash
# round 2, visit T: evidence pass with the current code: full GPU suite, smoke, default bench line (with cpu_baseline), 125k-agent line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/r2t_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/r2t_smoke.log
timeout 1200 python bench.py > gpurun_out/r2t_bench_c3.json 2> gpurun_out/r2t_bench_c3.err
echo "bench exit $?"; tail -n 2 gpurun_out/r2t_bench_c3.err
timeout 600 python bench.py --agents 125000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_c3_125k.json 2> gpurun_out/r2t_bench_c3_125k.err
echo "bench 125k exit $?"
for f in r2t_bench_c3 r2t_bench_c3_125k; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][-1])
r=d['roofline']
print('$f', round(d['ms_per_step'],1),'ms days/s', round(d['agent_days_per_s']), 'e2e', round(d['e2e']['agent_days_per_s']), 'roof', r['bound'], round(r['frac'],3), r['kernel'][:40], 'launches', d['gpu_launches'], 'mem', round(d['peak_mem_gb'],1), d.get('cpu_baseline',{}).get('value'))
for k in (r.get('stage_kernels') or []): print('   ', k['kernel'][:28], round(k['kernel_ms'],3), 'ms hbm', round(k['frac_hbm'],3), 'tensor', round(k['frac_tensor'],3), 'dram', round(k['dram_gbs'] or 0))"; done
