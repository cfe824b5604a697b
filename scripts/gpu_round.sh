#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --workload c2 > gpurun_out/bench_c2_f32.json 2> gpurun_out/bench_c2_f32.err
echo "bench exit: $?" >> gpurun_out/bench_c2_f32.err
timeout 600 python bench.py --steps 2 --warmup 3 --workload c3 --agents 65536 --no-cpu-baseline > gpurun_out/bench_c3s_f32.json 2> gpurun_out/bench_c3s_f32.err
echo "bench exit: $?" >> gpurun_out/bench_c3s_f32.err
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_c2_f32.json gpurun_out/bench_c2_f32.err gpurun_out/bench_c3s_f32.json gpurun_out/bench_c3s_f32.err
