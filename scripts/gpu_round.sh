#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
timeout 600 python bench.py --solver rk4 > gpurun_out/bench_c3_rk4.json 2> gpurun_out/bench_c3_rk4.err; echo "exit $?" >> gpurun_out/bench_c3_rk4.err
timeout 300 python bench.py --workload c2 --precision bf16 > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "exit $?" >> gpurun_out/bench_c2_bf16.err
timeout 300 python bench.py --workload c2 --precision f32 > gpurun_out/bench_c2_f32.json 2> gpurun_out/bench_c2_f32.err; echo "exit $?" >> gpurun_out/bench_c2_f32.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.err
tail -n 12 gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/smoke.log; tail -n 1 gpurun_out/*.err
