#!/bin/bash
# round 2, visit A: the split-activation forward kernel -- parity at 1e-5, bench line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_dopri5_parity.py -m gpu -q -s --durations=6 2>&1 | tail -60 > gpurun_out/pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_rk4.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_stage.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
timeout 300 python bench.py --impl reference-gpu --steps 3 --warmup 1 > gpurun_out/bench_c3_refgpu.json 2> gpurun_out/bench_c3_refgpu.err; echo "exit $?" >> gpurun_out/bench_c3_refgpu.err
tail -n 30 gpurun_out/pytest_parity.log; tail -n 6 gpurun_out/pytest_stage.log; tail -n 3 gpurun_out/*.err; cat gpurun_out/bench_c3_dopri5.json | head -c 3000
