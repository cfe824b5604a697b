#!/bin/bash
# GPU visit: tensor-core continuous adjoint (rk4): parity tests, whole GPU suite, configs[4] bench lines (1M and 8M agents, one solve each)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint_tc.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_adjtc.log
tail -n 5 gpurun_out/pytest_adjtc.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c5 --adjoint-mode continuous-rk4 --agents 1000000 --steps 2 --warmup 1 \
  > gpurun_out/bench_c5_contrk4_1M.json 2> gpurun_out/bench_c5_contrk4_1M.err; echo "exit $?" >> gpurun_out/bench_c5_contrk4_1M.err
tail -n 3 gpurun_out/bench_c5_contrk4_1M.err; cut -c1-400 gpurun_out/bench_c5_contrk4_1M.json
timeout 900 python bench.py --workload c5 --adjoint-mode continuous-rk4 --steps 2 --warmup 1 \
  > gpurun_out/bench_c5_contrk4_8M.json 2> gpurun_out/bench_c5_contrk4_8M.err; echo "exit $?" >> gpurun_out/bench_c5_contrk4_8M.err
tail -n 3 gpurun_out/bench_c5_contrk4_8M.err; cut -c1-400 gpurun_out/bench_c5_contrk4_8M.json
