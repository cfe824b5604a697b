#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
for wl in "c2 bf16" "c2 f32"; do
  set -- $wl
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $1 --precision $2 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "bench exit: $?" >> gpurun_out/bench_$1_$2.err
done
timeout 600 python bench.py --steps 2 --warmup 3 --workload c3 --agents 131072 --precision bf16 --no-cpu-baseline > gpurun_out/bench_c3s_bf16.json 2> gpurun_out/bench_c3s_bf16.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --workload c2 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err
tail -n 12 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/*.err
