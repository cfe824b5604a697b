#!/bin/bash
set -u
mkdir -p gpurun_out
for wl in "c2 bf16" "c2 f32"; do
  set -- $wl
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $1 --precision $2 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "bench exit: $?" >> gpurun_out/bench_$1_$2.err
done
timeout 900 python bench.py --steps 2 --warmup 3 --workload c3 --agents 131072 --precision bf16 --no-cpu-baseline > gpurun_out/bench_c3s_bf16.json 2> gpurun_out/bench_c3s_bf16.err
echo "bench exit: $?" >> gpurun_out/bench_c3s_bf16.err
tail -n 3 gpurun_out/bench_*.err
