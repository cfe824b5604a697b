#!/bin/bash
# round 2, visit C: full GPU suite after the VJP / adjoint / elementwise changes, bench lines
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
timeout 600 python bench.py --solver rk4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_rk4.json 2> gpurun_out/bench_c3_rk4.err; echo "exit $?" >> gpurun_out/bench_c3_rk4.err
timeout 300 python bench.py --workload c2 --precision bf16 > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "exit $?" >> gpurun_out/bench_c2_bf16.err
timeout 300 python bench.py --agents 125000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_125k.json 2> gpurun_out/bench_c3_125k.err; echo "exit $?" >> gpurun_out/bench_c3_125k.err
timeout 300 python scripts/prof_c3_step.py 333440 3 > gpurun_out/prof_c3_plain.log 2>&1
tail -n 30 gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/smoke.log; for f in gpurun_out/*.err; do echo $f; tail -n 3 $f; done; cat gpurun_out/prof_c3_plain.log; head -c 1500 gpurun_out/bench_c3_dopri5.json; echo; head -c 600 gpurun_out/bench_c3_125k.json
