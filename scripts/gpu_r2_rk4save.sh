#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_rk4.py tests/test_gpu_guards.py -x -q 2>&1 | tail -12 > gpurun_out/pytest_rk4.log
tail -n 8 gpurun_out/pytest_rk4.log
timeout 600 python bench.py --solver rk4 --no-cpu-baseline > gpurun_out/bench_c3_rk4.json 2> gpurun_out/bench_c3_rk4.err; echo "exit $?" >> gpurun_out/bench_c3_rk4.err
tail -n 2 gpurun_out/bench_c3_rk4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_rk4.json')); print(d['ms_per_step'], d['value'], d['agent_days_per_s'], d['peak_mem_gb'], d['config']['agent_chunks'])"
