#!/bin/bash
# round 2, visit B: full GPU suite, smoke, bench lines, reference arms, ncu launch list + full captures of the top kernels
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_dopri5.json 2> gpurun_out/bench_c3_dopri5.err; echo "exit $?" >> gpurun_out/bench_c3_dopri5.err
timeout 400 python bench.py --impl reference-gpu --steps 3 --warmup 1 > gpurun_out/bench_c3_refgpu.json 2> gpurun_out/bench_c3_refgpu.err; echo "exit $?" >> gpurun_out/bench_c3_refgpu.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.err; echo "exit $?" >> gpurun_out/bench_c3_reference.err
# launch list of one training step over one 333,440-agent chunk (2 reps: warm + measured)
timeout 300 python scripts/prof_c3_step.py 333440 2 > gpurun_out/prof_c3_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c3_step.csv python scripts/prof_c3_step.py 333440 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
for K in stage_fwd2_tc_kernel:8 stage_bwd_tc_kernel:3 wgrad_tc_kernel:3; do
  name=${K%%:*}; skip=${K##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o /tmp/prof_$name python scripts/prof_c3_step.py 333440 1 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/raw_$name.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page details > gpurun_out/details_$name.txt 2>/dev/null
done
tail -n 40 gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/smoke.log; tail -n 3 gpurun_out/*.err; cat gpurun_out/prof_c3_plain.log
