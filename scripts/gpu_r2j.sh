#!/bin/bash
# round 2, visit J: launch list of the chunk step after the forward-saved stage-input blobs
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage.py -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/r2j_pytest.log
timeout 600 python scripts/prof_c3_step.py 333440 3 > gpurun_out/r2j_step.log 2>&1
echo "step exit $?"; tail -n 4 gpurun_out/r2j_step.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2j_launches.csv python scripts/prof_c3_step.py 333440 2 > gpurun_out/r2j_ncu.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/r2j_ncu.log
