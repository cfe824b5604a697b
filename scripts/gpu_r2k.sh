#!/bin/bash
# round 2, visit K: forward-saved activations + masks (backward recomputes nothing)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage.py tests/test_gpu_dopri5_parity.py -x -q -m gpu -s > gpurun_out/r2k_pytest.log 2>&1
echo "pytest exit $?"; grep -E "saved activations|oracle|passed|failed|Error" gpurun_out/r2k_pytest.log | tail -n 12
for mode in inputs all; do
  timeout 600 python scripts/prof_c3_step.py 250112 4 $mode > gpurun_out/r2k_step_$mode.log 2>&1
  echo "step $mode exit $?"; grep "^rep" gpurun_out/r2k_step_$mode.log
done
