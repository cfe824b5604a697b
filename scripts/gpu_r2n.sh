#!/bin/bash
# round 2, visit N: level-2 saving with fp16 hi words (no conversions) + mixed-format weight gradient
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage.py tests/test_gpu_dopri5_parity.py -x -q -m gpu -s > gpurun_out/r2n_pytest.log 2>&1
echo "pytest exit $?"; grep -E "saved activations|oracle|passed|failed|Error" gpurun_out/r2n_pytest.log | tail -n 12
timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2n_step_all.log 2>&1
echo "step exit $?"; grep -A8 "^rep 2" gpurun_out/r2n_step_all.log
AB200_STAGE_TIMING_ONLY=1 AB200_STAGE_FLAGS=64 timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2n_step_all_f64.log 2>&1
echo "step (no stores) exit $?"; grep -A3 "^rep 2" gpurun_out/r2n_step_all_f64.log
