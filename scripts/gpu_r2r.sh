#!/bin/bash
# round 2, visit R: adjoint gather folded into the backward stage kernel (red.global), FSAL gradient as an upstream-only entry,
# dL/dy_path rows read in place by the step-level combine
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage.py tests/test_gpu_dopri5_parity.py tests/test_gpu_rk4.py tests/test_gpu_latent.py tests/test_gpu_losses.py -x -q -m gpu -s > gpurun_out/r2r_pytest.log 2>&1
echo "pytest exit $?"; grep -E "saved activations|oracle|passed|failed|Error" gpurun_out/r2r_pytest.log | tail -n 12
timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2r_step_all.log 2>&1
echo "step exit $?"; grep -A12 "^rep 2" gpurun_out/r2r_step_all.log
