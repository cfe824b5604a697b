#!/bin/bash
# round 2, visit I: forward-saved stage-input blobs (backward loads them instead of rebuilding from y0 + a_j)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dopri5_parity.py tests/test_gpu_stage.py tests/test_gpu_latent.py tests/test_gpu_guards.py -x -q -m gpu -s > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?"; tail -n 12 gpurun_out/r2i_pytest.log
timeout 600 python scripts/prof_c3_step.py 333440 4 > gpurun_out/r2i_step.log 2>&1
echo "step exit $?"; tail -n 6 gpurun_out/r2i_step.log
