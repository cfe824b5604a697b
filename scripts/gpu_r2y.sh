#!/bin/bash
# round 2, visit Y: side-stream error-norm read + 8x4 lane mapping of the step-level combine: parity tests, then the chunk-step timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage.py tests/test_gpu_dopri5_parity.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/r2y_pytest.log
timeout 600 python scripts/prof_c3_step.py 250112 3 all kineto > gpurun_out/r2y_idle_250k.txt 2>&1
echo "250k exit $?"; grep -E "rep 2|kernel time|pv_combine|device timeline|idle after Memcpy" gpurun_out/r2y_idle_250k.txt
timeout 600 python scripts/prof_c3_step.py 125056 3 all kineto > gpurun_out/r2y_idle_125k.txt 2>&1
echo "125k exit $?"; grep -E "rep 2|kernel time|pv_combine|device timeline|idle after Memcpy" gpurun_out/r2y_idle_125k.txt
