#!/bin/bash
# round 2, visit G: pipelined split-activation forward (default) vs the un-pipelined chain (AB200_STAGE_FLAGS=128)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dopri5_parity.py tests/test_gpu_stage.py -m gpu -q -x 2>&1 | tail -8
for F in 0 128; do
  AB200_STAGE_FLAGS=$F timeout 300 python scripts/prof_c3_step.py 333440 3 2>/dev/null | grep rep
  AB200_STAGE_FLAGS=$F timeout 300 python scripts/prof_c3_step.py 125056 3 2>/dev/null | grep rep
done
