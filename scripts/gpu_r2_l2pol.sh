#!/bin/bash
# L2 residency experiment (AB200_STAGE_FLAGS bit 128): evict_last on the a_j / gx working set of a fused launch, evict_first at the last use
set -u
mkdir -p gpurun_out
AB200_STAGE_FLAGS=128 timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_dopri5_parity.py -x -q 2>&1 | tail -4 > gpurun_out/l2pol_pytest.log
tail -n 2 gpurun_out/l2pol_pytest.log
for F in 0 128 0 128; do
  AB200_STAGE_FLAGS=$F timeout 600 python scripts/prof_c3_step.py 250112 4 all kineto > gpurun_out/l2pol_flags${F}_$RANDOM.txt 2>&1
done
grep -h "rep 3\|stage_fwd2\|stage_bwd_tc\|kernel time" gpurun_out/l2pol_flags*.txt
