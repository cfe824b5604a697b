#!/bin/bash
# round 2, visit E: L2 prefetch of the backward kernel's gather phases on / off
set -u
mkdir -p gpurun_out
for F in 0 64; do
  AB200_STAGE_FLAGS=$F timeout 300 python bench.py --steps 2 --warmup 2 --agents 333440 --no-cpu-baseline > gpurun_out/flag_$F.json 2> gpurun_out/flag_$F.err
  python -c "
import json
d=json.loads([x for x in open('gpurun_out/flag_$F.json').read().splitlines() if x.startswith('{')][-1]); print('flags=$F', 'step', d['ms_per_step'], 'ms; bwd kernel', d['roofline']['kernel_ms'], 'ms')"
done
timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_dopri5_parity.py -m gpu -q -x 2>&1 | tail -3
