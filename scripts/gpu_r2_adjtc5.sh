#!/bin/bash
set -u
mkdir -p gpurun_out
for S in fused linear linear1 staged; do
  timeout 600 python scripts/prof_c5_contrk4.py 1000000 2 $S > gpurun_out/c5_contrk4_$S.txt 2>&1
  echo "== $S"; grep -E "^rep 1|kernel time|%" gpurun_out/c5_contrk4_$S.txt | head -7
done
