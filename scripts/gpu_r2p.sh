#!/bin/bash
# round 2, visit P: source-level stall sampling of the activation-saving forward kernel
set -u
mkdir -p gpurun_out
name=stage_fwd2_tc_kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$name -s 8 -c 1 -f -o /tmp/prof_$name python scripts/prof_c3_step.py 250112 1 all > gpurun_out/r2p_ncu_$name.log 2>&1
echo "ncu $name exit $?"
ncu -i /tmp/prof_$name.ncu-rep --page source --csv > gpurun_out/r2p_source_$name.csv 2>/dev/null
ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/r2p_raw_$name.csv 2>/dev/null
ls -la gpurun_out | grep r2p
