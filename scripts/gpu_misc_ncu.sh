#!/bin/bash
# ncu --set full of the sparse / elementwise / head / fp32 kernels (1 GPU), only after the same program exited 0 without ncu.
# The report stays on the box (gpurun_out/ is capped at 64 MiB): only the raw-page CSV travels back.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/prof_misc.py 189440 2 > gpurun_out/prof_misc_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_misc_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none \
  -k 'regex:gat_|pv_combine|adjoint_gather|ga_assemble|rows_transpose|head_|combine_errnorm|stage_combine|emb_losses|rk4_bwd_f32|drift_eval|reduce_unpack' -c 140 \
  -f -o /tmp/prof_misc python scripts/prof_misc.py 189440 1 > gpurun_out/prof_misc_ncu.log 2>&1
echo "ncu exit: $?"; tail -2 gpurun_out/prof_misc_ncu.log
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/prof_misc_raw.csv 2>/dev/null
ls -la gpurun_out/prof_misc_raw.csv
python scripts/ncu_misc_summary.py gpurun_out/prof_misc_raw.csv | head -40
