"""Headline metrics per captured launch from an `ncu --page raw --csv` export.   usage: python scripts/ncu_raw_summary.py <raw.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
ix = {h: i for i, h in enumerate(H)}
want = [("duration", "gpu__time_duration.sum"),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM written", "dram__bytes_write.sum"),
        ("DRAM throughput % of ncu peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
        ("registers/thread", "launch__registers_per_thread"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("stall long_scoreboard (warps per issue)", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("stall lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("stall short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")]
for r in rows[2:]:
    if len(r) != len(H):
        continue
    print(f"== {r[ix['Kernel Name']]} grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}")
    for name, k in want:
        if k in ix:
            print(f"   {name:45s} {r[ix[k]]} {U[ix[k]]}")
    print()
