"""Summarise an ncu report: per-kernel headline metrics and the top stall locations (SASS) of one kernel.
usage: python scripts/ncu_top.py <report.ncu-rep> [kernel-regex] [launch-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "launch__registers_per_thread", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
for n, r in enumerate(rows[2:]):
    print(f"[{n}] {r[idx['Kernel Name']][:50]} grid {r[idx['Grid Size']] if 'Grid Size' in idx else ''}")
    for k in keys:
        if k in idx:
            print(f"      {k:72s} {r[idx[k]]} {units[idx[k]]}")
if kre:
    li = sys.argv[3] if len(sys.argv) > 3 else "0"
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", li, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[idx['# Samples']]) for r in data if r[idx['# Samples']].isdigit())
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {}
    for r in data:
        for c in stall_cols:
            if r[idx[c]].isdigit():
                agg[c] = agg.get(c, 0) + int(r[idx[c]])
    print("total samples", tot, sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    top = sorted(data, key=lambda r: -int(r[idx['# Samples']]) if r[idx['# Samples']].isdigit() else 0)[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]
    for r in top:
        s = int(r[idx['# Samples']])
        st = sorted(((int(r[idx[c]]), c) for c in stall_cols if r[idx[c]].isdigit() and int(r[idx[c]]) > 0), reverse=True)[:2]
        print(f"{100 * s / max(tot, 1):5.1f}% {r[idx['Source']].strip()[:80]:80s} {st}")
