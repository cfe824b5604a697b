"""One line per kernel from an `ncu --set full --page raw --csv` export: the longest of its captured launches, DRAM bytes, achieved GB/s.
usage: python scripts/ncu_misc_summary.py raw.csv [header comment ...]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0
units = rows[1]
def to_us(v, u): return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
def to_mb(v, u): return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
best, cnt = {}, collections.Counter()
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("ab200::", "").replace("void ", "")
    cnt[name] += 1
    d = to_us(num(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]])
    if name not in best or d > best[name][0]:
        rd = to_mb(num(r, "dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]])
        wr = to_mb(num(r, "dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]])
        best[name] = (d, rd, wr, r[ix["Grid Size"]], num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
for c in sys.argv[2:]:
    print("# " + c)
for name, (d, rd, wr, grid, pct, tens) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    print(f"{name[:52]:52s} n={cnt[name]:3d} grid={grid:14s} dur={d:9.1f}us dram r/w={rd:9.2f}/{wr:9.2f} MB -> {(rd + wr) / max(d, 1e-9) * 1e3:7.0f} GB/s ({pct:5.1f}% dram peak, tensor {tens:4.1f}%)")
