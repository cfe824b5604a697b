"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: python scripts/launch_shares.py <launches.csv> [header comment ...] > shares.csv"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, per = 0.0, defaultdict(lambda: [0.0, 0])
n = 0
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    k = r[ix["Kernel Name"]]
    per[k][0] += us
    per[k][1] += 1
    tot += us
    n += 1
for c in sys.argv[2:]:
    print("# " + c)
print(f"# total {tot:.1f} us over {n} launches")
print("share_pct,total_us,launches,avg_us,kernel")
for k, (us, c) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    print(f'{100 * us / tot:.2f},{us:.1f},{c},{us / c:.1f},"{k[:110]}"')
