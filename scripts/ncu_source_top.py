"""Top SASS instructions by warp-stall samples from an `ncu --page source --csv` export, with the dominant stall reason and
the two preceding instructions for context.   usage: python scripts/ncu_source_top.py <source.csv> [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ix = {h: i for i, h in enumerate(H)}
stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hdr + 1:] if len(r) == len(H)]
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0
tot = sum(num(r[ix["# Samples"]]) for r in body)
order = sorted(range(len(body)), key=lambda i: -num(body[i][ix["# Samples"]]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"total samples {tot:.0f}")
for i in order[:n]:
    r = body[i]
    s = num(r[ix["# Samples"]])
    reasons = sorted(((num(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    ctx = " <- ".join(body[j][ix["Source"]].split(";")[0].strip()[:44] for j in range(i - 1, max(-1, i - 3), -1))
    print(f"{100 * s / tot:5.1f}%  #{i:5d}  {r[ix['Source']].split(';')[0].strip()[:60]:60s} {reasons[0][1]}:{reasons[0][0]:.0f} {reasons[1][1]}:{reasons[1][0]:.0f}   [{ctx}]")
