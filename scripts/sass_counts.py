"""Static SASS instruction counts per tensor-core kernel of libananke_b200.so (cuobjdump -sass): tcgen05.mma (UTC*MMA), tcgen05.ld/st
(LDTM / STTM), bulk copies (UBLKCP), legacy HMMA.  usage: python scripts/sass_counts.py > profiles/r02_sass_counts.txt"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "ananke_abm_b200/libananke_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
print("# static SASS instruction counts per kernel of libananke_b200.so (cuobjdump -sass | grep); loops are rolled, so counts are per code path, not per launch")
print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA 1-D); regenerated with the final code of round 2 (scripts/sass_counts.py)")
for name, body in zip(names, blocks):
    c = {k: len(re.findall(p, body)) for k, p in (("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UBLKCP", r"\bUBLKCP\b"),
                                                   ("legacy-HMMA", r"\bHMMA\b"))}
    if c["UTC*MMA"] or c["LDTM"] or c["legacy-HMMA"]:
        print(f"{name[:100]:100s} " + "  ".join(f"{k} {v:4d}" for k, v in c.items()))
