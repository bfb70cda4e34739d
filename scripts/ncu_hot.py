"""Summarise an ncu source-page CSV: hottest SASS instructions by stall samples.
   ncu -i X.ncu-rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si = hdr.index("Warp Stall Sampling (All Samples)")
ei = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
total = sum(int(r[si] or 0) for r in body)
print(f"total samples {total}, instructions {len(body)}")
agg = {}
for r in body:
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            agg[h] = agg.get(h, 0) + v
print("stall reasons:", sorted(agg.items(), key=lambda x: -x[1])[:12])
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    reasons = sorted(((int(r[c] or 0), h) for c, h in stall_cols if int(r[c] or 0)), reverse=True)[:3]
    print(f"{i:6d} {int(r[si]):7d} {100.0 * int(r[si]) / max(total, 1):5.1f}% exec={r[ei]:>9} {r[1].strip()[:70]:70s} {reasons}")
