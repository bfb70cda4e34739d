"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: ncu_launches.py file.csv [last_n]"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]
kn, mv = cols.index("Kernel Name"), cols.index("Metric Value")
body = rows[hdr + 1:]
if len(sys.argv) > 2:
    body = body[-int(sys.argv[2]):]
agg = OrderedDict()
for r in body:
    name = r[kn].split("(")[0]
    t = float(r[mv].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
for name, (c, t) in agg.items():
    print(f"{name[:70]:70s} launches {c:4d}  mean {t / c / 1e3:9.2f} us")
