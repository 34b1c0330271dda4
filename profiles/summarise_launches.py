"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = r[ik].split("(")[0]
    a = agg[name]
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:48]:48s} n={a[0]:6d} total={a[1] / 1e3:10.2f} ms avg={a[1] / a[0]:10.1f} us max={a[2] / 1e3:8.3f} ms share={100 * a[1] / tot:5.1f}%")
print(f"total {tot / 1e3:.1f} ms over {sum(a[0] for a in agg.values())} launches")
