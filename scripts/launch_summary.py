"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_summary.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*", "", r[ik])[:110]
    tot[name][0] += v
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total {total / 1e3:.1f} ms over {sum(v[1] for v in tot.values())} launches\n")
print("| share | us | launches | kernel |\n|---|---|---|---|")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"| {100 * us / total:.1f}% | {us:.0f} | {n} | `{name}` |")
