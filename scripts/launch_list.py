"""Per-launch list (in launch order) + per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` log:
   python scripts/launch_list.py launches.csv [--all]"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
ig, ib = hdr.index("Grid Size"), hdr.index("Block Size")
tot, seq = defaultdict(lambda: [0.0, 0]), []
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
    name = re.sub(r"^void |b200ldm::|\(.*", "", r[ik])[:70]
    tot[name][0] += v; tot[name][1] += 1
    seq.append((name, v, r[ig], r[ib]))
total = sum(v[0] for v in tot.values())
print(f"total {total / 1e3:.3f} ms over {len(seq)} launches")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {100 * us / total:5.1f}%  {us:9.1f} us  x{n:3d}  avg {us / n:7.1f}  {name}")
if "--all" in sys.argv:
    for i, (name, v, g, b) in enumerate(seq):
        print(f"[{i:3d}] {v:8.1f} us  grid {g:>18s} block {b:>12s}  {name}")
