"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/r1_ncu_launches_v7_summary.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0]
    ms = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[iu], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(sys.argv[2], "w") as f:
    f.write("kernel,launches,total_ms,share\n")
    for n, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f'"{n}",{c},{ms:.3f},{ms / tot:.3f}\n')
        print(n[:72].ljust(72), c, round(ms, 3), round(ms / tot, 3))
    f.write(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.3f},1.0\n")
