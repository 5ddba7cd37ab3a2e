"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_summary.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, agg, total = None, collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
        agg[d["Kernel Name"]][0] += 1
        agg[d["Kernel Name"]][1] += v
        total += v
print(f"total {total:.3f} ms over {sum(c for c, _ in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t:10.3f} ms {100 * t / total:5.1f}% {c:5d}  {n[:120]}")
