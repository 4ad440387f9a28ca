"""Aggregates an `ncu --csv --metrics gpu__time_duration.sum[,...]` launch list per kernel name.

    python scripts/launch_summary.py gpurun_out/launches.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        if d["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
        agg[d["Kernel Name"].split("(")[0][-40:]][d["Metric Name"]].append(v)
tot = sum(sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]) for m in agg.values())
for k, m in agg.items():
    t = m["gpu__time_duration.sum"]
    line = f"{k:42s} n={len(t):3d} mean={sum(t) / len(t):8.1f} us  share={sum(t) / len(t) / tot * 100:5.1f}%"
    for name, vals in m.items():
        if name != "gpu__time_duration.sum":
            line += f"  {name}={sum(vals) / len(vals):.3g}"
    print(line)
