"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            k = d["Kernel Name"][:70]
            agg[k][0] += 1
            agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%4d %10.1f us %9.1f us/launch  %s" % (n, t, t / n, k))
