"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name launches, total and mean
duration, share. Usage: python tools/launch_summary.py launches.csv [skip_first_n_launches] > summary.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
    rows.append((name, ns))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for name, ns in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print("# %d launches, %.3f ms summed kernel time" % (len(rows), tot / 1e6))
print("share,total_us,launches,mean_us,kernel")
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%.4f,%.1f,%d,%.2f,%s" % (ns / tot, ns / 1e3, cnt, ns / 1e3 / cnt, name))
