"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: launchsum.py file.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"][:64]
    v = float(row["Metric Value"])
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-66s n=%4d total=%10.1f us  avg=%9.1f us  share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
print("total us %.1f" % tot)
