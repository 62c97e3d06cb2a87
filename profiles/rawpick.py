"""Pick metrics out of `ncu -i X.ncu-rep --page raw --csv` (stdin): one block per profiled launch.
usage: ncu -i rep --page raw --csv | python profiles/rawpick.py '<regex over metric names>'"""
import csv
import re
import sys

pat = re.compile(sys.argv[1])
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
if len(rows) < 3:
    sys.exit("no launches in the report")
names, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(names, r))
    print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:80], d.get("Grid Size", "?"), d.get("Block Size", "?")))
    for n, u, v in zip(names, units, r):
        if pat.search(n):
            print("  %-70s %16s %s" % (n, v, u))
