"""Summarise `ncu --page source --print-source cuda,sass --csv` output: stall samples and executed
instructions per CUDA source line, per kernel function.   usage: srcprof.py file.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sections, cur, fpath = [], None, ""
for r in rows:
    if r and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r and r[0] == "Function Name":
        cur = {"name": r[1], "file": fpath, "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
merged = collections.OrderedDict()
for sec in sections:
    m = merged.setdefault(sec["name"], {"agg": collections.Counter(), "smp": collections.Counter()})
    for r in sec["rows"]:
        if not r:
            continue
        try:
            line = int(r[0]); inst = int(r[7]) if r[7] else 0; s = int(r[6]) if r[6] else 0
        except (ValueError, IndexError):
            continue
        key = (sec["file"], line, r[1].strip()[:92])
        m["agg"][key] += inst
        m["smp"][key] += s
for name, m in merged.items():
    tot, ts = sum(m["agg"].values()), sum(m["smp"].values())
    print("=====", name[:70], "warp-inst", tot, "samples", ts)
    for k, v in m["smp"].most_common(top):
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * v / max(ts, 1), 100 * m["agg"][k] / max(tot, 1), k[0], k[1], k[2]))
