#!/usr/bin/env python
"""Development tool: per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv):
   tools/launch_list_summary.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, mi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[start + 2:]:
    if len(r) <= mi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("dqmc::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    try:
        d[name].append(float(r[mi].replace(",", "")) / 1000.0)
    except ValueError:
        pass
tot = sum(sum(v) for v in d.values())
print("%-52s %6s %10s %8s %8s %8s %8s %8s" % ("kernel", "n", "sum us", "avg", "p10", "p50", "p90", "max"))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print("%-52s %6d %10.1f %8.1f %8.1f %8.1f %8.1f %8.1f" % (k[:52], len(v), sum(v), sum(v) / len(v), v2[len(v) // 10],
                                                           v2[len(v) // 2], v2[len(v) * 9 // 10], v2[-1]))
print("total %.1f us in %d launches" % (tot, sum(len(v) for v in d.values())))
