#!/usr/bin/env python
"""Development tool: the hottest SASS lines of an .ncu-rep by warp-stall samples, with their stall reasons.
   tools/ncu_source_hot.py file.ncu-rep [top_n] [lo_addr hi_addr]"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data = []
for r in rows[1:]:
    try:
        n = int(r[ix["# Samples"]])
    except Exception:
        continue
    data.append((n, r))
total = sum(n for n, _ in data)
print("total samples", total)
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    sel = [(n, r) for n, r in data if lo <= int(r[ix["Address"]], 16) % 0x100000 <= hi]
    for n, r in sel:
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("%6d  %s  %-60s %s" % (n, r[ix["Address"]][-5:], r[ix["Source"]][:60], " ".join("%s=%d" % (c, v) for v, c in st if v)))
else:
    cum = 0
    for n, r in sorted(data, key=lambda t: -t[0])[:top]:
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print("%6d (%4.1f%%)  %s  %-56s %s" % (n, 100.0 * n / total, r[ix["Address"]][-5:], r[ix["Source"]][:56],
                                              " ".join("%s=%d" % (c, v) for v, c in st if v)))
