#!/usr/bin/env python
"""Development tool: print value / families of bench.py JSON lines (gpurun_out/*.json)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e)
        continue
    print(f, "value %.1f" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"])
    fam = d["roofline"].get("families", {})
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        print("   %-16s %8.2f ms  %6.0f launches  avg %7.1f us  frac %s" % (
            k, v["ms_per_step"], v["launches_per_step"], 1e3 * v["avg_launch_ms"],
            ("%.3f" % v["frac"]) if v.get("frac") is not None else None))
