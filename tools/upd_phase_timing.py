#!/usr/bin/env python
"""Development tool: per-phase clock counts of the update round kernel (replica 0 of the C3 batch).

Needs a library built with the counters compiled in (they cost registers, so the shipped build has none):

  python tools/upd_phase_timing.py --build        # here: nvcc -DDQMC_UPD_TIMING -> detqmc_b200/libdqmc_b200_timing.so
  python tools/upd_phase_timing.py > timing.txt   # on the GPU box; then  --summarise timing.txt

Each round prints, for the decision thread (tid 0), the stager (tid 32) and the first gather thread (tid 64),
the clocks spent in phase 1, waiting at barrier A, in phase 2, waiting at barrier B, and in the loop head.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
TIMING_LIB = os.path.join(ROOT, "detqmc_b200", "libdqmc_b200_timing.so")


def build():
    from detqmc_b200 import build as b
    objs = []
    for src in b.SOURCES:
        obj = os.path.join("/tmp", "timing_" + os.path.splitext(src)[0] + ".o")
        subprocess.check_call(["nvcc"] + b.NVCC_FLAGS + ["-DDQMC_UPD_TIMING", "-c", os.path.join(b.CSRC, src), "-o", obj])
        objs.append(obj)
    subprocess.check_call(["nvcc", "-shared", "-Wno-deprecated-gpu-targets", "-o", TIMING_LIB] + objs)
    print("built", TIMING_LIB)


def summarise(path):
    pat = re.compile(r"upd dbg round (\d+) tid +(\d+) sites (\d+): ph1 (\d+) waitA (\d+) ph2\(dec\) (\d+) waitB (\d+) Sred (\d+) top (\d+)")
    tot = {}
    for line in open(path):
        m = pat.search(line)
        if not m:
            continue
        tid, sites = int(m.group(2)), int(m.group(3))
        v = [int(x) for x in m.groups()[3:]]
        t = tot.setdefault(tid, [0] * 7)
        t[0] += sites
        for i, x in enumerate(v):
            t[1 + i] += x
    print("clocks per site (replica 0):  tid  sites   ph1  waitA   ph2  waitB  (Sred)   top   sum")
    for tid, t in sorted(tot.items()):
        n = max(t[0], 1)
        per = [x / n for x in t[1:]]
        print("  tid %3d  %6d  %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f  %6.0f" %
              (tid, t[0], per[0], per[1], per[2], per[3], per[4], per[5], per[0] + per[1] + per[2] + per[3] + per[5]))


def run():
    import detqmc_b200.lib as lib
    lib.LIB_PATH = TIMING_LIB
    from bench import WORKLOAD, ladder_values
    from detqmc_b200 import DetSDWBatch
    R = int(os.environ.get("TIMING_REPLICAS", "64"))
    b = DetSDWBatch(dict(WORKLOAD), n_replicas=R, rng_indices=[i + 1 for i in range(R)], r_values=ladder_values(R))
    for _ in range(3):
        b.sweepThermalization()
    b.synchronize()
    os.environ["DQMC_UPD_DEBUG"] = "1"
    for k in (1, 2, 3, 4):
        b.update_in_slice(k, True)
        b.synchronize()


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    elif "--summarise" in sys.argv:
        summarise(sys.argv[sys.argv.index("--summarise") + 1])
    else:
        run()
