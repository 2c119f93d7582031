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
    win = re.compile(r"win dbg round (\d+) sites (\d+) acc (\d+): prologue (\d+) site (\d+) post (\d+) tail (\d+) wait (\d+)")
    bxy = re.compile(r"gth dbg round (\d+) (J) (\d+): coef (\d+) gload (\d+) product (\d+)")
    w = [0] * 8
    bx = {0: [0] * 5, 1: [0] * 5}
    for line in open(path):
        m = win.search(line)
        if m:
            v = [int(x) for x in m.groups()]
            if v[1] == 0:
                continue
            w[0] += 1
            for i in range(1, 8):
                w[i] += v[i]
        m = bxy.search(line)
        if m:
            v = [0 if x == "J" else int(x) for x in m.groups()]
            t = bx[v[1]]
            t[0] += 1
            t[1] += v[2]
            for i in range(3):
                t[2 + i] += v[3 + i]
    if w[0]:
        sites, acc = w[1], w[2]
        print("window kernel (warp 0 of replica 0): %d rounds, %d sites, %d accepted" % (w[0], sites, acc))
        print("  per round: prologue %.0f tail %.0f clocks" % (w[3] / w[0], w[6] / w[0]))
        print("  per site: proposal look-up + decision (incl. waiting for the diagonal blocks) %.0f clocks" % (w[4] / sites))
        print("  per accepted site: acceptance work %.0f clocks, waiting for the previous block update %.0f clocks" % (w[5] / max(acc, 1), w[7] / max(acc, 1)))
        print("  total per site %.0f clocks" % (sum(w[3:8]) / sites))
    for side, t in bx.items():
        if t[0]:
            print("gather (side %d): %d launches, mean J %.1f: coef %.0f gload %.0f product %.0f clocks per launch"
                  % (side, t[0], t[1] / t[0], t[2] / t[0], t[3] / t[0], t[4] / t[0]))


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
