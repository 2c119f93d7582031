#!/usr/bin/env python
"""Development tool: cost of a measurement sweep (sweep(true): fermionic observables with the sparse
shiftGreenSymmetric) relative to a plain sweep, C3 batch.  tools/time_measure_sweep.py [replicas]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from bench import WORKLOAD, ladder_values  # noqa: E402
from detqmc_b200 import DetSDWBatch  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
b = DetSDWBatch(dict(WORKLOAD), n_replicas=R, rng_indices=[i + 1 for i in range(R)], r_values=ladder_values(R))
for _ in range(3):
    b.sweepThermalization()


def timed(meas, n=6):
    for _ in range(2):
        b.sweep(meas)
    b.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        b.sweep(meas)
    b.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


plain = timed(False)
meas = timed(True)
print("%d replicas: plain sweep %.1f ms, measurement sweep %.1f ms, ratio %.3f" % (R, plain, meas, meas / plain))
