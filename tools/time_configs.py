#!/usr/bin/env python
"""Sweep times of the other BASELINE configs (parity-test cases, not bench lines): C2 DetSDW O(2) L=8 beta=8 single
replica, C4 DetSDW O(3) L=14 beta=14, C5 DetHubbard L=20 U=8 beta=20 -- seconds per sweep after two warm-up sweeps."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from detqmc_b200 import DetSDWBatch, DetHubbardBatch  # noqa: E402


def timed(b, n=3):
    for _ in range(2):
        b.sweepThermalization()
    b.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        b.sweepThermalization()
    b.synchronize()
    return (time.perf_counter() - t0) / n


which = sys.argv[1:] or ["C2", "C4", "C5"]
if "C2" in which:
    for R in (1, 16):
        b = DetSDWBatch(dict(opdim=2, L=8, m=80, s=10), n_replicas=R, rng_indices=list(range(1, R + 1)))
        print("C2 DetSDW O(2) L=8 beta=8, %2d replica(s): %.3f s per sweep" % (R, timed(b)))
if "C4" in which:
    for R in (1, 4):
        b = DetSDWBatch(dict(opdim=3, L=14, m=140, s=10, weakZflux=False), n_replicas=R, rng_indices=list(range(1, R + 1)))
        print("C4 DetSDW O(3) L=14 beta=14, %2d replica(s): %.3f s per sweep" % (R, timed(b, 2)))
if "C5" in which:
    from dqmc_oracle import HubbardParams
    for R in (1, 4):
        b = DetHubbardBatch(HubbardParams(L=20, m=200, s=10, U=8.0), n_replicas=R, rng_indices=list(range(1, R + 1)))
        print("C5 DetHubbard L=20 U=8 beta=20, %2d replica(s): %.3f s per sweep" % (R, timed(b, 2)))
