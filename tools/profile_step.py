#!/usr/bin/env python
"""Small driver for ncu: build the C3 batch and run a few sweeps (+ exchange packing)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
from bench import WORKLOAD, ladder_values  # noqa: E402
from detqmc_b200 import DetSDWBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=64)
ap.add_argument("--sweeps", type=int, default=2)
ap.add_argument("--L", type=int, default=12)
ap.add_argument("--m", type=int, default=100)
args = ap.parse_args()
w = dict(WORKLOAD)
w.update(L=args.L, m=args.m)
vals = ladder_values(args.replicas)
b = DetSDWBatch(w, n_replicas=args.replicas, rng_indices=[i + 1 for i in range(args.replicas)], r_values=vals)
print("setup launches", b.launch_count)
for s in range(args.sweeps):
    b.sweepThermalization()
    print("sweep", s, "launches", b.launch_count, "acc", b.control_data(0).lastAccRatioLocal_phi)
