#!/usr/bin/env python
"""Development tool: a short run through every kernel family (DetSDW L = 12 with measurements and a global move,
DetSDW O(3) L = 4, DetHubbard L = 16) for a quick run of everything, or under `compute-sanitizer --tool memcheck` where that tool is available."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from detqmc_b200 import DetSDWBatch, DetHubbardBatch  # noqa: E402
from dqmc_oracle import HubbardParams  # noqa: E402

b = DetSDWBatch(dict(opdim=2, L=12, m=20, s=10, globalUpdateInterval=2), n_replicas=2, rng_indices=[1, 2])
b.sweepThermalization()
b.sweep(True)
b.sweepThermalization()
print("sdw o2 L12", b.control_data(0).lastAccRatioLocal_phi, b.green_consistency())
b.close()
b = DetSDWBatch(dict(opdim=3, L=4, m=20, s=10, weakZflux=False), n_replicas=1)
b.sweepThermalization()
b.sweep(True)
print("sdw o3 L4", b.control_data(0).lastAccRatioLocal_phi)
b.close()
h = DetHubbardBatch(HubbardParams(L=16, m=20, s=10, U=4.0), n_replicas=1)
h.sweepThermalization()
h.sweep(True)
print("hubbard L16", sorted(h.observables(0))[:5])
h.close()
