#!/usr/bin/env python
"""Driver for ncu: one slice of delayed local updates (round kernels + rank-K flushes) of the C3 batch, bracketed
by cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees only these launches (the flush kernel is
also used by the blocked QR, which must not be captured instead):

  DQMC_LANES=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:zgemm_rank_update_kernel -s 2 -c 1 -o prof_flush python tools/profile_update_slice.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from bench import WORKLOAD, ladder_values  # noqa: E402
from detqmc_b200 import DetSDWBatch  # noqa: E402

R = int(os.environ.get("DQMC_PROF_R", "64"))
b = DetSDWBatch(dict(WORKLOAD), n_replicas=R, rng_indices=[i + 1 for i in range(R)], r_values=ladder_values(R))
b.sweepThermalization()
b.synchronize()
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
for k_ in range(1, 1 + int(os.environ.get("DQMC_PROF_SLICES", "1"))):
    acc = b.update_in_slice(k_, True)
b.synchronize()
rt.cudaProfilerStop()
print("accepted", acc[:8])
