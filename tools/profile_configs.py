#!/usr/bin/env python
"""Development tool: per-kernel-family device time of one sweep of the non-headline BASELINE configs (C2, C4, C5),
one replica, through the library's CUDA-event profile (dqmc_profile_*)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from detqmc_b200 import DetSDWBatch, DetHubbardBatch  # noqa: E402
from detqmc_b200.lib import c_vp  # noqa: E402


def profile(b, label, sweeps=2):
    lib = b.lib
    for _ in range(2):
        b.sweepThermalization()
    b.synchronize()
    lib.dqmc_set_option(b.h, 1, 1)
    lib.dqmc_profile_enable(b.h, 1)
    for _ in range(sweeps):
        b.sweepThermalization()
    ncat = 10
    ms = np.zeros(ncat)
    cnt = np.zeros(ncat, dtype=np.uint64)
    lib.dqmc_profile_get(b.h, ms.ctypes.data_as(c_vp), cnt.ctypes.data_as(c_vp))
    print(label, "total %.1f ms per sweep" % (ms.sum() / sweeps))
    for i in np.argsort(-ms):
        if cnt[i]:
            print("   %-18s %9.2f ms %7d launches" % (lib.dqmc_profile_name(int(i)).decode(), ms[i] / sweeps, cnt[i] // sweeps))


which = sys.argv[1:] or ["C4", "C5"]
if "C2" in which:
    profile(DetSDWBatch(dict(opdim=2, L=8, m=80, s=10), n_replicas=1), "C2 DetSDW O(2) L=8 beta=8")
if "C4" in which:
    profile(DetSDWBatch(dict(opdim=3, L=14, m=140, s=10, weakZflux=False), n_replicas=1), "C4 DetSDW O(3) L=14 beta=14")
if "C5" in which:
    from dqmc_oracle import HubbardParams
    profile(DetHubbardBatch(HubbardParams(L=20, m=200, s=10, U=8.0), n_replicas=1), "C5 DetHubbard L=20 U=8 beta=20")
