#!/usr/bin/env python
"""Development tool (GPU box): print the actual deviations of the CUDA path from the full-size reference goldens
(tests/golden/*c2*, *c3*, *c4*, *c5*), to calibrate / document the tolerances of tests/test_gpu_parity.py."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import load_golden, sdw_params_of, hubbard_params_of      # noqa: E402


def dev(G, g, key):
    st = int(g["stride"])
    sc = float(g[key + "_maxabs"])
    return "sub %.2e trace %.2e" % (np.abs(G[::st, ::st] - g[key + "_sub"]).max() / sc,
                                    abs(np.trace(G) - g[key + "_trace"]) / max(abs(g[key + "_trace"]), sc))


def sdw(name):
    from detqmc_b200 import DetSDWBatch
    g = load_golden(name)
    p = sdw_params_of(g)
    t0 = time.time()
    b = DetSDWBatch(p, n_replicas=1)
    print(name, "setup %.1fs" % (time.time() - t0), dev(b.green(), g, "green0"),
          "logdet rel %.2e" % (abs(b.logdet() - float(g["logdet0"])) / abs(float(g["logdet0"]))), flush=True)
    for sw in range(int(g["n_sweeps"])):
        t0 = time.time()
        b.sweepThermalization()
        cd = b.control_data()
        line = "  sweep %d %.2fs acc equal %s shifts equal %s dphiDelta %.1e" % (
            sw + 1, time.time() - t0, cd.lastAccRatioLocal_phi == g["lastAccRatio"][sw],
            cd.acceptedGlobalShifts == g["acceptedGlobalShifts"][sw], abs(cd.phiDelta - g["phiDelta"][sw]))
        key = "phi_after_%d" % (sw + 1)
        if key in g.files:
            line += " |dphi| %.1e G: %s (ref selfdev %.1e) consistency %.1e" % (
                np.abs(b.phi()[1:] - g[key][1:]).max(), dev(b.green(), g, "green_after_%d" % (sw + 1)),
                float(g["ref_selfdev_after_%d" % (sw + 1)]) if "ref_selfdev_after_%d" % (sw + 1) in g.files else -1.0,
                b.green_consistency().max())
            print(line, flush=True)
        elif cd.lastAccRatioLocal_phi != g["lastAccRatio"][sw]:
            print(line, flush=True)
    print("  rng equal", np.array_equal(b.rng_draw(8), g["rng_next"]))


def hub(name):
    from detqmc_b200 import DetHubbardBatch
    g = load_golden(name)
    p = hubbard_params_of(g)
    b = DetHubbardBatch(p)
    print(name, "aux0 equal", np.array_equal(b.auxfield()[1:], g["aux0"]),
          [dev(b.green(0, gc), g, "green0_%d" % gc) for gc in (0, 1)],
          ["%.2e" % (abs(b.logdet(0, gc) - float(g["logdet0_%d" % gc])) / abs(float(g["logdet0_%d" % gc]))) for gc in (0, 1)])
    for sw in range(int(g["n_sweeps"])):
        t0 = time.time()
        b.sweep()
        aux = b.auxfield()[1:]
        print("  sweep %d %.2fs aux equal %s (%d differ)" % (sw + 1, time.time() - t0,
              np.array_equal(aux, g["aux_after_%d" % (sw + 1)]), int((aux != g["aux_after_%d" % (sw + 1)]).sum())),
              [dev(b.green(0, gc), g, "green_after_%d_%d" % (sw + 1, gc)) for gc in (0, 1)],
              "consistency %.1e" % b.green_consistency().max(), flush=True)
    print("  rng equal", np.array_equal(b.rng_draw(8), g["rng_next"]))


if __name__ == "__main__":
    for n in sys.argv[1:] or ["sdw_c2_L8_b8_traj100", "sdw_c3_L12_b10", "sdw_c4_o3_L14_b14", "hubbard_c5_L20_U8_b20"]:
        if not os.path.exists(os.path.join(ROOT, "tests", "golden", n + ".npz")):
            print(n, "golden missing")
            continue
        (hub if n.startswith("hub") else sdw)(n)
