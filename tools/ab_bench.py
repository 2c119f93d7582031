#!/usr/bin/env python
"""Development tool: same-box A/B of library builds (kernel variants) through bench.py.

Box-to-box variation on the pool is a few per cent, as large as most kernel changes: a variant is only compared
with the shipped kernel inside ONE gpurun call.

  here (no GPU):   git stash; python tools/ab_bench.py --snapshot head; git stash pop
                   python tools/ab_bench.py --snapshot new
  on the GPU box:  gpurun -- 'python tools/ab_bench.py head new head new'

`--snapshot NAME` builds the working tree and keeps the library as detqmc_b200/libdqmc_NAME.so (git-ignored, travels
with gpurun).  The run mode puts each named library in place of libdqmc_b200.so in turn, runs the quick parity
tests once per distinct variant (`--tests`, a pytest -k expression; "" to skip) and `bench.py --no-cpu-baseline`,
prints one line per run and restores the original library.  Remove the snapshots afterwards.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "detqmc_b200")
LIB = os.path.join(PKG, "libdqmc_b200.so")
QUICK = "update_in_slice or sweeps_vs_golden or trajectory or batch_of_replicas"


def variant_path(name):
    return os.path.join(PKG, "libdqmc_%s.so" % name)


def snapshot(name):
    sys.path.insert(0, ROOT)
    from detqmc_b200 import build
    build.build(force=True)
    shutil.copy2(LIB, variant_path(name))
    print("kept", variant_path(name))


def run(names, tests, steps, warmup, untested=()):
    keep = LIB + ".ab_keep"
    shutil.copy2(LIB, keep)
    tested = set(untested)
    try:
        for name in names:
            shutil.copy2(variant_path(name), LIB)
            if tests and name not in tested:
                tested.add(name)
                r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q", "-k", tests],
                                   cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                print("%-8s tests: %s" % (name, r.stdout.strip().splitlines()[-1]), flush=True)
                if r.returncode != 0:
                    continue
            r = subprocess.run([sys.executable, "bench.py", "--steps", str(steps), "--warmup", str(warmup),
                                "--no-cpu-baseline"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0 or not line:
                print("%-8s bench failed (rc %d)" % (name, r.returncode), flush=True)
                continue
            d = json.loads(line[-1])
            fam = d["roofline"]["families"]
            print("%-8s value %6.1f  e2e %6.1f  ms/step %6.1f  | %s" % (
                name, d["value"], d["e2e"]["value"], d["ms_per_step"],
                "  ".join("%s %.1f" % (k, v["ms_per_step"]) for k, v in sorted(fam.items()))), flush=True)
    finally:
        shutil.move(keep, LIB)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--snapshot")
    ap.add_argument("--tests", default=QUICK)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-tests-for", default="head", help="comma-separated variants whose parity is already known")
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    if a.snapshot:
        snapshot(a.snapshot)
    else:
        run(a.names, a.tests, a.steps, a.warmup, [n for n in a.no_tests_for.split(",") if n])
