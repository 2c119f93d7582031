#!/usr/bin/env python
"""Development tool: instruction mnemonic counts per hot kernel of the shipped library (no GPU needed):
   tools/sass_summary.py > profiles/sass_r02_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "detqmc_b200", "libdqmc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
HOT = ("cb_mult", "hub_update_slice", "qr_panel2", "update_gather", "update_window", "update_round", "fermion_measure",
       "zgemm_dmma", "zgemm_rank_update2")
COLS = ["DMMA", "FP64", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "LDS", "STS", "LDL", "STL", "BAR", "SHFL", "UCGABAR"]
print("# SASS evidence, round 2: `cuobjdump -sass detqmc_b200/libdqmc_b200.so` (sm_100a), instruction mnemonic counts per hot kernel.")
print("# DMMA = FP64 tensor-core MMA (mma.sync m8n8k4.f64; tcgen05 has no f64 kind); UBLKCP = cp.async.bulk (bulk-copy / TMA engine),")
print("# SYNCS = mbarrier operations (arrive.expect_tx / try_wait); LDGSTS = cp.async global->shared; LDL/STL = local-memory (spill)")
print("# traffic; BAR = CTA barriers; UCGABAR = thread-block-cluster barrier.  Kernel names are demangled prefixes.")
print("%-74s %6s" % ("kernel", "instr") + "".join(" %6s" % c for c in COLS))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
rows = []
for blk, name in zip(blocks, names):
    if not any(h in name for h in HOT):
        continue
    ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
    cnt = collections.Counter()
    for op in ops:
        base = op.split(".")[0]
        cnt[base] += 1
        if base.startswith("UCGABAR"):
            cnt["UCGABAR"] += 1
        if base in ("DFMA", "DMUL", "DADD"):
            cnt["FP64"] += 1
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("dqmc::", "").replace("void ", "")
    rows.append((short[:74], len(ops), [cnt[c] for c in COLS]))
for short, n, vals in sorted(rows):
    print("%-74s %6d" % (short, n) + "".join(" %6d" % v for v in vals))
