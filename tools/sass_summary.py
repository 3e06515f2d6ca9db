#!/usr/bin/env python
"""profiles/r2_sass_summary.txt: per kernel of libpydisort_b200.so -- registers, spills, shared memory (cuobjdump
-res-usage) and the counts of the SASS instructions that tell which hardware path a kernel uses (DFMA/DMUL/DADD =
FP64 pipe, DMMA = FP64 tensor cores, LDGSTS = cp.async, UBLKCP/UTMA = TMA bulk copies, SHFL, LDS/STS, MUFU).
    python tools/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "pythonic_disort_b200", "libpydisort_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
name = None
KEYS = ["DFMA", "DMUL", "DADD", "DMMA", "LDGSTS", "UBLKCP", "UTMALDG", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "MUFU", "BAR"]
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1).split(".")[0]
        counts[name]["total"] += 1
        if op in KEYS:
            counts[name][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':58s} {'regs':>4s} {'stack':>5s} {'smem':>6s} {'insts':>7s}  " + " ".join(f"{k:>6s}" for k in KEYS))
for mangled, pretty in sorted(zip(counts, demangle), key=lambda kv: kv[1]):
    reg, stack, shared, local = usage.get(mangled, (0, 0, 0, 0))
    short = re.sub(r"\(.*", "", pretty).replace("void ", "")
    c = counts[mangled]
    print(f"{short:58s} {reg:4d} {stack:5d} {shared:6d} {c['total']:7d}  " + " ".join(f"{c[k]:6d}" for k in KEYS))
print("\nstack > 0 means spills (STL/LDL); smem is the static part only (the production kernels size theirs at launch).")
print("No kernel issues DMMA or TMA instructions: the tiles are 4..16 wide FP64 blocks held in registers; cp.async (LDGSTS)")
print("stages the layer operands of k_stage_b_add.")
