#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions and stall samples) of one kernel from an
ncu report:  python tools/ncu_sass_hist.py <rep> <kernel regex> [launch index]"""
import collections
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw[1:]))
hdr = rows[0]
for i in range(1, len(rows)):          # keep the first launch's table only
    if rows[i] and rows[i][0] in ("Kernel Name", "Address"):
        rows = rows[:i]
        break
si, ei, ti, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
ops = collections.Counter()
smp = collections.Counter()
wav = collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= max(si, ei, ti, wi):
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2])
    n = int(r[ei] or 0)
    ops[op] += n
    smp[op] += int(r[ti] or 0)
    wav[op] += int(r[wi] or 0)
    tot += n
print(raw[0][:120])
print(f"total warp instructions {tot}, SASS lines {len(rows) - 1}")
for op, n in ops.most_common(28):
    print(f"{op:22s} {n:12d} {n / tot:6.1%}   samples {smp[op]:7d}   smem wavefronts {wav[op]}")
