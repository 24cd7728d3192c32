#!/usr/bin/env python
"""Per-source-line totals (executed warp instructions, stall samples, smem wavefronts) of one kernel from an
ncu report captured with --import-source on:  python tools/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{pat}"], capture_output=True, text=True).stdout.splitlines()
fname, fn, hdr, seen_fn = "?", None, None, set()
lines = []
for r in csv.reader(raw):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
        key = (fn,)
    elif hdr and r[0] not in ("", "...") and fn is not None:
        if len(seen_fn) and fn not in seen_fn and len(seen_fn) >= 1 and False:
            continue
        try:
            ie, sm, wv = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
            lines.append((fn, fname, int(r[0]), r[1].strip()[:90], int(r[ie] or 0), int(r[sm] or 0), int(r[wv] or 0)))
        except (ValueError, IndexError):
            pass
first = lines[0][0] if lines else None
lines = [l for l in lines if l[0] == first]           # first matching launch / instantiation only
# the same (file, line) can appear once per launch table: keep the first occurrence
uniq = {}
for l in lines:
    uniq.setdefault((l[1], l[2]), l)
lines = list(uniq.values())
tot_i = sum(l[4] for l in lines) or 1
tot_s = sum(l[5] for l in lines) or 1
print(first)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{l[1]:16s}:{l[2]:4d} inst {l[4]:10d} {l[4] / tot_i:6.1%}  smp {l[5] / tot_s:6.1%}  wav {l[6]:9d} | {l[3]}")
