#!/usr/bin/env python
"""Aggregate an ncu report's source page by CUDA source line: share of warp-stall samples and of
executed instructions per line, for kernels whose name contains a substring.
usage: ncu_lines.py report.ncu-rep <kernel-substring> [top]"""
import collections
import csv
import subprocess
import sys

rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn = fp = hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0])
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit() and fn and sub in fn:
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        key = (fp.split("/")[-1], int(r[0]))
        try:
            agg[key][0] += float(r[ie] or 0)
            agg[key][1] += float(r[ss] or 0)
            src[key] = r[1]
        except ValueError:
            pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"kernel~{sub}: instructions {ti:.3g}, samples {ts:.3g}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0][:18]:18s} {key[1]:4d} {100 * v[1] / ts:5.1f}% smp {100 * v[0] / ti:5.1f}% inst  {src[key][:96]}")
