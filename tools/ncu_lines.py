#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo, --import-source on).
usage: ncu_lines.py <rep> <kernel regex> [units (e.g. batches) to normalise instruction counts] [min share %]"""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
minshare = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
data, cur = [], "?"
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try:
            data.append((cur, int(r[0]), r[1], int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
toti = sum(d[4] for d in data) or 1
tots = sum(d[3] for d in data) or 1
print(f"warp instructions {toti}, samples {tots}" + (f", {toti / units:.1f} instructions per unit" if units else ""))
for d in data:
    if d[4] > minshare / 100 * toti or d[3] > minshare / 100 * tots:
        per = f"{d[4] / units:7.1f}/unit " if units else ""
        print(f"{d[0][:16]:16s} {d[1]:4d} {100 * d[4] / toti:5.1f}% inst {100 * d[3] / tots:5.1f}% smp  {per}{d[2][:100]}")
