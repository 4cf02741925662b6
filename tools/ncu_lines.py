#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an `ncu --set full --import-source on` report.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
lines = []
total = 0
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit():
        n = int(float(r[i_inst]))
        lines.append((int(r[0]), n, int(float(r[i_samp])), r[1].strip()))
        total += n
tot_s = sum(l[2] for l in lines) or 1
print(f"total warp instructions {total:.3e}, samples {tot_s}")
for ln, n, sm, src in sorted(lines, key=lambda l: -l[1])[:top]:
    print(f"{ln:5d} {100 * n / total:5.1f}% inst {100 * sm / tot_s:5.1f}% samples  {src[:110]}")
