#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: tools/ncu_sass.py report.ncu-rep [kernel-index] [min_share]
Prints address, executed warp instructions (share of kernel), avg active threads, stall samples, SASS."""
import csv
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[kidx]
h = {n: i for i, n in enumerate(b["hdr"])}
tot = sum(int(r[h["Instructions Executed"]] or 0) for r in b["rows"])
samples = sum(int(r[h["# Samples"]] or 0) for r in b["rows"])
print(b["name"][:150])
print(f"total warp instructions {tot}, samples {samples}")
for r in b["rows"]:
    n = int(r[h["Instructions Executed"]] or 0)
    if tot and n / tot >= min_share:
        print(f"{r[h['Address']][-5:]} {n:>11d} {100.0 * n / tot:5.2f}% thr={r[h['Avg. Threads Executed']]:>5s} smp={r[h['# Samples']]:>6s}  {r[h['Source']][:110]}")
