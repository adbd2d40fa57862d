#!/usr/bin/env python
"""Per-SOURCE-LINE view of an ncu report: tools/ncu_lines.py report.ncu-rep obj.o [kernel-index] [min_share]

ncu's CLI prints per-SASS-instruction counters (--page source) but not the line they belong to; nvdisasm -g
prints the line of every SASS instruction of the same binary.  Join the two on the instruction offset and
aggregate executed warp instructions, stall samples and shared-memory wavefronts per source line.
The object file must be the one the profiled library was linked from (same SASS)."""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, obj = sys.argv[1], sys.argv[2]
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.003

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

# mangled-name matching: find the function in nvdisasm output whose SASS matches instruction by instruction
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=tmp, capture_output=True)
cubin = next(Path(tmp).glob("*.cubin"))
dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
funcs, name, line = {}, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        name = m.group(1)
        funcs[name] = []
        line = None
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and name:
        funcs[name].append((int(m.group(1), 16), line, m.group(2).strip()))

base = int(b["rows"][0][h["Address"]], 16)
want = [re.sub(r"\s+", " ", r[h["Source"]].strip()) for r in b["rows"]]


def opcode(s):
    s = s.split()
    return s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "")


best = None
for fn, ins in funcs.items():
    if len(ins) != len(want):
        continue
    score = sum(opcode(a[2]) == opcode(w) for a, w in zip(ins, want))
    if best is None or score > best[0]:
        best = (score, fn, ins)
if best is None:
    sys.exit("no function with %d instructions in %s" % (len(want), obj))
score, fn, ins = best
print(b["name"][:120])
print(f"matched {fn[:80]} ({score}/{len(want)} opcodes equal)")
agg = defaultdict(lambda: [0, 0, 0, 0])
STALLS = [c for c in b["hdr"] if c.startswith("stall_") and "Not Issued" not in c]
stall_line = defaultdict(lambda: defaultdict(int))
stall_tot = defaultdict(int)
tot = smp = 0
for r, (off, line, txt) in zip(b["rows"], ins):
    n = int(r[h["Instructions Executed"]] or 0)
    s = int(r[h["# Samples"]] or 0)
    w = int(r[h["L1 Wavefronts Shared"]] or 0) if "L1 Wavefronts Shared" in h else 0
    a = agg[line]
    a[0] += n
    a[1] += s
    a[2] += w
    a[3] += 1
    tot += n
    smp += s
    for c in STALLS:
        v = int(r[h[c]] or 0)
        if v:
            stall_line[line][c[6:]] += v
            stall_tot[c[6:]] += v
print(f"total warp instructions {tot}, stall samples {smp}")
print("stalls: " + ", ".join(f"{k} {100.0 * v / max(smp, 1):.1f}%" for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:9]))
print(f"{'file:line':28s} {'warp inst':>12s} {'share':>7s} {'samples':>8s} {'smp%':>6s} {'smem wf':>11s} {'#sass':>5s}")
for line, (n, s, w, c) in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if tot and (n / tot >= min_share or (smp and s / smp >= min_share)):
        nm = f"{line[0]}:{line[1]}" if line else "?"
        top = ", ".join(f"{k} {v}" for k, v in sorted(stall_line[line].items(), key=lambda kv: -kv[1])[:3])
        print(f"{nm:28s} {n:>12d} {100.0 * n / tot:6.2f}% {s:>8d} {100.0 * s / max(smp, 1):5.1f}% {w:>11d} {c:>5d}  {top}")
