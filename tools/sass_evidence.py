#!/usr/bin/env python3
"""tools/sass_evidence.py [out.txt] — per hot kernel of libflashjoin_b200.so: register / shared-memory use and the counts
of the SASS mnemonics that show how it moves data (UBLKCP = TMA bulk copy, SYNCS = mbarrier, ATOMS = shared-memory
atomic, LDS/STS/LDG/STG widths, ATOMG/REDG = global atomics, VOTE/SHFL = warp collectives), plus the first lines of
each kernel's TMA / mbarrier instructions.  Runs on the build box (cuobjdump only, no GPU)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "flash_hash_join_b200" / "libflashjoin_b200.so"
HOT = ("k_part", "k_sjoin", "k_pairs_compact", "k_xsync", "k_count_dense_fused", "k_count_dense_peer", "k_mat_dense_fused", "k_probe_count", "k_probe_mat",
       "k_build", "k_scatter2", "k_join3", "k_djoin")
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
usage = {}
name = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        name = m.group(1)
    elif name and "REG:" in ln:
        usage[name] = ln.strip()
        name = None
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
cur, body = None, collections.defaultdict(list)
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if cur and m:
        body[cur].append(m.group(1).strip())
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print(f"SASS evidence for {LIB.name} (sm_100a), cuobjdump {subprocess.run(['cuobjdump', '--version'], capture_output=True, text=True).stdout.split('release')[-1].strip()}", file=out)
for fn in sorted(body):
    dn = demangle(fn)
    if not any(("fj::" + h + "<") in dn or ("fj::" + h + "(") in dn or dn.startswith(h + "(") or dn.startswith("void fj::" + h) for h in HOT):
        continue
    ins = body[fn]
    cnt = collections.Counter()
    for i in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", i).split()[0]
        cnt[op] += 1
    keys = [k for k in cnt if re.match(r"(UBLKCP|SYNCS|ATOMS|ATOMG|REDG|RED|ATOM|LDS|STS|LDG|STG|LDC|VOTE|SHFL|BAR|MEMBAR|FENCE|UTMA|CCTL|ERRBAR|NANOSLEEP|ELECT|R2UR|IDP|VIMNMX|POPC|MATCH)", k)]
    print("\n" + "=" * 110, file=out)
    print(dn, file=out)
    print(f"  {usage.get(fn, '')}", file=out)
    print(f"  {len(ins)} SASS instructions; " + ", ".join(f"{k} {cnt[k]}" for k in sorted(keys)), file=out)
    shown = 0
    for i in ins:
        if re.search(r"UBLKCP|SYNCS|UTMA", i) and shown < 6:
            print("    " + i, file=out)
            shown += 1
