#!/usr/bin/env python
"""A/B of k_part's warps per CTA on C3 (radix materialize): python tools/exp_part.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402
from flash_hash_join_b200.datagen import CONFIGS  # noqa: E402

N, ny, pct = CONFIGS["C3"]
bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
for kv, k in ((16, 16), (32, 16), (16, 32), (32, 32)):
    capi.config_set(part_warps_kv=kv, part_warps_k=k)
    for mode, flags in (("mat", capi.FLAG_MATERIALIZE), ("count", 0)):
        best = None
        for _ in range(6):
            n, sec, st = capi.join(2, flags, bk, bv, pk)
            if best is None or sec < best[0]:
                best = (sec, st)
        sec, st = best
        print(json.dumps({"warps_kv": kv, "warps_k": k, "mode": mode, "matches": n, "ms": round(sec * 1e3, 4), "part_build_us": st.get("part_build_us"),
                          "part_probe_us": st.get("part_probe_us"), "probe_ms": round(st["probe_s"] * 1e3, 4)}), flush=True)
