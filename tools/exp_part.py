#!/usr/bin/env python
"""A/B of k_part's input path on C3 (per-warp TMA rings vs 128-bit loads straight into registers): python tools/exp_part.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402
from flash_hash_join_b200.datagen import CONFIGS  # noqa: E402

N, ny, pct = CONFIGS["C3"]
bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
for kv, k in ((0, 0), (1, 0), (0, 1), (1, 1)):
    capi.config_set(part_direct_kv=kv, part_direct_k=k)
    for mode, flags in (("mat", capi.FLAG_MATERIALIZE), ("count", 0)):
        best = None
        for _ in range(6):
            n, sec, st = capi.join(2, flags, bk, bv, pk)
            if best is None or sec < best[0]:
                best = (sec, st)
        sec, st = best
        print(json.dumps({"direct_kv": kv, "direct_k": k, "mode": mode, "matches": n, "ms": round(sec * 1e3, 4), "part_build_us": st.get("part_build_us"),
                          "part_probe_us": st.get("part_probe_us"), "probe_ms": round(st["probe_s"] * 1e3, 4)}), flush=True)
