#!/usr/bin/env python
"""Experiment: C2 bloom count vs. filter size (shared-memory footprint -> CTAs per SM)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi
from flash_hash_join_b200.datagen import CONFIGS
N, ny, pct = CONFIGS["C2"]
bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
for bits in (16, 12, 10, 8, 7, 6, 5, 4):
    capi.config_set(bloom_bits_per_key=bits)
    best = None
    for _ in range(5):
        n, sec, st = capi.join(capi.ALGO_SCALAR, capi.FLAG_BLOOM, bk, bv, pk)
        if best is None or st["probe_s"] < best[2]["probe_s"]:
            best = (n, sec, st)
    n, sec, st = best
    print(json.dumps({"bits": bits, "matches": n, "ms": round(sec * 1e3, 4), "probe_ms": round(st["probe_s"] * 1e3, 4), "bloom": st["bloom_kind"], "table_bytes": st["table_bytes"]}), flush=True)
