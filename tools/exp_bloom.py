#!/usr/bin/env python
"""General-path C2 (hash table + shared-memory Bloom filter): filter bits per key vs CTAs per SM.  python tools/exp_bloom.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402

N, ny, pct = 100_000_000, 100_000, 10
bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
capi.config_set(dense=0)
for bits in (16, 12, 8, 6):
    capi.config_set(bloom_bits_per_key=bits)
    for mode, flags in (("count", capi.FLAG_BLOOM), ("mat", capi.FLAG_BLOOM | capi.FLAG_MATERIALIZE)):
        best = None
        for _ in range(6):
            n, sec, st = capi.join(capi.ALGO_SCALAR, flags, bk, bv, pk)
            if best is None or sec < best[0]:
                best = (sec, st)
        print(json.dumps({"bits_per_key": bits, "mode": mode, "matches": n, "ms": round(best[0] * 1e3, 4), "probe_ms": round(best[1]["probe_s"] * 1e3, 4),
                          "bloom": best[1]["bloom_kind"], "table_bytes": best[1]["table_bytes"]}), flush=True)
