#!/usr/bin/env python
"""Knob sweep for the dense-key-domain radix join (k_djoin) on C3: per-CTA look-ahead (ring, batch), pipeline
distances (delay_b, delay_p) and L2 group size.  Prints one JSON line per setting (best of --reps)."""
import argparse
import itertools
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402
from flash_hash_join_b200.datagen import CONFIGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
N, ny, pct = CONFIGS[a.config]
bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
M = capi.FLAG_MATERIALIZE
settings = []
for ring, batch in ((1, 1), (2, 1), (2, 2), (4, 1), (4, 2), (4, 4), (8, 2), (8, 4)):
    settings.append(dict(dense_ring=ring, dense_batch=batch, dense_delay_b=1, dense_delay_p=1, dense_group_mb=16))
for db, dp, mb in ((1, 2, 12), (1, 2, 16), (2, 2, 8), (2, 2, 12), (1, 3, 8), (1, 3, 12), (2, 3, 8), (1, 1, 24), (1, 1, 12)):
    for ring, batch in ((4, 2), (8, 4)):
        settings.append(dict(dense_ring=ring, dense_batch=batch, dense_delay_b=db, dense_delay_p=dp, dense_group_mb=mb))
for cfg in settings:
    capi.config_set(**cfg)
    best = None
    for _ in range(a.reps):
        n, sec, st = capi.join(capi.ALGO_RADIX, M, bk, bv, pk)
        if best is None or st["probe_s"] < best[2]["probe_s"]:
            best = (n, sec, st)
    n, sec, st = best
    print(json.dumps(dict(cfg, matches=n, ms=round(sec * 1e3, 4), join_ms=round(st["probe_s"] * 1e3, 4), dense=st["dense"])), flush=True)
