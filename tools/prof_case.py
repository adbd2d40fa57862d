#!/usr/bin/env python
"""Run ONE join case a few times (for ncu): python tools/prof_case.py C2 scalar count bloom [--reps 2] [--wide]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402
from flash_hash_join_b200.datagen import CONFIGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("algo", choices=["adaptive", "scalar", "radix"])
ap.add_argument("mode", choices=["count", "mat"])
ap.add_argument("bloom", nargs="?", default="nobloom", choices=["bloom", "nobloom"])
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--wide", action="store_true")
ap.add_argument("--set", action="append", default=[], help="config key=value")
ap.add_argument("--check", action="store_true", help="compare the count with the numpy oracle (small configs only)")
a = ap.parse_args()
for kv in a.set:
    k, v = kv.split("=")
    capi.config_set(**{k: int(v)})
N, ny, pct = CONFIGS[a.config]
bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
flags = (capi.FLAG_BLOOM if a.bloom == "bloom" else 0) | (capi.FLAG_MATERIALIZE if a.mode == "mat" else 0) | (capi.FLAG_FORCE_WIDE if a.wide else 0)
algo = {"adaptive": 0, "scalar": 1, "radix": 2}[a.algo]
if a.check:
    import numpy as np

    from flash_hash_join_b200.datagen import g2_slice
    from oracle import oracle as O

    hb, hv = g2_slice(N, ny, pct, 108, "build", 0, ny)
    hp = g2_slice(N, ny, pct, 108, "probe", 0, N)
    n0, k0, v0 = O.np_join(hb, hv, hp)
    n, sec, st = capi.join(algo, flags, bk, bv, pk)
    assert n == n0, (n, n0, st)
    if a.mode == "mat":
        assert np.array_equal(O.sorted_pairs(*capi.pairs()), O.sorted_pairs(k0, v0))
    print("check ok", n, "dense" if st["dense"] else "general", flush=True)
for _ in range(a.reps):
    n, sec, st = capi.join(algo, flags, bk, bv, pk)
    print(n, round(sec * 1e3, 4), st["path"], st["narrow"], st["bloom_kind"], "dense" if st["dense"] else "general", st["attempts"], flush=True)
