#!/usr/bin/env python
"""Beyond-BASELINE sizes on one GPU (no CPU oracle finishes in seconds there): the dense direct-address radix join, the
general two-pass radix join and the global-table path are three independent algorithms — their counts and the
checksums of their materialized pairs (sum / xor of keys, sum of values, mod 2^64) must agree.
Usage: python tools/big_check.py [rows ...]   (default 400000000)"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402


def checks(k, v):
    return {"n": int(k.size), "sum_keys": int(np.add.reduce(k, dtype=np.uint64)), "xor_keys": int(np.bitwise_xor.reduce(k)) if k.size else 0,
            "sum_vals": int(np.add.reduce(v, dtype=np.uint64))}


for N in [int(float(x)) for x in (sys.argv[1:] or ["4e8"])]:
    bk, bv = capi.generate_g2("build", N, N, 90, 108, 0, N)
    pk = capi.generate_g2("probe", N, N, 90, 108, 0, N)
    out = {"rows": N}
    ref = None
    for name, algo, cfg in (("radix dense", capi.ALGO_RADIX, {"dense": 1}), ("radix general", capi.ALGO_RADIX, {"dense": 0}),
                            ("scalar general", capi.ALGO_SCALAR, {"dense": 0})):
        capi.config_set(**cfg)
        n, sec, st = capi.join(algo, capi.FLAG_MATERIALIZE, bk, bv, pk)
        c = checks(*capi.pairs())
        c.update(matches=n, ms=round(sec * 1e3, 3), path=st["path"], dense=st["dense"], attempts=st["attempts"], bits=[st["radix_bits1"], st["radix_bits2"]])
        out[name] = c
        key = (n, c["n"], c["sum_keys"], c["xor_keys"], c["sum_vals"])
        ref = ref or key
        c["agrees"] = key == ref and c["n"] == n
        nc, _, stc = capi.join(algo, 0, bk, bv, pk)
        c["count_only"] = nc
        c["agrees"] = c["agrees"] and nc == n
    capi.config_set(dense=1)
    out["ok"] = all(out[k]["agrees"] for k in ("radix dense", "radix general", "scalar general"))
    print(json.dumps(out), flush=True)
    for x in (bk, bv, pk):
        x.free()
