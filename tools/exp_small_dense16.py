#!/usr/bin/env python
"""dense16 radix path (k_part + k_sjoin) on SMALL build sides against 1e8 probe rows, next to the global-table dense paths:
where should adaptive switch?  python tools/exp_small_dense16.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402

N = 100_000_000
for ny in (10_000, 100_000, 300_000, 600_000, 1_000_000, 2_000_000):
    bk, bv = capi.generate_g2("build", N, ny, 90, 108, 0, ny)
    pk = capi.generate_g2("probe", N, ny, 90, 108, 0, N)
    row = {"rows_build": ny}
    for logp in (0, 8, 9, 10, 11):
        capi.config_set(dense_min_rows=1024, dense16_logp=logp)
        for mode, flags in (("count", 0), ("mat", capi.FLAG_MATERIALIZE)):
            best = None
            for _ in range(4):
                n, sec, st = capi.join(capi.ALGO_RADIX, flags, bk, bv, pk)
                best = sec if best is None or sec < best else best
            row[f"{mode}_radix_logp{logp}_ms"] = round(best * 1e3, 4)
            row[f"{mode}_path_logp{logp}"] = f"dense{st['dense']} bits {st['radix_bits1']}"
    capi.config_set(dense_min_rows=1 << 20, dense16_logp=0)
    for mode, flags in (("count", 0), ("mat", capi.FLAG_MATERIALIZE)):
        best = None
        for _ in range(4):
            n, sec, st = capi.join(capi.ALGO_SCALAR, flags, bk, bv, pk)
            best = sec if best is None or sec < best else best
        row[f"{mode}_scalar_ms"] = round(best * 1e3, 4)
        row[f"{mode}_scalar_path"] = f"dense{st['dense']}"
    print(json.dumps(row), flush=True)
    for x in (bk, bv, pk):
        x.free()
