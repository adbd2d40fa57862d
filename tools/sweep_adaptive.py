#!/usr/bin/env python
"""Crossover sweep behind Engine::choose_path (adaptive_table_l2_pct): for build sides of 1e4 … 1e8 rows against 1e8
probe rows, the device time of the global-table path (algo = scalar) and of the radix path (algo = radix), count and
materialize, on the general hash paths (dense = 0) and with the dense-key-domain variants on, plus what adaptive picks.
    python tools/sweep_adaptive.py > gpurun_out/<tag>/sweep_adaptive.jsonl"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402

N = 100_000_000
pk_cache = {}
for ny in (10_000, 100_000, 300_000, 1_000_000, 2_000_000, 4_000_000, 8_000_000, 16_000_000, 30_000_000, 100_000_000):
    bk, bv = capi.generate_g2("build", N, ny, 90, 108, 0, ny)
    pk = capi.generate_g2("probe", N, ny, 90, 108, 0, N)
    for dense in (0, 1):
        capi.config_set(dense=dense)
        row = {"rows_build": ny, "rows_probe": N, "dense_paths": dense}
        for mode, flags in (("count", 0), ("mat", capi.FLAG_MATERIALIZE)):
            for name, algo in (("scalar", capi.ALGO_SCALAR), ("radix", capi.ALGO_RADIX), ("adaptive", capi.ALGO_ADAPTIVE)):
                if name == "scalar" and ny > 30_000_000 and not dense:
                    continue  # 1.6 GB table, 7 ms of build: far beyond the crossover
                best, st = None, None
                for _ in range(4):
                    n, sec, s = capi.join(algo, flags, bk, bv, pk)
                    if best is None or sec < best:
                        best, st = sec, s
                row[f"{mode}_{name}_ms"] = round(best * 1e3, 4)
                if name == "adaptive":
                    row[f"{mode}_adaptive_path"] = st["path"] + ("/dense%d" % st["dense"] if st["dense"] else "")
                row[f"{mode}_matches"] = n
        print(json.dumps(row), flush=True)
    capi.config_set(dense=1)
    for x in (bk, bv, pk):
        x.free()
