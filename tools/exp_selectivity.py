#!/usr/bin/env python
"""Where the match-rate threshold of the adaptive materialize lies (cfg dense16_sel_min_pct, k_sel_sample): 1e8 probe rows
against small dense build sides at match rates of 10 .. 90 %, timed on the dense16 radix path (sampling off), on the dense
table path reached THROUGH the abandoned dense16 attempt (threshold 100: what a low-rate join really pays), and with the
default threshold.
    python tools/exp_selectivity.py > gpurun_out/<tag>/exp_selectivity.jsonl"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402

N = 100_000_000
M = capi.FLAG_MATERIALIZE
for ny in (10_000, 100_000, 400_000, 800_000):
    for pct in (10, 30, 40, 50, 60, 90):
        bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
        pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
        row = {"rows_build": ny, "rows_probe": N, "match_pct": pct}
        for name, thr in (("dense16", 0), ("table_via_sample", 100), ("default", None)):
            if thr is not None:
                capi.config_set(dense16_sel_min_pct=thr)
            best, st, n = None, None, None
            for _ in range(4):
                capi.check(capi.lib().fj_flush_l2())
                n, sec, s = capi.join(capi.ALGO_ADAPTIVE, M, bk, bv, pk)
                if best is None or sec < best:
                    best, st = sec, s
            row[f"{name}_ms"] = round(best * 1e3, 4)
            row[f"{name}_path"] = f"{st['path']}/dense{st['dense']}/attempts{st['attempts']}"
            row["matches"] = n
            capi.config_set(dense16_sel_min_pct=50)
        print(json.dumps(row), flush=True)
        for x in (bk, bv, pk):
            x.free()
