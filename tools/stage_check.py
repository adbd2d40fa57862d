#!/usr/bin/env python
"""Large PAGEABLE host inputs (plain numpy arrays) through the user-facing calls: results against the numpy
restatement, and the host->device time with and without the multi-threaded pinned staging ring (Engine::h2d)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import flash_join  # noqa: E402
from flash_hash_join_b200.datagen import g1  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 30_000_000
flash_join.initialize()
bk, bv, pk = g1(N, 100_000, 10)
bk, bv, pk = bk.copy(), bv.copy(), pk.copy()  # plain pageable numpy memory
n0, k0, v0 = O.np_join(bk, bv, pk)
cs0 = O.checksums(k0, v0)
out = {"rows": N, "expected": n0}
for label, threads in (("staged (8 threads)", 8), ("plain cudaMemcpyAsync", 0), ("staged (4 threads)", 4)):
    flash_join.configure(stage_threads=threads)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        n, _ = flash_join.hash_join_count_bloom(bk, bv, pk)
        wall = time.perf_counter() - t0
        st = flash_join.last_stats()
        if best is None or wall < best[0]:
            best = (wall, st["h2d_s"])
    n_m, _ = flash_join.hash_join(bk, bv, pk)
    cs = O.checksums(*flash_join.last_pairs())
    out[label] = {"count_ok": n == n0, "mat_ok": n_m == n0 and all(cs[k] == cs0[k] for k in cs0), "wall_ms": round(best[0] * 1e3, 2),
                  "h2d_ms": round(best[1] * 1e3, 2), "h2d_GBps": round((2 * bk.size + pk.size) * 8 / best[1] * 1e-9, 1)}
flash_join.configure(stage_threads=8)
hp = flash_join.pinned_empty(pk.size); hp[:] = pk
t0 = time.perf_counter(); n, _ = flash_join.hash_join_count_bloom(bk, bv, hp); wall = time.perf_counter() - t0
out["pinned probe column"] = {"count_ok": n == n0, "wall_ms": round(wall * 1e3, 2), "h2d_ms": round(flash_join.last_stats()["h2d_s"] * 1e3, 2)}
out["ok"] = all(v.get("count_ok", True) and v.get("mat_ok", True) for v in out.values() if isinstance(v, dict))
print(json.dumps(out))
