#!/usr/bin/env python
"""Host turn-around of one fj_join_u64 call: wall time per call vs device time on a small radix case, with and without a
stats block.  python tools/exp_call_overhead.py"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402

L = capi.lib()
for N in (1 << 21, 100_000_000):
    bk, bv = capi.generate_g2("build", N, N, 90, 108, 0, N)
    pk = capi.generate_g2("probe", N, N, 90, 108, 0, N)
    n, sec, st = C.c_uint64(0), C.c_double(0), capi.Stats()
    flags = capi.FLAG_MATERIALIZE | capi.FLAG_DEVICE_INPUTS
    for with_stats in (True, False):
        p_st = C.byref(st) if with_stats else None
        for _ in range(5):
            capi.check(L.fj_join_u64(capi.ALGO_RADIX, flags, bk.ptr, bv.ptr, N, pk.ptr, N, C.byref(n), C.byref(sec), p_st))
        reps = 200 if N < 10**7 else 30
        dev = 0.0
        capi.check(L.fj_device_synchronize())
        t0 = time.perf_counter()
        for _ in range(reps):
            capi.check(L.fj_join_u64(capi.ALGO_RADIX, flags, bk.ptr, bv.ptr, N, pk.ptr, N, C.byref(n), C.byref(sec), p_st))
            dev += sec.value
        wall = (time.perf_counter() - t0) / reps
        print(json.dumps({"rows": N, "stats": with_stats, "wall_us_per_call": round(wall * 1e6, 1), "device_us_per_call": round(dev / reps * 1e6, 1),
                          "host_us_per_call": round((wall - dev / reps) * 1e6, 1), "matches": n.value}), flush=True)
    for x in (bk, bv, pk):
        x.free()
