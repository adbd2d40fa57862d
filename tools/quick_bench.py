#!/usr/bin/env python
"""Developer timing loop (not the contract bench — see bench.py): device-resident G2 data, best of N,
per-phase device times from fj_stats.  Usage: python tools/quick_bench.py [C2 C3 ...] [--reps 5]"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flash_hash_join_b200 import capi  # noqa: E402
from flash_hash_join_b200.datagen import CONFIGS  # noqa: E402

PEAK = 6468.3  # GB/s, MEASURED_PEAKS.json


def run(name, algo, flags, d, reps, N, cfg=None):
    if cfg:
        capi.config_set(**cfg)
    best = None
    for _ in range(reps):
        capi.check(capi.lib().fj_flush_l2())
        n, sec, st = capi.join(algo, flags, *d)
        if best is None or sec < best[1]:
            best = (n, sec, st)
    n, sec, st = best
    out = {"case": name, "matches": n, "ms": round(sec * 1e3, 4), "rows_per_s": round(N / sec, 1),
           "alg_GBps": round(st["algorithmic_bytes"] / sec * 1e-9, 1), "frac": round(st["algorithmic_bytes"] / sec * 1e-9 / PEAK, 4),
           "path": st["path"], "narrow": st["narrow"], "bloom": st["bloom_kind"], "dense": st["dense"], "attempts": st["attempts"],
           "clear_ms": round(st["clear_s"] * 1e3, 4), "build_ms": round(st["build_s"] * 1e3, 4),
           "part_ms": round(st["partition_s"] * 1e3, 4), "probe_ms": round(st["probe_s"] * 1e3, 4),
           "launches": st["kernel_launches"], "bits": [st["radix_bits1"], st["radix_bits2"]], "cfg": cfg or {}}
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C2", "C3"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--full", action="store_true", help="also the slower alternative paths of the large-build configs")
    a = ap.parse_args()
    S, R, A = capi.ALGO_SCALAR, capi.ALGO_RADIX, capi.ALGO_ADAPTIVE
    B, M, W = capi.FLAG_BLOOM, capi.FLAG_MATERIALIZE, capi.FLAG_FORCE_WIDE
    for c in a.configs:
        N, ny, pct = CONFIGS[c]
        bk, bv = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
        pk = capi.generate_g2("probe", N, ny, pct, 108, 0, N)
        d = (bk, bv, pk)
        if ny <= 10**7:
            run(f"{c} scalar count bloom (dense bitmap)", S, B, d, a.reps, N)
            run(f"{c} adaptive count (dense bitmap)", A, 0, d, a.reps, N)
            run(f"{c} scalar mat (dense bitmap + direct table)", S, M, d, a.reps, N)
            run(f"{c} adaptive mat (dense)", A, M, d, a.reps, N)
            capi.config_set(dense=0)
            run(f"{c} scalar count bloom", S, B, d, a.reps, N)
            run(f"{c} scalar count", S, 0, d, a.reps, N)
            run(f"{c} scalar count bloom(global)", S, B, d, a.reps, N, {"smem_bloom": 0})
            capi.config_set(smem_bloom=1)
            run(f"{c} scalar count wide", S, W, d, a.reps, N)
            run(f"{c} scalar mat", S, M, d, a.reps, N)
            run(f"{c} scalar mat bloom", S, M | B, d, a.reps, N)
            run(f"{c} radix count", R, 0, d, a.reps, N)
            run(f"{c} radix mat", R, M, d, a.reps, N)
            run(f"{c} adaptive count", A, 0, d, a.reps, N)
            capi.config_set(dense=1)
        else:
            run(f"{c} radix mat (dense16: k_part + k_sjoin)", R, M, d, a.reps, N)
            run(f"{c} radix count (dense16)", R, 0, d, a.reps, N)
            run(f"{c} adaptive mat (dense16)", A, M, d, a.reps, N)
            if a.full:
                run(f"{c} radix mat (dense, L2 direct-address k_djoin)", R, M, d, a.reps, N, {"dense16": 0})
                run(f"{c} radix count (dense, L2 direct-address k_djoin)", R, 0, d, a.reps, N)
                capi.config_set(dense16=1)
            capi.config_set(dense=0)
            if not a.full:
                run(f"{c} radix mat", R, M, d, a.reps, N)
                capi.config_set(dense=1)
                for x in d:
                    x.free()
                continue
            run(f"{c} radix mat", R, M, d, a.reps, N)
            run(f"{c} radix count", R, 0, d, a.reps, N)
            run(f"{c} radix mat 2^16 parts", R, M, d, a.reps, N, {"radix_sub_rows": 2400})
            capi.config_set(radix_sub_rows=0)
            run(f"{c} radix mat wide", R, M | W, d, a.reps, N)
            run(f"{c} scalar count", S, 0, d, max(1, a.reps // 2), N)
            run(f"{c} scalar mat", S, M, d, max(1, a.reps // 2), N)
            run(f"{c} adaptive mat", A, M, d, a.reps, N)
            capi.config_set(dense=1)
        for x in d:
            x.free()


if __name__ == "__main__":
    main()
