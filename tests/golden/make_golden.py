#!/usr/bin/env python
"""Regenerate tests/golden/*.json|*.npz from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (where oracle/build_ref.sh can see /root/reference):
    python tests/golden/make_golden.py [--big]
The reference binary is the source of truth: every count below is what the unmodified
hash_join.cpp returned (all 12 entry points must agree), and the pair checksums come from the
pairs-returning variant (hash_join.cpp:380/444/494 extended to return result_keys/result_values).
`--big` adds the 1e8-row cases (minutes, ~15 GB RAM).
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from flash_hash_join_b200.datagen import g1, g2  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent

G1_CASES = [(10**6, 10**3, 90), (10**6, 10**6, 90), (10**6, 10**4, 10), (10**7, 10, 90), (10**7, 10**4, 90), (10**7, 10**7, 90)]
G1_BIG = [(10**8, 10**5, 10), (10**8, 10**8, 90)]
G2_CASES = [(10**6, 10**5, 10), (10**6, 10**6, 90), (10**7, 10**4, 90)]
# the bench workloads (bench.py generates G2 in HBM): C2, C3, the C4 per-GPU slice, and the weak-scaling C2 totals
# at 2 / 4 / 8 GPUs (probe rows [0, G * 1e8) against the same 1e5-row build side; counts only)
G2_BIG = [(10**8, 10**5, 10), (10**8, 10**8, 90), (125 * 10**6, 10**6, 90)]
G2_WEAK = [(2 * 10**8, 10**5, 10), (4 * 10**8, 10**5, 10), (8 * 10**8, 10**5, 10)]


def run_case(gen, N, ny, pct, plain, pairs, skip_scalar_above=2 * 10**7, count_only=False):
    bk, bv, pk = gen(N, ny, pct)
    counts = {}
    for (algo, bloom, mat), name in O.ENTRY_POINTS.items():
        if algo == "scalar" and ny > skip_scalar_above:
            continue  # 8.6 GB table with a serial clear at 1e8 — skipped
        if count_only and (mat or algo == "radix"):
            continue
        counts[name] = int(getattr(plain, name)(bk, bv, pk)[0])
    assert len(set(counts.values())) == 1, counts
    if count_only:
        return {"N": N, "ny": ny, "match_pct": pct, "seed": 108, "count": next(iter(counts.values())), "sum_keys": None, "xor_keys": None,
                "sum_vals": None, "entry_points_agreed": sorted(counts), "pairs_from": None}
    entry = "hash_join_radix" if ny >= 10**6 else "hash_join"
    r = getattr(pairs, entry)(bk, bv, pk)
    cs = O.checksums(r[2], r[3])
    assert cs["count"] == r[0] == next(iter(counts.values()))
    return {"N": N, "ny": ny, "match_pct": pct, "seed": 108, **cs, "entry_points_agreed": sorted(counts), "pairs_from": entry}


def small_fixtures(pairs):
    """Full input/output fixtures small enough to commit."""
    out = {}
    rng = np.random.default_rng(7)
    # (a) h2o-shaped small case
    bk, bv, pk = g1(5000, 600, 90)
    r = pairs.hash_join(bk, bv, pk)
    out["g1_5000_600"] = dict(bk=bk, bv=bv, pk=pk, rk=r[2], rv=r[3])
    # (b) edge keys: 0, 2^64-1, >= 2^32, values needing 64 bits
    bk = np.array([0, 2**64 - 1, 2**32, 2**32 + 1, 2**63, 5, 2**64 - 2, 0xFFFFFFFF, 0xFFFFFFFE], dtype=np.uint64)
    bv = np.array([11, 22, 33, 2**40, 2**64 - 1, 0, 77, 88, 99], dtype=np.uint64)
    pk = np.array([0, 0, 2**64 - 1, 7, 2**32, 2**63, 2**63 + 1, 5, 5, 2**32 + 1, 2**64 - 2, 1, 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD], dtype=np.uint64)
    r = pairs.hash_join(bk, bv, pk)
    out["edge_keys"] = dict(bk=bk, bv=bv, pk=pk, rk=r[2], rv=r[3])
    # (c) duplicate build keys — keep-first; the radix path of the reference is deterministic
    bk = rng.integers(0, 400, 1500).astype(np.uint64)
    bv = np.arange(1500, dtype=np.uint64)
    pk = rng.integers(0, 500, 4000).astype(np.uint64)
    r = pairs.hash_join_radix(bk, bv, pk)
    out["dup_build_radix"] = dict(bk=bk, bv=bv, pk=pk, rk=r[2], rv=r[3])
    # (d) skew: all probe keys equal
    bk = np.arange(1, 301, dtype=np.uint64)
    bv = bk * np.uint64(3)
    pk = np.full(3000, 17, dtype=np.uint64)
    r = pairs.hash_join(bk, bv, pk)
    out["skew_probe"] = dict(bk=bk, bv=bv, pk=pk, rk=r[2], rv=r[3])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    plain, pairs = O.load_reference("plain"), O.load_reference("pairs")
    path = HERE / "g1_goldens.json"
    existing = {(-1,)}
    rows = []
    if path.exists():
        rows = json.loads(path.read_text())["cases"]
        existing = {(r["gen"], r["N"], r["ny"], r["match_pct"]) for r in rows}
    todo = [("g1", c) for c in G1_CASES] + [("g2", c) for c in G2_CASES] + ([("g1", c) for c in G1_BIG] if a.big else [])
    if a.big:
        todo += [("g2", c) for c in G2_BIG] + [("g2", c) for c in G2_WEAK]
    for gname, (N, ny, pct) in todo:
        if (gname, N, ny, pct) in existing:
            continue
        row = run_case(g1 if gname == "g1" else g2, N, ny, pct, plain, pairs, count_only=(N, ny, pct) in G2_WEAK)
        row["gen"] = gname
        print(row, flush=True)
        rows.append(row)
        path.write_text(json.dumps({"source": "compiled reference hash_join.cpp @4c756a9a via oracle/build_ref.sh", "cases": rows}, indent=1))
    fx = small_fixtures(pairs)
    for name, d in fx.items():
        np.savez_compressed(HERE / f"{name}.npz", **d)
    print("wrote", path, "and", sorted(fx))


if __name__ == "__main__":
    main()
