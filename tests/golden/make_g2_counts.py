#!/usr/bin/env python
"""Expected match counts of the multi-GPU bench workloads (G2 data): tests/golden/g2_counts.json.

    python tests/golden/make_g2_counts.py            # build container (oracle/_ref present)

Two independent sources per case:
  * "reference": the compiled reference (oracle/_ref/plain, unmodified hash_join.cpp) run on probe slices of 1.25e8
    rows against the whole build side — a probe row matches or not independently of the other probe rows, so the
    slice counts add up to the count of the whole job (only where the build side fits this container: C4);
  * "generator": the count the G2 generator implies — probe row j draws key id r_j mod ny, ids below
    c = ny * match_pct / 100 are the ones the (unique-key) build side holds — evaluated with numpy in chunks.
    It needs no join at all and is the only source for C5 (1e9 x 1e9 does not fit a CPU box).
bench.py checks `matches` of every multi-GPU run against this file."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from flash_hash_join_b200.datagen import _GOLD, _mix64, g2_slice  # noqa: E402

SEED = 108
# (name, total probe rows, build rows, match %, reference?)
CASES = [("C4 at %d GPUs (probe rows [0, %d))" % (g, g * 125_000_000), g * 125_000_000, 1_000_000, 90, True) for g in (1, 2, 4, 8)]
CASES += [("C3", 100_000_000, 100_000_000, 90, False), ("C5", 1_000_000_000, 1_000_000_000, 90, False),
          ("C2 at 8 GPUs", 800_000_000, 100_000, 10, False)]


def generator_count(N, ny, pct, chunk=20_000_000):
    c = (ny * pct) // 100
    n = 0
    for s in range(0, N, chunk):
        idx = np.arange(s, min(N, s + chunk), dtype=np.uint64)
        with np.errstate(over="ignore"):
            r = _mix64((idx + np.uint64(1)) * _GOLD + np.uint64(SEED))
        n += int(np.count_nonzero(r % np.uint64(ny) < np.uint64(c)))
    return n


def main():
    from oracle import oracle as O

    ref = O.load_reference("plain") if O.reference_available("plain") else None
    out = {"seed": SEED, "cases": []}
    slice_counts = {}
    for name, N, ny, pct, want_ref in CASES:
        e = {"name": name, "N": N, "ny": ny, "match_pct": pct, "generator_count": generator_count(N, ny, pct), "reference_count": None}
        if want_ref and ref is not None:
            bk, bv = g2_slice(N, ny, pct, SEED, "build", 0, ny)
            tot = 0
            for s in range(0, N, 125_000_000):
                if (ny, s) not in slice_counts:
                    pk = g2_slice(N, ny, pct, SEED, "probe", s, s + 125_000_000)
                    slice_counts[(ny, s)] = int(ref.adaptive_join_count(bk, bv, pk)[0])
                    del pk
                tot += slice_counts[(ny, s)]
            e["reference_count"] = tot
            assert tot == e["generator_count"], e
        out["cases"].append(e)
        print(e, flush=True)
    (Path(__file__).resolve().parent / "g2_counts.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
