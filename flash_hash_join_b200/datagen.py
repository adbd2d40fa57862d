"""Synthetic join inputs shaped like the h2o db-benchmark join tables.

G1 restates /root/reference/db-benchmark/_data/join-datagen.R (split_xlr :40-47, sample_all
:34-38, LHS column :102-104, RHS unique key columns :139/:155/:172, v2 :146/:162/:180) as a
deterministic numpy generator: key universe 1..U with U = 2*ny - c,
c = ny*match_pct/100 keys common to both sides, unique build keys, probe keys sampled with
replacement from the ny probe-side keys.  The order of RNG calls is part of the specification —
the golden values in tests/golden/ depend on it.

G2 is a counter-based generator with the same shape whose every element is a pure function of
(seed, index), so shards can be produced independently (per GPU / per rank) and at 1e9 rows
without a global permutation.  Build rows come in a pseudo-random order (a cycle-walking bijective
hash of the row index picks the key id), like the shuffled RHS tables of join-datagen.R: an
arithmetic progression of build keys would make a radix partition pass artificially free of
shared-memory bank conflicts.
"""
from __future__ import annotations

import numpy as np

__all__ = ["g1", "g2", "g2_slice", "CONFIGS"]

# BASELINE.json configs -> (N probe rows, ny build rows, match_pct)
CONFIGS = {
    "C1": (10_000_000, 10_000, 90),
    "C2": (100_000_000, 100_000, 10),
    "C3": (100_000_000, 100_000_000, 90),
    "C4": (1_000_000_000, 1_000_000, 90),
    "C5": (1_000_000_000, 1_000_000_000, 90),
    "C4s": (125_000_000, 1_000_000, 90),  # shape of one GPU's share of C4 on 8 GPUs (probe split, build replicated)
    "T": (200_000, 150_000, 90),  # tiny radix-shaped case for compute-sanitizer racecheck
    "S": (2_000_000, 1_500_000, 90),  # small radix-shaped case for sanitizer / sanity runs (not a BASELINE config)
}


def g1(N: int, ny: int, match_pct: int = 90, seed: int = 108):
    """Return (build_keys, build_values, probe_keys) as uint64 arrays (numpy PCG64, seed 108 =
    set.seed(108) of join-datagen.R:89)."""
    rng = np.random.default_rng(seed)
    c = (ny * match_pct) // 100
    U = 2 * ny - c
    key = rng.permutation(U).astype(np.int64) + 1
    x, l, r = key[:c], key[c:ny], key[ny:U]
    pd_ = np.concatenate([x, l])
    extra = rng.choice(pd_, size=N - ny, replace=True) if N > ny else pd_[:0]
    probe = rng.permutation(np.concatenate([pd_, extra]))
    build = rng.permutation(np.concatenate([x, r]))
    vals = rng.integers(0, 100, size=ny, dtype=np.int64)
    return build.view(np.uint64), vals.view(np.uint64), probe.view(np.uint64)


# ---- G2: counter-based -------------------------------------------------------------------------
_M1 = np.uint64(0xFF51AFD7ED558CCD)
_M2 = np.uint64(0xC4CEB9FE1A85EC53)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(x: np.ndarray) -> np.ndarray:
    """murmur3 fmix64 (a bijection on uint64), vectorised."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= _M1
        x ^= x >> np.uint64(33)
        x *= _M2
        x ^= x >> np.uint64(33)
    return x


def _perm_index(i: np.ndarray, U: int, seed: int) -> np.ndarray:
    """A bijection of [0, U) computed per element: multiply by an odd constant co-prime to U then
    add an offset, modulo U (affine permutation; cheap in numpy and in a CUDA kernel)."""
    a = 0x9E3779B1 | 1
    while np.gcd(a, U) != 1:
        a += 2
    b = (seed * 0x85EBCA6B + 12345) % U
    return ((i.astype(np.uint64) * np.uint64(a % U)) % np.uint64(U) + np.uint64(b)) % np.uint64(U)


def _shuffle_index(j: np.ndarray, n: int, seed: int) -> np.ndarray:
    """A pseudo-random bijection of [0, n), computed per element: a bijective mixer on k = ceil(log2 n)
    bits (add, xorshift, odd multiply — each a bijection modulo 2^k), applied again while the
    result is >= n (cycle walking; fewer than two rounds on average)."""
    k = max(int(n - 1).bit_length(), 2)
    mask = np.uint64((1 << k) - 1)
    sh = np.uint64((k + 1) // 2)
    add = np.uint64((seed * 0x9E3779B97F4A7C15 + 0x7F4A7C15) & ((1 << k) - 1))

    def f(x):
        with np.errstate(over="ignore"):
            x = (x + add) & mask
            x ^= x >> sh
            x = (x * np.uint64(0xD6E8FEB86659FD93)) & mask
            x ^= x >> sh
            x = (x * np.uint64(0xCA5A826395121157)) & mask
            x ^= x >> sh
        return x

    x = f(j.astype(np.uint64, copy=True))
    todo = np.flatnonzero(x >= np.uint64(n))
    while todo.size:
        x[todo] = f(x[todo])
        todo = todo[x[todo] >= np.uint64(n)]
    return x


def g2_slice(N: int, ny: int, match_pct: int, seed: int, side: str, start: int, stop: int):
    """Elements [start, stop) of one side of the G2 data set.

    side == 'build' -> (keys, values); side == 'probe' -> keys.
    Keys 1..U; build key j is key id perm(j') with j' in the build id range; probe row j draws a
    key id uniformly (hash of (seed, j)) from the ny probe-side ids, of which c are shared with
    the build side — the same x/l/r split as G1 without materialising a permutation."""
    c = (ny * match_pct) // 100
    U = 2 * ny - c
    idx = np.arange(start, stop, dtype=np.uint64)
    if side == "build":
        # build row idx holds build id j = shuffle(idx); build ids: the c common ids [0, c) and the build-only ids [ny, U)
        j = _shuffle_index(idx, ny, seed)
        ids = np.where(j < np.uint64(c), j, j - np.uint64(c) + np.uint64(ny))
        keys = _perm_index(ids, U, seed) + np.uint64(1)
        with np.errstate(over="ignore"):
            vals = _mix64(j + np.uint64(seed) * _GOLD) % np.uint64(100)
        return keys, vals
    if side == "probe":
        with np.errstate(over="ignore"):
            r = _mix64((idx + np.uint64(1)) * _GOLD + np.uint64(seed))
        ids = r % np.uint64(ny)  # probe-side ids [0, ny): [0,c) common, [c,ny) probe-only
        return _perm_index(ids, U, seed) + np.uint64(1)
    raise ValueError(side)


def g2(N: int, ny: int, match_pct: int = 90, seed: int = 108):
    bk, bv = g2_slice(N, ny, match_pct, seed, "build", 0, ny)
    pk = g2_slice(N, ny, match_pct, seed, "probe", 0, N)
    return bk, bv, pk
