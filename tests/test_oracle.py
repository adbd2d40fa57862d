"""CPU tests: the oracle (C restatement + numpy restatement) against the golden vectors produced
by the compiled reference, and — when oracle/_ref is present — against the reference itself."""
import numpy as np
import pytest

from flash_hash_join_b200.datagen import g1, g2
from oracle import oracle as O

ALL = sorted(O.ENTRY_POINTS)  # (algo, bloom, materialize)

# hash64(k, 0xAAAAAAAA) known answers probed from the reference binary
HASH_KAT = {
    0: 0x5D4DC3025D6589A6, 1: 0xA766855B1459F481, 2: 0x69999F78CF1D73E8, 42: 0xF4A24149802612AB,
    1000000: 0x1DA45A840B9BF8CC, 2**32: 0xADF8B4AA8020231E, 2**63: 0x1FCDEDEADF93B2DE,
    2**64 - 1: 0x19FFA209992A70EB, 0xDEADBEEFCAFEBABE: 0xE5E9AE1060ADEF30,
}
BLOOM_KAT = {0: (0x5D, 747, 0x4809), 1: (0xA7, 162, 0x8884), 2: (0x69, 1656, 0x0950), 42: (0xF4, 1025, 0x7200),
             1000000: (0x1D, 92, 0xD800), 0xDEADBEEFCAFEBABE: (0xE5, 773, 0x2804)}


def test_hash_known_answers():
    for k, h in HASH_KAT.items():
        assert O.hash64(k) == h, hex(k)
    for k, (part, bidx, mask) in BLOOM_KAT.items():
        h = O.hash64(k)
        assert O.partition_idx(k) == part == h >> 56
        assert (h & 0xFFFFFFFF) >> 21 == bidx
        assert O.bloom_tag(h) == mask


def test_capacity_rule():
    # hash_join.cpp:99 next_pow2(size_t(n*1.5 + 32)); SURVEY.md §8a2 sizes
    assert O.table_capacity(10**4) == 16384
    assert O.table_capacity(10**5) == 262144
    assert O.table_capacity(10**6) == 2097152
    assert O.table_capacity(10**8) == 268435456
    assert O.table_capacity(0) == 32


def test_adaptive_threshold():
    assert O._load().fjo_adaptive_path(999_999) == 1
    assert O._load().fjo_adaptive_path(1_000_000) == 2  # exactly 1e6 goes to radix (hash_join.cpp:580)


@pytest.mark.parametrize("case_idx", range(9))
def test_golden_counts_and_checksums(golden_cases, case_idx):
    cases = [c for c in golden_cases if c["N"] <= 10**7]
    if case_idx >= len(cases):
        pytest.skip("no such case")
    c = cases[case_idx]
    gen = g1 if c["gen"] == "g1" else g2
    bk, bv, pk = gen(c["N"], c["ny"], c["match_pct"], c["seed"])
    n, k, v = O.np_join(bk, bv, pk)
    assert n == c["count"]
    cs = O.checksums(k, v)
    for f in ("sum_keys", "xor_keys", "sum_vals"):
        assert cs[f] == c[f], f
    if c["N"] <= 10**6:  # the C restatement on every entry point (single-threaded: keep it small)
        for algo, bloom, mat in ALL:
            n2, k2, v2 = O.join(algo, bloom, mat, bk, bv, pk)
            assert n2 == c["count"], (algo, bloom, mat)
            if mat:
                assert O.checksums(k2, v2) == cs


def test_fixture_pairs_exact_order(fixtures_npz):
    """The committed fixtures hold the reference's own output arrays; the C restatement reproduces
    them element for element (probe order on the scalar path, partition-major on the radix path)."""
    f = fixtures_npz["g1_5000_600"]
    n, k, v = O.join("scalar", False, True, f["bk"], f["bv"], f["pk"])
    assert np.array_equal(k, f["rk"]) and np.array_equal(v, f["rv"])
    f = fixtures_npz["edge_keys"]
    n, k, v = O.join("scalar", True, True, f["bk"], f["bv"], f["pk"])
    assert np.array_equal(k, f["rk"]) and np.array_equal(v, f["rv"])
    assert n == 11 == len(f["rk"])
    f = fixtures_npz["dup_build_radix"]
    n, k, v = O.join("radix", False, True, f["bk"], f["bv"], f["pk"])
    assert np.array_equal(k, f["rk"]) and np.array_equal(v, f["rv"])
    # keep-first: the scalar restatement (1 thread) and numpy agree with the radix output as multisets
    n2, k2, v2 = O.join("scalar", False, True, f["bk"], f["bv"], f["pk"])
    n3, k3, v3 = O.np_join(f["bk"], f["bv"], f["pk"])
    assert np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs(k2, v2))
    assert np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs(k3, v3))
    f = fixtures_npz["skew_probe"]
    n, k, v = O.join("adaptive", False, True, f["bk"], f["bv"], f["pk"])
    assert n == 3000 and np.array_equal(k, f["rk"]) and np.array_equal(v, f["rv"])


@pytest.mark.parametrize("nb,np_", [(0, 0), (0, 10), (10, 0), (1, 1), (10, 2047), (1000, 2048), (1000, 2049), (4096, 10000)])
def test_c_oracle_vs_numpy_small(nb, np_):
    rng = np.random.default_rng(nb * 7919 + np_)
    bk = rng.integers(0, max(2 * nb, 4), nb).astype(np.uint64)  # duplicates on purpose
    bv = rng.integers(0, 2**63, nb).astype(np.uint64)
    pk = rng.integers(0, max(3 * nb, 4), np_).astype(np.uint64)
    n0, k0, v0 = O.np_join(bk, bv, pk)
    for algo, bloom, mat in ALL:
        n, k, v = O.join(algo, bloom, mat, bk, bv, pk)
        assert n == n0
        if mat:
            assert np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs(k0, v0))


def test_radix_partition_is_stable():
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 1000, 5000).astype(np.uint64)
    vals = np.arange(5000, dtype=np.uint64)
    ok, ov, off = O.radix_partition(keys, vals)
    assert off[0] == 0 and off[256] == 5000
    for p in range(256):
        seg = ov[int(off[p]):int(off[p + 1])]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)  # input order preserved inside a partition
        for kk in ok[int(off[p]):int(off[p + 1])][:3]:
            assert O.partition_idx(int(kk)) == p


@pytest.mark.skipif(not O.reference_available("plain"), reason="oracle/_ref not built (needs /root/reference)")
def test_c_oracle_matches_compiled_reference():
    plain, pairs = O.load_reference("plain"), O.load_reference("pairs")
    rng = np.random.default_rng(11)
    bk = rng.permutation(300000)[:80000].astype(np.uint64)
    bv = rng.integers(0, 100, 80000).astype(np.uint64)
    pk = rng.integers(0, 300000, 400000).astype(np.uint64)
    for algo, bloom, mat in ALL:
        name = O.entry_point_name(algo, bloom, mat)
        n, k, v = O.join(algo, bloom, mat, bk, bv, pk)
        if mat:
            r = getattr(pairs, name)(bk, bv, pk)
            assert r[0] == n and np.array_equal(r[2], k) and np.array_equal(r[3], v), name
        else:
            assert getattr(plain, name)(bk, bv, pk)[0] == n, name
    # edge keys 0 and 2^64-1 join correctly on the reference
    bk = np.array([0, 2**64 - 1, 7], dtype=np.uint64)
    bv = np.array([1, 2, 3], dtype=np.uint64)
    pk = np.array([0, 2**64 - 1, 8, 7, 7], dtype=np.uint64)
    for name in ("hash_join_count", "hash_join_count_bloom", "hash_join_count_radix"):
        assert getattr(plain, name)(bk, bv, pk)[0] == 4 == O.join("scalar", False, False, bk, bv, pk)[0]


@pytest.mark.skipif(not O.reference_available("pairs"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(8))
def test_numpy_restatement_vs_compiled_reference_randomized(seed):
    """np_join is the checker of most GPU parity tests: pin IT against the compiled reference on seeded random shapes
    and key domains (dense / sparse 32- and 64-bit, edge keys 0 and 2^64-1).  Unique build keys: every reference
    entry point must agree.  Duplicate build keys: only the reference's radix path is deterministic (keep-first,
    hash_join.cpp:125 + :226-234), so duplicates are checked against that path alone."""
    plain, pairs = O.load_reference("plain"), O.load_reference("pairs")
    rng = np.random.default_rng(500 + seed)
    for _ in range(4):
        nb = int(np.exp(rng.uniform(0, np.log(60_000))))
        np_ = int(np.exp(rng.uniform(0, np.log(300_000))))
        U = [max(nb + 1, 2 * nb), 2**32 - 1, 2**64 - 1][int(rng.integers(0, 3))]
        if U <= 4 * nb + 8:
            bk = rng.permutation(U)[:nb].astype(np.uint64)
        else:
            bk = np.unique(rng.integers(0, U, nb, dtype=np.uint64, endpoint=True))
        if rng.random() < 0.5:
            bk = np.concatenate([bk, np.array([0, 2**64 - 1], dtype=np.uint64)])
            bk = np.unique(bk)
        dup = rng.random() < 0.4 and bk.size > 1
        if dup:
            bk = np.concatenate([bk, rng.choice(bk, max(1, bk.size // 7))])
        rng.shuffle(bk)
        bv = rng.integers(0, 2**64 - 1, bk.size, dtype=np.uint64)
        n_hit = int(np_ * rng.uniform(0, 1))
        pk = np.concatenate([rng.choice(bk, n_hit), rng.integers(0, min(2**64 - 1, max(2 * U, 16)), np_ - n_hit, dtype=np.uint64, endpoint=True)])
        rng.shuffle(pk)
        n0, k0, v0 = O.np_join(bk, bv, pk)
        sp0 = O.sorted_pairs(k0, v0)
        for algo, bloom, mat in ALL:
            if dup and algo != "radix":
                continue
            name = O.entry_point_name(algo, bloom, mat)
            if mat:
                r = getattr(pairs, name)(bk, bv, pk)
                assert r[0] == n0, (name, nb, np_, U)
                assert np.array_equal(O.sorted_pairs(r[2], r[3]), sp0), (name, nb, np_, U)
            else:
                assert getattr(plain, name)(bk, bv, pk)[0] == n0, (name, nb, np_, U)
