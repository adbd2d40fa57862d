"""GPU parity tests (run with `-m gpu` on a B200): the CUDA engine, called through the reference-
facing pybind11 module and through the C ABI, against the oracle on the same seeded inputs.

Bar: bit-exact — identical match counts and identical sorted multiset of (probe key, build value)
pairs.  All arithmetic on the path is integer; there is no tolerance anywhere in this file.
Nothing here reads /root/reference (absent on the GPU box): the oracle is oracle/join_oracle.c,
the numpy restatement, and the committed golden vectors produced by the reference binary."""
import os

import numpy as np
import pytest

from flash_hash_join_b200.datagen import g1, g2, g2_slice
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ALL = sorted(O.ENTRY_POINTS)  # (algo, bloom, materialize)


@pytest.fixture(scope="module")
def fj():
    from flash_hash_join_b200 import flash_join

    flash_join.initialize()
    return flash_join


@pytest.fixture(scope="module")
def capi():
    from flash_hash_join_b200 import capi as c

    return c


def run_entry(fj, algo, bloom, mat, bk, bv, pk):
    name = O.entry_point_name(algo, bloom, mat)
    n, sec = getattr(fj, name)(bk, bv, pk)
    assert isinstance(n, int) and isinstance(sec, float) and sec >= 0.0
    pairs = fj.last_pairs() if mat else None
    return n, pairs


def check_all_entry_points(fj, bk, bv, pk, expect=None, algos=("adaptive", "scalar", "radix")):
    if expect is None:
        expect = O.np_join(bk, bv, pk)
    n0, k0, v0 = expect
    sp0 = O.sorted_pairs(k0, v0)
    for algo, bloom, mat in ALL:
        if algo not in algos:
            continue
        n, pairs = run_entry(fj, algo, bloom, mat, bk, bv, pk)
        assert n == n0, (algo, bloom, mat, n, n0, fj.last_stats())
        if mat:
            k, v = pairs
            assert k.size == n0 and v.size == n0
            assert np.array_equal(O.sorted_pairs(k, v), sp0), (algo, bloom, mat, fj.last_stats())


# ------------------------------------------------------------------------------------------------
def test_smoke_first_kernel(fj):
    bk = np.array([5, 7, 9, 11], dtype=np.uint64)
    bv = np.array([50, 70, 90, 110], dtype=np.uint64)
    pk = np.array([7, 7, 1, 11, 12, 5], dtype=np.uint64)
    n, sec = fj.hash_join_count(bk, bv, pk)
    assert n == 4
    st = fj.last_stats()
    assert st["kernel_launches"] >= 1 and st["path"] == "scalar"
    n, sec = fj.hash_join(build_keys=bk, build_values=bv, probe_keys=pk)
    k, v = fj.last_pairs()
    assert n == 4 and np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs([5, 7, 7, 11], [50, 70, 70, 110]))


@pytest.mark.parametrize("nb,np_", [(0, 0), (0, 100), (100, 0), (1, 1), (1, 5000), (10, 2047), (1000, 2048), (1000, 2049),
                                    (4095, 4097), (10**4, 10**5), (65536, 300001)])
def test_shape_matrix_unique_build(fj, nb, np_):
    rng = np.random.default_rng(nb * 31 + np_)
    bk = (rng.permutation(max(3 * nb, 8))[:nb] + 1).astype(np.uint64)
    bv = rng.integers(0, 1000, nb).astype(np.uint64)
    pk = rng.integers(0, max(3 * nb, 8) + 2, np_).astype(np.uint64)
    check_all_entry_points(fj, bk, bv, pk)


@pytest.mark.parametrize("pct", [0, 10, 90, 100])
def test_match_rates(fj, pct):
    bk, bv, pk = g1(200_000, 20_000, pct)
    n0 = O.np_join(bk, bv, pk)
    assert abs(n0[0] / 200_000 - pct / 100) < 0.02
    check_all_entry_points(fj, bk, bv, pk, expect=n0)


def test_c_oracle_agrees_on_gpu_case(fj):
    """Same seeded input through the C restatement of the reference (not only numpy)."""
    bk, bv, pk = g1(300_000, 30_000, 90)
    for algo, bloom, mat in ALL:
        n0, k0, v0 = O.join(algo, bloom, mat, bk, bv, pk)
        n, pairs = run_entry(fj, algo, bloom, mat, bk, bv, pk)
        assert n == n0
        if mat:
            assert np.array_equal(O.sorted_pairs(*pairs), O.sorted_pairs(k0, v0))


@pytest.mark.parametrize("idx", range(9))
def test_golden_vectors(fj, golden_cases, idx):
    """Counts and pair checksums the unmodified reference produced (tests/golden/g1_goldens.json)."""
    cases = [c for c in golden_cases if c["N"] <= 10**7]
    if idx >= len(cases):
        pytest.skip("no such case")
    c = cases[idx]
    gen = g1 if c["gen"] == "g1" else g2
    bk, bv, pk = gen(c["N"], c["ny"], c["match_pct"], c["seed"])
    for algo, bloom, mat in ALL:
        n, pairs = run_entry(fj, algo, bloom, mat, bk, bv, pk)
        assert n == c["count"], (algo, bloom, mat)
        if mat:
            cs = O.checksums(*pairs)
            assert (cs["sum_keys"], cs["xor_keys"], cs["sum_vals"]) == (c["sum_keys"], c["xor_keys"], c["sum_vals"])


def test_reference_fixtures(fj, fixtures_npz):
    """Committed input/output arrays of the reference binary, incl. edge keys (0, 2^64-1, >= 2^32,
    64-bit values), duplicate build keys (keep-first = the reference's radix path) and probe skew."""
    for name, f in fixtures_npz.items():
        expect = (len(f["rk"]), f["rk"], f["rv"])
        check_all_entry_points(fj, f["bk"], f["bv"], f["pk"], expect=expect)


def test_edge_keys_and_sentinel(fj):
    E = 2**64 - 1
    bk = np.array([E, 0, 1, 2**32 - 1, 2**32, E - 1, 2**63], dtype=np.uint64)
    bv = np.array([E, 0, 2**32, 5, 6, 7, E - 1], dtype=np.uint64)
    pk = np.array([E, E, 0, 3, 2**32 - 1, 2**32, 2**63, E - 1, E - 2, 1, E], dtype=np.uint64)
    check_all_entry_points(fj, bk, bv, pk)
    # sentinel key present in the probe side only / build side only
    check_all_entry_points(fj, bk[1:], bv[1:], pk)
    check_all_entry_points(fj, bk, bv, pk[2:10])
    # duplicate sentinel keys on the build side: keep-first
    bk2 = np.array([E, 5, E, E], dtype=np.uint64)
    bv2 = np.array([10, 20, 30, 40], dtype=np.uint64)
    check_all_entry_points(fj, bk2, bv2, np.array([E, 5, E, 6], dtype=np.uint64))


def test_wide_keys_and_values(fj):
    rng = np.random.default_rng(5)
    nb, np_ = 50_000, 400_000
    bk = np.unique(rng.integers(0, 2**64 - 1, nb, dtype=np.uint64))
    bv = rng.integers(0, 2**64 - 1, bk.size, dtype=np.uint64)
    pk = np.concatenate([rng.choice(bk, np_ // 2), rng.integers(0, 2**64 - 1, np_ // 2, dtype=np.uint64)])
    rng.shuffle(pk)
    check_all_entry_points(fj, bk, bv, pk)
    assert fj.last_stats()["narrow"] is False
    # narrow keys but one 64-bit value: the optimistic packed attempt must be abandoned and re-run
    bk, bv, pk = g1(100_000, 10_000, 90)
    bv = bv.copy(); bv[1234] = 2**40
    n0 = O.np_join(bk, bv, pk)
    for algo in ("scalar", "radix"):
        n, _ = run_entry(fj, algo, False, True, bk, bv, pk)
        st = fj.last_stats()
        if st["dense"]:  # global-table path on a dense key domain: the direct-address table holds 64-bit values
            assert algo == "scalar" and n == n0[0] and st["attempts"] == 1
        else:
            assert n == n0[0] and st["attempts"] == 2 and st["narrow"] is False
        assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(n0[1], n0[2]))
    fj.configure(dense=0)
    try:
        n, _ = run_entry(fj, "scalar", False, True, bk, bv, pk)
        st = fj.last_stats()
        assert n == n0[0] and st["attempts"] == 2 and st["narrow"] is False
        assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(n0[1], n0[2]))
    finally:
        fj.configure(dense=1)


def test_force_wide_flag(fj):
    bk, bv, pk = g1(100_000, 10_000, 90)
    n0 = O.np_join(bk, bv, pk)
    for algo in ("scalar", "radix"):
        for bloom in (False, True):
            n, _ = fj.join_flags(algo, bloom, True, bk, bv, pk, force_wide=True)
            assert n == n0[0] and fj.last_stats()["narrow"] is False
            assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(n0[1], n0[2]))
        n, _ = fj.join_flags(algo, False, True, bk, bv, pk)
        assert n == n0[0] and fj.last_stats()["narrow"] is True


def test_duplicate_build_keys_keep_first(fj):
    """Duplicate build keys: the contract is keep-first (the reference's radix path and its
    1-thread scalar path, hash_join.cpp:125/:191); the engine detects duplicates and re-runs the
    exact path."""
    rng = np.random.default_rng(9)
    bk = rng.integers(0, 5000, 60_000).astype(np.uint64)
    bv = np.arange(60_000, dtype=np.uint64)  # value identifies the build row
    pk = rng.integers(0, 6000, 200_000).astype(np.uint64)
    n0, k0, v0 = O.join("radix", False, True, bk, bv, pk)
    assert (n0, ) == (O.np_join(bk, bv, pk)[0], )
    for algo, bloom, mat in ALL:
        n, pairs = run_entry(fj, algo, bloom, mat, bk, bv, pk)
        assert n == n0
        st = fj.last_stats()
        # the exact bitmap count (dense key domain) is a set: duplicates need no keep-first pass there
        assert st["dedup_exact"] is True or (not mat and st["dense"] == 1), st
        if mat:
            assert np.array_equal(O.sorted_pairs(*pairs), O.sorted_pairs(k0, v0))


def test_probe_skew(fj):
    bk = np.arange(1, 50_001, dtype=np.uint64)
    bv = bk * np.uint64(7)
    pk = np.full(1_000_000, 4242, dtype=np.uint64)  # one partition receives every probe row
    check_all_entry_points(fj, bk, bv, pk)
    pk2 = np.full(1_000_000, 99_999_999, dtype=np.uint64)  # ...and none of them match
    check_all_entry_points(fj, bk, bv, pk2)


def test_bloom_invariance_and_kinds(fj):
    bk, bv, pk = g1(500_000, 50_000, 10)
    n_dense, _ = fj.hash_join_count_bloom(bk, bv, pk)
    assert fj.last_stats()["bloom_kind"] == "bitmap" and fj.last_stats()["dense"] == 1
    fj.configure(dense=0)  # the general hash path: table + register-blocked Bloom filter
    n_plain, _ = fj.hash_join_count(bk, bv, pk)
    n_bloom, _ = fj.hash_join_count_bloom(bk, bv, pk)
    assert fj.last_stats()["bloom_kind"] == "smem" and fj.last_stats()["dense"] == 0
    assert n_dense == n_plain == n_bloom == O.np_join(bk, bv, pk)[0]
    fj.configure(smem_bloom=0, bloom_guard=0)
    try:
        n_g, _ = fj.hash_join_count_bloom(bk, bv, pk)
        assert fj.last_stats()["bloom_kind"] == "global" and n_g == n_plain
        n_g, _ = fj.hash_join_bloom(bk, bv, pk)
        assert n_g == n_plain
        # a build side too large for the shared-memory filter takes the global (L2) filter ...
        fj.configure(smem_bloom=1)
        bk, bv, pk = g1(600_000, 400_000, 10)
        n_b, _ = fj.hash_join_count_bloom(bk, bv, pk)
        assert fj.last_stats()["bloom_kind"] == "global" and n_b == O.np_join(bk, bv, pk)[0]
        # ... unless the guard is on (the default): next to an L2-resident table a filter read through L2 only costs
        fj.configure(bloom_guard=1)
        n_b2, _ = fj.hash_join_count_bloom(bk, bv, pk)
        assert fj.last_stats()["bloom_kind"] == "none" and n_b2 == n_b
    finally:
        fj.configure(smem_bloom=1, dense=1, bloom_guard=1)


def test_radix_bloom_partition_filter(fj):
    """hash_join_radix_bloom / hash_join_count_radix_bloom on the general radix path: k_join builds a per-partition
    register-blocked filter in shared memory and checks it before the table (the FlashHashTable<true> per partition of
    hash_join.cpp:344, :518); k_join3's membership bitmap is reported as the (exact) filter.  Results never change."""
    fj.configure(dense=0)
    try:
        for N, ny, pct, kind in ((600_000, 150_000, 10, "partition"), (600_000, 150_000, 90, "partition"), (3_000_000, 3_000_000, 50, None)):
            bk, bv, pk = g1(N, ny, pct)
            n0, k0, v0 = O.np_join(bk, bv, pk)
            n, _ = fj.hash_join_count_radix_bloom(bk, bv, pk)
            st = fj.last_stats()
            assert n == n0 and st["path"] == "radix", st
            if kind:
                assert st["bloom_kind"] == kind, st
            else:
                assert st["bloom_kind"] in ("partition", "bitmap"), st
            n, _ = fj.hash_join_radix_bloom(bk, bv, pk)
            k, v = fj.last_pairs()[:2]
            assert n == n0 and np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs(k0, v0))
            # wide rows take k_join as well
            bkw = bk | (np.uint64(1) << np.uint64(40))
            pkw = pk | (np.uint64(1) << np.uint64(40))
            n, _ = fj.hash_join_count_radix_bloom(bkw, bv, pkw)
            assert n == n0
    finally:
        fj.configure(dense=1)


def test_radix_two_pass_small_partitions(fj):
    """Force tiny shared-memory partitions so that the two-pass scatter and many partitions are
    exercised at a size the oracle checks in seconds."""
    bk, bv, pk = g1(400_000, 300_000, 90)
    expect = O.np_join(bk, bv, pk)
    fj.configure(radix_sub_rows=256)
    try:
        check_all_entry_points(fj, bk, bv, pk, expect=expect, algos=("radix",))
        st = fj.last_stats()
        assert st["path"] == "radix" and st["radix_bits"][1] > 0, st
        fj.configure(radix_sub_rows=2048)
        check_all_entry_points(fj, bk, bv, pk, expect=expect, algos=("radix",))
        assert fj.last_stats()["path"] == "radix"
        # wide rows through the same machinery
        n, _ = fj.join_flags("radix", False, True, bk, bv, pk, force_wide=True)
        assert n == expect[0] and fj.last_stats()["path"] == "radix"
        assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(expect[1], expect[2]))
    finally:
        fj.configure(radix_sub_rows=0)


def test_adaptive_prefers_dense16_for_big_probe_materialize(fj):
    """Adaptive policy on a dense key domain (profiles/r02t_sweep_adaptive.jsonl, r02u_exp_small_dense16.jsonl): a
    materialize with a big probe side takes the dense16 radix path whatever the build size, a count keeps the
    shared-memory bitmap while it fits; a key outside the domain costs one abandoned attempt and never the result."""
    N = (1 << 24) + 12_345
    bk, bv, pk = g1(N, 50_000, 90)
    expect = O.np_join(bk, bv, pk)
    n, _ = fj.adaptive_join(bk, bv, pk)
    st = fj.last_stats()
    assert n == expect[0] and st["path"] == "radix" and st["dense"] == 2 and st["attempts"] == 1, st
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()[:2]), O.sorted_pairs(expect[1], expect[2]))
    n, _ = fj.adaptive_join_count(bk, bv, pk)
    st = fj.last_stats()
    assert n == expect[0] and st["path"] == "scalar" and st["dense"] == 1, st
    bk2 = bk.copy()
    bk2[7] = np.uint64(3_000_000_000)  # outside every dense domain, still a packed row
    expect2 = O.np_join(bk2, bv, pk)
    n, _ = fj.adaptive_join(bk2, bv, pk)
    st = fj.last_stats()
    assert n == expect2[0] and st["attempts"] >= 2 and st["dense"] == 0, st
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()[:2]), O.sorted_pairs(expect2[1], expect2[2]))


def test_adaptive_materialize_samples_the_match_rate(fj):
    """Small dense build side, big probe side: k_sel_sample measures the match rate in front of the dense16 attempt; few
    matches abandon it for the dense table path (one extra attempt), many keep it — the pairs are the oracle's either way,
    and the threshold is a config key (0 = never sample)."""
    N = (1 << 24) + 4_321
    for pct, low in ((10, True), (90, False)):
        bk, bv, pk = g1(N, 60_000, pct)
        expect = O.np_join(bk, bv, pk)
        n, _ = fj.adaptive_join(bk, bv, pk)
        st = fj.last_stats()
        assert n == expect[0], (pct, st)
        if low:
            assert st["path"] == "scalar" and st["dense"] == 1 and st["attempts"] == 2, st
        else:
            assert st["path"] == "radix" and st["dense"] == 2 and st["attempts"] == 1, st
        assert np.array_equal(O.sorted_pairs(*fj.last_pairs()[:2]), O.sorted_pairs(expect[1], expect[2]))
    fj.configure(dense16_sel_min_pct=0)
    try:
        n, _ = fj.adaptive_join(bk, bv, pk)  # the 90 % inputs
        assert n == expect[0] and fj.last_stats()["dense"] == 2
        bk, bv, pk = g1(N, 60_000, 10)
        expect = O.np_join(bk, bv, pk)
        n, _ = fj.adaptive_join(bk, bv, pk)
        st = fj.last_stats()
        assert n == expect[0] and st["path"] == "radix" and st["dense"] == 2 and st["attempts"] == 1, st
        assert np.array_equal(O.sorted_pairs(*fj.last_pairs()[:2]), O.sorted_pairs(expect[1], expect[2]))
    finally:
        fj.configure(dense16_sel_min_pct=50)


def test_result_word_mapped_and_copied_agree(fj):
    """The attempt's control block comes back through mapped pinned memory (k_publish_ctl + host spin) by default and through
    cudaMemcpyAsync + cudaStreamSynchronize with mapped_result = 0: same counts, same pairs, on a path with one attempt and
    on one that abandons its first attempt."""
    bk, bv, pk = g1(400_000, 40_000, 90)
    bk2 = bk.copy()
    bk2[11] = np.uint64(2**40)  # forces a second attempt (wide rows)
    for keys in (bk, bk2):
        expect = O.np_join(keys, bv, pk)
        got = {}
        for mapped in (1, 0):
            fj.configure(mapped_result=mapped)
            try:
                for name in ("hash_join_count", "hash_join_radix", "adaptive_join"):
                    n, _ = getattr(fj, name)(keys, bv, pk)
                    assert n == expect[0], (name, mapped, fj.last_stats())
                    if name != "hash_join_count":
                        assert np.array_equal(O.sorted_pairs(*fj.last_pairs()[:2]), O.sorted_pairs(expect[1], expect[2])), (name, mapped)
                    got[(name, mapped)] = fj.last_stats()["attempts"]
            finally:
                fj.configure(mapped_result=1)
        assert all(got[(nm, 1)] == got[(nm, 0)] for nm in ("hash_join_count", "hash_join_radix", "adaptive_join")), got


    for ny in (20_000, 3_000_000):
        bk, bv, pk = g1(3_000_000, ny, 90)
        n_a, _ = fj.adaptive_join_count(bk, bv, pk)
        path = fj.last_stats()["path"]
        n_s, _ = fj.hash_join_count(bk, bv, pk)
        n_r, _ = fj.hash_join_count_radix(bk, bv, pk)
        assert n_a == n_s == n_r == O.np_join(bk, bv, pk)[0]
        assert path in ("scalar", "radix")


def test_probe_idx_output(fj):
    bk, bv, pk = g1(200_000, 20_000, 90)
    n, _ = fj.join_flags("scalar", False, True, bk, bv, pk, probe_idx=True)
    k, v, ix = fj.last_pairs(with_probe_idx=True)
    assert n == k.size == ix.size
    assert np.array_equal(pk[ix.astype(np.int64)], k)  # idx maps back to the probe key
    assert np.unique(ix).size == n  # each probe row at most once
    assert np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs(*O.np_join(bk, bv, pk)[1:]))


def test_int64_and_forcecast_inputs(fj):
    bk = np.array([-1, 5, 7], dtype=np.int64)  # -1 is the same bits as 2^64-1
    bv = np.array([1, 2, 3], dtype=np.int64)
    pk = np.array([-1, 7, 8], dtype=np.int64)
    assert fj.hash_join_count(bk, bv, pk)[0] == 2
    assert fj.hash_join_count([5, 7], [1, 2], [7, 7, 9])[0] == 2  # lists are converted like forcecast
    assert fj.hash_join_count(np.array([5, 7], dtype=np.int32), np.array([1, 2], dtype=np.int32), np.array([7], dtype=np.uint8))[0] == 1
    assert fj.hash_join_count(np.arange(10, dtype=np.uint64)[::2], np.arange(5, dtype=np.uint64), np.array([4, 5], dtype=np.uint64))[0] == 1


def test_c_abi_device_resident_inputs(capi):
    bk, bv, pk = g1(1_000_000, 100_000, 10)
    n0, k0, v0 = O.np_join(bk, bv, pk)
    dbk, dbv, dpk = (capi.DeviceArray.from_host(x) for x in (bk, bv, pk))
    for algo in (capi.ALGO_ADAPTIVE, capi.ALGO_SCALAR, capi.ALGO_RADIX):
        for flags in (0, capi.FLAG_BLOOM, capi.FLAG_MATERIALIZE, capi.FLAG_MATERIALIZE | capi.FLAG_BLOOM):
            n, sec, st = capi.join(algo, flags, dbk, dbv, dpk)
            assert n == n0 and st["h2d_bytes"] == 0 and sec > 0
            if flags & capi.FLAG_MATERIALIZE:
                assert np.array_equal(O.sorted_pairs(*capi.pairs()), O.sorted_pairs(k0, v0))
                assert st["algorithmic_bytes"] == 16 * bk.size + 8 * pk.size + 16 * n0
            else:
                assert st["algorithmic_bytes"] == 8 * (bk.size + pk.size)
    # unaligned device pointer (8-byte aligned only): the probe kernel must not use 128-bit loads
    import ctypes as C

    class View(capi.DeviceArray):
        def __init__(self, base, off, n):
            self.ptr, self.n, self._base = base.ptr + 8 * off, n, base

        def free(self):
            pass

    n, _, _ = capi.join(capi.ALGO_SCALAR, 0, dbk, dbv, View(dpk, 1, pk.size - 1))
    assert n == O.np_join(bk, bv, pk[1:])[0]
    with pytest.raises(capi.FlashJoinError):
        capi.join(7, 0, dbk, dbv, dpk)
    with pytest.raises(capi.FlashJoinError):
        capi.join(capi.ALGO_SCALAR, 1 << 12, dbk, dbv, dpk)


def test_pairs_state_errors(fj, capi):
    a = np.arange(100, dtype=np.uint64)
    fj.hash_join_count(a, a, a)
    with pytest.raises(RuntimeError):
        fj.last_pairs()  # the last call was not a materialize call
    fj.hash_join(a, a, a)
    k, v = fj.last_pairs()
    assert k.size == 100


def test_device_generator_matches_numpy(capi):
    N, ny, pct = 300_000, 40_000, 90
    keys, vals = capi.generate_g2("build", N, ny, pct, 108, 0, ny)
    bk, bv = g2_slice(N, ny, pct, 108, "build", 0, ny)
    assert np.array_equal(keys.to_host(), bk) and np.array_equal(vals.to_host(), bv)
    pk_d = capi.generate_g2("probe", N, ny, pct, 108, 1000, 50_000)
    assert np.array_equal(pk_d.to_host(), g2_slice(N, ny, pct, 108, "probe", 1000, 51_000))


def test_idempotence_and_arena_reuse(fj):
    """Same call twice gives the same answer (the arena is reused, the table is re-cleared)."""
    bk, bv, pk = g1(500_000, 200_000, 90)
    r = [fj.hash_join_radix(bk, bv, pk)[0] for _ in range(3)] + [fj.hash_join(bk, bv, pk)[0] for _ in range(3)]
    assert len(set(r)) == 1
    small = np.arange(10, dtype=np.uint64)
    assert fj.hash_join_count(small, small, small)[0] == 10  # a smaller join after a larger one


# ------------------------------------------------------------------------------------------------ dense key domain
def test_dense_bitmap_count(fj):
    """Count entry points on a dense key domain: exact membership bitmap, no table."""
    # fused = 1: one persistent launch (zero | grid barrier | build | grid barrier | probe); 0: three kernels
    for fused in (1, 0, 1):
        fj.configure(dense_fused=fused)
        for N, ny, pct in ((500_000, 50_000, 10), (300_000, 1_000, 90), (2_000_000, 700_000, 90), (3_000, 40, 90)):
            bk, bv, pk = g1(N, ny, pct)
            n0 = O.np_join(bk, bv, pk)[0]
            for name in ("hash_join_count", "hash_join_count_bloom", "adaptive_join_count", "adaptive_join_count_bloom"):
                n, _ = getattr(fj, name)(bk, bv, pk)
                st = fj.last_stats()
                assert n == n0 and st["dense"] == 1 and st["bloom_kind"] == "bitmap" and st["attempts"] == 1, (name, st)
                assert st["kernel_launches"] == (1 if fused else 3), st
    # key 0, duplicates in the build side (a set: no keep-first pass needed), probe keys far outside the domain
    bk = np.array([0, 3, 3, 7, 200, 0], dtype=np.uint64)
    pk = np.array([0, 0, 3, 4, 7, 200, 2**40, 2**64 - 1, 255, 256], dtype=np.uint64)
    n, _ = fj.hash_join_count(bk, bk, pk)
    assert n == O.np_join(bk, bk, pk)[0] == 5 and fj.last_stats()["dense"] == 1
    # one build key outside the optimistic domain: the attempt is abandoned and the hash path answers
    bk, bv, pk = g1(500_000, 50_000, 10)
    bk = bk.copy(); bk[777] = 10**9
    n, _ = fj.hash_join_count_bloom(bk, bv, pk)
    st = fj.last_stats()
    assert n == O.np_join(bk, bv, pk)[0] and st["dense"] == 0 and st["attempts"] == 2 and st["bloom_kind"] == "smem", st
    # values play no role in a count: 64-bit values do not disturb the bitmap path
    bk, bv, pk = g1(200_000, 20_000, 90)
    bv = bv.copy(); bv[5] = 2**50
    n, _ = fj.hash_join_count(bk, bv, pk)
    assert n == O.np_join(bk, bv, pk)[0] and fj.last_stats()["dense"] == 1


def test_dense_direct_materialize(fj):
    """Materialize entry points of the global-table path on a dense key domain: exact bitmap in shared memory +
    L2-resident direct-address value table, one persistent launch (k_mat_dense_fused)."""
    for N, ny, pct in ((500_000, 50_000, 10), (300_000, 1_000, 90), (2_000_000, 700_000, 90), (3_000, 40, 90), (5, 3, 100)):
        # 700 000 build rows: the bitmap is clipped to what two CTAs per SM can hold (896 k keys >= 1.1 * 700 000)
        bk, bv, pk = g1(N, ny, pct)
        bv = bv.copy(); bv[ny // 2] = 2**63 + 12345  # 64-bit values are fine: the table holds 8-byte values
        n0, k0, v0 = O.np_join(bk, bv, pk)
        sp0 = O.sorted_pairs(k0, v0)
        for name in ("hash_join", "hash_join_bloom", "adaptive_join", "adaptive_join_bloom"):
            n, _ = getattr(fj, name)(bk, bv, pk)
            st = fj.last_stats()
            assert n == n0 and st["dense"] == 1 and st["bloom_kind"] == "bitmap" and st["attempts"] == 1 and st["kernel_launches"] == 1, (name, st)
            assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), sp0), name
    # probe row indices ride along
    bk, bv, pk = g1(400_000, 30_000, 50)
    n0, k0, v0 = O.np_join(bk, bv, pk)
    n, _ = fj.join_flags("scalar", False, True, bk, bv, pk, probe_idx=True)
    k, v, idx = fj.last_pairs(with_probe_idx=True)
    assert n == n0 and fj.last_stats()["dense"] == 1 and np.array_equal(pk[idx.astype(np.int64)], k)
    assert np.array_equal(O.sorted_pairs(k, v), O.sorted_pairs(k0, v0)) and np.unique(idx).size == n0
    # key 0 and probe keys far outside the domain
    bk = np.array([0, 3, 7, 200], dtype=np.uint64)
    pk = np.array([0, 0, 3, 4, 7, 200, 2**40, 2**64 - 1, 255, 256], dtype=np.uint64)
    n, _ = fj.hash_join(bk, bk + np.uint64(10), pk)
    assert n == 5 and fj.last_stats()["dense"] == 1
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(*O.np_join(bk, bk + np.uint64(10), pk)[1:]))
    # one build key outside the optimistic domain: the hash path answers
    bk, bv, pk = g1(500_000, 50_000, 10)
    bk = bk.copy(); bk[777] = 10**9
    n0, k0, v0 = O.np_join(bk, bv, pk)
    n, _ = fj.hash_join(bk, bv, pk)
    st = fj.last_stats()
    assert n == n0 and st["dense"] == 0 and st["attempts"] == 2, st
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(k0, v0))
    # a domain too large for two bitmaps per SM keeps the hash-table kernel (faster at high match rates there)
    bk, bv, pk = g1(2_000_000, 1_000_000, 90)
    n0, k0, v0 = O.np_join(bk, bv, pk)
    n, _ = fj.hash_join(bk, bv, pk)
    st = fj.last_stats()
    assert n == n0 and st["dense"] == 0 and st["attempts"] == 1, st
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(k0, v0))
    # duplicate build keys: an already-set bit -> exact keep-first path
    rng = np.random.default_rng(3)
    bk = rng.integers(0, 5000, 60_000).astype(np.uint64)
    bv = np.arange(60_000, dtype=np.uint64)
    pk = rng.integers(0, 6000, 200_000).astype(np.uint64)
    n0, k0, v0 = O.join("scalar", False, True, bk, bv, pk)
    n, _ = fj.hash_join(bk, bv, pk)
    st = fj.last_stats()
    assert n == n0 and st["dedup_exact"] is True and st["dense"] == 0, st
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(k0, v0))


@pytest.fixture()
def dense_small(fj):
    """The round-1 dense radix path (k_scatter2 by the low 8 key bits + L2-resident k_djoin); the round-2 path
    (k_part + k_sjoin) is switched off here and has its own tests below."""
    fj.configure(dense_min_rows=1024, dense16=0)
    yield fj
    fj.configure(dense_min_rows=1 << 20, dense_group_mb=8, dense_ring=4, dense_batch=2, dense_delay_b=1, dense_delay_p=3, dense=1, dense16=1)


@pytest.mark.parametrize("group_mb,ring,batch,delay_b,delay_p", [(8, 4, 2, 1, 3), (1, 4, 2, 1, 3), (1, 1, 1, 1, 1), (1, 16, 8, 2, 4),
                                                                (16, 8, 4, 1, 1)])
def test_dense_radix_direct_join(dense_small, group_mb, ring, batch, delay_b, delay_p):
    """Radix entry points on a dense key domain: one scatter pass by the low key bits + direct-address join
    (k_djoin); group_mb = 1 forces many L2 groups (the whole Z/B/P/C item pipeline with its waits); the other
    knobs are the dispatcher look-ahead and the pipeline distances (results must not depend on any of them)."""
    fj = dense_small
    fj.configure(dense_group_mb=group_mb, dense_ring=ring, dense_batch=batch, dense_delay_b=delay_b, dense_delay_p=delay_p)
    for N, ny, pct in ((400_000, 300_000, 90), (3_000_000, 2_000_000, 90), (1_000_000, 70_000, 10)):
        bk, bv, pk = g1(N, ny, pct)
        expect = O.np_join(bk, bv, pk)
        check_all_entry_points(fj, bk, bv, pk, expect=expect, algos=("radix",))
        st = fj.last_stats()
        assert st["dense"] == 1 and st["path"] == "radix" and st["radix_bits"] == (8, 0) and st["attempts"] == 1, st
    fj.configure(dense=0)
    n, _ = fj.hash_join_radix(bk, bv, pk)
    assert n == expect[0] and fj.last_stats()["dense"] == 0
    fj.configure(dense=1)


def test_dense_radix_edges(dense_small):
    fj = dense_small
    rng = np.random.default_rng(11)
    # key 0, probe keys beyond the domain
    bk = rng.permutation(60_000).astype(np.uint64)  # 60 000 rows -> optimistic key bound 2^17
    bv = rng.integers(0, 2**32 - 1, bk.size).astype(np.uint64)
    pk = np.concatenate([rng.integers(0, 140_000, 500_000).astype(np.uint64),
                         np.array([2**17 - 1, 2**17, 2**31, 2**32 - 1, 2**32, 2**63, 2**64 - 1], dtype=np.uint64)])
    check_all_entry_points(fj, bk, bv, pk, algos=("radix",))
    assert fj.last_stats()["dense"] == 1
    # the largest value a packed row can hold does not fit value + 1: general packed path
    bv2 = bv.copy(); bv2[100] = 2**32 - 1
    check_all_entry_points(fj, bk, bv2, pk, algos=("radix",))
    st = fj.last_stats()
    assert st["dense"] == 0 and st["narrow"] is True and st["attempts"] == 2, st
    # a build key outside the optimistic bound 2^(ceil(log2 nb) + 1)
    bk3 = bk.copy(); bk3[5] = 3_000_000
    check_all_entry_points(fj, bk3, bv, pk, algos=("radix",))
    assert fj.last_stats()["dense"] == 0
    # 64-bit key: dense attempt, packed attempt, wide rows
    bk4 = bk.copy(); bk4[5] = 2**40
    n, _ = fj.hash_join_radix(bk4, bv, pk)
    st = fj.last_stats()
    assert n == O.np_join(bk4, bv, pk)[0] and st["narrow"] is False and st["attempts"] == 3, st
    # duplicate build keys: fewer direct-address slots than rows -> exact keep-first path
    bk5 = rng.integers(0, 50_000, 120_000).astype(np.uint64)
    bv5 = np.arange(bk5.size, dtype=np.uint64)
    pk5 = rng.integers(0, 60_000, 300_000).astype(np.uint64)
    n0, k0, v0 = O.join("radix", False, True, bk5, bv5, pk5)
    n, _ = fj.hash_join_radix(bk5, bv5, pk5)
    assert n == n0 and fj.last_stats()["dedup_exact"] is True
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(k0, v0))
    n, _ = fj.hash_join_count_radix(bk5, bv5, pk5)  # a count is a set operation: the dense path answers directly
    assert n == n0 and fj.last_stats()["dense"] == 1
    # skewed low key bits (only 100 of the 256 residues occur, inside the optimistic bound 2^18): the partitions of
    # those residues overflow their fixed-capacity regions -> hash partitioning instead
    allk = np.arange(262_144, dtype=np.uint64)
    bk6 = rng.permutation(allk[(allk & np.uint64(255)) < 100])[:100_000]
    pk6 = rng.choice(allk, 400_000)
    check_all_entry_points(fj, bk6, bk6 + np.uint64(1), pk6, algos=("radix",))
    st = fj.last_stats()
    assert st["dense"] == 0 and st["path"] == "radix" and st["attempts"] == 2, st
    # all probe rows carry one key
    bk7 = np.arange(1, 50_001, dtype=np.uint64)
    check_all_entry_points(fj, bk7, bk7 * np.uint64(7), np.full(600_000, 4242, dtype=np.uint64), algos=("radix",))


# ------------------------------------------------------------------------------------------------ dense key domain, round 2
@pytest.fixture()
def dense16_small(fj):
    fj.configure(dense_min_rows=1024, dense16=1, dense16_logp=0)
    yield fj
    fj.configure(dense_min_rows=1 << 20, dense16=1, dense16_logp=0, dense=1)


@pytest.mark.parametrize("logp", [0, 8, 10, 11])
def test_dense16_partition_join(dense16_small, logp):
    """Radix entry points on a dense key domain, round 2: ONE partition pass by the low key bits with per-SM
    write-combining sector buffers (k_part; rows shrink to idx16 | value16 and idx16) + the direct-address join in
    shared memory (k_sjoin).  The partition count must not change the result."""
    fj = dense16_small
    fj.configure(dense16_logp=logp)
    for N, ny, pct in ((400_000, 300_000, 90), (3_000_000, 2_000_000, 90), (1_000_000, 70_000, 10), (5_001, 2_049, 50)):
        bk, bv, pk = g1(N, ny, pct)
        expect = O.np_join(bk, bv, pk)
        check_all_entry_points(fj, bk, bv, pk, expect=expect, algos=("radix",))
        st = fj.last_stats()
        assert st["dense"] == 2 and st["path"] == "radix" and st["attempts"] == 1 and st["radix_bits"][1] == 0, st
        if logp:
            assert st["radix_bits"] == (logp, 0), st


def test_dense16_edges(dense16_small):
    fj = dense16_small
    rng = np.random.default_rng(12)
    # key 0, the largest value the 16-bit slot holds, probe keys beyond the domain and beyond 32 bits
    bk = rng.permutation(60_000).astype(np.uint64)
    bv = rng.integers(0, 65535, bk.size).astype(np.uint64)
    bv[:3] = (0, 65534, 65533)
    pk = np.concatenate([rng.integers(0, 140_000, 500_000).astype(np.uint64),
                         np.array([2**17 - 1, 2**17, 2**27, 2**31, 2**32 - 1, 2**32, 2**63, 2**64 - 1], dtype=np.uint64)])
    check_all_entry_points(fj, bk, bv, pk, algos=("radix",))
    assert fj.last_stats()["dense"] == 2
    # a value that does not fit 16 bits (value + 1 is stored): the round-1 layout (32-bit values) answers
    bv2 = bv.copy(); bv2[100] = 65535
    check_all_entry_points(fj, bk, bv2, pk, algos=("radix",))
    st = fj.last_stats()
    assert st["dense"] == 1 and st["attempts"] == 2, st
    n, _ = fj.hash_join_count_radix(bk, bv2, pk)  # a count never reads the values
    assert n == O.np_join(bk, bv2, pk)[0] and fj.last_stats()["dense"] == 2
    # a build key outside the optimistic domain
    bk3 = bk.copy(); bk3[5] = 2**30
    check_all_entry_points(fj, bk3, bv, pk, algos=("radix",))
    assert fj.last_stats()["dense"] == 0
    # duplicate build keys: the slot is already taken -> exact keep-first path (materialize); a count is a set operation
    bk5 = rng.integers(0, 50_000, 120_000).astype(np.uint64)
    bv5 = np.arange(bk5.size, dtype=np.uint64) % np.uint64(60_000)
    pk5 = rng.integers(0, 60_000, 300_000).astype(np.uint64)
    n0, k0, v0 = O.join("radix", False, True, bk5, bv5, pk5)
    n, _ = fj.hash_join_radix(bk5, bv5, pk5)
    assert n == n0 and fj.last_stats()["dedup_exact"] is True
    assert np.array_equal(O.sorted_pairs(*fj.last_pairs()), O.sorted_pairs(k0, v0))
    n, _ = fj.hash_join_count_radix(bk5, bv5, pk5)
    assert n == n0 and fj.last_stats()["dense"] == 2
    # skewed low key bits (only 100 of 512 residues occur): those partitions overflow their regions -> other layouts
    allk = np.arange(262_144, dtype=np.uint64)
    bk6 = rng.permutation(allk[(allk & np.uint64(511)) < 100])[:50_000]
    pk6 = rng.choice(allk, 400_000)
    check_all_entry_points(fj, bk6, bk6 % np.uint64(999), pk6, algos=("radix",))
    assert fj.last_stats()["dense"] != 2
    # all probe rows carry one key: one partition's ring fills again and again within a round (retry iterations)
    bk7 = np.arange(1, 50_001, dtype=np.uint64)
    check_all_entry_points(fj, bk7, bk7 % np.uint64(65_000), np.full(600_000, 4242, dtype=np.uint64), algos=("radix",))
    # empty intersection; build side of one partition only
    check_all_entry_points(fj, bk7, bk7 % np.uint64(7), bk7 + np.uint64(100_000), algos=("radix",))
    bk8 = (np.arange(3_000, dtype=np.uint64) << np.uint64(9)) + np.uint64(5)
    check_all_entry_points(fj, bk8, bk8 % np.uint64(11), rng.integers(0, 3_000 << 9, 100_000).astype(np.uint64), algos=("radix",))


def test_dense16_ragged_and_unaligned(capi):
    """C ABI, device-resident inputs whose base is only 8-byte aligned and whose lengths are not multiples of the
    2048-row round: the TMA key ring needs 16-byte alignment, so these rounds take the direct-load path."""
    capi.config_set(dense_min_rows=1024, dense16=1)
    try:
        for N, ny in ((10_001, 4_097), (300_001, 200_003), (2_047, 1_025)):
            bk, bv, pk = g1(N, ny, 90)
            n0, k0, v0 = O.np_join(bk, bv, pk)
            big = [capi.DeviceArray.from_host(np.concatenate([[np.uint64(7)], x])) for x in (bk, bv, pk)]
            views = []
            for d, x in zip(big, (bk, bv, pk)):
                v = capi.DeviceArray.__new__(capi.DeviceArray)
                v.n, v.ptr = x.size, d.ptr + 8
                views.append(v)
            for flags in (0, capi.FLAG_MATERIALIZE):
                n, _, st = capi.join(capi.ALGO_RADIX, flags, *views)
                assert n == n0 and st["dense"] == 2, (N, ny, flags, n, n0, st)
                if flags:
                    assert np.array_equal(O.sorted_pairs(*capi.pairs()), O.sorted_pairs(k0, v0))
            for v in views:
                v.ptr = None  # borrowed
            for d in big:
                d.free()
    finally:
        capi.config_set(dense_min_rows=1 << 20)


# ------------------------------------------------------------------------------------------------ full sizes
@pytest.mark.big
def test_c2_full_size_golden(fj, golden_cases):
    """BASELINE.json configs[1]: flash_join_bloom count, 1e8 x 1e5, ~10 % match."""
    c = next(x for x in golden_cases if (x["N"], x["ny"], x["match_pct"], x["gen"]) == (10**8, 10**5, 10, "g1"))
    bk, bv, pk = g1(c["N"], c["ny"], c["match_pct"])
    for name in ("hash_join_count_bloom", "hash_join_count", "adaptive_join_count", "hash_join_count_radix"):
        n, sec = getattr(fj, name)(bk, bv, pk)
        assert n == c["count"], name
    n, _ = fj.hash_join_bloom(bk, bv, pk)
    cs = O.checksums(*fj.last_pairs())
    assert (n, cs["sum_keys"], cs["xor_keys"], cs["sum_vals"]) == (c["count"], c["sum_keys"], c["xor_keys"], c["sum_vals"])


@pytest.mark.big
def test_c3_full_size_golden(fj, golden_cases):
    """BASELINE.json configs[2]: flash_join_radix materialize, 1e8 x 1e8 (single GPU)."""
    c = next(x for x in golden_cases if (x["N"], x["ny"], x["gen"]) == (10**8, 10**8, "g1"))
    bk, bv, pk = g1(c["N"], c["ny"], c["match_pct"])
    for name in ("hash_join_radix", "adaptive_join"):
        n, sec = getattr(fj, name)(bk, bv, pk)
        assert n == c["count"], name
        k, v = fj.last_pairs()
        cs = O.checksums(k, v)
        assert (cs["sum_keys"], cs["xor_keys"], cs["sum_vals"]) == (c["sum_keys"], c["xor_keys"], c["sum_vals"]), name
        # size-independent properties: every pair key is a build key with that build value; N == ny
        # makes the probe a permutation, so matched keys are distinct
        assert np.unique(k).size == n
    n, _ = fj.hash_join_count_radix(bk, bv, pk)
    assert n == c["count"]
    n, _ = fj.hash_join_count(bk, bv, pk)
    assert n == c["count"]


def test_pageable_inputs_staged(fj):
    """Plain (pageable) numpy columns above `stage_min_mb` travel through the multi-threaded pinned staging ring
    (Engine::h2d): three 8 MB chunks, the last one partial, three host threads; page-locked columns bypass it."""
    bk, bv, pk = g1(3_000_000, 50_000, 50)  # 24 MB probe column
    expect = O.np_join(bk, bv, pk)
    fj.configure(stage_min_mb=1, stage_threads=3)
    try:
        check_all_entry_points(fj, bk, bv, pk, expect=expect, algos=("scalar",))
        hp = fj.pinned_empty(pk.size)
        hp[:] = pk
        n, _ = fj.hash_join_count(bk, bv, hp)
        assert n == expect[0]
        fj.configure(stage_threads=0)  # plain cudaMemcpyAsync
        n, _ = fj.hash_join_count(bk, bv, pk)
        assert n == expect[0]
    finally:
        fj.configure(stage_min_mb=64, stage_threads=8)


@pytest.mark.parametrize("seed", range(12))
def test_randomized_differential(fj, seed):
    """Seeded random shapes and key domains through all twelve entry points against the numpy restatement: sizes that
    straddle tile boundaries, dense and sparse 32- and 64-bit key domains, keys at the edge of the optimistic dense
    bound, duplicate build keys, 32- and 64-bit values, probe sides with few / many matches.  Whatever layout the
    engine picks (dense, packed, wide, exact), the result must be the same."""
    rng = np.random.default_rng(1000 + seed)
    fj.configure(dense_min_rows=int(rng.choice([1024, 1 << 20])))
    try:
        for _ in range(6):
            nb = int(np.exp(rng.uniform(0, np.log(300_000))))
            np_ = int(np.exp(rng.uniform(0, np.log(1_500_000))))
            domain = rng.choice(["dense", "dense_edge", "sparse32", "sparse64"])
            if domain == "dense":
                U = max(nb + 1, int(nb * rng.uniform(1.0, 1.9)) + 1)
            elif domain == "dense_edge":  # the largest key sits exactly at / just past the power-of-two bound above 2 nb
                U = 1024
                while U < 2 * nb:
                    U <<= 1
                U += int(rng.integers(-1, 2))
            elif domain == "sparse32":
                U = 2**32 - 1
            else:
                U = 2**64 - 1
            if U <= 4 * nb + 8:
                pool = rng.permutation(U)[:nb].astype(np.uint64)  # unique keys in [0, U)
                if domain == "dense_edge":
                    pool[0] = U - 1
            else:
                pool = np.unique(rng.integers(0, U, nb, dtype=np.uint64, endpoint=True))
            bk = pool
            if rng.random() < 0.3 and bk.size > 1:  # duplicate build keys: keep-first decides the value
                bk = np.concatenate([bk, rng.choice(bk, max(1, bk.size // 10))])
                rng.shuffle(bk)
            bv = (rng.integers(0, 2**64 - 1, bk.size, dtype=np.uint64) if rng.random() < 0.3
                  else rng.integers(0, 100, bk.size).astype(np.uint64))
            hit = rng.uniform(0, 1)
            n_hit = int(np_ * hit)
            miss_hi = int(min(2**64 - 1, max(2 * int(U), 16)))
            pk = np.concatenate([rng.choice(bk, n_hit), rng.integers(0, miss_hi, np_ - n_hit, dtype=np.uint64, endpoint=True)])
            rng.shuffle(pk)
            check_all_entry_points(fj, bk, bv, pk)
    finally:
        fj.configure(dense_min_rows=1 << 20)


@pytest.mark.big
def test_beyond_two_pass_plan_takes_dense_radix(capi):
    """3e8 x 3e8 rows (G2, generated in HBM): two general scatter passes cannot cut the build side down to shared-memory
    partitions any more (plan_radix gives up above ~2.6e8 rows); on a dense key domain the direct-address radix join
    still applies instead of one huge global table.  No CPU oracle finishes in seconds here: the dense radix join and
    the global-table path are independent algorithms — counts and pair checksums must agree (tools/big_check.py)."""
    N = 300_000_000
    bk, bv = capi.generate_g2("build", N, N, 90, 108, 0, N)
    pk = capi.generate_g2("probe", N, N, 90, 108, 0, N)
    try:
        got = []
        for algo, cfg in ((capi.ALGO_RADIX, {"dense": 1}), (capi.ALGO_SCALAR, {"dense": 0})):
            capi.config_set(**cfg)
            n, _, st = capi.join(algo, capi.FLAG_MATERIALIZE, bk, bv, pk)
            cs = O.checksums(*capi.pairs())
            got.append((n, cs["sum_keys"], cs["xor_keys"], cs["sum_vals"], st["path"], st["dense"]))
        assert got[0][:4] == got[1][:4], got
        assert got[0][4:] == ("radix", 1) and got[1][4:] == ("scalar", 0), got
        n_adaptive, _, st = capi.join(capi.ALGO_ADAPTIVE, 0, bk, bv, pk)
        assert n_adaptive == got[0][0]
    finally:
        capi.config_set(dense=1)
        for x in (bk, bv, pk):
            x.free()


# ------------------------------------------------------------------------------------------------ shuffle (1 GPU)
@pytest.fixture(scope="module")
def comm1(capi):
    """A single-rank communicator: the whole FJ_DIST_SHUFFLE machinery (destination scatter, self send/recv,
    join from received partition-format rows) runs on one GPU; `shuffle_virtual_ranks` > 1 gives the rank
    several destinations so that the multi-destination layout is exercised too."""
    capi.check(capi.lib().fj_init(0))
    capi.comm_init(0, 1)
    yield capi
    capi.config_set(shuffle_virtual_ranks=1)
    capi.comm_destroy()


def _shuffle_check(capi, bk, bv, pk, algos=None, expect=None):
    n0, k0, v0 = expect if expect is not None else O.np_join(bk, bv, pk)
    sp0 = O.sorted_pairs(k0, v0)
    for algo in algos or (capi.ALGO_ADAPTIVE, capi.ALGO_SCALAR, capi.ALGO_RADIX):
        for flags in (0, capi.FLAG_MATERIALIZE):
            g, l, sec, st = capi.join_dist(capi.DIST_SHUFFLE, algo, flags, 0, bk, bv, pk)
            assert g == l == n0, (algo, flags, g, l, n0, st)
            if flags & capi.FLAG_MATERIALIZE:
                assert np.array_equal(O.sorted_pairs(*capi.pairs()), sp0), (algo, st)
    return st


@pytest.mark.parametrize("virtual", [1, 4, 7])
def test_shuffle_single_rank(comm1, virtual):
    capi = comm1
    capi.config_set(shuffle_virtual_ranks=virtual)
    bk, bv, pk = g1(300_000, 60_000, 90)
    st = _shuffle_check(capi, bk, bv, pk)
    assert st["narrow"] == 1
    # empty sides
    e = np.empty(0, dtype=np.uint64)
    for a, b, c in ((e, e, pk), (bk, bv, e), (e, e, e)):
        g, l, _, _ = capi.join_dist(capi.DIST_SHUFFLE, capi.ALGO_ADAPTIVE, capi.FLAG_MATERIALIZE, 0, a, b, c)
        assert g == 0 and capi.pairs()[0].size == 0
    # rows that do not fit the packed format: the narrow attempt is abandoned on every rank
    bv2 = bv.copy(); bv2[17] = 2**45
    st = _shuffle_check(capi, bk, bv2, pk)
    assert st["narrow"] == 0 and st["attempts"] == 2
    # 64-bit keys incl. the out-of-band key 2^64-1 and key 0
    E = 2**64 - 1
    rng = np.random.default_rng(3)
    wk = np.unique(np.concatenate([rng.integers(0, E, 40_000, dtype=np.uint64), np.array([0, E, 2**32], dtype=np.uint64)]))
    wv = rng.integers(0, E, wk.size, dtype=np.uint64)
    wp = np.concatenate([rng.choice(wk, 100_000), rng.integers(0, E, 50_000, dtype=np.uint64), np.array([E, E, 0], dtype=np.uint64)])
    _shuffle_check(capi, wk, wv, wp)


def test_shuffle_duplicate_build_keys_and_skew(comm1):
    capi = comm1
    capi.config_set(shuffle_virtual_ranks=4)
    rng = np.random.default_rng(9)
    bk = rng.integers(0, 5000, 60_000).astype(np.uint64)
    bv = np.arange(60_000, dtype=np.uint64)
    pk = rng.integers(0, 6000, 200_000).astype(np.uint64)
    st = _shuffle_check(capi, bk, bv, pk, expect=O.join("radix", False, True, bk, bv, pk))
    assert st["dedup_exact"] == 1
    # all probe rows carry one key: one destination / one partition receives everything
    bk = np.arange(1, 50_001, dtype=np.uint64)
    _shuffle_check(capi, bk, bk * np.uint64(7), np.full(500_000, 4242, dtype=np.uint64))


def test_shuffle_large_two_pass(comm1):
    """Enough rows that the local join after the exchange takes two radix passes + k_join3."""
    capi = comm1
    capi.config_set(shuffle_virtual_ranks=2)
    bk, bv, pk = g2(4_000_000, 4_000_000, 90)
    g, l, sec, st = capi.join_dist(capi.DIST_SHUFFLE, capi.ALGO_ADAPTIVE, capi.FLAG_MATERIALIZE, 0, bk, bv, pk)
    n0, k0, v0 = O.np_join(bk, bv, pk)
    assert g == n0 and st["path"] == "radix" and st["radix_bits2"] > 0
    assert np.array_equal(O.sorted_pairs(*capi.pairs()), O.sorted_pairs(k0, v0))


def test_shuffle_dest_mirror(capi):
    """flash_hash_join_b200.dist.shuffle_dest (numpy) is the device destination function: every key lands on
    exactly one rank, so per-destination oracle joins add up to the global join."""
    from flash_hash_join_b200.dist import shuffle_dest

    bk, bv, pk = g1(200_000, 50_000, 90)
    world = 4
    db, dp = shuffle_dest(bk, world), shuffle_dest(pk, world)
    total = sum(O.np_join(bk[db == r], bv[db == r], pk[dp == r])[0] for r in range(world))
    assert total == O.np_join(bk, bv, pk)[0]
    assert db.min() >= 0 and db.max() < world and np.bincount(db, minlength=world).min() > 0.2 * bk.size / world
