"""CPU tests of the synthetic data generators (host logic)."""
import numpy as np

from flash_hash_join_b200.datagen import CONFIGS, g1, g2, g2_slice


def test_g1_shape_properties():
    for N, ny, pct in [(20000, 3000, 90), (20000, 3000, 10), (5000, 5000, 90)]:
        bk, bv, pk = g1(N, ny, pct)
        assert bk.dtype == bv.dtype == pk.dtype == np.uint64
        assert bk.size == bv.size == ny and pk.size == N
        assert np.unique(bk).size == ny  # unique RHS keys (join-datagen.R:147)
        c = ny * pct // 100
        assert np.intersect1d(bk, pk).size == c  # every probe-side key appears at least once
        assert bk.min() >= 1 and max(bk.max(), pk.max()) <= 2 * ny - c
        assert bv.max() < 100
        if N == ny:
            assert np.isin(pk, bk).sum() == c


def test_g1_is_deterministic():
    a = g1(10000, 1000, 90)
    b = g1(10000, 1000, 90)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_g2_slices_compose():
    N, ny, pct = 50000, 4000, 90
    bk, bv, pk = g2(N, ny, pct)
    assert np.unique(bk).size == ny
    parts = [g2_slice(N, ny, pct, 108, "probe", s, min(s + 7001, N)) for s in range(0, N, 7001)]
    assert np.array_equal(np.concatenate(parts), pk)
    k2, v2 = g2_slice(N, ny, pct, 108, "build", 1000, 2500)
    assert np.array_equal(k2, bk[1000:2500]) and np.array_equal(v2, bv[1000:2500])
    rate = np.isin(pk, bk).mean()
    assert abs(rate - pct / 100) < 0.02


def test_configs_match_baseline():
    assert CONFIGS["C2"] == (10**8, 10**5, 10)
    assert CONFIGS["C3"] == (10**8, 10**8, 90)


def test_g2_build_rows_come_in_pseudo_random_order():
    """The build row order is a cycle-walked bijective hash of the row index (like the shuffled RHS tables of
    join-datagen.R): a bijection for every size, and no arithmetic progression of keys (which made a radix partition
    pass conflict-free and the round-1 G2 numbers too good)."""
    from flash_hash_join_b200.datagen import _shuffle_index

    for n in (1, 2, 3, 7, 1000, 4096, 4097, 150_000):
        x = _shuffle_index(np.arange(n, dtype=np.uint64), n, 108)
        assert np.array_equal(np.sort(x), np.arange(n, dtype=np.uint64)), n
    bk, bv, pk = g2(50_000, 40_000, 90)
    d = np.diff(bk.astype(np.int64))
    assert np.unique(d).size > 20_000  # an affine order has a handful of distinct steps
    low = (bk[:4096] & np.uint64(31)).reshape(-1, 32)
    assert np.mean([np.unique(r).size for r in low]) < 24  # 32 consecutive rows collide in the low key bits, like random keys (~20.4 distinct)
    # the (key, value) SET is what the goldens pin: value is a function of the key id, not of the row
    k2, v2 = g2_slice(50_000, 40_000, 90, 108, "build", 0, 40_000)
    assert np.array_equal(k2, bk) and np.array_equal(v2, bv)


def test_bench_expected_counts_cover_the_multi_gpu_workloads():
    """bench.py checks `matches` at every GPU count against tests/golden: C3 / C2 / C4 slices from the compiled reference
    (g1_goldens.json, gen == g2) and the C4 / C5 totals (g2_counts.json), where the reference's count on probe slices and the
    generator-implied count agree."""
    import json
    from pathlib import Path

    import bench

    assert bench.expected_matches(100_000_000, 100_000_000, 90)[0] == 89_999_578
    assert bench.expected_matches(100_000_000, 100_000, 10)[0] == 9_999_927
    for g, want in ((1, 112_499_618), (2, 225_002_176), (4, 449_999_272), (8, 900_003_207)):
        n, src = bench.expected_matches(g * 125_000_000, 1_000_000, 90)
        assert n == want and "reference" in src, (g, n, src)
    assert bench.expected_matches(1_000_000_000, 1_000_000_000, 90) == (899_998_249, "generator-implied count (tests/golden/g2_counts.json)")
    assert bench.expected_matches(12345, 678, 90) == (None, None)
    cases = json.loads((Path(bench.__file__).parent / "tests" / "golden" / "g2_counts.json").read_text())["cases"]
    assert all(c["reference_count"] in (None, c["generator_count"]) for c in cases)
