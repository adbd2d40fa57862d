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
