"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/*.h
declares, the pybind11 module mirrors the reference's names, and input validation happens before
any device work.  No compute calls are made here (there is no GPU in the build container)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT

REF_NAMES = [  # hash_join.cpp:603-639
    "adaptive_join", "adaptive_join_bloom", "adaptive_join_count", "adaptive_join_count_bloom",
    "hash_join_radix", "hash_join", "hash_join_radix_bloom", "hash_join_bloom",
    "hash_join_count_radix", "hash_join_count", "hash_join_count_radix_bloom", "hash_join_count_bloom",
    "initialize",
]


def declared_symbols():
    syms = []
    for h in sorted((ROOT / "include").glob("*.h")):
        text = h.read_text()
        syms += re.findall(r"FJ_API\s+[\w\s\*]+?\b(fj_\w+)\s*\(", text)
    return sorted(set(syms))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("fj_init", "fj_shutdown", "fj_join_u64", "fj_pairs_fetch", "fj_last_error", "fj_join_dist_u64", "fj_comm_init"):
        assert s in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    from flash_hash_join_b200 import capi

    L = capi.lib()
    for s in declared_symbols():
        assert hasattr(L, s), f"libflashjoin_b200.so does not export {s}"


def test_stats_struct_layout_matches_header():
    from flash_hash_join_b200 import capi

    # 8 doubles + 4 u64 + 16 int32 = 64 + 32 + 64
    assert C.sizeof(capi.Stats) == 160
    text = (ROOT / "include" / "flashjoin_b200.h").read_text()
    body = text[text.index("typedef struct fj_stats {"): text.index("} fj_stats;")]
    fields = re.findall(r"\b(?:double|uint64_t|int32_t)\s+([^;]+);", body)
    names = [n.strip().split("[")[0] for f in fields for n in f.split(",")]
    assert names == [n for n, _ in capi.Stats._fields_]


def test_pybind_module_mirrors_reference_api():
    from flash_hash_join_b200 import flash_join

    for n in REF_NAMES:
        assert callable(getattr(flash_join, n)), n
    for n in ("last_stats", "last_pairs", "configure", "pinned_empty", "join_flags"):
        assert callable(getattr(flash_join, n)), n
    assert "sm_100a" in flash_join.version()


def test_input_validation_before_device_work():
    from flash_hash_join_b200 import flash_join

    a = np.arange(8, dtype=np.uint64)
    with pytest.raises(ValueError):
        flash_join.hash_join_count(a.reshape(2, 4), a, a)  # ndim != 1
    with pytest.raises(ValueError):
        flash_join.hash_join_count(a, a[:4], a)  # length mismatch
    with pytest.raises(TypeError):
        flash_join.hash_join_count(a, a)  # missing argument
    # keyword names of the reference
    with pytest.raises(ValueError):
        flash_join.hash_join_count(build_keys=a, build_values=a[:3], probe_keys=a)


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from flash_hash_join_b200 import capi, flash_join

    a = np.arange(8, dtype=np.uint64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        flash_join.hash_join_count(a, a, a)
    with pytest.raises(capi.FlashJoinError) as e:
        capi.join(capi.ALGO_SCALAR, 0, a, a, a)
    assert e.value.status in (capi.ERR_NO_DEVICE, capi.ERR_CUDA)


def test_config_roundtrip_without_device():
    from flash_hash_join_b200 import capi

    old = capi.config_get("load_pct")
    capi.config_set(load_pct=40)
    assert capi.config_get("load_pct") == 40
    capi.config_set(load_pct=old)
    with pytest.raises(capi.FlashJoinError):
        capi.config_set(no_such_key=1)


def test_product_does_not_import_oracle_or_torch():
    """The product path must not route through oracle/ (or torch): scan the package sources."""
    pkg = ROOT / "flash_hash_join_b200"
    for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, p
        assert "import torch" not in text, p
        assert "join_oracle" not in text, p
