"""ctypes front-end of oracle/join_oracle.c, numpy restatement, and oracle/_ref loader.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sysconfig
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_LIB_PATH = _BUILD / "libjoin_oracle.so"
_lib = None

ALGO = {"adaptive": 0, "scalar": 1, "radix": 2}

# (algo, bloom, materialize) -> name in the reference's pybind11 module (hash_join.cpp:603-637)
ENTRY_POINTS = {
    ("adaptive", False, True): "adaptive_join",
    ("adaptive", True, True): "adaptive_join_bloom",
    ("adaptive", False, False): "adaptive_join_count",
    ("adaptive", True, False): "adaptive_join_count_bloom",
    ("scalar", False, True): "hash_join",
    ("scalar", True, True): "hash_join_bloom",
    ("scalar", False, False): "hash_join_count",
    ("scalar", True, False): "hash_join_count_bloom",
    ("radix", False, True): "hash_join_radix",
    ("radix", True, True): "hash_join_radix_bloom",
    ("radix", False, False): "hash_join_count_radix",
    ("radix", True, False): "hash_join_count_radix_bloom",
}


def entry_point_name(algo: str, bloom: bool, materialize: bool) -> str:
    return ENTRY_POINTS[(algo, bool(bloom), bool(materialize))]


def build(force: bool = False) -> Path:
    """Compile join_oracle.c -> oracle/_build/libjoin_oracle.so (gcc, a second or two)."""
    src = _HERE / "join_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        _BUILD.mkdir(exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-fvisibility=hidden", str(src), "-o", str(_LIB_PATH)]
        )
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        u64p = ctypes.POINTER(ctypes.c_uint64)
        lib.fjo_hash64.restype = ctypes.c_uint64
        lib.fjo_hash64.argtypes = [ctypes.c_uint64, ctypes.c_uint32]
        lib.fjo_bloom_tag.restype = ctypes.c_uint16
        lib.fjo_bloom_tag.argtypes = [ctypes.c_uint64]
        lib.fjo_table_capacity.restype = ctypes.c_size_t
        lib.fjo_table_capacity.argtypes = [ctypes.c_size_t]
        lib.fjo_partition_idx.restype = ctypes.c_uint32
        lib.fjo_partition_idx.argtypes = [ctypes.c_uint64]
        lib.fjo_adaptive_path.restype = ctypes.c_int
        lib.fjo_adaptive_path.argtypes = [ctypes.c_size_t]
        lib.fjo_join.restype = ctypes.c_int64
        lib.fjo_join.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p, u64p]
        lib.fjo_radix_partition.restype = ctypes.c_int
        lib.fjo_radix_partition.argtypes = [u64p, u64p, ctypes.c_size_t, u64p, u64p, u64p]
        _lib = lib
    return _lib


def _u64(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype == np.int64:
        a = a.view(np.uint64)
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def hash64(key: int, seed: int = 0xAAAAAAAA) -> int:
    return int(_load().fjo_hash64(ctypes.c_uint64(key & (2**64 - 1)), ctypes.c_uint32(seed)))


def bloom_tag(h: int) -> int:
    return int(_load().fjo_bloom_tag(ctypes.c_uint64(h)))


def table_capacity(nb: int) -> int:
    return int(_load().fjo_table_capacity(nb))


def partition_idx(key: int) -> int:
    return int(_load().fjo_partition_idx(ctypes.c_uint64(key & (2**64 - 1))))


def join(algo: str, bloom: bool, materialize: bool, build_keys, build_values, probe_keys):
    """Run the C restatement.  Returns (count, keys|None, values|None); pairs are in the
    reference's own output order (probe order on the scalar path, partition-major on radix)."""
    lib = _load()
    bk, bv, pk = _u64(build_keys), _u64(build_values), _u64(probe_keys)
    if bv.size < bk.size:
        raise ValueError("build_values shorter than build_keys")
    if materialize:
        ok = np.empty(max(pk.size, 1), dtype=np.uint64)
        ov = np.empty(max(pk.size, 1), dtype=np.uint64)
        n = lib.fjo_join(ALGO[algo], int(bool(bloom)), _ptr(bk), _ptr(bv), bk.size, _ptr(pk), pk.size, _ptr(ok), _ptr(ov))
        if n < 0:
            raise MemoryError("oracle join failed")
        return int(n), ok[:n].copy(), ov[:n].copy()
    n = lib.fjo_join(ALGO[algo], int(bool(bloom)), _ptr(bk), _ptr(bv), bk.size, _ptr(pk), pk.size, None, None)
    if n < 0:
        raise MemoryError("oracle join failed")
    return int(n), None, None


def radix_partition(keys, values=None):
    lib = _load()
    k = _u64(keys)
    v = _u64(values) if values is not None else None
    ok = np.empty(max(k.size, 1), dtype=np.uint64)
    ov = np.empty(max(k.size, 1), dtype=np.uint64)
    off = np.empty(257, dtype=np.uint64)
    rc = lib.fjo_radix_partition(_ptr(k), _ptr(v) if v is not None else None, k.size, _ptr(ok), _ptr(ov), _ptr(off))
    if rc:
        raise MemoryError
    return ok[: k.size], (ov[: k.size] if v is not None else None), off


def np_join(build_keys, build_values, probe_keys):
    """Independent numpy restatement of the join semantics: de-duplicate the build
    side keeping the first occurrence (hash_join.cpp:125), each probe row matches at most one
    build row (:176), pairs are (probe key, build value) in probe order (:435-436).
    Returns (count, keys, values)."""
    bk, bv, pk = _u64(build_keys), _u64(build_values), _u64(probe_keys)
    if bk.size == 0 or pk.size == 0:
        e = np.empty(0, dtype=np.uint64)
        return 0, e, e.copy()
    uk, first = np.unique(bk, return_index=True)  # sorted unique keys + index of first occurrence
    uv = bv[first]
    pos = np.searchsorted(uk, pk)
    pos_c = np.minimum(pos, uk.size - 1)
    hit = uk[pos_c] == pk
    return int(hit.sum()), pk[hit], uv[pos_c[hit]]


def sorted_pairs(keys, values) -> np.ndarray:
    """Canonical form of a multiset of (key, value) pairs: an (n, 2) uint64 array sorted by
    (key, value).  Parity on materialized output is equality of this array."""
    k, v = _u64(keys), _u64(values)
    order = np.lexsort((v, k))
    return np.stack([k[order], v[order]], axis=1)


def checksums(keys, values) -> dict:
    """Size-independent digest of a pair multiset (order independent): count, sum and xor of keys,
    sum of values, all mod 2^64 — the quantities tabulated in SURVEY.md §8(c)."""
    k, v = _u64(keys), _u64(values)
    with np.errstate(over="ignore"):
        return {
            "count": int(k.size),
            "sum_keys": int(np.add.reduce(k, dtype=np.uint64)) if k.size else 0,
            "xor_keys": int(np.bitwise_xor.reduce(k)) if k.size else 0,
            "sum_vals": int(np.add.reduce(v, dtype=np.uint64)) if v.size else 0,
        }


# ------------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref, built by oracle/build_ref.sh from /root/reference)
# ------------------------------------------------------------------------------------------------
_ref_cache: dict = {}
_REF_MODNAME = {"plain": "flash_join", "pairs": "flash_join_pairs"}


def _ref_path(kind: str) -> Path:
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return _HERE / "_ref" / kind / f"{_REF_MODNAME[kind]}{ext}"


def reference_available(kind: str = "plain") -> bool:
    return _ref_path(kind).exists()


def load_reference(kind: str = "plain"):
    """Import the compiled reference module (kind 'plain' or 'pairs') without putting it on sys.path / sys.modules,
    so it can never shadow the engine's own ``flash_join`` module."""
    if kind not in _ref_cache:
        p = _ref_path(kind)
        if not p.exists():
            raise FileNotFoundError(f"{p} missing — run oracle/build_ref.sh where /root/reference exists")
        spec = importlib.util.spec_from_file_location(_REF_MODNAME[kind], str(p))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_cache[kind] = mod
    return _ref_cache[kind]


def mimalloc_preload_path():
    """The reference's vendored allocator as an LD_PRELOAD library (oracle/build_ref.sh), or None.  Only for timing
    the reference in a numpy-only subprocess (bench.py's CPU worker)."""
    p = _HERE / "_ref" / "mimalloc" / "libmimalloc_override.so"
    return p if p.exists() else None


def build_reference() -> None:
    """Run oracle/build_ref.sh (a no-op on machines without /root/reference)."""
    subprocess.check_call(["bash", str(_HERE / "build_ref.sh")], env=dict(os.environ))
