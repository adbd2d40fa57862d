"""ctypes binding of the C ABI (include/flashjoin_b200.h) — the same calls a cgo/JNI/FFI host makes.

Used by bench.py (device-resident inputs, per-phase statistics) and by the parity tests that go
through the C ABI directly instead of the pybind11 module.  No torch, no numpy requirement beyond
array pointers.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "libflashjoin_b200.so"

ALGO_ADAPTIVE, ALGO_SCALAR, ALGO_RADIX = 0, 1, 2
FLAG_BLOOM, FLAG_MATERIALIZE, FLAG_DEVICE_INPUTS, FLAG_FORCE_WIDE, FLAG_PROBE_IDX = 1, 2, 4, 8, 16
DIST_BROADCAST, DIST_SHUFFLE = 0, 1
OK, ERR_BAD_ARG, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_NO_DEVICE, ERR_STATE = 0, -1, -2, -3, -4, -5, -6


class Stats(C.Structure):
    _fields_ = [
        ("h2d_s", C.c_double), ("clear_s", C.c_double), ("build_s", C.c_double), ("partition_s", C.c_double),
        ("probe_s", C.c_double), ("comm_s", C.c_double), ("device_s", C.c_double), ("wall_s", C.c_double),
        ("matches", C.c_uint64), ("table_bytes", C.c_uint64), ("algorithmic_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64),
        ("path", C.c_int32), ("narrow", C.c_int32), ("bloom_kind", C.c_int32), ("attempts", C.c_int32),
        ("dedup_exact", C.c_int32), ("kernel_launches", C.c_int32), ("radix_bits1", C.c_int32), ("radix_bits2", C.c_int32),
        ("n_gpus", C.c_int32), ("dense", C.c_int32), ("part_build_us", C.c_int32), ("part_probe_us", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["path"] = {1: "scalar", 2: "radix"}.get(d["path"], str(d["path"]))
        d["bloom_kind"] = {0: "none", 1: "smem", 2: "global", 3: "bitmap", 4: "partition"}[d["bloom_kind"]]
        return d


class FlashJoinError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[fj_status {status}] {msg}")
        self.status = status


_lib = None


def lib():
    """Load libflashjoin_b200.so (raises OSError when it is not built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise OSError(f"{_LIB_PATH} is not built — run `python setup.py build_ext --inplace`")
        L = C.CDLL(str(_LIB_PATH))
        u64p, vp = C.POINTER(C.c_uint64), C.c_void_p
        sig = {
            "fj_init": [C.c_int], "fj_shutdown": [], "fj_device_count": [C.POINTER(C.c_int)],
            "fj_join_u64": [C.c_int, C.c_uint, vp, vp, C.c_size_t, vp, C.c_size_t, u64p, C.POINTER(C.c_double), C.POINTER(Stats)],
            "fj_pairs_count": [u64p], "fj_pairs_fetch": [vp, vp, vp, C.c_size_t],
            "fj_pairs_device": [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), u64p],
            "fj_config_set": [C.c_char_p, C.c_int64], "fj_config_get": [C.c_char_p, C.POINTER(C.c_int64)],
            "fj_dev_alloc": [C.POINTER(vp), C.c_size_t], "fj_dev_free": [vp],
            "fj_memcpy_h2d": [vp, vp, C.c_size_t], "fj_memcpy_d2h": [vp, vp, C.c_size_t],
            "fj_host_alloc_pinned": [C.POINTER(vp), C.c_size_t], "fj_host_free_pinned": [vp],
            "fj_device_synchronize": [],
            "fj_generate_g2": [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp],
            "fj_flush_l2": [], "fj_timer_start": [], "fj_timer_stop": [C.POINTER(C.c_double)],
            "fj_comm_unique_id": [vp], "fj_comm_init": [C.c_int, C.c_int, vp], "fj_comm_destroy": [],
            "fj_join_dist_u64": [C.c_int, C.c_int, C.c_uint, C.c_int, vp, vp, C.c_size_t, vp, C.c_size_t, u64p, u64p,
                                 C.POINTER(C.c_double), C.POINTER(Stats)],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        L.fj_last_error.restype = C.c_char_p
        L.fj_last_error.argtypes = []
        L.fj_version.restype = C.c_char_p
        L.fj_version.argtypes = []
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise FlashJoinError(status, lib().fj_last_error().decode())


def _u64(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype == np.int64:
        a = a.view(np.uint64)
    return np.ascontiguousarray(a, dtype=np.uint64)


class DeviceArray:
    """A uint64 array in HBM owned through fj_dev_alloc / fj_dev_free."""

    def __init__(self, n: int):
        self.n = int(n)
        p = C.c_void_p()
        check(lib().fj_dev_alloc(C.byref(p), max(self.n, 1) * 8))
        self.ptr = p.value

    @classmethod
    def from_host(cls, a) -> "DeviceArray":
        a = _u64(a)
        d = cls(a.size)
        if a.size:
            check(lib().fj_memcpy_h2d(d.ptr, a.ctypes.data, a.size * 8))
        return d

    def to_host(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint64)
        if self.n:
            check(lib().fj_memcpy_d2h(out.ctypes.data, self.ptr, self.n * 8))
        return out

    def free(self):
        if self.ptr:
            lib().fj_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _ptr(x):
    if isinstance(x, DeviceArray):
        return x.ptr, x.n, True
    a = _u64(x)
    return a.ctypes.data, a.size, False, a


def join(algo: int, flags: int, build_keys, build_values, probe_keys):
    """fj_join_u64 on numpy (host) or DeviceArray (HBM-resident) inputs.  Returns (matches, seconds, stats dict)."""
    keep = []
    ptrs, sizes, dev = [], [], []
    for x in (build_keys, build_values, probe_keys):
        r = _ptr(x)
        ptrs.append(r[0]); sizes.append(r[1]); dev.append(r[2])
        if len(r) > 3:
            keep.append(r[3])
    if len(set(dev)) != 1:
        raise ValueError("inputs must be all host arrays or all DeviceArray")
    if sizes[0] != sizes[1]:
        raise ValueError("build_values and build_keys differ in length")
    if dev[0]:
        flags |= FLAG_DEVICE_INPUTS
    n = C.c_uint64(0)
    sec = C.c_double(0.0)
    st = Stats()
    check(lib().fj_join_u64(algo, flags, ptrs[0], ptrs[1], sizes[0], ptrs[2], sizes[2], C.byref(n), C.byref(sec), C.byref(st)))
    return int(n.value), float(sec.value), st.as_dict()


def pairs(with_probe_idx: bool = False):
    n = C.c_uint64(0)
    check(lib().fj_pairs_count(C.byref(n)))
    k = np.empty(n.value, dtype=np.uint64)
    v = np.empty(n.value, dtype=np.uint64)
    ix = np.empty(n.value if with_probe_idx else 0, dtype=np.uint64)
    check(lib().fj_pairs_fetch(k.ctypes.data, v.ctypes.data, ix.ctypes.data if with_probe_idx else None, n.value))
    return (k, v, ix) if with_probe_idx else (k, v)


def config_set(**kw) -> None:
    for k, v in kw.items():
        check(lib().fj_config_set(k.encode(), int(v)))


def config_get(key: str) -> int:
    v = C.c_int64(0)
    check(lib().fj_config_get(key.encode(), C.byref(v)))
    return int(v.value)


def generate_g2(side: str, N: int, ny: int, match_pct: int, seed: int, start: int, count: int):
    """Generate a slice of data set G2 directly in HBM.  Returns DeviceArray keys (and values for 'build')."""
    keys = DeviceArray(count)
    vals = DeviceArray(count) if side == "build" else None
    check(lib().fj_generate_g2(0 if side == "build" else 1, N, ny, match_pct, seed, start, count, keys.ptr, vals.ptr if vals else None))
    return (keys, vals) if vals is not None else keys


# ---- multi-GPU (one process per GPU) -----------------------------------------------------------------
def comm_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    check(lib().fj_comm_unique_id(buf))
    return bytes(buf)


def comm_init(rank: int, world: int, ident: bytes | None = None) -> None:
    """fj_comm_init.  `ident` is the 128-byte NCCL id produced by rank 0 (comm_unique_id) and carried to the
    other ranks by the caller (torch.distributed, MPI, a file ...); a single-rank communicator makes its own."""
    if ident is None:
        if world != 1:
            raise ValueError("ident is required when world > 1")
        ident = comm_unique_id()
    check(lib().fj_comm_init(rank, world, ident))


def comm_destroy() -> None:
    check(lib().fj_comm_destroy())


def join_dist(mode: int, algo: int, flags: int, root: int, build_keys, build_values, probe_keys):
    """fj_join_dist_u64 on this rank's slices (numpy or DeviceArray).  Returns (global matches, local matches,
    seconds, stats dict)."""
    keep, ptrs, sizes, dev = [], [], [], []
    for x in (build_keys, build_values, probe_keys):
        r = _ptr(x)
        ptrs.append(r[0]); sizes.append(r[1]); dev.append(r[2])
        if len(r) > 3:
            keep.append(r[3])
    if len(set(dev)) != 1:
        raise ValueError("inputs must be all host arrays or all DeviceArray")
    if sizes[0] != sizes[1]:
        raise ValueError("build_values and build_keys differ in length")
    if dev[0]:
        flags |= FLAG_DEVICE_INPUTS
    g, l = C.c_uint64(0), C.c_uint64(0)
    sec = C.c_double(0.0)
    st = Stats()
    check(lib().fj_join_dist_u64(mode, algo, flags, root, ptrs[0], ptrs[1], sizes[0], ptrs[2], sizes[2], C.byref(g), C.byref(l),
                                 C.byref(sec), C.byref(st)))
    return int(g.value), int(l.value), float(sec.value), st.as_dict()
