"""flash_hash_join_b200 — B200-native (sm_100a) equi-join engine behind flash_join's Python API.

    from flash_hash_join_b200 import flash_join
    n, seconds = flash_join.hash_join_count(build_keys, build_values, probe_keys)

The compiled pieces live next to this file (built in-tree by ``python setup.py build_ext
--inplace`` or ``__graft_entry__.build()``):
    libflashjoin_b200.so          the C ABI (include/flashjoin_b200.h) over the CUDA kernels
    flash_join.<abi>.so           the pybind11 module mirroring the reference's hash_join.cpp:598-640
There is no CPU fallback: importing ``flash_join`` without the built extension raises ImportError,
and every join call without a usable B200 raises RuntimeError.
"""
from __future__ import annotations

__version__ = "0.1.0"


def __getattr__(name):
    if name == "flash_join":
        import importlib

        try:
            mod = importlib.import_module(".flash_join", __name__)
        except ImportError as e:  # loud: no fallback path exists
            raise ImportError(
                "flash_hash_join_b200.flash_join is not built — run `python setup.py build_ext --inplace` "
                "(needs nvcc; sm_100a) — there is no CPU fallback"
            ) from e
        globals()["flash_join"] = mod
        return mod
    raise AttributeError(name)
