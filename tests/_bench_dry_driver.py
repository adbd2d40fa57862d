"""Test helper (not a test): runs bench.main() with the C library, the pybind11 module and the CPU-reference worker
replaced by stand-ins, so that the JSON contract of the bench line can be checked on a box without a GPU.
Nothing here is a product path: it only exists for tests/test_bench_contract_cpu.py."""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import flash_hash_join_b200 as pkg  # noqa: E402
from flash_hash_join_b200 import capi  # noqa: E402


class FakeDev:
    def __init__(self):
        self.ptr = 0x1000

    def free(self):
        pass


def _fill(st, world):
    st.kernel_launches = 1
    st.probe_s = 1.3e-4
    st.matches = 12345
    st.path = 1
    st.dense = 1
    st.narrow = 1
    st.n_gpus = world


class FakeLib:
    def __getattr__(self, name):
        def f(*a):
            if name == "fj_join_u64":  # algo, flags, bk, bv, nb, pk, np, p_n, p_sec, p_st
                a[7]._obj.value = 12345
                if a[9] is not None:
                    _fill(a[9]._obj, 1)
            elif name == "fj_join_dist_u64":  # mode, algo, flags, root, bk, bv, nb, pk, np, p_n, p_nl, p_sec, p_st
                a[9]._obj.value = 24690
                if a[10] is not None:
                    a[10]._obj.value = 12345
                if a[12] is not None:
                    _fill(a[12]._obj, 2)
            elif name == "fj_timer_stop":
                a[0]._obj.value = 0.01
            return 0

        return f


capi.lib = lambda: FakeLib()
capi.generate_g2 = lambda side, *a: (FakeDev(), FakeDev()) if side == "build" else FakeDev()
capi.config_set = lambda **kw: None
fj = types.ModuleType("flash_hash_join_b200.flash_join")
fj.pinned_empty = lambda n: np.zeros(min(n, 1000), dtype=np.uint64)
fj.last_pairs = lambda: (np.zeros(12345, dtype=np.uint64), np.zeros(12345, dtype=np.uint64))
for nm in ("hash_join_count_bloom", "hash_join_radix", "adaptive_join_count", "adaptive_join"):
    setattr(fj, nm, lambda bk, bv, pk: (12345, 0.001))
sys.modules["flash_hash_join_b200.flash_join"] = fj
pkg.flash_join = fj
import bench  # noqa: E402

bench.run_cpu_worker = lambda *a, **k: {"rows": 100000000, "core_s_mean": 0.05, "cores": 8, "kind": "reference", "matches": 12345}
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
