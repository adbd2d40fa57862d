import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "big: full BASELINE.json sizes (1e8 rows); set FJ_SKIP_BIG=1 to skip")


def _gpu_available() -> bool:
    try:
        from flash_hash_join_b200 import capi
        import ctypes as C

        n = C.c_int(0)
        capi.lib().fj_device_count(C.byref(n))
        return n.value > 0
    except Exception:
        return False


HAVE_GPU = _gpu_available()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not silently skip: leave the tests alone.
    skip_big = pytest.mark.skip(reason="FJ_SKIP_BIG=1")
    for item in items:
        if "big" in item.keywords and os.environ.get("FJ_SKIP_BIG") == "1":
            item.add_marker(skip_big)


@pytest.fixture(scope="session")
def golden_cases():
    return json.loads((GOLDEN / "g1_goldens.json").read_text())["cases"]


@pytest.fixture(scope="session")
def fixtures_npz():
    out = {}
    for name in ("g1_5000_600", "edge_keys", "dup_build_radix", "skew_probe"):
        with np.load(GOLDEN / f"{name}.npz") as z:
            out[name] = {k: z[k] for k in z.files}
    return out
