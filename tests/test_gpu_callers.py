"""GPU tests of the callers either side of the hot path and of the multi-GPU entry point:

  * f2 — the benchmark harness (`python -m flash_hash_join_b200.benchmark --synthetic ...`, the workflow of
    /root/reference/benchmark.py:183-300) run on the device, every `RESULT,` count checked against the oracle;
  * f3 — the ingest step (/root/reference/benchmark.py:200-237): h2o-shaped J1_* CSV and Parquet files ->
    ingest.load_case -> pinned uint64 columns -> join -> the oracle's numpy join of the same tables;
  * e  — tests/dist_gpu_check.py (both distributed modes against the oracle, one process per GPU) under pytest,
    skipped when the box has fewer than two GPUs.
"""
import ctypes as C
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _device_count() -> int:
    from flash_hash_join_b200 import capi

    n = C.c_int(0)
    capi.lib().fj_device_count(C.byref(n))
    return n.value


def test_benchmark_harness_on_device(tmp_path):
    """The harness end to end in its own process, as a user runs it; counts of every label x task x case must be the
    oracle's (the six labels only differ in the path taken, never in the result)."""
    from flash_hash_join_b200 import benchmark as B

    n = 2_000_000
    out = subprocess.run([sys.executable, "-m", "flash_hash_join_b200.benchmark", "--synthetic", "2e6", "--no-duckdb", "--json",
                          str(tmp_path / "res.json")], cwd=str(ROOT), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln.strip() for ln in out.stdout.splitlines() if ln.strip().startswith("RESULT,")]
    tabs = B.synthetic_suite(n)
    expect = {}
    for case, arrays in B.iter_cases_from_tables(tabs, pinned=False):
        if arrays is not None:
            expect[case.id] = O.np_join(*arrays)[0]
    assert set(expect) == {"Q1", "Q2", "Q5"}  # Q4 joins on a factor column the synthetic tables do not carry
    rows = json.loads((tmp_path / "res.json").read_text())
    rows = rows["results"] if isinstance(rows, dict) else rows
    assert len(rows) == len(lines) == len(expect) * len(B.IMPLEMENTATIONS) * len(B.TASKS)
    for r in rows:
        assert r["result"] == expect[r["case"].split("-")[-1]], r
        assert r["stats"]["kernel_launches"] > 0  # the CUDA path ran (no CPU fallback exists)
    for ln in lines:
        assert re.fullmatch(r"RESULT,Library=\w+,Task=join_(count|materialize),Threads=\d+,Time=\d+\.\d{4},Result=\d+", ln), ln


@pytest.mark.parametrize("fmt", ["csv", "parquet"])
def test_ingest_files_to_join(tmp_path, fmt):
    """J1_* files written with the reference's naming rule -> discover_suites -> load_case (pinned columns) -> every
    entry point -> the oracle's join of the in-memory tables."""
    from flash_hash_join_b200 import benchmark as B
    from flash_hash_join_b200 import flash_join, ingest

    tabs = B.synthetic_suite(1_000_000)
    B.write_suite(tabs, "1e6", str(tmp_path), fmt)
    suites = ingest.discover_suites(str(tmp_path), lhs="na")
    assert len(suites) == 1 and suites[0]["group_name"] == "1e6"
    seen = 0
    for case in ingest.CASES:
        arrays = ingest.load_case(suites[0], case, pinned=True)
        if arrays is None:
            assert case.id == "Q4"
            continue
        bk, bv, pk = arrays
        assert bk.dtype == bv.dtype == pk.dtype == np.uint64
        right, left = tabs[case.right], tabs["x"]
        n0, k0, v0 = O.np_join(right[case.key], right["v2"], left[case.key])
        ref = O.sorted_pairs(k0, v0)
        for cnt, mat in B.IMPLEMENTATIONS.values():
            assert getattr(flash_join, cnt)(bk, bv, pk)[0] == n0, (case.id, cnt)
            assert getattr(flash_join, mat)(bk, bv, pk)[0] == n0, (case.id, mat)
            k, v = flash_join.last_pairs()[:2]
            assert np.array_equal(O.sorted_pairs(k, v), ref), (case.id, mat)
        seen += 1
    assert seen == 3


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_modes_against_oracle(world):
    """BROADCAST and SHUFFLE (peer-memory shuffle on the dense key domain, count and materialize) at `world` GPUs,
    plus the retry cases, one process per GPU under torchrun; the script exits non-zero on any mismatch."""
    have = _device_count()
    if have < world:
        pytest.skip(f"needs {world} GPUs, this box has {have}")
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ)
    env.pop("LOCAL_RANK", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(ROOT / "tests" / "dist_gpu_check.py"), "--rows", "3000000"],
                         cwd=str(ROOT), capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    rep = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["ok"] and rep["world"] == world
    shuffled = [c for c in rep["cases"] if c["mode"] == "shuffle"]
    assert len(shuffled) == 2 and all(p == "radix/dense2" for c in shuffled for p in c["paths"]), shuffled  # the peer-memory path answered
