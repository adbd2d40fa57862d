"""CPU tests of the ingest step and the benchmark harness: file discovery with the
reference's naming rule, uint64 conversion with `astype(np.uint64)` semantics, the RESULT line format of
/root/reference/benchmark.py:83, and the label -> entry point map of :240-247 — driven through a stub module
with flash_join's 12 entry points (the oracle's numpy join; no GPU, no CUDA extension needed here)."""
import io
import json
import re

import numpy as np
import pytest

from flash_hash_join_b200 import benchmark as B
from flash_hash_join_b200 import ingest
from oracle import oracle as O


class StubModule:
    """flash_join's call surface over the numpy oracle (tests only)."""

    def __init__(self):
        self.calls = []
        self.cache = {}
        for label, (cnt, mat) in B.IMPLEMENTATIONS.items():
            setattr(self, cnt, self._make(cnt))
            setattr(self, mat, self._make(mat))

    def _make(self, name):
        def f(build_keys, build_values, probe_keys):
            self.calls.append(name)
            key = (build_keys.ctypes.data, probe_keys.ctypes.data, build_keys.size, probe_keys.size)
            if key not in self.cache:
                self.cache[key] = O.np_join(build_keys, build_values, probe_keys)[0]
            return self.cache[key], 0.001

        return f


def test_label_map_is_the_reference_one():
    assert B.IMPLEMENTATIONS == {
        "adaptive_join": ("adaptive_join_count", "adaptive_join"),
        "adaptive_bloom": ("adaptive_join_count_bloom", "adaptive_join_bloom"),
        "flash_join": ("hash_join_count", "hash_join"),
        "flash_join_radix": ("hash_join_count_radix", "hash_join_radix"),
        "flash_join_bloom": ("hash_join_count_bloom", "hash_join_bloom"),
        "flash_join_radix_bloom": ("hash_join_count_radix_bloom", "hash_join_radix_bloom"),
    }
    assert sorted(x for pair in B.IMPLEMENTATIONS.values() for x in pair) == sorted(O.entry_point_name(*e) for e in O.ENTRY_POINTS)


def test_result_line_format():
    line = B.result_line("flash_join", "join_count", 8, 0.123456, 42)
    assert re.fullmatch(r"\s+RESULT,Library=flash_join,Task=join_count,Threads=8,Time=0\.1235,Result=42", line)


def test_to_uint64_semantics():
    assert ingest.to_uint64(np.array([-1, 5], dtype=np.int64)).tolist() == [2**64 - 1, 5]
    assert ingest.to_uint64(np.array([3.999, 0.2, 99.5])).tolist() == [3, 0, 99]  # truncation like astype(np.uint64)
    assert ingest.to_uint64([1, 2, 3]).dtype == np.uint64
    with pytest.raises(TypeError):
        ingest.to_uint64(np.array(["id1", "id2"], dtype=object))


@pytest.mark.parametrize("fmt", ["csv", "parquet"])
def test_suite_roundtrip_and_harness(tmp_path, fmt, capsys):
    n = 2_000_000  # "2e6": small = 2 rows, medium = 2 000 rows, big = 2 000 000 rows
    tabs = B.synthetic_suite(n)
    assert tabs["x"]["id3"].size == n and tabs["medium"]["id2"].size == 2000 and tabs["big"]["v2"].size == n
    paths = B.write_suite(tabs, "2e6", str(tmp_path), fmt)
    # file names follow the reference's discovery rule (J1_<N>_<ny>_0_0)
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted(
        f"J1_2e6_{x}_0_0.{fmt}" for x in ("NA", "2e0", "2e3", "2e6"))
    assert set(paths) == {"x", "small", "medium", "big"}
    for lhs, xname in (("reference", "2e6"), ("na", "NA")):
        suites = ingest.discover_suites(str(tmp_path), lhs)
        assert len(suites) == 1 and suites[0]["group_name"] == "2e6" and suites[0]["x"].endswith(f"J1_2e6_{xname}_0_0.{fmt}")
    s = ingest.discover_suites(str(tmp_path), "na")[0]
    got = {c.id: ingest.load_case(s, c) for c in ingest.CASES}
    assert got["Q4"] is None  # id5 absent: skipped like benchmark.py:217-219
    for cid, rt, key in (("Q1", "small", "id1"), ("Q2", "medium", "id2"), ("Q5", "big", "id3")):
        bk, bv, pk = got[cid]
        assert np.array_equal(bk, tabs[rt][key]) and np.array_equal(bv, tabs[rt]["v2"]) and np.array_equal(pk, tabs["x"][key])
    # the harness over the stub: every label x task once per case, RESULT lines parse, counts are the oracle's
    stub = StubModule()
    out = io.StringIO()
    rows = []
    for case, arrays in B.iter_cases_from_tables(tabs, pinned=False):
        if arrays is not None:
            rows += B.run_case(stub, f"2e6-{case.id}", arrays, 4, 0, list(B.IMPLEMENTATIONS), out)
    assert len(rows) == 3 * 6 * 2 and len(stub.calls) == 36
    lines = [l for l in out.getvalue().splitlines() if "RESULT," in l]
    assert len(lines) == 36
    for l, r in zip(lines, rows):
        m = re.search(r"RESULT,Library=(\w+),Task=(\w+),Threads=4,Time=([0-9.]+),Result=(\d+)", l)
        assert m and m.group(1) == r["implementation"] and m.group(2) == r["task"] and int(m.group(4)) == r["result"]
    q5 = [r for r in rows if r["case"] == "2e6-Q5"]
    assert {r["result"] for r in q5} == {O.np_join(tabs["big"]["id3"], tabs["big"]["v2"], tabs["x"]["id3"])[0]}
    json.dumps(rows)
