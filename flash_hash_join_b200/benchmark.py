#!/usr/bin/env python
"""Join benchmark harness with the workflow of the reference's benchmark.py.

    python -m flash_hash_join_b200.benchmark --data-dir ./data            # J1_*.csv / .parquet suites
    python -m flash_hash_join_b200.benchmark --synthetic 1e7              # h2o-shaped tables made in memory
    python -m flash_hash_join_b200.benchmark --synthetic 1e7 --write-data ./data   # ... and written as J1_*.csv

Same six implementation labels x {join_count, join_materialize} as /root/reference/benchmark.py:240-247, the
same cases Q1 / Q2 / Q4 / Q5 (:202-207), and the same machine-readable line per run (:83):

    RESULT,Library=<label>,Task=<task>,Threads=<n>,Time=<total seconds>,Result=<count>

so anything that parses the reference's output keeps working.  What differs: tables are read through
flash_hash_join_b200.ingest (pyarrow, projected columns, pinned host buffers), a warm-up call precedes the
timed one (CUDA context / arena growth is not join time), per-call fj_stats are kept (H2D seconds, device
seconds, path taken), results go to a JSON file, and the DuckDB / matplotlib columns are optional (neither
is installed in the build image; they are used when importable).  `--module` runs the very same harness
over any other module with flash_join's 12 entry points (e.g. a build of the reference) for side-by-side
numbers.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import json
import os
import sys
import time
from typing import Callable, Dict, List, Optional

import numpy as np

from . import ingest

# label -> (count entry point, materialize entry point)            benchmark.py:240-247
IMPLEMENTATIONS = {
    "adaptive_join": ("adaptive_join_count", "adaptive_join"),
    "adaptive_bloom": ("adaptive_join_count_bloom", "adaptive_join_bloom"),
    "flash_join": ("hash_join_count", "hash_join"),
    "flash_join_radix": ("hash_join_count_radix", "hash_join_radix"),
    "flash_join_bloom": ("hash_join_count_bloom", "hash_join_bloom"),
    "flash_join_radix_bloom": ("hash_join_count_radix_bloom", "hash_join_radix_bloom"),
}
TASKS = ("join_count", "join_materialize")


def result_line(label: str, task: str, threads: int, total_s: float, result) -> str:
    return f"    RESULT,Library={label},Task={task},Threads={threads},Time={total_s:.4f},Result={result}"


def run_benchmark(label: str, task: str, threads: int, func: Callable[[], object], out=sys.stdout) -> dict:
    """Time one call; `func` returns a count or the (count, core_seconds) tuple of a flash_join entry point."""
    t0 = time.perf_counter()
    raw = func()
    total = time.perf_counter() - t0
    core = None
    if isinstance(raw, tuple) and len(raw) == 2:
        result, core = raw
    else:
        result = raw
    print(f"  {label} ({task}): total {total:.4f} s" + (f", core {core:.6f} s" if core is not None else "") +
          f", count {int(result):,}", file=out)
    print(result_line(label, task, threads, total, result), file=out)
    d = {"total_time": total, "result": int(result)}
    if core is not None:
        d["core_time"] = float(core)
    return d


def load_module(spec: Optional[str]):
    """None -> this package's flash_join; 'name' -> import name; 'path/to/file.so' -> load that extension."""
    if spec is None:
        from . import flash_join

        return flash_join
    if os.path.sep in spec or spec.endswith(".so"):
        name = os.path.basename(spec).split(".")[0]
        s = importlib.util.spec_from_file_location(name, spec)
        mod = importlib.util.module_from_spec(s)
        s.loader.exec_module(mod)
        return mod
    return importlib.import_module(spec)


# ---- synthetic h2o-shaped suite (no Rscript in the image; db-benchmark/_data/join-datagen.R restated) ----
def synthetic_suite(n: int, seed: int = 108) -> Dict[str, Dict[str, np.ndarray]]:
    """Tables x / small / medium / big with the numeric columns the cases use: x has id1, id2, id3 (N rows; N/1e6,
    N/1e3 and N distinct-key domains, 90 % of which exist on the right-hand side), the right-hand tables have
    their unique key plus v2 (join-datagen.R:95-105, :134-184; generator G1 of SURVEY.md appendix B per key)."""
    from .datagen import g1

    sizes = {"small": max(1, n // 10**6), "medium": max(1, n // 10**3), "big": n}
    keys = {"small": "id1", "medium": "id2", "big": "id3"}
    tabs: Dict[str, Dict[str, np.ndarray]] = {"x": {}}
    for i, (name, ny) in enumerate(sizes.items()):
        bk, bv, pk = g1(n, ny, 90, seed + i)
        tabs[name] = {keys[name]: bk, "v2": bv}
        tabs["x"][keys[name]] = pk
    # the h2o big table also carries id1 / id2 (join-datagen.R:172-180); the reference probes IT as 'x'
    # (benchmark.py:167 picks J1_N_N), so give it the same key columns
    rng = np.random.default_rng(seed + 17)
    for c in ("id1", "id2"):
        tabs["big"][c] = rng.permutation(tabs["x"][c])
    tabs["big"] = {c: tabs["big"][c] for c in ("id1", "id2", "id3", "v2")}
    return tabs


def write_suite(tabs: Dict[str, Dict[str, np.ndarray]], n_key: str, data_dir: str, fmt: str = "csv") -> Dict[str, str]:
    """Write the tables under the reference's file names (J1_<N>_<ny>_0_0.<fmt>)."""
    import pyarrow as pa
    import pyarrow.csv as pcsv
    import pyarrow.parquet as pq

    d, e = n_key[0], int(n_key.split("e")[1])
    names = {"x": "NA", "small": f"{d}e{e - 6}", "medium": f"{d}e{e - 3}", "big": f"{d}e{e}"}
    os.makedirs(data_dir, exist_ok=True)
    out = {}
    for t, cols in tabs.items():
        path = os.path.join(data_dir, f"J1_{n_key}_{names[t]}_0_0.{fmt}")
        tab = pa.table({c: pa.array(v.view(np.int64)) for c, v in cols.items()})
        if fmt == "parquet":
            pq.write_table(tab, path)
        else:
            pcsv.write_csv(tab, path)
        out[t] = path
    return out


def iter_cases_from_tables(tabs, pinned: bool):
    for case in ingest.CASES:
        right, left = tabs[case.right], tabs["x"]
        if case.key not in right or case.key not in left or "v2" not in right:
            yield case, None
            continue
        yield case, tuple(ingest.to_uint64(a, pinned) for a in (right[case.key], right["v2"], left[case.key]))


def run_case(mod, case_id: str, arrays, threads: int, warmup: int, labels: List[str], out=sys.stdout) -> List[dict]:
    bk, bv, pk = arrays
    rows = []
    stats = getattr(mod, "last_stats", None)
    for label in labels:
        for task, fname in zip(TASKS, IMPLEMENTATIONS[label]):
            fn = getattr(mod, fname)
            for _ in range(warmup):
                fn(bk, bv, pk)
            d = run_benchmark(label, task, threads, lambda: fn(bk, bv, pk), out)
            d.update({"case": case_id, "implementation": label, "task": task, "rows_build": int(bk.size), "rows_probe": int(pk.size)})
            if stats is not None:
                st = stats()
                d["stats"] = {k: (v if not isinstance(v, tuple) else list(v)) for k, v in st.items()}
                d["probe_rows_per_s_device"] = pk.size / st["device_s"] if st.get("device_s") else None
            rows.append(d)
    return rows


def run_duckdb(case_id: str, arrays, threads: int, out=sys.stdout) -> List[dict]:
    """The reference's comparison column (benchmark.py:263-288); only when duckdb is importable."""
    try:
        import duckdb
        import pandas as pd
    except ImportError:
        return []
    bk, bv, pk = arrays
    build_df = pd.DataFrame({"key": bk, "value": bv})  # noqa: F841  (referenced by name from SQL)
    probe_df = pd.DataFrame({"key": pk})  # noqa: F841
    con = duckdb.connect(database=":memory:")
    con.execute(f"PRAGMA THREADS={threads}")
    con.execute("CREATE TABLE build_native AS SELECT * FROM build_df;")
    con.execute("CREATE TABLE probe_native AS SELECT * FROM probe_df;")
    rows = []
    q = "SELECT count(*) FROM build_native b JOIN probe_native p ON b.key = p.key;"
    d = run_benchmark("duckdb", "join_count", threads, lambda: con.execute(q).fetchone()[0], out)
    rows.append(dict(d, case=case_id, implementation="duckdb", task="join_count"))

    def mat():
        con.execute("CREATE OR REPLACE TEMPORARY TABLE temp AS SELECT p.key, b.value FROM build_native b JOIN probe_native p ON b.key = p.key;")
        return con.execute("SELECT count(*) FROM temp").fetchone()[0]

    d = run_benchmark("duckdb", "join_materialize", threads, mat, out)
    rows.append(dict(d, case=case_id, implementation="duckdb", task="join_materialize"))
    con.close()
    return rows


def plot_results(rows: List[dict], task: str, path: str) -> bool:
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return False
    sel = [r for r in rows if r["task"] == task]
    if not sel:
        return False
    cases = sorted({r["case"] for r in sel})
    impls = sorted({r["implementation"] for r in sel})
    fig, ax = plt.subplots(figsize=(16, 9))
    w = 0.8 / max(1, len(impls))
    for i, impl in enumerate(impls):
        ys = [next((r.get("core_time", r["total_time"]) for r in sel if r["case"] == c and r["implementation"] == impl), 0.0) for c in cases]
        ax.bar([k + i * w for k in range(len(cases))], ys, w, label=impl)
    ax.set_xticks([k + 0.4 for k in range(len(cases))])
    ax.set_xticklabels(cases, rotation=45)
    ax.set_ylabel("Time (seconds)")
    ax.set_title(f"Benchmark Performance: {task.replace('_', ' ').title()}")
    ax.legend(title="Implementation")
    fig.tight_layout()
    fig.savefig(path)
    plt.close(fig)
    return True


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="flash_join benchmark suite (h2o db-benchmark join shapes) on the B200 engine")
    ap.add_argument("--data-dir", default="./data", help="directory with J1_*.csv / .parquet tables")
    ap.add_argument("--synthetic", default=None, help="make the tables in memory instead, e.g. 1e7")
    ap.add_argument("--write-data", default=None, help="with --synthetic: also write the tables there as J1_* files")
    ap.add_argument("--format", default="csv", choices=["csv", "parquet"])
    ap.add_argument("--threads", type=int, default=os.cpu_count(), help="reported in RESULT lines; DuckDB uses it (flash_join ignores it, as in the reference)")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--labels", default=",".join(IMPLEMENTATIONS), help="comma-separated subset of: " + ", ".join(IMPLEMENTATIONS))
    ap.add_argument("--cases", default="Q1,Q2,Q4,Q5")
    ap.add_argument("--module", default=None, help="module (import name or path to a .so) exposing flash_join's entry points; default: this engine")
    ap.add_argument("--lhs", default="reference", choices=["reference", "na"],
                    help="probe table of a suite: 'reference' = J1_N_N like benchmark.py:167 (the big right-hand table), 'na' = the real h2o left-hand table J1_N_NA")
    ap.add_argument("--no-pinned", action="store_true", help="keep the columns in pageable numpy arrays")
    ap.add_argument("--no-duckdb", action="store_true")
    ap.add_argument("--json", default="benchmark_results.json")
    a = ap.parse_args(argv)

    mod = load_module(a.module)
    if hasattr(mod, "initialize"):
        mod.initialize()
    pinned = not a.no_pinned and a.module is None
    labels = [x for x in a.labels.split(",") if x]
    for x in labels:
        if x not in IMPLEMENTATIONS:
            ap.error(f"unknown label {x}")
    want_cases = set(a.cases.split(","))
    rows: List[dict] = []

    def do_suite(group: str, case_iter):
        for case, arrays in case_iter:
            if case.id not in want_cases:
                continue
            case_id = f"{group}-{case.id}"
            print("-" * 80)
            print(f"Benchmark case {case_id}: {case.desc}")
            if arrays is None:
                print("  - WARNING: required columns not found or not numeric. Skipping case.")
                continue
            rows.extend(run_case(mod, case_id, arrays, a.threads, a.warmup, labels))
            if not a.no_duckdb:
                rows.extend(run_duckdb(case_id, arrays, a.threads))

    if a.synthetic:
        n = int(float(a.synthetic))
        n_key = f"{str(n)[0]}e{len(str(n)) - 1}"
        t0 = time.perf_counter()
        tabs = synthetic_suite(n)
        print(f"synthetic suite '{n_key}' generated in {time.perf_counter() - t0:.2f} s")
        if a.write_data:
            print("written:", write_suite(tabs, n_key, a.write_data, a.format))
        do_suite(n_key, iter_cases_from_tables(tabs, pinned))
    else:
        suites = ingest.discover_suites(a.data_dir, a.lhs)
        if not suites:
            print(f"Error: no complete benchmark suites found in {a.data_dir}", file=sys.stderr)
            return 1
        for s in suites:
            print("=" * 80)
            print(f"Suite '{s['group_name']}'")

            def it(s=s):
                for case in ingest.CASES:
                    t0 = time.perf_counter()
                    arrays = ingest.load_case(s, case, pinned) if case.id in want_cases else None
                    if arrays is not None:
                        print(f"  ingest {case.id}: {time.perf_counter() - t0:.3f} s ({'pinned' if pinned else 'pageable'} host buffers)")
                    yield case, arrays

            do_suite(s["group_name"], it())

    print("=" * 80)
    print("All benchmark cases finished.")
    with open(a.json, "w") as f:
        json.dump(rows, f, indent=1, default=str)
    print(f"results: {a.json}")
    for task in TASKS:
        if plot_results(rows, task, f"benchmark_{task}.png"):
            print(f"plot: benchmark_{task}.png")
    return 0


if __name__ == "__main__":
    sys.exit(main())
