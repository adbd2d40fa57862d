"""The bench line's JSON contract, checked without a GPU: bench.main() runs against stand-ins for the C library
(tests/_bench_dry_driver.py), so only the host logic of bench.py is exercised — argument handling, the W >= 3 rule,
aggregation of the per-step stats, and the presence / shape of every key the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "_bench_dry_driver.py"), *args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]  # rank 0 prints ONE JSON line
    return json.loads(lines[0])


def test_bench_line_contract_single_gpu():
    d = _run("--steps", "4", "--warmup", "1")
    assert d["warmup"] == 3 and d["steps"] == 4  # W >= 3 is enforced
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    # the default workload is C3 (1e8 x 1e8 materialize), a fixed total job: strong scaling
    assert d["unit"] == "rows/s" and d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert "C3" in d["config"]["workload"] and d["config"]["entry_point"] == "hash_join_radix" and d["metric"].endswith("(join materialize)")
    assert d["dtype"] == "u64" and d["n_gpus"] == 1 and "workload" in d["config"] and "model" not in d["config"]
    assert "l2" in d["config"]  # says how L2 reuse between timed iterations is ruled out
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] == (2 * 100_000_000 + 100_000_000) * 8
    assert "with_pairs_fetch" in d["e2e"]  # materialize: the e2e with the device->host copy of the pairs rides along
    assert {"matches", "matches_expected", "matches_ok"} <= set(d) and d["matches_expected"] == 89_999_578
    assert set(d["other_configs"]) == {"C2", "C1", "C4c", "C4"} and d["other_configs"]["C2"]["scaling"] == "weak"
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["gpu_launches"] == 4  # one launch per timed step, summed from the per-step stats blocks
    assert d["roofline"]["kernels"] and abs(d["roofline"]["kernels"][0]["ms"] - d["roofline"]["kernel_ms"]) < 1e-12
    assert abs(d["value"] - 100_000_000 * 4 / 0.01) < 1e-3 and abs(d["ms_per_step"] - 2.5) < 1e-9
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_bench_line_other_config():
    d = _run("--config", "C2", "--steps", "3", "--no-e2e", "--no-cpu-baseline", "--no-other")
    assert d["metric"].endswith("(join count)") and d["e2e"] is None and d["cpu_baseline"] is None and d["other_configs"] is None
    assert "C2" in d["config"]["workload"] and d["config"]["entry_point"] == "hash_join_count_bloom" and d["scaling"] == "weak"


def test_bench_line_two_ranks_gloo():
    """Launched like the driver does for N > 1 (torchrun, one rank per GPU): rank 0 alone prints the line, the value is
    the whole-job aggregate, the time is the max over ranks (gloo all-reduce)."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(ROOT / "tests" / "_bench_dry_driver.py"), "--gpus", "2", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    # C3 at 2 GPUs: the same 1e8 x 1e8 job, both sides split (strong scaling through FJ_DIST_SHUFFLE)
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["rows_probe_total"] == 100_000_000
    assert d["config"]["rows_probe_per_gpu"] == 50_000_000
    assert abs(d["value"] - 100_000_000 * 3 / 0.01) < 1e-3
    assert set(d["other_configs"]) == {"C4c", "C4", "C2"} and d["other_configs"]["C4"]["rows_probe_total"] == 250_000_000
    assert d["cpu_baseline"] is None  # the CPU reference is timed at N = 1 only
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
