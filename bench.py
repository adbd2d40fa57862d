#!/usr/bin/env python
"""bench.py — the contract benchmark.

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation

Workload (config.workload): BASELINE.json configs[2] — `flash_join_radix` join MATERIALIZE, x = 1e8 probe rows vs
y = 1e8 build rows, 90 % match (C3: the north_star target and the largest single-GPU configuration).  At N > 1 the
same job is split over the GPUs (strong scaling): every rank holds 1/N of both sides and the engine shuffles them
(peer-memory partition pass on a dense key domain, NCCL all-to-all-v otherwise).  One "step" = one whole join:
control block + partition pass of both sides (+ exchange) + partition joins + pair compaction (+ result exchange).
A default run also times C2 (`flash_join_bloom` count 1e8 x 1e5), C1 and the C4 per-GPU slice (N = 1), or C4
count + materialize with the build side broadcast and C2 weak scaling (N > 1), under "other_configs".

  value      probe rows / second, whole job, inputs resident in HBM, device time (CUDA events on the engine's
             stream, bracketed by barrier + synchronize, max over ranks)
  e2e        same metric through the reference-facing call with HOST (pinned) buffers: the host->device copy of
             every input and the device->host read of the result are inside the timed region
  roofline   the dominant kernel of the step (the longest one: k_part over the build side): algorithmic bytes it
             is responsible for (16 B per build row) / its CUDA-event duration, against
             MEASURED_PEAKS.json hbm_gbs; whole_step_frac = the step's algorithmic bytes / step time / peak
  cpu_baseline  the reference's own CPU code (oracle/_ref, built from /root/reference/hash_join.cpp by
             oracle/build_ref.sh, linked against its vendored mimalloc when that build exists) on this box's host
             cores, same workload, in a numpy-only subprocess
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # N = probe rows (per GPU when dist == "broadcast": weak scaling; in total when dist == "shuffle": strong scaling)
    "C3": dict(N=100_000_000, ny=100_000_000, pct=90, entry="hash_join_radix", algo="radix", bloom=False, mat=True, dist="shuffle",
               desc="C3: flash_join_radix materialize, x=1e8 vs y=1e8, 90% match (BASELINE.json configs[2])"),
    "C2": dict(N=100_000_000, ny=100_000, pct=10, entry="hash_join_count_bloom", algo="scalar", bloom=True, mat=False, dist="broadcast",
               desc="C2: flash_join_bloom count, x=1e8 probe rows vs y=1e5 build rows, ~10% match (BASELINE.json configs[1])"),
    "C1": dict(N=10_000_000, ny=10_000, pct=90, entry="adaptive_join_count", algo="adaptive", bloom=False, mat=False, dist="broadcast",
               desc="C1: adaptive_join count, x=1e7 vs y=1e4, 90% match (BASELINE.json configs[0])"),
    "C4": dict(N=125_000_000, ny=1_000_000, pct=90, entry="adaptive_join", algo="adaptive", bloom=False, mat=True, dist="broadcast",
               desc="C4: adaptive_join materialize, x=1.25e8 probe rows per GPU (1e9 at 8 GPUs) vs y=1e6, build broadcast, probe split "
                    "(BASELINE.json configs[3])"),
    "C4c": dict(N=125_000_000, ny=1_000_000, pct=90, entry="adaptive_join_count", algo="adaptive", bloom=False, mat=False, dist="broadcast",
                desc="C4: adaptive_join count, x=1.25e8 probe rows per GPU (1e9 at 8 GPUs) vs y=1e6, build broadcast, probe split "
                     "(BASELINE.json configs[3])"),
    "C5": dict(N=1_000_000_000, ny=1_000_000_000, pct=90, entry="hash_join_radix", algo="radix", bloom=False, mat=True, dist="shuffle",
               desc="C5: flash_join_radix materialize, x=1e9 vs y=1e9, 90% match, radix shuffle (BASELINE.json configs[4])"),
}
SEED = 108
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
NVLINK_GBS = 900.0         # per direction and GPU (NVLink 5 nominal)


def expected_matches(n_total: int, ny: int, pct: int):
    """The count the reference (or, for C5, the generator itself) gives for probe rows [0, n_total) of the G2 data set:
    tests/golden/g1_goldens.json (gen == 'g2') and tests/golden/g2_counts.json.  None when no golden exists."""
    try:
        for c in json.loads((ROOT / "tests" / "golden" / "g1_goldens.json").read_text())["cases"]:
            if c.get("gen") == "g2" and (c["N"], c["ny"], c["match_pct"]) == (n_total, ny, pct):
                return int(c["count"]), "reference (tests/golden/g1_goldens.json)"
        for c in json.loads((ROOT / "tests" / "golden" / "g2_counts.json").read_text())["cases"]:
            if (c["N"], c["ny"], c["match_pct"]) == (n_total, ny, pct):
                if c.get("reference_count") is not None:
                    return int(c["reference_count"]), "reference on probe slices (tests/golden/g2_counts.json)"
                return int(c["generator_count"]), "generator-implied count (tests/golden/g2_counts.json)"
    except Exception:
        pass
    return None, None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        rows = []
        try:
            for line in Path(self.f.name).read_text().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) >= 9:
                    rows.append(c)
        finally:
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons), "samples": len(rows), "power_w_max": max((float(r[3]) for r in rows if r[3].replace(".", "").isdigit()), default=None)}


# ------------------------------------------------------------------------------------------------
# CPU reference worker (separate process: numpy + the compiled reference only)
def cpu_worker(args) -> None:
    import numpy as np

    from flash_hash_join_b200.datagen import g2_slice
    from oracle import oracle as O

    w = WORKLOADS[args.config]
    build = None
    if O.reference_available("plain"):
        preloaded = "mimalloc" in os.environ.get("LD_PRELOAD", "")
        build = ("plain", "unmodified hash_join.cpp, g++ -O3 -msse4.2 -mavx2, " +
                 ("all allocations through its vendored mimalloc (LD_PRELOAD of oracle/_ref/mimalloc/libmimalloc_override.so), as its "
                  "CMakeLists.txt arranges with MI_OVERRIDE" if preloaded else "glibc malloc (mimalloc stubbed out)"))
    mod = O.load_reference(build[0]) if build is not None else None
    t0 = time.perf_counter()
    bk, bv = g2_slice(w["N"], w["ny"], w["pct"], SEED, "build", 0, w["ny"])
    pk = g2_slice(w["N"], w["ny"], w["pct"], SEED, "probe", 0, w["N"])
    gen_s = time.perf_counter() - t0
    if build is not None:
        kind = "reference"
        fn = getattr(mod, w["entry"])
        cores = os.cpu_count()

        def step():
            n, sec = fn(bk, bv, pk)
            return n, sec  # the reference's own SimpleTimer seconds (hash_join.cpp:45-55)
    else:
        kind = "port"
        cores = 1

        def step():
            t = time.perf_counter()
            n, _, _ = O.join(w["algo"], w["bloom"], w["mat"], bk, bv, pk)
            return n, time.perf_counter() - t
    for _ in range(args.warmup):
        step()
    times, n = [], None
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        n, s = step()
        times.append(s)
    wall = time.perf_counter() - wall0
    print(json.dumps({"kind": kind, "cores": cores, "matches": int(n), "core_s_mean": sum(times) / len(times), "core_s_best": min(times),
                      "wall_s_per_step": wall / len(times), "gen_s": gen_s, "rows": w["N"],
                      "build": build[1] if build else "oracle/join_oracle.c (single-threaded port)"}))


def run_cpu_worker(config: str, steps: int, warmup: int) -> dict:
    cmd = [sys.executable, str(ROOT / "bench.py"), "--_cpu_worker", "--config", config, "--steps", str(steps), "--warmup", str(warmup)]
    env = dict(os.environ)
    mi = ROOT / "oracle" / "_ref" / "mimalloc" / "libmimalloc_override.so"  # the reference's own allocator, when oracle/build_ref.sh built it
    if mi.exists():
        env["LD_PRELOAD"] = str(mi)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env)
    if out.returncode != 0:
        raise RuntimeError("cpu worker failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


# ------------------------------------------------------------------------------------------------
def dist_setup(n_gpus: int):
    """torch.distributed (gloo) is plumbing only: rendezvous, barrier, max-over-ranks, and carrying the
    NCCL unique id to every rank.  The data path uses the engine's own NCCL communicator."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return rank, world, local, None
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    return rank, world, local, dist


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="headline workload only (skip other_configs / general_path)")
    ap.add_argument("--no-dense", action="store_true", help="switch the dense-key-domain fast paths off (general hash path only)")
    ap.add_argument("--_cpu_worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3 and not args._cpu_worker and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args._cpu_worker:
        cpu_worker(args)
        return
    w = WORKLOADS[args.config]
    metric = "probe rows/sec (join materialize)" if w["mat"] else "probe rows/sec (join count)"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        if w["N"] > 200_000_000:
            print(json.dumps({"impl": "reference", "unavailable": f"{args.config} does not fit a CPU box (SURVEY.md §8d); use --config C3"}))
            return
        # one step = the whole workload (C3 on 16 host cores: a few seconds), so a K-step run stays within minutes
        steps = max(1, min(args.steps, 10))
        r = run_cpu_worker(args.config, steps, max(1, min(args.warmup, 2)))
        value = r["rows"] / r["core_s_mean"]
        exp, exp_src = expected_matches(w["N"], w["ny"], w["pct"])
        line = {
            "impl": "reference", "metric": metric, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": r["core_s_mean"] * 1e3, "higher_is_better": True,
            "scaling": "strong" if w["dist"] == "shuffle" else "weak",
            "vs_baseline": None, "dtype": "u64", "data": f"synthetic (G2 counter-based h2o join shape, seed {SEED})",
            "config": {"workload": w["desc"], "entry_point": w["entry"], "rows_probe_total": w["N"], "rows_build": w["ny"], "match_pct": w["pct"],
                       "note": "reference CPU implementation on the host cores of this box (one process, all cores, whatever --gpus says); "
                               "time = its own core seconds (SimpleTimer)"},
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": f"full workload ({w['N']} probe rows x {w['ny']} build rows) x {steps} steps", "build": r.get("build")},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "matches": r["matches"], "matches_expected": exp, "matches_ok": (r["matches"] == exp) if exp is not None else None,
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ this engine
    rank, world, local, dist = dist_setup(args.gpus)
    os.environ["LOCAL_RANK"] = str(local)
    import ctypes as C

    import numpy as np

    from flash_hash_join_b200 import capi, flash_join

    L = capi.lib()
    capi.check(L.fj_init(local))
    if world > 1:
        ident = [None]
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            capi.check(L.fj_comm_unique_id(buf))
            ident = [bytes(buf)]
        dist.broadcast_object_list(ident, src=0)
        capi.check(L.fj_comm_init(rank, world, ident[0]))

    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        capi.check(L.fj_device_synchronize())
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = measured_peak()
    traffic_db = {}
    try:
        traffic_db = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    except Exception:
        pass

    def kernels_of(last: dict, wl: dict, nb_local: int, N: int, matches_local: int):
        """The step's kernels with their CUDA-event durations and the algorithmic bytes each is responsible for
        (inputs read once, outputs written once); the first entry is the dominant (longest) one."""
        ks = []
        if last["path"] == "radix" and last["dense"] == 2:
            ks.append(("k_part<VAL> (one partition pass over the build side: TMA input rings -> 2048-way shared-memory write-combining)",
                       last["part_build_us"] * 1e-6, (16.0 if wl["mat"] else 8.0) * nb_local))
            ks.append(("k_part (partition pass over the probe side)", last["part_probe_us"] * 1e-6, 8.0 * N))
            ks.append(("k_sjoin + k_pairs_compact (shared-memory direct-address join of the partitions, pair output)" if wl["mat"]
                       else "k_sjoin (shared-memory direct-address join of the partitions)", last["probe_s"], 16.0 * matches_local if wl["mat"] else 0.0))
        elif last["path"] == "radix":
            name = ("k_djoin (L2-resident direct-address join)" if last["dense"] else
                    ("k_join3 (shared-memory partition join)" if last["radix_bits2"] else "k_join (shared-memory partition join)"))
            ks.append(("k_scatter2 (radix passes over both sides)", last["partition_s"], (16.0 if wl["mat"] else 8.0) * nb_local + 8.0 * N))
            ks.append((name, last["probe_s"], 16.0 * matches_local if wl["mat"] else 0.0))
        else:
            fused = last["dense"] and last["kernel_launches"] == 1
            if fused:
                name = ("k_count_dense_peer (bitmap build + probe + count exchange over NVLink peer memory, one launch per GPU)" if world > 1
                        else ("k_mat_dense_fused (bitmap + direct-address values, one persistent launch)" if wl["mat"]
                              else "k_count_dense_fused (bitmap build + probe, one persistent launch)"))
                ks.append((name, last["probe_s"], 8.0 * (N + nb_local) + (16.0 * matches_local + 8.0 * nb_local if wl["mat"] else 0.0)))
            else:
                name = "k_probe_mat" if wl["mat"] else ("k_probe_count_dense (exact bitmap)" if last["dense"] else "k_probe_count")
                ks.append((name, last["probe_s"], 8.0 * N + (16.0 * matches_local if wl["mat"] else 0.0)))
                if last["build_s"] > 0:
                    ks.append(("k_build", last["build_s"], (16.0 if wl["mat"] else 8.0) * nb_local))
        ks = [k for k in ks if k[1] > 0]
        ks.sort(key=lambda k: -k[1])
        return ks

    def measure(cfg_name: str, steps: int, warmup: int, want_e2e: bool, dense: bool = True, fetch_pairs: bool = False) -> dict:
        from flash_hash_join_b200.dist import row_slice

        capi.config_set(dense=1 if dense else 0)
        wl = WORKLOADS[cfg_name]
        ny, pct = wl["ny"], wl["pct"]
        shuffle = world > 1 and wl["dist"] == "shuffle"  # large build side: both sides split (strong scaling)
        if shuffle or (world == 1 and wl["dist"] == "shuffle"):
            n_total = wl["N"]
            p0, p1 = row_slice(n_total, world, rank)
            b0, b1 = row_slice(ny, world, rank)
            N = p1 - p0
        else:  # weak scaling: every GPU gets its own probe slice of the named size, the build side is replicated
            N = wl["N"]
            n_total = N * world
            p0, b0, b1 = rank * N, 0, ny
        nb_local = b1 - b0
        algo = {"adaptive": capi.ALGO_ADAPTIVE, "scalar": capi.ALGO_SCALAR, "radix": capi.ALGO_RADIX}[wl["algo"]]
        flags = (capi.FLAG_BLOOM if wl["bloom"] else 0) | (capi.FLAG_MATERIALIZE if wl["mat"] else 0)
        mode = capi.DIST_SHUFFLE if shuffle else capi.DIST_BROADCAST
        # inputs resident in HBM
        d_bk, d_bv = capi.generate_g2("build", n_total, ny, pct, SEED, b0, nb_local)
        d_pk = capi.generate_g2("probe", n_total, ny, pct, SEED, p0, N)

        # argument objects are made once (one stats block per timed step): the timed loop is the C-ABI call and nothing else
        n, nl, sec = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
        sts = [capi.Stats() for _ in range(max(steps, 1))]
        p_n, p_nl, p_sec = C.byref(n), C.byref(nl), C.byref(sec)
        p_sts = [C.byref(x) for x in sts]
        dflags = flags | capi.FLAG_DEVICE_INPUTS
        a_bk, a_bv, a_pk = d_bk.ptr, d_bv.ptr, d_pk.ptr

        def step_device(i=0):
            if world == 1:
                rc = L.fj_join_u64(algo, dflags, a_bk, a_bv, ny, a_pk, N, p_n, p_sec, p_sts[i])
            else:
                rc = L.fj_join_dist_u64(mode, algo, dflags, 0, a_bk, a_bv, nb_local, a_pk, N, p_n, p_nl, p_sec, p_sts[i])
            if rc:
                capi.check(rc)
            return n.value

        for _ in range(warmup):
            matches = step_device()
        barrier()
        capi.check(L.fj_timer_start())
        for i in range(steps):
            matches = step_device(i)
        t = C.c_double(0)
        capi.check(L.fj_timer_stop(C.byref(t)))
        barrier()
        elapsed = max_over_ranks(t.value)
        value = n_total * steps / elapsed
        launches = sum(x.kernel_launches for x in sts[:steps])
        phases = {k: sum(getattr(x, k) for x in sts[:steps]) for k in ("clear_s", "build_s", "partition_s", "probe_s", "comm_s")}
        # per-kernel durations: mean over the timed steps
        mean = {k: statistics.mean(getattr(x, k) for x in sts[:steps]) for k in ("part_build_us", "part_probe_us", "probe_s", "partition_s", "build_s")}
        last = sts[steps - 1].as_dict()
        last.update(mean)
        matches_local = nl.value if world > 1 else matches

        ks = kernels_of(last, wl, nb_local, N, matches_local)
        dom = ks[0]
        achieved = dom[2] / dom[1] * 1e-9
        tkey = cfg_name + ("" if last["dense"] else "_general")
        tinfo = traffic_db.get(tkey, {}) if world == 1 else {}
        alg_step = (16.0 * ny + 8.0 * n_total + 16.0 * matches) if wl["mat"] else 8.0 * (ny + n_total)
        step_s = elapsed / steps
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": tinfo.get("dram_bytes_per_launch"), "traffic_source": tinfo.get("source"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dom[2], "kernel_ms": dom[1] * 1e3,
                    "kernels": [{"kernel": k[0], "ms": k[1] * 1e3, "share_of_step": k[1] / step_s, "algorithmic_bytes": k[2],
                                 "frac": (k[2] / k[1] * 1e-9 / peak) if k[2] else None} for k in ks],
                    "whole_step_frac": (alg_step / step_s * 1e-9) / (peak * world), "whole_step_algorithmic_bytes": alg_step,
                    "whole_step_frac_of_8000_GBps_nominal": (alg_step / step_s * 1e-9) / (8000.0 * world)}
        if shuffle and last["dense"] == 2:
            out_bytes = (world - 1) / world * ((4.0 if wl["mat"] else 2.0) * nb_local + 2.0 * N)  # partition rows stored into peers
            roofline["nvlink"] = {"bytes_leaving_each_gpu_per_step": out_bytes, "bound_ms_at_900_GBps": out_bytes / (NVLINK_GBS * 1e9) * 1e3,
                                  "partition_phase_ms": last["partition_s"] * 1e3}
        elif shuffle:
            out_bytes = (world - 1) / world * (8.0 * nb_local + 4.0 * N)  # packed rows through ncclSend/ncclRecv
            roofline["nvlink"] = {"bytes_leaving_each_gpu_per_step": out_bytes, "bound_ms_at_900_GBps": out_bytes / (NVLINK_GBS * 1e9) * 1e3}

        # e2e: host (pinned) buffers through the reference-facing call
        e2e = None
        if want_e2e:
            h_bk, h_bv, h_pk = flash_join.pinned_empty(nb_local), flash_join.pinned_empty(nb_local), flash_join.pinned_empty(N)
            capi.check(L.fj_memcpy_d2h(h_bk.ctypes.data, d_bk.ptr, nb_local * 8))
            capi.check(L.fj_memcpy_d2h(h_bv.ctypes.data, d_bv.ptr, nb_local * 8))
            capi.check(L.fj_memcpy_d2h(h_pk.ctypes.data, d_pk.ptr, N * 8))
            e_steps = max(3, min(steps, 10))

            def step_host():
                if world == 1:
                    n, _sec = getattr(flash_join, wl["entry"])(h_bk, h_bv, h_pk)  # the call a flash_join user makes
                    return n
                n = C.c_uint64(0)
                capi.check(L.fj_join_dist_u64(mode, algo, flags, 0, h_bk.ctypes.data, h_bv.ctypes.data, nb_local, h_pk.ctypes.data, N,
                                              C.byref(n), None, None, None))
                return n.value

            step_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                m2 = step_host()
            capi.check(L.fj_device_synchronize())
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            assert m2 == matches, (m2, matches)
            e2e = {"value": n_total * e_steps / dt, "unit": "rows/s", "h2d_bytes_per_step": (2 * nb_local + N) * 8, "d2h_bytes_per_step": 48,
                   "steps": e_steps, "ms_per_step": dt / e_steps * 1e3, "host_memory": "pinned (flash_join.pinned_empty)",
                   "api": f"flash_join.{wl['entry']}" if world == 1 else "fj_join_dist_u64 (C ABI, host buffers)",
                   "note": "result pairs of a materialize call stay in HBM (the reference drops them, hash_join.cpp:380); the count is read back. "
                           "with_pairs_fetch adds flash_join.last_pairs(): the device->host copy of every pair"}
            if wl["mat"] and fetch_pairs and world == 1:
                try:
                    fn = getattr(flash_join, wl["entry"])
                    t0 = time.perf_counter()
                    for _ in range(2):
                        m4, _sec = fn(h_bk, h_bv, h_pk)
                        pk_out, pv_out = flash_join.last_pairs()[:2]
                    dtf = (time.perf_counter() - t0) / 2
                    e2e["with_pairs_fetch"] = {"value": N / dtf, "unit": "rows/s", "ms_per_step": dtf * 1e3, "d2h_bytes_per_step": int(16 * m4 + 48),
                                               "pairs": int(len(pk_out)), "pairs_ok": bool(len(pk_out) == m4 == matches)}
                    del pk_out, pv_out
                except Exception as ex:  # informational only
                    e2e["with_pairs_fetch"] = {"error": str(ex)[:200]}
            # the same call with plain (pageable) numpy columns, as a reference user would pass them: the engine stages
            # them through pinned buffers with a few host threads (informational; the headline e2e uses pinned memory)
            if world == 1:
                try:
                    p_bk, p_bv, p_pk = np.array(h_bk), np.array(h_bv), np.array(h_pk)
                    fn = getattr(flash_join, wl["entry"])
                    fn(p_bk, p_bv, p_pk)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        m3, _sec = fn(p_bk, p_bv, p_pk)
                    dtp = (time.perf_counter() - t0) / 3
                    e2e["pageable_inputs"] = {"value": N / dtp, "unit": "rows/s", "ms_per_step": dtp * 1e3, "matches_ok": bool(m3 == matches),
                                              "note": "plain numpy columns, staged host->device by Engine::h2d (8 threads, pinned ring)"}
                    del p_bk, p_bv, p_pk
                except Exception as ex:  # informational only
                    e2e["pageable_inputs"] = {"error": str(ex)[:200]}
            del h_bk, h_bv, h_pk
        for x in (d_bk, d_bv, d_pk):
            x.free()
        capi.config_set(dense=1)
        if shuffle and last["dense"] == 2:
            par = (f"both sides split over {world} GPUs; ONE local partition pass per side into the GPU's own IPC-mapped partition buffer "
                   f"(k_part), device-side count push / barrier / result exchange (k_xsync), shared-memory join of 2048/{world} partitions per "
                   "GPU whose TMA producer pulls the partition rows from every peer's buffer over NVLink (k_sjoin); no NCCL call in the step")
        elif shuffle:
            par = f"both sides split over {world} GPUs, rows hash-partitioned by destination, NCCL all-to-all-v, local radix join, count ncclAllReduce"
        elif world > 1 and last["dense"] and last["kernel_launches"] == 1 and not wl["mat"]:
            par = (f"probe side split over {world} GPUs; every GPU reads the build keys from rank 0 and exchanges its count through "
                   "IPC-mapped peer memory inside the one count kernel (no NCCL call in the step)")
        elif world > 1:
            par = f"build side ncclBroadcast from rank 0, probe side split over {world} GPUs, count ncclAllReduce"
        else:
            par = "single GPU"
        exp, exp_src = expected_matches(n_total, ny, pct)
        return {"w": wl, "N": N, "n_total": n_total, "value": value, "elapsed": elapsed, "steps": steps, "matches": matches, "last": last,
                "launches": launches, "roofline": roofline, "e2e": e2e, "scaling": "strong" if wl["dist"] == "shuffle" else "weak",
                "parallelism": par, "phases": {k: v / steps * 1e3 for k, v in phases.items()},
                "expected": exp, "expected_source": exp_src}

    def summary(o: dict) -> dict:
        d = {"workload": o["w"]["desc"], "value": o["value"], "unit": "rows/s", "ms_per_step": o["elapsed"] / o["steps"] * 1e3,
             "rows_probe_total": o["n_total"], "scaling": o["scaling"], "parallelism": o["parallelism"],
             "matches": o["matches"], "matches_expected": o["expected"], "matches_expected_source": o["expected_source"],
             "matches_ok": (o["matches"] == o["expected"]) if o["expected"] is not None else None,
             "path": o["last"]["path"], "dense_key_domain": o["last"]["dense"],
             "filter": o["last"]["bloom_kind"], "radix_bits": [o["last"]["radix_bits1"], o["last"]["radix_bits2"]],
             "narrow_rows": bool(o["last"]["narrow"]), "phases_ms_per_step": o["phases"], "roofline": o["roofline"],
             "gpu_launches": o["launches"]}
        if o["e2e"]:
            d["e2e"] = o["e2e"]
        return d

    m = measure(args.config, args.steps, args.warmup, not args.no_e2e, dense=not args.no_dense, fetch_pairs=True)
    w, matches = m["w"], m["matches"]
    # the same workload with the data-dependent dense-key-domain fast paths switched off: the general hash path
    # (table + register-blocked Bloom filter / two radix passes + shared-memory join)
    general = None
    if world == 1 and m["last"]["dense"] and not args.no_other:
        try:
            g = measure(args.config, max(5, args.steps // 2), 3, False, dense=False)
            general = summary(g)
            if g["matches"] != matches:
                general["MISMATCH"] = f"general path counted {g['matches']}, dense path {matches}"
        except Exception as e:
            general = {"error": str(e)[:300]}
    other = None
    if not args.no_other and args.config == "C3":
        other = {}
        names = ("C2", "C1", "C4c", "C4") if world == 1 else ("C4c", "C4", "C2")
        for name in names:
            try:
                o = measure(name, 10 if world == 1 else 8, 3, False, dense=not args.no_dense)
                other[name] = summary(o)
                if world == 1 and o["last"]["dense"] and name in ("C2", "C4"):
                    og = measure(name, 5, 3, False, dense=False)
                    other[name]["general_path"] = {k: v for k, v in summary(og).items() if k in ("value", "ms_per_step", "matches", "path", "filter", "roofline")}
                    if og["matches"] != o["matches"]:
                        other[name]["MISMATCH"] = f"general path counted {og['matches']}, dense path {o['matches']}"
            except Exception as e:  # never lose the headline line
                other[name] = {"error": str(e)[:300]}
    clocks = sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and w["N"] <= 200_000_000:
        try:
            r = run_cpu_worker(args.config, 3, 1)
            cpu = {"value": r["rows"] / r["core_s_mean"], "unit": "rows/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": f"full workload ({r['rows']} probe rows x {w['ny']} build rows) x 3 steps after 1 warm-up, time = reference core seconds",
                   "ms_per_step": r["core_s_mean"] * 1e3, "matches": r["matches"], "build": r.get("build")}
            if r["matches"] != matches:
                cpu["MISMATCH"] = f"reference counted {r['matches']}, engine counted {matches}"
        except Exception as e:  # the baseline is reported, never required for the engine number
            cpu = {"value": None, "unit": "rows/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)[:200]}

    if rank == 0:
        last = m["last"]
        dense_note = {2: "dense-key-domain radix path (keys < 2^16 * partitions: one 2048-way partition pass by the low key bits, rows shrink to "
                         "idx16|value16 / idx16, direct-address join in shared memory); the general hash path on the same inputs is under general_path",
                      1: "dense-key-domain fast path (keys < 2*rows: exact membership bitmap / direct addressing); the general hash path on the same "
                         "inputs is under general_path", 0: "general hash path"}[int(last["dense"])]
        line = {
            "metric": metric, "value": m["value"], "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["elapsed"] / args.steps * 1e3, "higher_is_better": True, "scaling": m["scaling"], "vs_baseline": None, "dtype": "u64",
            "data": f"synthetic (G2 counter-based h2o join shape, build rows in pseudo-random order, seed {SEED}; generated in HBM)",
            "config": {"workload": w["desc"], "entry_point": w["entry"], "rows_probe_per_gpu": m["N"], "rows_probe_total": m["n_total"],
                       "rows_build": w["ny"], "match_pct": w["pct"],
                       "l2": f"inputs {(8 * m['N'] + 16 * (w['ny'] // world if m['scaling'] == 'strong' else w['ny'])) / 1e6:.0f} MB per GPU and step > 126 MB L2 "
                             "(no flush needed); partitions / table are rebuilt every step",
                       "parallelism": m["parallelism"], "path": last["path"], "narrow_slots": bool(last["narrow"]), "bloom": last["bloom_kind"],
                       "dense_key_domain": int(last["dense"]), "note": dense_note},
            "clocks": dict(clocks, window="sampled every 100 ms from before warm-up to the end of the last measurement"),
            "e2e": m["e2e"], "gpu_launches": m["launches"], "roofline": m["roofline"], "cpu_baseline": cpu,
            "matches": matches, "matches_expected": m["expected"], "matches_expected_source": m["expected_source"],
            "matches_ok": (matches == m["expected"]) if m["expected"] is not None else None,
            "phases_ms_per_step": m["phases"], "general_path": general, "other_configs": other,
        }
        print(json.dumps(line))
    if dist is not None:
        capi.check(L.fj_comm_destroy())
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
