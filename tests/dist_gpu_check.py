#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU (run under torchrun on a box with >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py [--rows 4000000]

Every rank generates its slices of the G2 data set, runs fj_join_dist_u64 in BROADCAST and SHUFFLE mode (count
and materialize) and the pairs of all ranks are gathered on rank 0 and compared, as a sorted multiset, with the
oracle's join of the whole data set.  Results must be identical for every world size."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--build", type=int, default=0, help="build rows (default: = rows for shuffle, rows/100 for broadcast)")
    a = ap.parse_args()
    import torch.distributed as dist

    from flash_hash_join_b200 import capi
    from flash_hash_join_b200.datagen import g2_slice
    from flash_hash_join_b200.dist import rendezvous_comm, row_slice
    from oracle import oracle as O

    dist.init_process_group("gloo")
    rank, world = rendezvous_comm(dist)
    report = {"world": world, "cases": []}
    ok = True
    for mode, name, N, ny in ((capi.DIST_BROADCAST, "broadcast", a.rows, a.build or a.rows // 100),
                              (capi.DIST_SHUFFLE, "shuffle", a.rows, a.build or a.rows)):
        p0, p1 = row_slice(N, world, rank)
        pk = g2_slice(N, ny, 90, 108, "probe", p0, p1)
        if mode == capi.DIST_BROADCAST:
            bk, bv = g2_slice(N, ny, 90, 108, "build", 0, ny) if rank == 0 else (np.empty(0, np.uint64), np.empty(0, np.uint64))
            nb_arg = ny
        else:
            b0, b1 = row_slice(ny, world, rank)
            bk, bv = g2_slice(N, ny, 90, 108, "build", b0, b1)
        for flags in (0, capi.FLAG_MATERIALIZE):
            if mode == capi.DIST_BROADCAST and rank != 0:
                # non-root ranks pass nb (the size) but no data
                import ctypes as C

                g, l, sec = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
                st = capi.Stats()
                pkc = np.ascontiguousarray(pk)
                capi.check(capi.lib().fj_join_dist_u64(mode, capi.ALGO_ADAPTIVE, flags, 0, None, None, ny, pkc.ctypes.data, pkc.size,
                                                       C.byref(g), C.byref(l), C.byref(sec), C.byref(st)))
                g, l, sec, st = g.value, l.value, sec.value, st.as_dict()
            else:
                g, l, sec, st = capi.join_dist(mode, capi.ALGO_ADAPTIVE, flags, 0, bk, bv, pk)
            k, v = capi.pairs() if flags & capi.FLAG_MATERIALIZE else (np.empty(0, np.uint64), np.empty(0, np.uint64))
            gathered = [None] * world
            dist.all_gather_object(gathered, (l, k, v, sec, st["comm_s"], st["path"] + ("/dense%d" % st["dense"] if st["dense"] else "")))
            if rank == 0:
                fbk, fbv = g2_slice(N, ny, 90, 108, "build", 0, ny)
                fpk = g2_slice(N, ny, 90, 108, "probe", 0, N)
                n0, k0, v0 = O.np_join(fbk, fbv, fpk)
                case_ok = g == n0 and sum(x[0] for x in gathered) == n0
                if flags & capi.FLAG_MATERIALIZE:
                    sp = O.sorted_pairs(np.concatenate([x[1] for x in gathered]), np.concatenate([x[2] for x in gathered]))
                    case_ok = case_ok and np.array_equal(sp, O.sorted_pairs(k0, v0))
                ok = ok and case_ok
                report["cases"].append({"mode": name, "materialize": bool(flags), "rows": N, "build_rows": ny, "matches": g, "expected": n0,
                                        "ok": bool(case_ok), "device_ms_max": max(x[3] for x in gathered) * 1e3,
                                        "comm_ms_max": max(x[4] for x in gathered) * 1e3, "paths": [x[5] for x in gathered]})
    # ---- broadcast-mode COUNT with retries.  The all-reduce of the control block is enqueued behind the first
    # attempt (one host sync per step); when any rank has to retry, every rank must notice and take the second
    # all-reduce.  (a) symmetric: one build key far outside every optimistic domain -> all ranks re-run;
    # (b) asymmetric: radix count where only the LAST rank's probe slice is skewed onto one key -> only that rank's
    # partitions overflow and only it re-runs on the global table.
    import ctypes as C

    def bcast_count(algo, bk, bv, nb, pk):
        g, l, sec, st = C.c_uint64(0), C.c_uint64(0), C.c_double(0), capi.Stats()
        pkc = np.ascontiguousarray(pk)
        if rank == 0:
            bkc, bvc = np.ascontiguousarray(bk), np.ascontiguousarray(bv)
            capi.check(capi.lib().fj_join_dist_u64(capi.DIST_BROADCAST, algo, 0, 0, bkc.ctypes.data, bvc.ctypes.data, nb, pkc.ctypes.data,
                                                   pkc.size, C.byref(g), C.byref(l), C.byref(sec), C.byref(st)))
        else:
            capi.check(capi.lib().fj_join_dist_u64(capi.DIST_BROADCAST, algo, 0, 0, None, None, nb, pkc.ctypes.data, pkc.size,
                                                   C.byref(g), C.byref(l), C.byref(sec), C.byref(st)))
        return g.value, l.value, st.as_dict()

    N, ny = 2_000_000, 500_000  # the membership bitmap of the peer-memory count kernel must fit shared memory
    fbk, fbv = g2_slice(N, ny, 90, 108, "build", 0, ny)
    fpk = g2_slice(N, ny, 90, 108, "probe", 0, N)
    p0, p1 = row_slice(N, world, rank)
    # every case with the plain peer kernel (every rank pulls all keys from the root) and with the relay (key slices +
    # partial bitmaps; forced here, by default only from 2^18 build rows): the verdict on a key outside the domain must
    # reach every rank in both
    for relay_min in (1 << 30, 1):
        capi.config_set(peer_relay_min_rows=relay_min)
        for label, algo, bk, pk_local in (
                ("symmetric retry (key outside the domain)", capi.ALGO_SCALAR, np.concatenate([fbk[:-1], np.array([10**12], np.uint64)]), fpk[p0:p1]),
                ("asymmetric retry (one rank's probe slice skewed)", capi.ALGO_RADIX, fbk,
                 np.full(p1 - p0, fbk[5], np.uint64) if rank == world - 1 else fpk[p0:p1]),
                ("no retry", capi.ALGO_ADAPTIVE, fbk, fpk[p0:p1]),
                ("no retry, odd build size", capi.ALGO_SCALAR, fbk[:-3], fpk[p0:p1])):
            g, l, st = bcast_count(algo, bk, fbv[:bk.size], bk.size, pk_local)
            gathered = [None] * world
            dist.all_gather_object(gathered, (g, l, st["attempts"], pk_local))
            if rank == 0:
                n0 = O.np_join(bk, fbv[:bk.size], np.concatenate([x[3] for x in gathered]))[0]
                case_ok = all(x[0] == n0 for x in gathered) and sum(x[1] for x in gathered) == n0
                ok = ok and case_ok
                report["cases"].append({"mode": "broadcast count, " + label + (" [relay]" if relay_min == 1 else ""), "matches": g, "expected": n0,
                                        "ok": bool(case_ok), "attempts": [x[2] for x in gathered]})
    capi.config_set(peer_relay_min_rows=1 << 18)
    if rank == 0:
        report["ok"] = bool(ok)
        print(json.dumps(report))
    capi.comm_destroy()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
