"""CPU (gloo, world_size 2) tests of the host-side logic of the multi-GPU joins: row slicing, the shuffle
destination function and the way per-rank results compose.  The local join on every rank is the oracle here
(allowed in tests/ only); on GPUs it is fj_join_dist_u64 — tests/dist_gpu_check.py is the same scenario on real
devices (run with `gpurun --gpus 2 -- python -m torch.distributed.run ... tests/dist_gpu_check.py`)."""
import os
import socket

import numpy as np
import pytest

from flash_hash_join_b200.datagen import g1, g2_slice
from flash_hash_join_b200.dist import hash32, row_slice, shuffle_dest
from oracle import oracle as O


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, ny, pct, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank generates ITS slices of both sides (G2 is counter based: no global array anywhere)
        b0, b1 = row_slice(ny, world, rank)
        p0, p1 = row_slice(N, world, rank)
        bk, bv = g2_slice(N, ny, pct, 108, "build", b0, b1)
        pk = g2_slice(N, ny, pct, 108, "probe", p0, p1)
        # --- broadcast mode: build side replicated, probe side split, counts summed
        all_b = [None] * world
        dist.all_gather_object(all_b, (bk, bv))
        fbk = np.concatenate([x[0] for x in all_b]); fbv = np.concatenate([x[1] for x in all_b])
        n_b, kb, vb = O.np_join(fbk, fbv, pk)
        # --- shuffle mode: rows travel to shuffle_dest(key); equal keys meet on one rank
        db, dp = shuffle_dest(bk, world), shuffle_dest(pk, world)
        send = [(bk[db == r], bv[db == r], pk[dp == r]) for r in range(world)]
        everything = [None] * world
        dist.all_gather_object(everything, send)
        mine = [everything[src][rank] for src in range(world)]
        rbk = np.concatenate([m[0] for m in mine]); rbv = np.concatenate([m[1] for m in mine]); rpk = np.concatenate([m[2] for m in mine])
        n_s, ks, vs = O.np_join(rbk, rbv, rpk)
        import torch

        t = torch.tensor([n_b, n_s], dtype=torch.int64)
        dist.all_reduce(t)
        pairs = [None] * world
        dist.all_gather_object(pairs, (kb, vb, ks, vs))
        if rank == 0:
            out["counts"] = t.tolist()
            out["pairs_b"] = O.sorted_pairs(np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs]))
            out["pairs_s"] = O.sorted_pairs(np.concatenate([p[2] for p in pairs]), np.concatenate([p[3] for p in pairs]))
    finally:
        dist.destroy_process_group()


def test_row_slice_covers_everything():
    for n in (0, 1, 7, 100, 101, 10**9 + 3):
        for world in (1, 2, 3, 8):
            s = [row_slice(n, world, r) for r in range(world)]
            assert s[0][0] == 0 and s[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(s, s[1:]))
            assert max(b - a for a, b in s) <= (n + world - 1) // world


def test_hash32_matches_known_answers():
    # lowbias32 of the low word xor the high word * 0x9E3779B1 (csrc/fj_common.cuh); the finaliser is a
    # bijection on 32-bit values (k_join3 relies on it)
    x = hash32(np.arange(1 << 16, dtype=np.uint64))
    assert np.unique(x).size == 1 << 16
    assert int(hash32(np.array([0], dtype=np.uint64))[0]) == 0
    a, b = hash32(np.array([5, 5 + (1 << 32)], dtype=np.uint64))
    assert a != b


def test_shuffle_dest_partitions_keys():
    bk, bv, pk = g1(200_000, 50_000, 90)
    for world, virtual in ((2, 1), (4, 1), (8, 1), (2, 4), (3, 5)):
        db, dp = shuffle_dest(bk, world, virtual), shuffle_dest(pk, world, virtual)
        assert db.min() >= 0 and db.max() < world
        total = sum(O.np_join(bk[db == r], bv[db == r], pk[dp == r])[0] for r in range(world))
        assert total == O.np_join(bk, bv, pk)[0]
        assert np.bincount(db, minlength=world).min() > 0.5 * bk.size / world


@pytest.mark.timeout(300)
def test_two_rank_broadcast_and_shuffle_compose():
    import torch.multiprocessing as mp

    N, ny, pct = 120_001, 30_000, 90
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, N, ny, pct, out), nprocs=2, join=True)
    bk, bv = g2_slice(N, ny, pct, 108, "build", 0, ny)
    pk = g2_slice(N, ny, pct, 108, "probe", 0, N)
    n0, k0, v0 = O.np_join(bk, bv, pk)
    assert out["counts"] == [n0, n0]
    sp0 = O.sorted_pairs(k0, v0)
    assert np.array_equal(out["pairs_b"], sp0) and np.array_equal(out["pairs_s"], sp0)
