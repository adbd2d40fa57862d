"""Host-side helpers of the one-process-per-GPU joins (no reference counterpart: hash_join.cpp is
single-process).  torch.distributed is plumbing only — rendezvous, barriers and carrying the 128-byte NCCL id
of the engine's own communicator; the data path is fj_join_dist_u64 (NCCL inside libflashjoin_b200.so).
"""
from __future__ import annotations

import os

import numpy as np

__all__ = ["row_slice", "hash32", "shuffle_dest", "init_from_env", "rendezvous_comm"]


def row_slice(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [start, stop) of n rows owned by `rank`: ceil(n / world) rows per rank, the last ranks
    may get fewer or none (probe split into G contiguous slices)."""
    per = (n + world - 1) // world
    start = min(n, rank * per)
    return start, min(n, start + per)


def hash32(keys) -> np.ndarray:
    """numpy mirror of the engine's key hash (csrc/fj_common.cuh: hash32): lowbias32 over lo ^ hi * odd."""
    k = np.asarray(keys).astype(np.uint64, copy=False)
    with np.errstate(over="ignore"):
        x = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ ((k >> np.uint64(32)).astype(np.uint32) * np.uint32(0x9E3779B1))
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return x


def shuffle_dest(keys, world: int, virtual: int = 1) -> np.ndarray:
    """Destination rank of every key in FJ_DIST_SHUFFLE (csrc/fj_radix.cu: scatter_digit, shift < 0): the low
    16 hash bits range-reduced to world * virtual destinations, `virtual` consecutive destinations per rank."""
    fan = world * virtual
    d = ((hash32(keys) & np.uint32(0xFFFF)).astype(np.uint64) * np.uint64(fan)) >> np.uint64(16)
    return (d // np.uint64(virtual)).astype(np.int64)


def init_from_env():
    """(rank, world, local_rank) from the torchrun environment; a single process gives (0, 1, 0)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def rendezvous_comm(dist=None) -> tuple[int, int]:
    """Create the engine's NCCL communicator for this process.  `dist` is an initialised torch.distributed
    module (any backend; gloo is enough) or None for a single process.  Returns (rank, world)."""
    from . import capi

    rank, world, local = init_from_env()
    capi.check(capi.lib().fj_init(local))
    if world == 1:
        capi.comm_init(0, 1)
        return 0, 1
    ident = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    capi.comm_init(rank, world, ident[0])
    return rank, world
