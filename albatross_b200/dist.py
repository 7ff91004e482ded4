"""Host-side plumbing for one-process-per-GPU runs.

`torch.distributed` is only the launcher-side transport here (rendezvous, a 128-byte broadcast, the
max-over-ranks of timings); the data path runs inside ``libalbatross_b200.so`` on its own NCCL
communicator (``ab_dist_*`` in include/albatross_b200.h).  Everything in this module is integer /
index logic and works on CPU with the ``gloo`` backend, which is how tests/test_dist_host.py covers it.

Sharding rules (the integer contract of the multi-GPU path, all bit-exact):
  * block-column-cyclic factor: block column j lives on rank j % world at local index j // world;
  * Gram row blocks: rows are split into ceil(n / world) blocks rounded up to the 64-row Gram tile;
  * CV folds and sparse-GP observation groups: group g (std::map key order, as
    IndexerBuilder::build include/albatross/src/indexing/group_by.hpp:349-376 yields them) belongs
    to rank g % world; pure leave-one-out deals 2048-column chunks the same way.
"""
from __future__ import annotations

import numpy as np

GRAM_TILE = 64
LOO_CHUNK = 2048


def block_owner(block: int, world: int):
    """(rank, local index) of block column `block`."""
    return block % world, block // world


def local_blocks(nblocks: int, rank: int, world: int):
    return list(range(rank, nblocks, world))


def gram_row_block(n: int, rank: int, world: int):
    """(row0, rows) of this rank's Gram row block (mirrors ab_dist_gram_rows)."""
    per = -(-n // world)
    per = -(-per // GRAM_TILE) * GRAM_TILE
    row0 = min(n, per * rank)
    return row0, min(per, n - row0)


def shard_groups(offsets, indices, rank: int, world: int):
    """CSR indexer restricted to the groups g with g % world == rank.

    Returns (local_offsets, local_indices, group_ids): `local_indices` still refer to the ORIGINAL
    observation numbering."""
    offsets = np.asarray(offsets, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    gids = np.arange(rank, len(offsets) - 1, world, dtype=np.int64)
    loc_off = [0]
    loc_idx = []
    for g in gids:
        members = indices[offsets[g]:offsets[g + 1]]
        loc_idx.append(members)
        loc_off.append(loc_off[-1] + len(members))
    loc_idx = np.concatenate(loc_idx) if loc_idx else np.empty(0, dtype=np.int64)
    return np.asarray(loc_off, dtype=np.int64), loc_idx, gids


def shard_sparse_inputs(x, y, yvar, offsets, indices, rank: int, world: int):
    """Observations of this rank's groups for ab_sparse_fit on a distributed handle.

    Returns (x_local, y_local, yvar_local, local_offsets, local_indices) where local_indices index
    into the LOCAL arrays (groups stay contiguous and in key order)."""
    x = np.asarray(x, dtype=np.float64)
    x2 = x.reshape(len(x), -1)
    loc_off, members, _ = shard_groups(offsets, indices, rank, world)
    xl = np.ascontiguousarray(x2[members])
    yl = np.ascontiguousarray(np.asarray(y, dtype=np.float64)[members])
    vl = None if yvar is None else np.ascontiguousarray(np.asarray(yvar, dtype=np.float64)[members])
    return xl, yl, vl, loc_off, np.arange(len(members), dtype=np.int64)


def loo_chunks(n: int, rank: int, world: int):
    """Column ranges of the inverse diagonal this rank computes in sharded pure leave-one-out."""
    out = []
    for c, j0 in enumerate(range(0, n, LOO_CHUNK)):
        if c % world == rank:
            out.append((j0, min(LOO_CHUNK, n - j0)))
    return out


def exchange_unique_id(make_id):
    """Rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes.  Works on any
    initialised torch.distributed backend."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    if rank == 0:
        raw = bytes(make_id())
        assert len(raw) == 128
        t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def bootstrap(handle):
    """Initialises `handle`'s NCCL communicator from the ambient torch.distributed group."""
    import torch.distributed as dist

    from . import capi

    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        handle.dist_init(0, 1, bytes(128))
        return 0, 1
    uid = exchange_unique_id(capi.dist_unique_id)
    handle.dist_init(rank, world, uid)
    return rank, world


def max_over_ranks(value: float) -> float:
    """Timing reduction of the bench contract (device times, max over ranks)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
