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


def panel_schedule(nblk: int, world: int, rank: int, last_full: bool = True, pairing: bool = True):
    """Host-side model of the update schedule of the distributed factorisation: which stream applies
    which packed panel to which block column of this rank, in enqueue order, and which event each
    operation waits for / records.  Mirrors the pipelined loop of dist_fit_impl
    (albatross_b200/csrc/dist.cu): the panel stream PS applies every panel k to the first column the
    rank owns after k and factors a column once panel (column - 1) is in; the update stream S applies
    the panels to the columns beyond that one — one panel per launch, or, with `pairing`, panels
    2q and 2q + 1 in one launch of depth 2 nb at the odd step, plus one single-column launch at the
    even step for the column PS takes over in between.  `last_full`: the last block column has the
    full width (a ragged last panel is never paired).

    Returns a list of dicts {stream, iter, panels, cols, waits, records} (and {stream: 'PS',
    iter, factor: j} entries); tests/test_dist_host.py checks the integer contract on it: every
    (column, earlier panel) pair exactly once, take-over of a column only behind S's last update of it,
    buffers reused only behind their readers."""
    W, me = world, rank
    owned = list(range(me, nblk, W))

    def full(k):
        return k < nblk - 1 or last_full

    def paired(k):
        head = k & ~1
        return pairing and head + 1 < nblk and full(head + 1)

    def first_owned_after(k):
        return k + 1 + ((me - (k + 1)) % W + W) % W

    def cols_from(j0):
        return [j for j in owned if j >= j0]

    ops = []
    for k in range(nblk):
        jstar = first_owned_after(k)
        p_takes = jstar < nblk
        if p_takes:
            waits = [f"arrived[{k}]"]
            handover = jstar - W - 1
            if k == max(jstar - W, 0) and handover >= 0:
                head = paired(handover) and handover % 2 == 0
                waits.append(f"{'coldone' if head else 'bulkdone'}[{handover}]")
            ops.append(dict(stream="PS", iter=k, panels=(k,), cols=[jstar], waits=waits, records=[]))
        ops.append(dict(stream="PS", iter=k, panels=(), cols=[], waits=[], records=[f"pdone[{k}]"]))
        if jstar == k + 1 and jstar < nblk:
            ops.append(dict(stream="PS", iter=k, factor=jstar))
        if not paired(k):
            ops.append(dict(stream="S", iter=k, panels=(k,), cols=cols_from(jstar + W if p_takes else jstar),
                            waits=[f"arrived[{k}]"], records=[f"bulkdone[{k}]"]))
        elif k % 2 == 0:
            cols = [jstar + W] if jstar == k + 1 and jstar + W < nblk else []
            ops.append(dict(stream="S", iter=k, panels=(k,), cols=cols, waits=[f"arrived[{k}]"],
                            records=[f"coldone[{k}]"]))
        else:
            ops.append(dict(stream="S", iter=k, panels=(k - 1, k), cols=cols_from(jstar + W),
                            waits=[f"arrived[{k}]"], records=[f"bulkdone[{k - 1}]", f"bulkdone[{k}]"]))
    return ops
