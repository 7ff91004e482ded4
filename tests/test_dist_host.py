"""Host-side logic of the one-process-per-GPU path on CPU: world_size-2 `gloo` process groups.

What runs here without a GPU: the sharding rules (bit-exact integer contract of SURVEY.md §8e), the
unique-id exchange that bootstraps the library's NCCL communicator, the max-over-ranks timing
reduction, and that per-shard CV / sparse inputs assemble to the unsharded ones."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from albatross_b200 import capi
from albatross_b200 import dist as abd
from oracle.oracle import Restate, group_keys
from tests.helpers import features, prog, targets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = abd.exchange_unique_id(capi.dist_unique_id)  # ncclGetUniqueId works without a GPU
        t = abd.max_over_ranks(10.0 + rank)
        # group-sharded leave-one-group-out, emulated on the host with the oracle per shard
        ops, pp = prog(6)
        x = features(120, 1, 5).ravel()
        y = targets(x)
        keys = group_keys(x, 1, 5.0)
        _, offsets, indices = capi.group_indexers(keys)
        loc_off, loc_idx, gids = abd.shard_groups(offsets, indices, rank, world)
        mean, var, _, _ = Restate.gp_cv(ops, pp, x, y, keys, what=1)
        part = np.zeros(len(x))
        part[loc_idx] = mean[loc_idx]
        tt = torch.from_numpy(part.copy())
        dist.all_reduce(tt)
        q.put((rank, uid, t, gids.tolist(), float(np.max(np.abs(tt.numpy() - mean)))))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bootstrap_and_shard_assembly():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, uid0, t0, g0, e0), (r1, uid1, t1, g1, e1) = out
    assert uid0 == uid1 and len(uid0) == capi.DIST_ID_BYTES
    assert t0 == t1 == 11.0
    assert sorted(g0 + g1) == list(range(len(g0) + len(g1))) and not set(g0) & set(g1)
    assert e0 == 0.0 and e1 == 0.0  # disjoint shards: the sum over ranks is exact


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_block_cyclic_owner_contract(world):
    nblk = 37
    seen = []
    for r in range(world):
        for lb, j in enumerate(abd.local_blocks(nblk, r, world)):
            assert abd.block_owner(j, world) == (r, lb)
            assert capi.dist_block_owner(j, world) == (r, lb)  # the C ABI agrees
            seen.append(j)
    assert sorted(seen) == list(range(nblk))


@pytest.mark.parametrize("n,world", [(1, 2), (63, 2), (1000, 8), (32768, 8), (65, 4), (0, 2)])
def test_gram_row_blocks_cover(n, world):
    rows = [abd.gram_row_block(n, r, world) for r in range(world)]
    covered = 0
    for r0, nr in rows:
        assert nr >= 0 and (nr == 0 or r0 == covered)
        assert r0 % abd.GRAM_TILE == 0 or nr == 0
        covered += nr
    assert covered == n


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sparse_shards_partition_observations(world):
    x = features(501, 1, 2).ravel()
    y = targets(x)
    keys = group_keys(x, 2, 3.0)
    _, offsets, indices = capi.group_indexers(keys)
    total = 0
    seen = []
    for r in range(world):
        xl, yl, vl, lo, li = abd.shard_sparse_inputs(x, y, None, offsets, indices, r, world)
        assert vl is None and lo[0] == 0 and lo[-1] == len(xl) == len(yl)
        assert np.array_equal(li, np.arange(len(xl)))
        _, members, gids = abd.shard_groups(offsets, indices, r, world)
        assert np.array_equal(xl.ravel(), x[members]) and np.array_equal(yl, y[members])
        # groups stay whole and in key order
        for k, g in enumerate(gids):
            assert np.array_equal(members[lo[k]:lo[k + 1]], indices[offsets[g]:offsets[g + 1]])
        seen.append(members)
        total += len(xl)
    assert total == len(x)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(len(x)))


def test_loo_chunks_cover():
    n = 5000
    got = sorted(c for r in range(3) for c in abd.loo_chunks(n, r, 3))
    assert got[0][0] == 0 and sum(c[1] for c in got) == n
    for (a, la), (b, _) in zip(got, got[1:]):
        assert a + la == b


# ---- the update schedule of the distributed factorisation (dist.cu, pipelined loop) -------------------------

@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("nblk", [1, 2, 3, 5, 8, 9, 16, 33])
@pytest.mark.parametrize("pairing,last_full", [(True, True), (True, False), (False, True)])
def test_panel_schedule_contract(world, nblk, pairing, last_full):
    """Integer contract of the panel pipeline, on the host-side model of the loop in dist.cu:
    (1) every owned block column receives every earlier panel exactly once;
    (2) the panel stream takes a column over only behind the update stream's last launch on it — the event it
        waits for is the one that launch records, and it is recorded earlier in enqueue order;
    (3) on the panel stream a column's panels arrive in increasing order and its factorisation follows the last;
    (4) no launch uses a panel before the step in which it arrives;
    (5) with at least four packed-panel buffers (two pair buffers) a buffer is overwritten only after the
        update-stream launch that read its previous panel has been enqueued."""
    for rank in range(world):
        ops = abd.panel_schedule(nblk, world, rank, last_full=last_full, pairing=pairing)
        owned = list(range(rank, nblk, world))
        seen = {}
        first_ps, last_s, factor_pos, recorded_at = {}, {}, {}, {}
        for pos, op in enumerate(ops):
            for e in op.get("records", []):
                recorded_at[e] = (pos, op["iter"])
            if "factor" in op:
                factor_pos[op["factor"]] = pos
                continue
            for j in op["cols"]:
                assert j in owned and j < nblk
                for k in op["panels"]:
                    assert k < j, (j, k)
                    assert op["iter"] >= k                                    # (4)
                    seen[(j, k)] = seen.get((j, k), 0) + 1
                if op["stream"] == "PS":
                    first_ps.setdefault(j, pos)
                else:
                    last_s[j] = pos
        assert seen == {(j, k): 1 for j in owned for k in range(j)}, (world, nblk, rank)   # (1)
        for j in owned:
            ps_ops = [(pos, op) for pos, op in enumerate(ops)
                      if op["stream"] == "PS" and j in op.get("cols", [])]
            ks = [op["panels"][0] for _, op in ps_ops]
            assert ks == sorted(ks) and (not ks or ks[-1] == j - 1)            # (3)
            if j > 0:
                assert factor_pos[j] > ps_ops[-1][0]
            if j in last_s:                                                    # (2)
                pos0, op0 = ps_ops[0]
                assert last_s[j] < pos0
                gate = [w for w in op0["waits"] if not w.startswith("arrived")]
                assert len(gate) == 1 and gate[0] in ops[last_s[j]]["records"], (j, gate, ops[last_s[j]])
                assert recorded_at[gate[0]][0] < pos0
        nbuf = 4
        for k in range(nbuf, nblk):                                            # (5)
            prev = k - nbuf
            # factor_and_pack_p(k) is enqueued in the panel-stream part of step k - 1, ahead of that step's S part
            assert recorded_at[f"bulkdone[{prev}]"][1] <= k - 2
            assert recorded_at[f"pdone[{prev}]"][1] <= k - 2
