"""Multi-GPU legs of BASELINE.json's configs, run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29513 tools/bench_configs_dist.py [tag] [--only 3,4,5] [--n3 131072]

  config 3  exact GP, one matrix block-column-cyclic over the ranks (ab_dist_gp_fit): phase times and the
            per-rank breakdown of the factorisation (waiting for a panel / panel chain / DMMA updates)
  config 4  LOO-CV N = 32 768: rank 0 fits, L is replicated by ONE ncclBroadcast (ab_dist_factor_broadcast),
            folds / inverse-diagonal chunks are sharded (ab_dist_gp_cv); checked against the 1-GPU ab_gp_cv
  config 5  sparse GP N = 2^20, M = 4096: observation groups sharded over the ranks (ab_sparse_fit on a
            distributed handle); the log-likelihood must equal the 1-GPU value

One JSON line per config to stdout (rank 0) and gpurun_out/<tag>_configs_dist.jsonl."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from albatross_b200 import capi, dist as abd  # noqa: E402
from albatross_b200.capi import MARGINAL  # noqa: E402

FP64_PEAK = 35.5  # cuBLAS DGEMM 8192^3 on this pool's B200 (bench.py measures it live)


def bench_dataset(n, seed):
    x = np.random.default_rng(seed).uniform(0.0, 10.0, size=n)
    return x, np.sin(x) + 0.1 * np.cos(10.0 * x)


def gather(value):
    """list over ranks of a python float."""
    world = dist.get_world_size()
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[dist.get_rank()] = value
    dist.all_reduce(t)
    return [float(v) for v in t.cpu()]


def config3_schedules(h, rank, world, n, schedules):
    """The same fit under different schedules of the distributed factorisation (environment switches of
    dist.cu, read at every call): one line per schedule."""
    out = []
    for spec in schedules:
        for kv in spec.split("+"):
            k, v = kv.split("=")
            os.environ[k] = v
        line = config3(h, rank, world, n)
        line["schedule"] = spec
        out.append(line)
        for kv in spec.split("+"):
            os.environ.pop(kv.split("=")[0], None)
    return out


def config3(h, rank, world, n):
    x = np.random.default_rng(0).uniform(0, 10, size=(n, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
    ops, pp = capi.bench_program("se_noise")
    best = None
    for rep in range(2):
        dist.barrier()
        t0 = time.perf_counter()
        f, info, nll = h.dist_gp_fit(ops, pp, x, y)
        wall = (time.perf_counter() - t0) * 1e3
        t = h.timings()
        wait_ms, panel_ms, steps = h.dist_fit_breakdown()
        f.free()
        cur = {"wall_ms": abd.max_over_ranks(wall), "gram_ms": abd.max_over_ranks(t["gram_ms"]),
               "factor_ms": abd.max_over_ranks(t["factor_ms"]), "solve_ms": abd.max_over_ranks(t["solve_ms"]),
               "per_rank_factor_ms": gather(t["factor_ms"]), "per_rank_wait_for_panel_ms": gather(wait_ms),
               "per_rank_panel_chain_ms": gather(panel_ms), "steps": steps, "nll": nll}
        if best is None or cur["factor_ms"] < best["factor_ms"]:
            best = cur
    fl = n ** 3 / 3.0
    best["factor_TFLOPs_aggregate"] = fl / best["factor_ms"] * 1e-9
    best["factor_frac_of_world_x_cublas"] = best["factor_TFLOPs_aggregate"] / (world * FP64_PEAK)
    best["fit_TFLOPs_aggregate_wall"] = fl / best["wall_ms"] * 1e-9
    return {"config": f"3 exact GP N={n} block-column-cyclic 1x{world}", "device": best}


def config4(h, rank, world, n=32768):
    x, y = bench_dataset(n, 27)
    ops, pp = capi.bench_program("se_noise")
    dev = {}
    f = info = None
    if rank == 0:
        f, info = h.gp_fit(ops, pp, x, y)
        dev["fit_factor_ms_rank0"] = h.timings()["factor_ms"]
    info_t = torch.zeros(n, dtype=torch.float64, device="cuda")
    if rank == 0:
        info_t.copy_(torch.from_numpy(info))
    dist.broadcast(info_t, src=0)
    info = info_t.cpu().numpy()
    dist.barrier()
    t0 = time.perf_counter()
    f = h.dist_factor_broadcast(f, root=0)
    dev["L_broadcast_wall_ms"] = abd.max_over_ranks((time.perf_counter() - t0) * 1e3)
    dev["L_broadcast_device_ms"] = abd.max_over_ranks(h.timings()["h2d_ms"])
    dev["L_bytes"] = 8.0 * n * (n + 16)
    keys8 = x.astype(np.int64) % 8
    _, off8, idx8 = capi.group_indexers(keys8)
    _, off1, idx1 = capi.group_indexers(np.arange(n, dtype=np.int64))
    for name, off, idx in (("grouped8", off8, idx8), ("loo", off1, idx1)):
        best = 1e30
        for _ in range(2):
            dist.barrier()
            t0 = time.perf_counter()
            mean, var, score = h.dist_gp_cv(f, y, info, off, idx, MARGINAL, want_score=True)
            best = min(best, abd.max_over_ranks((time.perf_counter() - t0) * 1e3))
        dev[f"{name}_wall_ms"] = best
        dev[f"{name}_score"] = score
        if rank == 0:  # parity with the single-GPU path on the same factor
            m1, v1, _, s1 = h.gp_cv(f, y, info, off, idx, MARGINAL, want_score=True)
            dev[f"{name}_vs_1gpu"] = {
                "mean_rel": float(np.max(np.abs(mean - m1)) / np.max(np.abs(m1))),
                "var_rel": float(np.max(np.abs(var - v1)) / np.max(np.abs(v1))),
                "score_rel": abs(score - s1) / abs(s1), "one_gpu_wall_ms": None}
            t0 = time.perf_counter()
            h.gp_cv(f, y, info, off, idx, MARGINAL, want_score=True)
            dev[f"{name}_vs_1gpu"]["one_gpu_wall_ms"] = (time.perf_counter() - t0) * 1e3
        dist.barrier()
    f.free()
    h.trim()
    return {"config": f"4 LOO-CV N={n}, L replicated by one broadcast, folds over {world} GPUs", "device": dev}


def config5(h, rank, world, n=1 << 20, m=4096):
    x, y = bench_dataset(n, 0)
    u = np.linspace(x.min(), x.max(), m)
    t = np.linspace(0.0, 10.0, 512)
    ops, pp = capi.bench_program("se_noise")
    dev = {}
    for name, keys in (("fitc", np.arange(n, dtype=np.int64)),
                       ("pitc1024", (x * (n / 10.0 / 1024.0)).astype(np.int64))):
        _, off, idx = capi.group_indexers(keys)
        xl, yl, vl, lo, li = abd.shard_sparse_inputs(x, y, None, off, idx, rank, world)
        best = None
        for _ in range(2):
            dist.barrier()
            t0 = time.perf_counter()
            f, info, ll = h.sparse_fit(ops, pp, xl, yl, u, lo, li)
            wall = abd.max_over_ranks((time.perf_counter() - t0) * 1e3)
            tf = h.timings()
            cur = {"fit_wall_ms": wall, "fit_device_ms": abd.max_over_ranks(tf["total_ms"]),
                   "factor_ms": abd.max_over_ranks(tf["factor_ms"]), "ll": ll,
                   "local_observations": gather(float(len(xl)))}
            mean, var, _ = f.predict(ops, pp, t, MARGINAL)
            cur["mean_checksum"] = float(np.sum(mean))
            f.free()
            h.trim()
            if best is None or cur["fit_device_ms"] < best["fit_device_ms"]:
                best = cur
        dev[name] = best
    return {"config": f"5 sparse GP N={n} M={m}, groups over {world} GPUs", "device": dev}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = args[0] if args else "dist"
    only, n3 = None, 131072
    for i, a in enumerate(sys.argv):
        if a == "--only":
            only = set(int(v) for v in sys.argv[i + 1].split(","))
        if a == "--n3":
            n3 = int(sys.argv[i + 1])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = capi.Handle(local)
    rank, world = abd.bootstrap(h)
    os.makedirs("gpurun_out", exist_ok=True)
    fh = open(f"gpurun_out/{tag}_configs_dist.jsonl", "a") if rank == 0 else None
    for i, a_ in enumerate(sys.argv):
        if a_ == "--schedules":  # e.g. AB_DIST_SCHEDULE=lookahead1,AB_DIST_NBUF=4,AB_DIST_NBUF=8+AB_DIST_PCOL=0
            for line in config3_schedules(h, rank, world, n3, sys.argv[i + 1].split(",")):
                line["n_gpus"] = world
                if rank == 0:
                    dev = line["device"]
                    brief = {k: dev[k] for k in ("factor_ms", "solve_ms", "factor_TFLOPs_aggregate")}
                    brief["wait_ms_max"] = max(dev["per_rank_wait_for_panel_ms"])
                    print(json.dumps({"schedule": line["schedule"], **brief}), flush=True)
                    fh.write(json.dumps(line) + "\n")
                    fh.flush()
            only = set()
    for k, fn in ((3, lambda: config3(h, rank, world, n3)), (4, lambda: config4(h, rank, world)),
                  (5, lambda: config5(h, rank, world))):
        if only is not None and k not in only:
            continue
        line = fn()
        line["n_gpus"] = world
        if rank == 0:
            s = json.dumps(line)
            print(s, flush=True)
            fh.write(s + "\n")
            fh.flush()
    h.dist_finalize()
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
