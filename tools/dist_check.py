"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py [n]

Checks ab_dist_gp_fit (block-column-cyclic Cholesky + NCCL panel broadcasts) against the single-GPU
path and, at small n, the oracle; sharded Gram rows; sharded CV; sharded sparse GP."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from albatross_b200 import capi, dist as abd  # noqa: E402
from albatross_b200.capi import MARGINAL  # noqa: E402
from oracle.oracle import Restate, group_keys, menu_program  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def main():
    n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = capi.Handle(local)
    rank, world = abd.bootstrap(h)
    ok = True
    ops, pp = menu_program(8, [2.0, 1.5, 3.0, 0.7, 0.1])
    rng = np.random.default_rng(0)

    def report(name, err, tol):
        nonlocal ok
        good = err <= tol
        ok = ok and good
        if rank == 0:
            print(f"[dist_check world={world}] {name}: err={err:.3e} tol={tol:.1e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)

    # 1. distributed fit vs oracle (small) and vs the single-GPU path (large)
    for n, nb in ((1500, 128), (3000, 256)):
        x = rng.uniform(0, 10, size=(n, 3))
        y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
        f, info, nll = h.dist_gp_fit(ops, pp, x, y, nb=nb)
        want = Restate.gp_fit(ops, pp, x, y)["information"]
        want_nll = Restate.gp_nll(ops, pp, x, y)
        report(f"fit n={n} nb={nb} information vs oracle", rel(info, want), 1e-9)
        report(f"fit n={n} nb={nb} nll vs oracle", abs(nll - want_nll) / abs(want_nll), 1e-9)
        f.free()
    x = rng.uniform(0, 10, size=(n_big, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
    f, info, nll = h.dist_gp_fit(ops, pp, x, y)
    t = h.timings()
    f1, info1 = h.gp_fit(ops, pp, x, y)
    nll1 = h.gp_nll(ops, pp, x, y)
    report(f"fit n={n_big} information vs single GPU", rel(info, info1), 1e-9)
    report(f"fit n={n_big} nll vs single GPU", abs(nll - nll1) / abs(nll1), 1e-10)
    if rank == 0:
        tf = n_big ** 3 / 3.0 / (t["factor_ms"] * 1e-3) * 1e-12
        print(f"[dist_check] n={n_big}: gram {t['gram_ms']:.1f} ms factor {t['factor_ms']:.1f} ms "
              f"({tf:.1f} TFLOP/s aggregate) solve {t['solve_ms']:.1f} ms", flush=True)
    f.free()

    # 2. sharded Gram rows
    r0, K = h.dist_gram_rows(ops, pp, x[:2000])
    want = Restate.gram_sym(ops, pp, x[:2000])
    rows = K.shape[0]
    report("gram rows", rel(K.download(), want[r0:r0 + rows]) if rows else 0.0, 1e-14)
    K.free()

    # 3. sharded CV (groups and pure LOO)
    xc = rng.uniform(0, 10, size=3000)
    yc = np.sin(xc) + 0.1 * np.cos(10 * xc)
    ops6, pp6 = menu_program(6, [1.0, 1.0, 0.1])
    fc, infoc = h.gp_fit(ops6, pp6, xc, yc)
    for name, keys in (("int(f)%8", group_keys(xc, 1, 8.0)), ("loo", np.arange(len(xc)))):
        _, offsets, indices = capi.group_indexers(keys)
        m1, v1, _, s1 = h.gp_cv(fc, yc, infoc, offsets, indices, MARGINAL, want_score=True)
        m2, v2, s2 = h.dist_gp_cv(fc, yc, infoc, offsets, indices, MARGINAL, want_score=True)
        report(f"cv {name} mean", rel(m2, m1), 1e-9)
        report(f"cv {name} var", rel(v2, v1), 1e-9)
        report(f"cv {name} score", abs(s2 - s1) / abs(s1), 1e-9)
    fc.free()

    # 4. sharded sparse GP: each rank passes its groups; result replicated
    xs = rng.uniform(0, 10, size=4000)
    ys = np.sin(xs) + 0.1 * np.cos(10 * xs)
    u = np.linspace(xs.min(), xs.max(), 96)
    tt = np.linspace(0.1, 9.9, 41)
    for name, keys in (("pitc", group_keys(xs, 2, 2.0)), ("fitc", np.arange(len(xs)))):
        _, offsets, indices = capi.group_indexers(keys)
        want = Restate.sparse_gp(ops6, pp6, xs, ys, u, keys, test=tt, what=1, want_ll=True)
        xl, yl, vl, lo, li = abd.shard_sparse_inputs(xs, ys, None, offsets, indices, rank, world)
        sf, v, ll = h.sparse_fit(ops6, pp6, xl, yl, u, lo, li)
        mean, var, _ = sf.predict(ops6, pp6, tt, MARGINAL)
        report(f"sparse {name} mean", rel(mean, want["mean"]), 1e-9)
        report(f"sparse {name} var", rel(var, want["var"]), 1e-9)
        report(f"sparse {name} ll", abs(ll - want["ll"]) / abs(want["ll"]), 1e-9)
        sf.free()

    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    h.dist_finalize()
    h.close()
    dist.destroy_process_group()
    if rank == 0:
        print("[dist_check] " + ("ALL OK" if flag.item() == 0 else "FAILED"), flush=True)
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
