"""Tuner-loop throughput: objective evaluations per second with the dataset resident in HBM,
on 1 GPU and — one evaluation per GPU, as the finite-difference gradient of tune() deals them
(finite_difference.hpp:25-31) — on all GPUs of the box.

    python tools/tune_bench.py [tag] [--n 16384] [--evals 100] [--gpus 8]

One process drives every GPU: one handle and one host thread per device, features / targets uploaded once per
device, every evaluation = ab_gp_nll_d (Gram + factorisation + solve + reductions; nothing but the three
hyper-parameters crosses PCIe).  The objective values of the first candidates are compared with the compiled
reference on the host at a size it can run (N = 2048)."""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from albatross_b200 import capi  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def candidates(count, seed=0):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0.5, 2.0, count), rng.uniform(0.5, 1.5, count), rng.uniform(0.05, 0.3, count)], axis=1)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = args[0] if args and not args[0].isdigit() else "tune"
    n, evals = arg("--n", 16384), arg("--evals", 100)
    ngpus = min(arg("--gpus", capi.device_count()), capi.device_count())
    x = np.random.default_rng(0).uniform(0, 10, size=(n, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
    cands = candidates(evals)
    ops = [capi.SE, capi.NOISE, capi.SUM]
    out = {"n": n, "evaluations": evals, "objective": "negative log marginal likelihood (ab_gp_nll_d)"}
    handles = [capi.Handle(d) for d in range(ngpus)]
    data = [(h.upload_features(x), h.upload(y)) for h in handles]

    def run(lanes):
        values = np.zeros(evals)

        def worker(lane):
            h, (fd, yd) = handles[lane], data[lane]
            for i in range(lane, evals, lanes):
                l, s, sn = cands[i]
                values[i] = h.gp_nll_d(ops, [l, s, sn, 0.0, 0.0, 0.0], fd, yd)

        for lane in range(lanes):  # warm every device (first launch, workspace allocation)
            handles[lane].gp_nll_d(ops, [1.0, 1.0, 0.1, 0.0, 0.0, 0.0], *data[lane])
        t0 = time.perf_counter()
        threads = [threading.Thread(target=worker, args=(lane,)) for lane in range(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        return values, time.perf_counter() - t0

    v1, s1 = run(1)
    out["1_gpu"] = {"seconds": s1, "evaluations_per_second": evals / s1,
                    "TFLOPs": evals * n ** 3 / 3.0 / s1 * 1e-12}
    if ngpus > 1:
        vn, sn = run(ngpus)
        out[f"{ngpus}_gpus"] = {"seconds": sn, "evaluations_per_second": evals / sn, "speedup": s1 / sn,
                                "TFLOPs": evals * n ** 3 / 3.0 / sn * 1e-12,
                                "max_rel_diff_vs_1_gpu": float(np.max(np.abs(vn - v1) / np.abs(v1)))}
    # parity of the objective with the reference (and its host cost) at a size it can run
    try:
        from oracle.oracle import Ref
        if Ref.available():
            nc = 2048
            errs, t_ref = [], 0.0
            for i in range(3):
                l, s, sn_ = cands[i]
                got = handles[0].gp_nll(ops, [l, s, sn_, 0, 0, 0], x[:nc], y[:nc])
                t0 = time.perf_counter()
                want, _ = Ref.gp_nll(6, [l, s, sn_], x[:nc], y[:nc])
                t_ref += time.perf_counter() - t0
                errs.append(abs(got - want) / abs(want))
            out["reference_host"] = {"n": nc, "objective_rel_err_max": max(errs), "seconds_per_evaluation": t_ref / 3,
                                     "extrapolated_seconds_per_evaluation_at_n": t_ref / 3 * (n / nc) ** 3}
    except Exception as exc:
        out["reference_host"] = {"error": repr(exc)}
    os.makedirs("gpurun_out", exist_ok=True)
    line = json.dumps(out)
    print(line, flush=True)
    with open(f"gpurun_out/{tag}_tune_bench.json", "w") as fh:
        fh.write(line + "\n")


if __name__ == "__main__":
    main()
