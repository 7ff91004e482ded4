"""Per-config phase timings of every BASELINE.json config on one B200, with the
compiled reference timed beside each on a bounded sample.  Development / evidence tool: one JSON
line per config to stdout and to gpurun_out/<tag>_configs.jsonl.

    python tools/bench_configs.py [tag] [--only 1,2,3,4,5] [--no-cpu]

Device times are the library's own CUDA-event phase times (ab_timings) — gram / factor / solve /
reduce / predict — plus the wall time of the host-pointer call (h2d + d2h included)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from albatross_b200 import capi  # noqa: E402
from albatross_b200.capi import JOINT, MARGINAL, MEAN  # noqa: E402
from oracle.oracle import Ref, menu_program  # noqa: E402

HBM = 6535.4
if os.path.exists("MEASURED_PEAKS.json"):
    HBM = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
FP64_PEAK = 35.9  # cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r01_microbench_fp64_peaks.txt)


def wall(fn):
    t0 = time.perf_counter()
    out = fn()
    return out, (time.perf_counter() - t0) * 1e3


def emit(fh, line):
    s = json.dumps(line)
    print(s, flush=True)
    fh.write(s + "\n")
    fh.flush()


def bench_dataset(n, seed):
    """benchmarks/bench_utils.h:76-85: x ~ U[0,10], y = sin x + 0.1 cos 10x."""
    x = np.random.default_rng(seed).uniform(0.0, 10.0, size=n)
    return x, np.sin(x) + 0.1 * np.cos(10.0 * x)


def config1(h, cpu):
    """sinc_example: 1-D SE(3.5, 5.7) + IndependentNoise(1.0), N = 1000, marginal on a 161-pt grid."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-10.0, 23.0, size=1000)
    sinc = np.sinc((x - 3.0) / np.pi)
    y = np.sqrt(2.0) * x + 3.14159 + 10.0 * sinc + rng.normal(size=1000)
    t = np.linspace(-20.0, 33.0, 161)
    ops, pp = menu_program(6, [3.5, 5.7, 1.0])
    best = {}
    for _ in range(5):
        (f, info), ms_fit = wall(lambda: h.gp_fit(ops, pp, x, y))
        tf = h.timings()
        (mean, var, _), ms_pred = wall(lambda: h.gp_predict(f, ops, pp, x, info, t, MARGINAL))
        tp = h.timings()
        nll, ms_nll = wall(lambda: h.gp_nll(ops, pp, x, y))
        f.free()
        cur = {"fit_wall_ms": ms_fit, "predict_wall_ms": ms_pred, "nll_wall_ms": ms_nll,
               "fit_device_ms": tf["total_ms"], "predict_device_ms": tp["total_ms"]}
        for k, v in cur.items():
            best[k] = min(best.get(k, 1e30), v)
    line = {"config": "1 sinc_example N=1000 P=161", "device": best, "nll": nll}
    if cpu and Ref.available():
        (_, s_fit) = wall(lambda: Ref.gp_fit(6, [3.5, 5.7, 1.0], x, y))
        (pr, s_pred) = wall(lambda: Ref.gp_predict(6, [3.5, 5.7, 1.0], x, y, t, 1))
        ((rn, _), s_nll) = wall(lambda: Ref.gp_nll(6, [3.5, 5.7, 1.0], x, y))
        line["cpu_reference_ms"] = {"fit": s_fit, "fit+predict": s_pred, "nll": s_nll, "cores": 1}
        line["parity"] = {"mean_rel": float(np.max(np.abs(mean - pr[0])) / np.max(np.abs(pr[0]))),
                          "var_rel": float(np.max(np.abs(var - pr[1])) / np.max(np.abs(pr[1]))),
                          "nll_rel": abs(nll - rn) / abs(rn)}
    return line


def config2(h, cpu):
    """Gram build N = 32 768, 3-D, SE(2,1.5) + Matern52(3,0.7); full symmetric, lower only, cross."""
    n = 32768
    x = np.random.default_rng(0).uniform(0, 10, size=(n, 3))
    fd = h.upload_features(x)
    flush = h.alloc(8192, 8192)
    out = {}
    for name, cid, pp_, flags in (("se+m52 full", 7, [2.0, 1.5, 3.0, 0.7], 0),
                                  ("se+m52 lower", 7, [2.0, 1.5, 3.0, 0.7], 1),
                                  ("se+noise full", 6, [1.0, 1.0, 0.1], 0),
                                  ("se+noise lower", 6, [1.0, 1.0, 0.1], 1),
                                  ("const full (store pattern only)", 4, [1.3], 0)):
        ops, pp = menu_program(cid, pp_)
        best = 1e30
        for rep in range(5):
            flush.add_diag(np.full(8192, float(rep)))
            K = h.gram_sym_d(ops, pp, fd, flags=flags)
            ms = h.timings()["gram_ms"]
            K.free()
            if rep:
                best = min(best, ms)
        gb = (8.0 * n * n * (0.5 if flags else 1.0) + 8.0 * n * 3) * 1e-9
        out[name] = {"ms": best, "GB/s": gb / best * 1e3, "frac_hbm": gb / best * 1e3 / HBM}
    # cross Gram N x 512 (predict)
    tdev = h.upload_features(np.random.default_rng(1).uniform(0, 10, size=(512, 3)))
    ops, pp = menu_program(7, [2.0, 1.5, 3.0, 0.7])
    best = 1e30
    for rep in range(5):
        K = h.gram_cross_d(ops, pp, fd, tdev)
        best = min(best, h.timings()["gram_ms"])
        K.free()
    out["se+m52 cross 32768x512"] = {"ms": best, "GB/s": 8.0 * n * 512 * 1e-9 / best * 1e3}
    line = {"config": "2 Gram N=32768 D=3", "device": out, "hbm_peak": HBM}
    fd.free()
    tdev.free()
    flush.free()
    h.trim()
    if cpu and Ref.available():
        nc = 4096
        xc = x[:nc]
        cores = os.cpu_count() or 1
        _, s1 = wall(lambda: Ref.gram_sym(7, [2.0, 1.5, 3.0, 0.7], xc, nthreads=1))
        _, sp = wall(lambda: Ref.gram_sym(7, [2.0, 1.5, 3.0, 0.7], xc, nthreads=cores))
        gbc = 8.0 * nc * nc * 1e-9
        line["cpu_reference"] = {"sample": f"N={nc}", "serial_GB/s": gbc / s1 * 1e3,
                                 "threaded_GB/s": gbc / sp * 1e3, "cores": cores}
    return line


def config3(h, cpu, n=65536):
    """exact GP N = 65 536: gram, factor, solve, nll, predict P = 512 (mean / marginal / joint)."""
    x = np.random.default_rng(0).uniform(0, 10, size=(n, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
    t = np.random.default_rng(1).uniform(0, 10, size=(512, 3))
    ops, pp = menu_program(6, [1.0, 1.0, 0.1])
    h.gp_fit(ops, pp, x[:4096], y[:4096])[0].free()  # warm
    (f, info), ms_fit = wall(lambda: h.gp_fit(ops, pp, x, y))
    tf = h.timings()
    fl = n ** 3 / 3.0
    dev = {"fit_wall_ms": ms_fit, "gram_ms": tf["gram_ms"], "factor_ms": tf["factor_ms"],
           "solve_ms": tf["solve_ms"], "factor_TFLOPs": fl / tf["factor_ms"] * 1e-9,
           "factor_frac_fp64_peak": fl / tf["factor_ms"] * 1e-9 / FP64_PEAK,
           "gram_lower_GB/s": 4.0 * n * n / tf["gram_ms"] * 1e-6}
    for what, name in ((MEAN, "mean"), (MARGINAL, "marginal"), (JOINT, "joint")):
        best = 1e30
        for _ in range(2):
            _, ms = wall(lambda: h.gp_predict(f, ops, pp, x, info, t, what))
            best = min(best, h.timings()["total_ms"])
        dev[f"predict_{name}_device_ms"] = best
    # marginal: one TRSM N^2 P flops; joint the same + N P^2
    dev["predict_marginal_TFLOPs"] = n * n * 512.0 / dev["predict_marginal_device_ms"] * 1e-9
    f.free()
    h.trim()
    nll, ms_nll = wall(lambda: h.gp_nll(ops, pp, x, y))
    tn = h.timings()
    dev.update({"nll_wall_ms": ms_nll, "nll_reduce_ms": tn["reduce_ms"], "nll": nll})
    h.trim()
    line = {"config": f"3 exact GP N={n} D=3 P=512", "device": dev, "fp64_peak": FP64_PEAK}
    if cpu and Ref.available():
        nc = 4096
        _, s_fit = wall(lambda: Ref.gp_fit(6, [1.0, 1.0, 0.1], x[:nc], y[:nc]))
        _, s_pred = wall(lambda: Ref.gp_predict(6, [1.0, 1.0, 0.1], x[:nc], y[:nc], t, 2))
        line["cpu_reference"] = {"sample": f"N={nc}", "fit_ms": s_fit, "fit+joint_predict_ms": s_pred,
                                 "fit_TFLOPs": nc ** 3 / 3.0 / s_fit * 1e-9, "cores": 1,
                                 "extrapolated_fit_s_at_N": s_fit * 1e-3 * (n / nc) ** 3}
    return line


def config4(h, cpu, n=32768):
    """LOO-CV (bench_loo_cv shape): N = 32 768 1-D, groups int(f) % 8 and pure leave-one-out."""
    x, y = bench_dataset(n, 27)
    ops, pp = menu_program(6, [1.0, 1.0, 0.1])
    f, info = h.gp_fit(ops, pp, x, y)
    tf = h.timings()
    dev = {"fit_factor_ms": tf["factor_ms"]}
    keys8 = (x.astype(np.int64)) % 8
    _, off8, idx8 = capi.group_indexers(keys8)
    _, off1, idx1 = capi.group_indexers(np.arange(n, dtype=np.int64))
    for name, off, idx in (("grouped8", off8, idx8), ("loo", off1, idx1)):
        best, bw = 1e30, 1e30
        for _ in range(2):
            (mean, var, _, score), ms = wall(
                lambda: h.gp_cv(f, y, info, off, idx, MARGINAL, want_score=True))
            best = min(best, h.timings()["total_ms"])
            bw = min(bw, ms)
        dev[f"{name}_device_ms"] = best
        dev[f"{name}_wall_ms"] = bw
        dev[f"{name}_score"] = score
    # explicit triangular inverse: N^3/3 flops (+ SYRK per group)
    dev["loo_TFLOPs_(N^3/3)"] = n ** 3 / 3.0 / dev["loo_device_ms"] * 1e-9
    f.free()
    h.trim()
    line = {"config": f"4 LOO-CV N={n}", "device": dev}
    if cpu and Ref.available():
        nc = 2048
        xc, yc = bench_dataset(nc, 27)
        cores = os.cpu_count() or 1
        _, s8 = wall(lambda: Ref.gp_cv(6, [1.0, 1.0, 0.1], xc, yc, 1, 8.0, what=1, nthreads=cores))
        _, s1 = wall(lambda: Ref.gp_cv(6, [1.0, 1.0, 0.1], xc, yc, 0, 0.0, what=1, nthreads=cores))
        line["cpu_reference"] = {"sample": f"N={nc}", "grouped8_ms": s8, "loo_ms": s1, "cores": cores,
                                 "extrapolated_loo_s_at_N": s1 * 1e-3 * (n / nc) ** 3}
    return line


def config5(h, cpu, n=1 << 20, m=4096):
    """Sparse GP N = 2^20, M = 4096 uniformly spaced inducing points: FITC and PITC (1024-pt groups)."""
    x, y = bench_dataset(n, 0)
    u = np.linspace(x.min(), x.max(), m)
    t = np.linspace(0.0, 10.0, 512)
    ops, pp = menu_program(6, [1.0, 1.0, 0.1])
    dev = {}
    for name, keys in (("fitc", np.arange(n, dtype=np.int64)),
                       ("pitc1024", (x * (n / 10.0 / 1024.0)).astype(np.int64))):
        _, off, idx = capi.group_indexers(keys)
        best = None
        for _ in range(2):
            (f, info, ll), ms = wall(lambda: h.sparse_fit(ops, pp, x, y, u, off, idx))
            tf = h.timings()
            if best is None or tf["total_ms"] < best["fit_device_ms"]:
                best = {"fit_wall_ms": ms, "fit_device_ms": tf["total_ms"], "gram_ms": tf["gram_ms"],
                        "factor_ms": tf["factor_ms"], "solve_ms": tf["solve_ms"],
                        "h2d_ms": tf["h2d_ms"], "ll": ll, "groups": int(len(off) - 1)}
            (mean, var, _), msp = wall(lambda: f.predict(ops, pp, t, MARGINAL))
            best["predict_marginal_wall_ms"] = min(best.get("predict_marginal_wall_ms", 1e30), msp)
            f.free()
            h.trim()
        # CholQR2: 2 x (SYRK (N+M) M^2 + TRSM (N+M) M^2) + P = L_u^-1 K_uf (N M^2) flops
        best["TFLOPs_(5NM^2)"] = 5.0 * n * m * m / best["fit_device_ms"] * 1e-9
        dev[name] = best
    line = {"config": f"5 sparse GP N={n} M={m}", "device": dev}
    if cpu and Ref.available():
        nc, mc = 65536, 256
        xc, yc = bench_dataset(nc, 0)
        uc = np.linspace(xc.min(), xc.max(), mc)
        _, s = wall(lambda: Ref.sparse_gp(6, [1.0, 1.0, 0.1], xc, yc, uc, 0, 0.0))
        line["cpu_reference"] = {"sample": f"N={nc} M={mc} FITC", "fit_ms": s, "cores": 1,
                                 "extrapolated_fit_s_at_N_M": s * 1e-3 * (n / nc) * (m / mc) ** 2}
    return line


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = args[0] if args else "configs"
    only = None
    for i, a in enumerate(sys.argv):
        if a == "--only":
            only = set(int(v) for v in sys.argv[i + 1].split(","))
    cpu = "--no-cpu" not in sys.argv
    os.makedirs("gpurun_out", exist_ok=True)
    h = capi.Handle(0)
    with open(f"gpurun_out/{tag}_configs.jsonl", "w") as fh:
        for k, fn in ((1, config1), (2, config2), (3, config3), (4, config4), (5, config5)):
            if only is not None and k not in only:
                continue
            try:
                emit(fh, fn(h, cpu))
            except Exception as exc:  # keep going: the other configs are independent
                emit(fh, {"config": str(k), "error": repr(exc)})
                h.trim()


if __name__ == "__main__":
    main()
