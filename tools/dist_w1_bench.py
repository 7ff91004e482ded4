"""World-1 run of the distributed factorisation next to the single-GPU one (development aid): isolates what the
block-column algorithm itself costs (panel width = DMMA k-depth, per-step launches, packing) from communication.

    python tools/dist_w1_bench.py [n] [ENV=value+ENV=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from albatross_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    h = capi.Handle(0)
    h.dist_init(0, 1, bytes(128))
    x = np.random.default_rng(0).uniform(0, 10, size=(n, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
    ops, pp = capi.bench_program("se_noise")
    fl = n ** 3 / 3.0
    for _ in range(2):
        f, info = h.gp_fit(ops, pp, x, y)
        t = h.timings()
        f.free()
    print(f"single-GPU potrf      n={n}: factor {t['factor_ms']:.1f} ms = {fl / t['factor_ms'] * 1e-9:.2f} TFLOP/s, "
          f"solve {t['solve_ms']:.1f} ms", flush=True)
    h.trim()
    specs = sys.argv[2:] or ["AB_DIST_NB=1024", "AB_DIST_NB=512", "AB_DIST_NB=2048",
                             "AB_DIST_NB=1024+AB_DIST_SCHEDULE=lookahead1"]
    for spec in specs:
        for kv in spec.split("+"):
            k, v = kv.split("=")
            os.environ[k] = v
        best = None
        for _ in range(2):
            df, dinfo, nll = h.dist_gp_fit(ops, pp, x, y)
            t = h.timings()
            w, p, steps = h.dist_fit_breakdown()
            df.free()
            if best is None or t["factor_ms"] < best[0]:
                best = (t["factor_ms"], t["solve_ms"], w, p, steps)
        for kv in spec.split("+"):
            os.environ.pop(kv.split("=")[0], None)
        print(f"dist world=1 {spec:45s}: factor {best[0]:.1f} ms = {fl / best[0] * 1e-9:.2f} TFLOP/s, solve "
              f"{best[1]:.1f} ms, wait {best[2]:.1f} ms, panel chain {best[3]:.1f} ms, steps {best[4]}", flush=True)
        h.trim()
    err = float(np.max(np.abs(dinfo - info)) / np.max(np.abs(info)))
    print(f"information vs single GPU: {err:.2e}")


if __name__ == "__main__":
    main()
