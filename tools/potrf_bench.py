"""Times Gram (lower) + blocked Cholesky + solve at one size (development aid):
python tools/potrf_bench.py [n] [reps]."""
import sys

import numpy as np

sys.path.insert(0, ".")
from albatross_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    h = capi.Handle(0)
    ops, pp = capi.bench_program("se_noise")
    x = np.random.default_rng(0).uniform(0, 10, size=(n, 3))
    y = np.sin(x[:, 0])
    best = {}
    for rep in range(reps):
        f, info = h.gp_fit(ops, pp, x, y)
        t = h.timings()
        f.free()
        for k in ("gram_ms", "factor_ms", "solve_ms", "total_ms"):
            best[k] = min(best.get(k, 1e30), t[k])
    fl = n ** 3 / 3.0
    print(f"fit n={n}: factor {best['factor_ms']:.2f} ms = {fl / best['factor_ms'] * 1e-9:.2f} TFLOP/s, "
          f"gram {best['gram_ms']:.2f} ms, solve {best['solve_ms']:.2f} ms, total {best['total_ms']:.2f} ms, "
          f"info[0]={info[0]:.15g}")


if __name__ == "__main__":
    main()
