"""Quick device timings of the individual phases (development aid, not the bench contract)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from albatross_b200 import capi  # noqa: E402
from oracle.oracle import menu_program  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768]
    h = capi.Handle(0)
    rng = np.random.default_rng(0)
    ops7, p7 = menu_program(7, [2.0, 1.5, 3.0, 0.7])
    ops6, p6 = menu_program(6, [1.0, 1.0, 0.1])
    for n in sizes:
        x = rng.uniform(0, 10, size=(n, 3))
        y = np.sin(x[:, 0]) + 0.1 * np.cos(10 * x[:, 0])
        fd = h.upload_features(x)
        yd = h.upload(y)
        for rep in range(3):
            K = h.gram_sym_d(ops7, p7, fd)
            t = h.timings()
            K.free()
        gbs = (8.0 * n * n + 8.0 * n * 3) / (t["gram_ms"] * 1e-3) * 1e-9
        print(f"n={n} gram SE+M52 full: {t['gram_ms']:.3f} ms  {gbs:.1f} GB/s", flush=True)
        for rep in range(2):
            t0 = time.time()
            f, info = h.gp_fit_d(ops6, p6, fd, yd)
            t = h.timings()
            wall = time.time() - t0
            f.free()
            info.free()
        tf = n ** 3 / 3.0 / (t["factor_ms"] * 1e-3) * 1e-12
        print(f"n={n} fit: gram {t['gram_ms']:.2f} ms factor {t['factor_ms']:.2f} ms "
              f"({tf:.2f} TFLOP/s) solve {t['solve_ms']:.2f} ms total {t['total_ms']:.2f} ms "
              f"wall {wall*1e3:.1f} ms launches {t['kernel_launches']}", flush=True)
        nll = h.gp_nll_d(ops6, p6, fd, yd)
        t = h.timings()
        print(f"n={n} nll={nll:.6f}: gram {t['gram_ms']:.2f} factor {t['factor_ms']:.2f} "
              f"solve {t['solve_ms']:.2f} reduce {t['reduce_ms']:.2f} total {t['total_ms']:.2f} ms",
              flush=True)
        fd.free()
        yd.free()
        h.trim()


if __name__ == "__main__":
    main()
