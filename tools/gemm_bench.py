"""Times the DMMA GEMM kernel in isolation (development aid)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from albatross_b200 import capi  # noqa: E402


def main():
    h = capi.Handle(0)
    shapes = [(8192, 8192, 8192, False, True, False), (8192, 8192, 8192, False, False, False),
              (8192, 8192, 8192, True, False, False), (16384, 16384, 1024, False, True, False),
              (16384, 16384, 1024, False, True, True), (16384, 16384, 512, False, True, True),
              (65536, 8192, 512, False, True, False), (16384, 16384, 256, False, True, True),
              (16384, 16384, 64, False, True, True), (4096, 4096, 4096, False, True, True),
              (2048, 2048, 2048, False, True, True), (1024, 1024, 1024, False, True, True),
              (512, 512, 512, False, True, True), (256, 256, 256, False, True, True),
              (128, 128, 128, False, True, True), (64, 64, 64, False, True, False),
              (16384, 64, 64, False, True, False), (16384, 128, 128, False, True, False),
              (16384, 512, 512, False, True, False), (16384, 1, 16384, False, False, False),
              (64, 1, 64, False, False, False)]
    if len(sys.argv) > 1:
        shapes = shapes[: int(sys.argv[1])]
    for m, n, k, ta, tb, lower in shapes:
        A = h.alloc(k if ta else m, m if ta else k)
        B = h.alloc(n if tb else k, k if tb else n)
        Cm = h.alloc(m, n)
        best = 1e30
        for rep in range(4):
            h.gemm(A, B, Cm, alpha=-1.0, beta=1.0, trans_a=ta, trans_b=tb, lower=lower)
            best = min(best, h.timings()["factor_ms"])
        flops = 2.0 * m * n * k * (0.5 if lower else 1.0)
        print(f"gemm m={m} n={n} k={k} ta={int(ta)} tb={int(tb)} lower={int(lower)}: "
              f"{best:.4f} ms  {flops / best * 1e-9:.2f} TFLOP/s", flush=True)
        A.free(); B.free(); Cm.free()
        h.trim()


if __name__ == "__main__":
    main()
