"""Times the pieces of one panel-chain step in isolation on an idle GPU (development aid): the factorisation of
a diagonal block (ab_potrf at n = 64 .. 2048: the dependent chain of leaf kernels and small GEMMs) and the
GEMM-shaped pieces (column update, TRSM-sized product) at the row counts of the chain-bound tail.

    python tools/chain_bench.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from albatross_b200 import capi  # noqa: E402


def main():
    h = capi.Handle(0)
    rng = np.random.default_rng(0)
    for n in (64, 128, 256, 512, 1024, 2048):
        x = rng.uniform(0, 10, size=(n, 3))
        ops, pp = capi.bench_program("se_noise")
        best = 1e30
        for _ in range(5):
            K = h.gram_sym(ops, pp, x)
            f = h.potrf(K)
            best = min(best, h.timings()["factor_ms"])
            f.free()
        print(f"potrf n={n:5d}: {best * 1e3:8.1f} us", flush=True)
    for m, n, k in ((30000, 512, 512), (60000, 512, 512), (120000, 512, 512), (30000, 64, 64), (30000, 256, 256),
                    (30000, 512, 1024)):
        A = h.alloc(m, k)
        B = h.alloc(n, k)
        Cm = h.alloc(m, n)
        best = 1e30
        for _ in range(5):
            h.gemm(A, B, Cm, alpha=-1.0, beta=1.0, trans_b=True)
            best = min(best, h.timings()["factor_ms"])
        print(f"gemm NT m={m} n={n} k={k}: {best * 1e3:8.1f} us = {2.0 * m * n * k / best * 1e-9:.1f} TFLOP/s", flush=True)
        A.free(); B.free(); Cm.free()


if __name__ == "__main__":
    main()
