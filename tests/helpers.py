"""Shared test data: the reference's covariance menu (oracle/ref_shim/ref_common.h) and generators."""
import os

import numpy as np

from oracle.oracle import menu_program

PARAMS = {0: [2.0, 1.5], 1: [2.0, 1.5], 2: [3.0, 0.7], 3: [3.0, 0.7], 4: [1.3], 5: [0.2],
          6: [1.0, 1.0, 0.1], 7: [2.0, 1.5, 3.0, 0.7], 8: [2.0, 1.5, 3.0, 0.7, 0.1],
          9: [2.0, 1.5, 3.0, 0.7, 1.1, 0.9, 1.2, 0.3]}
GP_COVS = (6, 8, 9)


def prog(cov_id):
    return menu_program(cov_id, PARAMS[cov_id])


def features(n, dim, seed=0):
    """U[0,10]^dim features, the shape of benchmarks/bench_utils.h:25-35."""
    return np.random.default_rng(seed).uniform(0.0, 10.0, size=(n, dim))


def targets(x):
    """benchmarks/bench_utils.h:76-85: y = sin x + 0.1 cos 10x on the first coordinate."""
    x0 = np.asarray(x).reshape(len(x), -1)[:, 0]
    return np.sin(x0) + 0.1 * np.cos(10.0 * x0)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / scale) if b.size else 0.0


def assert_close(a, b, rtol, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = rel_err(a, b)
    log = os.environ.get("AB_ERR_LOG")  # achieved errors, recorded for profiles/ (one line per check)
    if log:
        with open(log, "a") as fh:
            fh.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '?').split(' ')[0]}\t{what}\t{err:.3e}\t{rtol:.1e}\n")
    assert err <= rtol, f"{what}: relative error {err:.3e} > {rtol:.1e}"
