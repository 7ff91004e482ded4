"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run
them: the reference's LDLT needs 6.4 h and 96 GiB at N = 65 536).  Everything goes through the C ABI.

  configs[1]  Gram build, N = 32 768, 3-D, SE + Matern52: sampled rows against the oracle, bit-identical
              mirror, lower-only build identical to the full one where it is defined.
  configs[2]  exact GP, N = 65 536, 3-D, SE + IndependentNoise: K alpha = y, the NLL assembled from its
              pieces, and the value the 1-GPU recursion and the 2-GPU block-cyclic factorisation agreed on
              to 3e-13 (profiles/r01_bench_n65536.json, profiles/r01_bench_2gpu_n65536.json).
  configs[3]  LOO / leave-one-group-out CV, N = 32 768 (bench_loo_cv shape): the closed forms of
              evaluation/cross_validation_utils.hpp:132-286 checked through independent solves (first
              green run on a B200: profiles/r02a_pytest.log).
"""
import os

import numpy as np
import pytest

from albatross_b200 import capi
from oracle.oracle import Restate
from tests.helpers import assert_close, features, prog, targets

pytestmark = pytest.mark.gpu

ULP_TOL = 4e-15   # Gram entries (relative), as in test_gpu_gram.py
RTOL = 1e-9       # north_star tolerance for information / NLL


def test_gram_config2_full_size(handle):
    n = 32768
    handle.trim()  # release the buffers earlier tests left in the handle's pool
    ops, pp = prog(7)
    x = features(n, 3, 0)
    fd = handle.upload_features(x)
    K = handle.gram_sym_d(ops, pp, fd)
    assert K.shape == (n, n)
    rows = [0, 1, 63, 64, 65, 1023, 1024, 16383, 16384, 20000, n - 65, n - 64, n - 1]
    for r in rows:
        got = K.download_block(r, 0, 1, n).ravel()
        want = Restate.gram_cross(ops, pp, x[r:r + 1], x).ravel()
        assert_close(got, want, ULP_TOL, f"row {r}")
        col = K.download_block(0, r, n, 1).ravel()
        assert np.array_equal(col, got), f"mirror of row {r} must be bit-identical"
    # the lower-only build (what fit uses) writes the same values on and below the diagonal
    Kl = handle.gram_sym_d(ops, pp, fd, flags=capi.GRAM_LOWER_ONLY)
    for c0 in (0, 4096, 20032, n - 128):
        a = K.download_block(c0, c0, n - c0, 64)
        b = Kl.download_block(c0, c0, n - c0, 64)
        tri = np.tril(np.ones((n - c0, 64), dtype=bool))
        assert np.array_equal(a[tri], b[tri]), f"lower-only build differs at column block {c0}"
    K.free()
    Kl.free()
    fd.free()


def test_exact_gp_config3_full_size(handle):
    n = 65536
    handle.trim()
    ops, pp = prog(6)
    x = features(n, 3, 0)
    y = targets(x)
    f, info = handle.gp_fit(ops, pp, x, y)
    assert f.is_positive_definite()
    assert np.all(np.isfinite(info))
    # K alpha = y with the independently rebuilt full symmetric K (GEMV on the device)
    fd = handle.upload_features(x)
    K = handle.gram_sym_d(ops, pp, fd)
    a_dev = handle.upload(info)
    r_dev = handle.alloc(n, 1)
    handle.gemm(K, a_dev, r_dev)
    resid = r_dev.download().ravel() - y
    assert np.linalg.norm(resid) <= RTOL * np.linalg.norm(y), np.linalg.norm(resid) / np.linalg.norm(y)
    # the diagonal carries sigma_se^2 + sigma_noise^2
    d = K.download_block(0, 0, 1, 1)[0, 0]
    assert abs(d - 1.01) <= 4e-16 * 1.01, d
    K.free()
    a_dev.free()
    r_dev.free()
    fd.free()
    logdet = f.log_determinant()
    f.free()
    # log-likelihood = a second Gram + factorisation (models/gp.hpp:443-451); consistent with the fit
    nll = handle.gp_nll(ops, pp, x, y)
    want = 0.5 * (logdet + y @ info + n * np.log(2.0 * np.pi))
    assert abs(nll - want) <= 1e-10 * abs(want), (nll, want)
    # value on which the recursive 1-GPU and the block-cyclic 2-GPU factorisations agreed to 3e-13
    assert abs(nll - (-64227.1208743)) <= RTOL * 64227.0, nll


def test_loo_cv_config4_full_size(handle):
    n = 32768
    handle.trim()
    ops, pp = prog(6)
    x = np.random.default_rng(27).uniform(0.0, 10.0, size=n)  # bench_utils.h:76-85 shape, seed 27
    y = targets(x)
    f, info = handle.gp_fit(ops, pp, x.reshape(-1, 1), y)
    assert f.is_positive_definite()
    # pure leave-one-out (LeaveOneOutGrouper): mean_i = y_i - alpha_i / (K^-1)_ii, var_i = 1 / (K^-1)_ii
    _, off, idx = capi.group_indexers(np.arange(n))
    mean, var, _, _ = handle.gp_cv(f, y, info, off, idx, capi.MARGINAL)
    sample = np.array([0, 1, 63, 64, 4095, 12345, 20000, n - 2, n - 1])
    E = np.zeros((n, len(sample)))
    E[sample, np.arange(len(sample))] = 1.0
    Kinv_cols = f.solve(E)
    d = Kinv_cols[sample, np.arange(len(sample))]
    assert_close(var[sample], 1.0 / d, 1e-9, "LOO variance")
    assert_close(mean[sample], y[sample] - info[sample] / d, 1e-9, "LOO mean")
    # leave-one-group-out, grouper int(x) % 8 (bench_loo_cv.cc:95-105): for each group g
    #   (K^-1)_gg (y_g - mean_g) = alpha_g, checked with one solve per group
    keys = x.astype(np.int64) % 8
    gk, goff, gidx = capi.group_indexers(keys)
    assert list(gk) == sorted(set(keys.tolist()))               # std::map key order
    assert np.array_equal(np.sort(gidx), np.arange(n))           # a permutation
    for g in range(len(gk)):                                     # encounter order inside a group
        members = gidx[goff[g]:goff[g + 1]]
        assert np.array_equal(members, np.flatnonzero(keys == gk[g]))
    gmean, gvar, _, _ = handle.gp_cv(f, y, info, goff, gidx, capi.MARGINAL)
    assert np.all(gvar > 0.0)
    for g in (0, len(gk) - 1):
        members = gidx[goff[g]:goff[g + 1]]
        r = np.zeros(n)
        r[members] = y[members] - gmean[members]
        z = f.solve(r.reshape(-1, 1)).ravel()
        assert_close(z[members], info[members], 1e-9, f"group {g} conditional mean")
    f.free()


# ---- configs[4]: sparse GP, N = 2^20, M = 4096 -------------------------------------------------------
# The oracle cannot run this size (its QR alone is 3.5e13 flops on one core; SURVEY.md §8d extrapolates
# ~3 h).  Size-independent properties instead, all through the C ABI:
#   * R^T R = B^T B is additive over observation groups: B^T B = K_uf A^-1 K_fu + K_uu with a block-diagonal A
#     (sparse_gp.hpp:368-375, :632-706), so fits of two disjoint halves with the SAME inducing points satisfy
#     R^T R = R1^T R1 + R2^T R2 - (K_uu + nugget I), and the normal equations R^T R v = B^T y_aug add up the
#     same way: R^T R v = R1^T R1 v1 + R2^T R2 v2;
#   * a repeated fit is bit-identical (no atomics, fixed reduction orders);
#   * ab_sparse_log_likelihood returns the fit's value;
#   * the nested sub-problem the reference CAN run (N = 131 072, M = 512; fixture generated from the
#     compiled reference by tests/golden/make_golden_r2.py) agrees to 1e-9.

def _sparse_data(n):
    x = np.random.default_rng(0).uniform(0.0, 10.0, size=n)
    return x, np.sin(x) + 0.1 * np.cos(10.0 * x)


@pytest.mark.parametrize("mode", ["fitc", "pitc1024"])
def test_sparse_config5_full_size(handle, mode):
    n, m = 1 << 20, 4096
    handle.trim()
    ops, pp = prog(6)
    x, y = _sparse_data(n)
    u = np.linspace(x.min(), x.max(), m)
    scale = n / 10.0 / 1024.0
    keys = np.arange(n, dtype=np.int64) if mode == "fitc" else (x * scale).astype(np.int64)

    def fit(sel):
        _, off, idx = capi.group_indexers(keys[sel])
        f, v, ll = handle.sparse_fit(ops, pp, x[sel], y[sel], u, off, idx)
        R = f.export_R()
        f.free()
        handle.trim()
        return R.T @ R, v, ll

    everything = np.ones(n, dtype=bool)
    G, v, ll = fit(everything)
    G_again, v_again, ll_again = fit(everything)
    assert np.array_equal(v, v_again) and ll == ll_again and np.array_equal(G, G_again)
    _, off, idx = capi.group_indexers(keys)
    assert handle.sparse_log_likelihood(ops, pp, x, y, u, off, idx) == ll
    assert np.isfinite(ll) and np.all(np.isfinite(v))
    # two halves at a group boundary (x = 5 is a multiple of the 1024-point group width 10 / 1024 ... exactly
    # key 512 for the PITC grouper; singletons split anywhere)
    left = x < 5.0
    G1, v1, _ = fit(left)
    G2, v2, _ = fit(~left)
    Kuu = Restate.gram_sym(ops, pp, u) + 1e-8 * np.eye(m)
    gscale = np.max(np.abs(G))
    err_G = np.max(np.abs(G - (G1 + G2 - Kuu))) / gscale
    rhs, rhs12 = G @ v, G1 @ v1 + G2 @ v2
    err_ne = np.max(np.abs(rhs - rhs12)) / np.max(np.abs(rhs))
    print(f"config 5 {mode}: ll {ll:.6f}  additivity of R^T R {err_G:.2e}  of the normal equations {err_ne:.2e}")
    assert err_G <= 1e-9 and err_ne <= 1e-9, (err_G, err_ne)


def test_sparse_nested_subproblem_vs_reference_fixture(handle, golden):
    _, ref = golden
    n, m = (int(v) for v in ref["spn_n_m"])
    ops, pp = prog(6)
    x, y = _sparse_data(n)
    assert np.array_equal(np.array([x.sum(), x[12345], y.sum()]), ref["spn_x_checksum"]), "RNG stream changed"
    u = np.linspace(x.min(), x.max(), m)
    t = ref["spn_test"]
    for tag, keys in (("fitc", np.arange(n, dtype=np.int64)),
                      ("pitc", np.floor(x * (n / 10.0 / 1024.0)).astype(np.int64))):
        _, off, idx = capi.group_indexers(keys)
        f, v, ll = handle.sparse_fit(ops, pp, x, y, u, off, idx)
        mean, var, _ = f.predict(ops, pp, t, capi.MARGINAL)
        f.free()
        want = float(ref[f"spn_{tag}_ll"])
        assert abs(ll - want) <= RTOL * abs(want), (tag, ll, want)
        assert_close(mean, ref[f"spn_{tag}_mean"], RTOL, f"nested {tag} mean")
        assert np.max(np.abs(var - ref[f"spn_{tag}_var"])) <= RTOL * 1.01, tag  # prior scale: 1 + 0.1^2
    handle.trim()
