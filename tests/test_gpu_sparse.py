"""GPU parity: sparse (FITC / PITC) GP through the C ABI vs the oracle and the reference fixtures.

Reference tests mirrored: tests/test_sparse_gp.cc (fit/predict for several groupers :60-123, the
log-likelihood against the dense one :172-221, sparse == dense when every point is inducing
:125-164).  Tolerances: 1e-9 relative on means and the log-likelihood (BASELINE.json north_star);
1e-9 of the prior scale on variances and covariances, which the reference itself forms by cancellation
(sparse_gp.hpp:481-536)."""
import numpy as np
import pytest

from albatross_b200 import capi
from albatross_b200.capi import JOINT, MARGINAL, MEAN
from oracle.oracle import Restate, group_keys, menu_program, menu_program_plain
from tests.helpers import assert_close, features, prog, rel_err, targets

pytestmark = pytest.mark.gpu


def fit(handle, cid, x, y, u, keys, **kw):
    ops, pp = prog(cid)
    _, offsets, indices = capi.group_indexers(keys)
    return handle.sparse_fit(ops, pp, x, y, u, offsets, indices, **kw)


def test_sparse_fixture(handle, golden):
    """Outputs of the compiled reference (tests/golden/make_golden.py), FITC and PITC groups."""
    _, ref = golden
    ops, pp = prog(6)
    x, y, u, t = ref["sp_x"], ref["sp_y"], ref["sp_u"], ref["sp_test"]
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        f, info, ll = fit(handle, 6, x, y, u, group_keys(x, gk, ga))
        mean, var, _ = f.predict(ops, pp, t, MARGINAL)
        mean2, _, cov = f.predict(ops, pp, t, JOINT)
        mean3, _, _ = f.predict(ops, pp, t, MEAN)
        assert_close(mean, ref[f"sp_{tag}_mean"], 1e-9, f"{tag} mean")
        assert np.array_equal(mean, mean2) and np.array_equal(mean, mean3)
        assert_close(var, ref[f"sp_{tag}_var"], 1e-9, f"{tag} var")
        assert_close(cov, ref[f"sp_{tag}_cov"], 1e-9, f"{tag} cov")
        want = float(ref[f"sp_{tag}_ll"])
        assert abs(ll - want) <= 1e-9 * abs(want), (tag, ll, want)
        assert f.log_likelihood == ll and f.m == len(u)
        f.free()


@pytest.mark.parametrize("n,m,gk,ga", [(300, 30, 0, 0.0), (777, 64, 2, 1.0), (2048, 64, 2, 1.0),
                                        (1500, 129, 1, 7.0), (513, 200, 0, 0.0)])
def test_sparse_vs_oracle(handle, n, m, gk, ga):
    """benchmarks/bench_gram.cc:47-71 shape (groups `int(f)`), odd sizes, FITC and grouped."""
    ops, pp = prog(6)
    x = features(n, 1, n).ravel()
    y = targets(x)
    u = Restate.linspace(x.min(), x.max(), m)
    t = np.linspace(0.1, 9.9, 37)
    keys = group_keys(x, gk, ga)
    want = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=2, want_ll=True)
    want_var = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=1)["var"]
    f, info, ll = fit(handle, 6, x, y, u, keys)
    mean, var, _ = f.predict(ops, pp, t, MARGINAL)
    _, _, cov = f.predict(ops, pp, t, JOINT)
    assert_close(mean, want["mean"], 1e-9, "mean")
    assert abs(ll - want["ll"]) <= 1e-9 * abs(want["ll"]), (ll, want["ll"])
    scale = np.max(np.abs(want["cov"]))
    assert np.max(np.abs(var - want_var)) <= 1e-9 * scale
    assert np.max(np.abs(cov - want["cov"])) <= 1e-9 * scale
    # K_*u v is what the caller sees; v itself is conditioned like K_uu (1e-8 nugget)
    assert_close(handle.sparse_log_likelihood(ops, pp, x, y, u, *capi.group_indexers(keys)[1:]),
                 ll, 1e-12)
    f.free()


def test_sparse_measurement_variance_and_nuggets(handle):
    ops, pp = prog(6)
    x = features(400, 1, 9).ravel()
    y = targets(x)
    yvar = 0.01 + 0.05 * np.random.default_rng(2).uniform(size=len(x))
    u = Restate.linspace(0.0, 10.0, 25)
    keys = group_keys(x, 2, 2.0)
    t = np.linspace(0.0, 10.0, 11)
    kw = dict(yvar=yvar, measurement_nugget=1e-6, inducing_nugget=1e-7)
    want = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=1, want_ll=True, **kw)
    f, info, ll = fit(handle, 6, x, y, u, keys, **kw)
    mean, var, _ = f.predict(ops, pp, t, MARGINAL)
    assert_close(mean, want["mean"], 1e-9)
    assert_close(var, want["var"], 1e-9, "variances with measurement noise")
    assert abs(ll - want["ll"]) <= 1e-9 * abs(want["ll"])
    f.free()


def test_sparse_R_reproduces_normal_matrix(handle):
    """sigma_R is representation-internal (the reference's is column-pivoted); its Gram is not:
    R^T R = B^T B = K_uf A^-1 K_fu + K_uu  (sparse_gp.hpp:368-375)."""
    ops, pp = prog(6)
    x = np.sort(features(500, 1, 4).ravel())
    y = targets(x)
    m = 40
    u = Restate.linspace(0.0, 10.0, m)
    keys = group_keys(x, 0)  # FITC
    f, info, ll = fit(handle, 6, x, y, u, keys)
    R = f.export_R()
    assert np.allclose(np.tril(R, -1), 0.0)
    Kuu = Restate.gram_sym(ops, pp, u) + 1e-8 * np.eye(m)
    Kfu = Restate.gram_cross(ops, pp, x, u)
    Lu = np.linalg.cholesky(Kuu)
    P = np.linalg.solve(Lu, Kfu.T)
    a = Restate.gram_diag(ops, pp, x) - np.sum(P * P, axis=0) + 1e-8
    want = Kfu.T @ (Kfu / a[:, None]) + Kuu
    assert_close(R.T @ R, want, 1e-9, "R^T R")
    # v solves the normal equations
    rhs = Kfu.T @ (y / a)
    assert_close(want @ info, rhs, 1e-7, "normal equations")
    f.free()


def test_sparse_equals_dense_when_all_points_induce(handle):
    """tests/test_sparse_gp.cc:125-164: inducing points == training points => the exact GP."""
    ops, pp = prog(6)
    x = np.sort(features(200, 1, 8).ravel())
    y = targets(x)
    t = np.linspace(0.5, 9.5, 9)
    f, info, ll = fit(handle, 6, x, y, x.copy(), group_keys(x, 0), inducing_nugget=0.0,
                      measurement_nugget=1e-12)
    mean, var, _ = f.predict(ops, pp, t, MARGINAL)
    dense_mean, dense_var, _ = Restate.gp_predict(ops, pp, x, y, t, 1)
    # with IndependentNoise on K_uu the inducing model carries the noise itself; means agree
    assert rel_err(mean, dense_mean) < 1e-5
    f.free()


def test_sparse_large_fitc_consistency(handle):
    """n = 20 000 FITC: deterministic reductions (two fits are bit-identical) and parity with the
    oracle at a size where the row-wise FITC kernels run many CTAs."""
    ops, pp = prog(6)
    n, m = 20000, 256
    x = features(n, 1, 21).ravel()
    y = targets(x)
    u = Restate.linspace(0.0, 10.0, m)
    keys = np.arange(n, dtype=np.int64)
    t = np.linspace(0.0, 10.0, 101)
    want = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=1, want_ll=True)
    f1, info1, ll1 = fit(handle, 6, x, y, u, keys)
    f2, info2, ll2 = fit(handle, 6, x, y, u, keys)
    assert ll1 == ll2 and np.array_equal(info1, info2)  # deterministic reductions
    mean, var, _ = f1.predict(ops, pp, t, MARGINAL)
    assert_close(mean, want["mean"], 1e-9, "mean")
    assert np.max(np.abs(var - want["var"])) <= 1e-9 * np.max(np.abs(want["var"]))
    assert abs(ll1 - want["ll"]) <= 1e-9 * abs(want["ll"])
    assert np.all(var > 0.0)
    f1.free()
    f2.free()


# ---- covariances with a MeasurementOnly term: the reference's standard sparse configuration --------
# (tests/lib/albatross/test/test_models.h:26-30; K_ff / K_fu / K_uu programs differ, sparse_gp.hpp:646-679)

P10 = [1.0, 1.0, 0.1]


def fit_mo(handle, params, x, y, u, keys, **kw):
    ops, pp = menu_program(10, params)          # k(Measurement, Measurement): SE + noise
    plain = menu_program_plain(10, params)      # any other pairing: SE alone
    _, offsets, indices = capi.group_indexers(keys)
    return handle.sparse_fit(ops, pp, x, y, u, offsets, indices, fu=plain, uu=plain, **kw), plain


def test_sparse_measurement_only_fixture(handle, golden):
    """Menu entry 10 (SE + measurement_only(IndependentNoise)) against the compiled reference."""
    _, ref = golden
    x, y, u, t = ref["spmo_x"], ref["spmo_y"], ref["spmo_u"], ref["spmo_test"]
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        for vtag, yv in (("", None), ("_yvar", ref["spmo_yvar"])):
            (f, info, ll), plain = fit_mo(handle, P10, x, y, u, group_keys(x, gk, ga), yvar=yv)
            mean, var, _ = f.predict(*plain, t, MARGINAL)
            _, _, cov = f.predict(*plain, t, JOINT)
            k = f"spmo_{tag}{vtag}"
            assert_close(mean, ref[f"{k}_mean"], 1e-9, f"{k} mean")
            scale = np.max(np.abs(ref[f"{k}_cov"]))
            assert np.max(np.abs(var - ref[f"{k}_var"])) <= 1e-9 * scale, k
            assert np.max(np.abs(cov - ref[f"{k}_cov"])) <= 1e-9 * scale, k
            want = float(ref[f"{k}_ll"])
            assert abs(ll - want) <= 1e-9 * abs(want), (k, ll, want)
            f.free()


def test_sparse_reference_test_model(handle, golden):
    """MakeSparseGaussianProcess (test_models.h:44-57): SE(100, 100) + measurement_only(noise 0.1) on
    make_toy_linear_data(), 25 uniformly spaced inducing points for 10 observations.  K_uu has condition
    number ~1e12 relative to its 1e-8 nugget: means and the log-likelihood agree with the pivoted reference
    to 1e-6 / 1e-7 here (the reference's own tolerance for this model against the dense GP is 1e-2 .. 1e-6,
    tests/test_sparse_gp.cc:107-122,158-163)."""
    _, ref = golden
    x, y, u, t = ref["sptoy_x"], ref["sptoy_y"], ref["sptoy_u"], ref["sptoy_test"]
    (f, info, ll), plain = fit_mo(handle, [100.0, 100.0, 0.1], x, y, u, np.arange(len(x)))
    mean, var, _ = f.predict(*plain, t, MARGINAL)
    err_mean = rel_err(mean, ref["sptoy_mean"])
    err_var = float(np.max(np.abs(var - ref["sptoy_var"])))
    err_ll = abs(ll - float(ref["sptoy_ll"])) / abs(float(ref["sptoy_ll"]))
    print(f"sparse toy model: mean {err_mean:.2e} var(abs) {err_var:.2e} ll {err_ll:.2e}")
    assert err_mean <= 1e-6 and err_var <= 1e-6 and err_ll <= 1e-6, (err_mean, err_var, err_ll)
    f.free()


@pytest.mark.parametrize("n,m,gk,ga", [(500, 40, 0, 0.0), (1500, 96, 2, 1.0)])
def test_sparse_measurement_only_vs_oracle(handle, n, m, gk, ga):
    x = features(n, 1, 3 * n).ravel()
    y = targets(x)
    u = Restate.linspace(x.min(), x.max(), m)
    t = np.linspace(0.1, 9.9, 23)
    keys = group_keys(x, gk, ga)
    ops, pp = menu_program(10, P10)
    plain = menu_program_plain(10, P10)
    want = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=2, want_ll=True, fu=plain, uu=plain)
    (f, info, ll), _ = fit_mo(handle, P10, x, y, u, keys)
    mean, _, cov = f.predict(*plain, t, JOINT)
    assert_close(mean, want["mean"], 1e-9, "mean")
    assert abs(ll - want["ll"]) <= 1e-9 * abs(want["ll"]), (ll, want["ll"])
    # the posterior covariance is prior - Q** + S** (sparse_gp.hpp:509-536): its error scales with the prior
    assert np.max(np.abs(cov - want["cov"])) <= 1e-9 * P10[1] ** 2
    _, offsets, indices = capi.group_indexers(keys)
    assert_close(handle.sparse_log_likelihood(ops, pp, x, y, u, offsets, indices, fu=plain, uu=plain),
                 ll, 1e-12)
    f.free()
