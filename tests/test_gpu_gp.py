"""GPU parity: factorisation, solves, reductions, exact-GP fit / predict / NLL / cross-validation
through the C ABI vs the oracle.  Tolerance: 1e-9 relative (BASELINE.json north_star) unless a
tighter one is written; index outputs are bit-exact."""
import numpy as np
import pytest

from albatross_b200 import capi
from albatross_b200.capi import JOINT, MARGINAL, MEAN
from oracle.oracle import Ref, Restate, group_keys
from tests.helpers import GP_COVS, PARAMS, assert_close, features, prog, rel_err, targets

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def spd(n, seed, dim=3, cid=8):
    ops, pp = prog(cid)
    return Restate.gram_sym(ops, pp, features(n, dim, seed))


# ---- SerializableLDLT surface (tests/test_serializable_ldlt.cc:34-85) ---------------------------

@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 129, 200, 333, 1000])
def test_factor_solve_logdet(handle, n):
    A = spd(n, seed=n)
    rhs = np.random.default_rng(n).standard_normal((n, 3))
    f = handle.potrf(handle.upload(A))
    assert f.is_positive_definite() and f.n == n
    want = Restate.ldlt(A, rhs=rhs)
    assert_close(f.solve(rhs), want["solve"], RTOL, f"solve n={n}")
    assert_close(f.solve(rhs[:, 0]), want["solve"][:, 0], RTOL)
    assert abs(f.log_determinant() - want["logdet"]) <= RTOL * max(1.0, abs(want["logdet"]))
    # sqrt_solve is representation dependent (pivoting); its Gram is not:
    s = f.sqrt_solve(rhs)
    assert_close(s.T @ s, rhs.T @ want["solve"], RTOL, "sqrt_solve identity")
    # L L^T reproduces A
    LD, tr = f.export_packed()
    assert np.array_equal(tr, np.arange(n))
    L = np.tril(LD, -1) + np.eye(n)
    assert_close((L * np.diag(LD)) @ L.T, A, 1e-12, "L D L^T")


@pytest.mark.parametrize("n", [1, 5, 64, 65, 200, 333, 1000])
def test_sqrt_surface_identities(handle, n):
    """sqrt_product / sqrt_transpose / sqrt_transpose_solve / diagonal_sqrt
    (src/eigen/serializable_ldlt.hpp:58-126).  The square root is representation dependent (the reference's
    is P^T L D^1/2 with its pivot order), so parity is what the reference's own test checks
    (tests/test_serializable_ldlt.cc:58-85): S S^T = A, sqrt_solve(S) = I, sqrt_transpose_solve(S^T) = I."""
    A = spd(n, seed=7 * n + 1)
    f = handle.potrf(handle.upload(A))
    eye = np.eye(n)
    S = f.sqrt_product(eye)
    assert np.array_equal(S, np.tril(S))                      # unpivoted: the square root is lower triangular
    assert_close(S @ S.T, A, 1e-12, "S S^T = A")
    ST = f.sqrt_transpose()
    assert np.array_equal(ST, S.T)                            # the same numbers, transposed
    assert np.linalg.norm(eye - f.sqrt_solve(S)) <= 1e-12 * n
    assert np.linalg.norm(eye - f.sqrt_transpose_solve(np.ascontiguousarray(ST))) <= 1e-12 * n
    d = f.diagonal_sqrt()
    assert np.array_equal(d, np.diag(S)) and np.all(d > 0.0)
    rhs = np.random.default_rng(n).standard_normal((n, 4))
    assert_close(f.sqrt_product(rhs), S @ rhs, 1e-13, "sqrt_product rhs")
    assert_close(f.sqrt_product(rhs[:, 0]), S @ rhs[:, 0], 1e-13, "sqrt_product vector")
    # K^-1 = S^-T S^-1: the two half solves compose to the full one
    assert_close(f.sqrt_transpose_solve(f.sqrt_solve(rhs)), f.solve(rhs), 1e-12, "two half solves")


def test_export_packed_streams_in_panels(handle, monkeypatch):
    """ab_factor_export_packed goes through a bounded device buffer (no n^2 scratch); forced to 37-column
    panels here, the result is the single-panel one bit for bit."""
    n = 500
    A = spd(n, seed=99)
    f = handle.potrf(handle.upload(A))
    LD1, _ = f.export_packed()
    monkeypatch.setenv("AB_EXPORT_PANEL_COLS", "37")
    LD2, tr = f.export_packed()
    ST = f.sqrt_transpose()
    monkeypatch.delenv("AB_EXPORT_PANEL_COLS")
    assert np.array_equal(LD1, LD2) and np.array_equal(tr, np.arange(n))
    assert np.array_equal(ST, f.sqrt_transpose())
    L = np.tril(LD2, -1) + np.eye(n)
    assert_close((L * np.diag(LD2)) @ L.T, A, 1e-12, "L D L^T")


def test_ldlt_wrapper_fixture(handle, golden):
    _, ref = golden
    A, rhs = ref["ldlt_A"], ref["ldlt_rhs"]
    f = handle.potrf(handle.upload(A))
    assert_close(f.solve(rhs), ref["ldlt_solve"], RTOL)
    assert abs(f.log_determinant() - float(ref["ldlt_logdet"])) < RTOL * abs(float(ref["ldlt_logdet"]))
    assert_close(f.inverse_diagonal(), ref["ldlt_inverse_diagonal"], RTOL)
    groups = [[0, 5, 9], [1], [100, 101, 102, 127], [64, 63]]
    got = np.concatenate([b.ravel(order="F") for b in f.inverse_blocks(groups)])
    assert_close(got, ref["ldlt_inverse_blocks"], RTOL)


@pytest.mark.parametrize("n", [5, 64, 300, 700])
def test_inverse_diagonal_and_blocks(handle, n):
    A = spd(n, seed=3 * n)
    f = handle.potrf(handle.upload(A))
    inv = np.linalg.inv(A)
    # against numpy's explicit inverse: the reference's own tolerance (tests/test_serializable_ldlt.cc:41-46)
    assert_close(f.inverse_diagonal(), np.diag(inv), 1e-8, "inverse diagonal vs numpy inv")
    rng = np.random.default_rng(n)
    perm = rng.permutation(n)
    groups = [perm[: n // 3], perm[n // 3: n // 3 + 1], perm[n // 3 + 1:]]
    for g, b in zip(groups, f.inverse_blocks(groups)):
        assert_close(b, inv[np.ix_(g, g)], 1e-8)
    want = Restate.inverse_blocks(A, groups)
    for b, w in zip(f.inverse_blocks(groups), want):
        assert_close(b, w, RTOL)


def test_nll_known_answer(handle, golden):
    """tests/test_evaluate.cc:34-63 through the device factor."""
    t, _ = golden
    k = t["nll_known_answer"]
    f = handle.potrf(handle.upload(np.array(k["cov"])))
    assert abs(f.nll(np.array(k["x"])) - k["value"]) < 1e-12


def test_not_positive_definite_is_reported(handle):
    A = spd(130, seed=1)
    A[70, 70] = -1.0
    with pytest.raises(capi.AbError) as e:
        handle.potrf(handle.upload(A))
    assert e.value.status == 4
    f = handle.potrf(handle.upload(A), allow_not_pd=True)
    assert not f.is_positive_definite() and f.info() == 70
    with pytest.raises(capi.AbError):
        f.solve(np.ones(130))
    # NaN anywhere in the lower triangle -> reported, never UB (gp.hpp:66)
    B = spd(200, seed=2)
    B[150, 20] = np.nan
    f = handle.potrf(handle.upload(B), allow_not_pd=True)
    assert not f.is_positive_definite()


def test_loo_fast_path_validates_indices(handle):
    """n singleton groups take an element-wise fast path only when their indices are a permutation of
    0..n-1; out-of-range indices are rejected, duplicates go through the general (range-checked) path."""
    ops, pp = prog(6)
    n = 90
    x = features(n, 1, 5).ravel()
    y = targets(x)
    f, info = handle.gp_fit(ops, pp, x, y)
    offsets = np.arange(n + 1, dtype=np.int64)
    good = np.arange(n, dtype=np.int64)
    m0, v0, _, _ = handle.gp_cv(f, y, info, offsets, good, MARGINAL)
    bad = good.copy()
    bad[5] = n + 3
    with pytest.raises(capi.AbError) as err:
        handle.gp_cv(f, y, info, offsets, bad, MARGINAL)
    assert err.value.status == 1  # AB_ERR_INVALID
    dup = good.copy()
    dup[7] = 6                    # observation 6 held out twice, observation 7 never
    m1, v1, _, _ = handle.gp_cv(f, y, info, offsets, dup, MARGINAL)
    keep = np.arange(n) != 7
    assert_close(m1[keep], m0[keep], 1e-12, "duplicate index: the other observations")
    assert m1[7] == 0.0 and v1[7] == 0.0


# ---- exact GP (tests/test_gp.cc, tests/lib/albatross/test/test_models.cc) -----------------------

@pytest.mark.parametrize("cid", GP_COVS)
@pytest.mark.parametrize("tag", ["1", "3"])
def test_gp_vs_reference_fixture(handle, golden, cid, tag):
    _, ref = golden
    ops, pp = prog(cid)
    x, y, t = ref[f"gp_x{tag}"], ref[f"gp_y{tag}"], ref[f"gp_test{tag}"]
    f, info = handle.gp_fit(ops, pp, x, y)
    assert_close(info, ref[f"info{tag}_{cid}"], RTOL, "information")
    nll = handle.gp_nll(ops, pp, x, y)
    assert abs(nll - float(ref[f"nll{tag}_{cid}"])) <= RTOL * abs(float(ref[f"nll{tag}_{cid}"]))
    assert_close(handle.gp_predict(f, ops, pp, x, info, t, MEAN)[0], ref[f"mean{tag}_{cid}"], RTOL)
    m, v, _ = handle.gp_predict(f, ops, pp, x, info, t, MARGINAL)
    assert_close(m, ref[f"mean{tag}_{cid}"], RTOL)
    assert_close(v, ref[f"var{tag}_{cid}"], RTOL, "marginal variance")
    m, _, c = handle.gp_predict(f, ops, pp, x, info, t, JOINT)
    assert_close(c, ref[f"cov{tag}_{cid}"], RTOL, "joint covariance")
    # leave-one-out and leave-one-group-out
    _, offsets, indices = capi.group_indexers(group_keys(x, 0))
    m, v, _, s = handle.gp_cv(f, y, info, offsets, indices, MARGINAL, want_score=True)
    assert_close(m, ref[f"loo_mean{tag}_{cid}"], RTOL, "loo mean")
    assert_close(v, ref[f"loo_var{tag}_{cid}"], RTOL, "loo var")
    assert abs(s - float(ref[f"loo_score{tag}_{cid}"])) <= RTOL * abs(float(ref[f"loo_score{tag}_{cid}"]))
    keys, offsets, indices = capi.group_indexers(group_keys(x, 1, 8))
    m, v, _, s = handle.gp_cv(f, y, info, offsets, indices, MARGINAL, want_score=True)
    assert_close(m, ref[f"logo_mean{tag}_{cid}"], RTOL, "logo mean")
    assert_close(v, ref[f"logo_var{tag}_{cid}"], RTOL, "logo var")
    assert abs(s - float(ref[f"logo_score{tag}_{cid}"])) <= RTOL * abs(float(ref[f"logo_score{tag}_{cid}"]))
    m2, _, j, _ = handle.gp_cv(f, y, info, offsets, indices, JOINT)
    assert_close(j, ref[f"logo_joint{tag}_{cid}"], RTOL, "logo joint")
    assert_close(m2, m, 1e-13)
    m3 = handle.gp_cv(f, y, info, offsets, indices, MEAN)[0]
    assert_close(m3, m, 1e-13)


def test_fit_adds_target_variance_but_nll_does_not(handle, golden):
    """gp.hpp:65 vs gp.hpp:447-448."""
    _, ref = golden
    ops, pp = prog(6)
    x, y, yvar, t = ref["gp_x1"], ref["gp_y1"], ref["gp_yvar"], ref["gp_test1"]
    f, info = handle.gp_fit(ops, pp, x, y, yvar=yvar)
    assert_close(info, ref["info1_6_yvar"], RTOL)
    assert_close(handle.gp_predict(f, ops, pp, x, info, t, MARGINAL)[1], ref["var1_6_yvar"], RTOL)
    assert abs(handle.gp_nll(ops, pp, x, y) - float(ref["nll1_6"])) <= RTOL * abs(float(ref["nll1_6"]))


@pytest.mark.parametrize("n,dim", [(1, 1), (2, 3), (65, 1), (500, 3), (1500, 3)])
def test_gp_against_restatement(handle, n, dim):
    ops, pp = prog(8)
    x = features(n, dim, seed=n)
    y = targets(x)
    t = features(9, dim, seed=99)
    f, info = handle.gp_fit(ops, pp, x, y)
    assert_close(info, Restate.gp_fit(ops, pp, x, y)["information"], RTOL)
    want_nll = Restate.gp_nll(ops, pp, x, y)
    assert abs(handle.gp_nll(ops, pp, x, y) - want_nll) <= RTOL * max(1.0, abs(want_nll))
    f2, info2, nll2 = handle.gp_fit_nll(ops, pp, x, y)
    assert abs(nll2 - want_nll) <= RTOL * max(1.0, abs(want_nll))
    assert_close(info2, info, 1e-13)
    for what in (MEAN, MARGINAL, JOINT):
        got = handle.gp_predict(f, ops, pp, x, info, t, what)
        want = Restate.gp_predict(ops, pp, x, y, t, what)
        for g, w in zip(got, want):
            if w is not None:
                assert_close(g, w, RTOL, f"predict {what}")


def test_predict_training_points_and_order(handle):
    """tests/test_gp.cc: predictions preserve order; mean at training points ~ targets."""
    ops, pp = prog(6)
    x = features(400, 1, 7)
    y = targets(x)
    f, info = handle.gp_fit(ops, pp, x, y)
    perm = np.random.default_rng(0).permutation(400)[:50]
    a = handle.gp_predict(f, ops, pp, x, info, x[perm], MARGINAL)
    b = handle.gp_predict(f, ops, pp, x, info, x[np.sort(perm)], MARGINAL)
    order = np.argsort(perm)
    assert_close(a[0][order], b[0], 1e-12)
    assert_close(a[1][order], b[1], 1e-12)
    assert np.max(np.abs(a[0] - y[perm])) < 0.2


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not shipped")
def test_gp_against_live_reference_n2000(handle):
    """The reference's own Eigen path at a size it finishes in seconds."""
    cid = 6
    ops, pp = prog(cid)
    x = Ref.random_features(2000, 3, 0)
    y = Ref.random_targets(x)
    t = Ref.random_features(32, 3, 5)
    f, info = handle.gp_fit(ops, pp, x, y)
    assert_close(info, Ref.gp_fit(cid, PARAMS[cid], x, y)["information"], RTOL)
    want = Ref.gp_nll(cid, PARAMS[cid], x, y)[0]
    assert abs(handle.gp_nll(ops, pp, x, y) - want) <= RTOL * abs(want)
    m, v, _ = handle.gp_predict(f, ops, pp, x, info, t, MARGINAL)
    rm, rv, _ = Ref.gp_predict(cid, PARAMS[cid], x, y, t, 1)
    assert_close(m, rm, RTOL)
    assert_close(v, rv, RTOL)


# ---- size-independent properties at scale -------------------------------------------------------

def test_large_fit_residual_and_consistency(handle):
    """N=8192: K alpha = y to backward-error level; NLL pieces consistent with the factor."""
    ops, pp = prog(6)
    n = 8192
    x = features(n, 3, 0)
    y = targets(x)
    f, info = handle.gp_fit(ops, pp, x, y)
    K = handle.gram_sym(ops, pp, x).download()
    resid = np.linalg.norm(K @ info - y) / (np.linalg.norm(K, 2) * np.linalg.norm(info))
    assert resid < 1e-14, resid
    nll = handle.gp_nll(ops, pp, x, y)
    want = 0.5 * (f.log_determinant() + y @ info + n * np.log(2 * np.pi))
    assert abs(nll - want) <= 1e-10 * abs(want)
    sign, logdet = np.linalg.slogdet(K)
    assert sign > 0 and abs(f.log_determinant() - logdet) <= 1e-9 * abs(logdet)
    # leave-one-out vs the closed form on a subset of points
    _, offsets, indices = capi.group_indexers(np.arange(n))
    m, v, _, _ = handle.gp_cv(f, y, info, offsets, indices, MARGINAL)
    Kinv_diag = np.diag(np.linalg.inv(K))
    assert_close(v, 1.0 / Kinv_diag, 1e-9, "LOO variance vs dense algebra")
    assert_close(m, y - info / Kinv_diag, 1e-9, "LOO mean vs dense algebra")


@pytest.mark.parametrize("n", [4096, 4097, 6000])
def test_lookahead_factorisation_matches_recursion(handle, n, monkeypatch):
    """The two-stream look-ahead schedule (linalg.cu potrf_lookahead, panels of 2048) factors the same
    matrix as the plain recursion: same L to rounding, L L^T = K, same solve; ragged last panel."""
    ops, pp = prog(8)
    x = features(n, 3, n)
    K = handle.gram_sym(ops, pp, x).download()
    rhs = np.random.default_rng(n).standard_normal((n, 2))
    # force the look-ahead schedule at these sizes (default threshold: 4 panels of 2048)
    monkeypatch.setenv("AB_POTRF_LOOKAHEAD_MIN", "1")
    f = handle.potrf(handle.upload(K))
    LD, _ = f.export_packed()
    sol = f.solve(rhs)
    monkeypatch.delenv("AB_POTRF_LOOKAHEAD_MIN")
    monkeypatch.setenv("AB_POTRF_RECURSIVE", "1")
    f2 = handle.potrf(handle.upload(K))
    LD2, _ = f2.export_packed()
    monkeypatch.delenv("AB_POTRF_RECURSIVE")
    assert f.is_positive_definite() and f2.is_positive_definite()
    assert_close(np.tril(LD), np.tril(LD2), 1e-11, "L vs recursion")
    L = np.tril(LD, -1) + np.eye(n)
    assert_close((L * np.diag(LD)) @ L.T, K, 1e-12, "L D L^T")
    assert_close(sol, f2.solve(rhs), 1e-10, "solve")
    resid = np.linalg.norm(K @ sol - rhs) / (np.linalg.norm(K, 2) * np.linalg.norm(sol))
    assert resid < 1e-14, resid
    # a non-positive pivot in a late panel is still reported with its global index
    Kb = K.copy()
    Kb[n - 3, n - 3] = -1.0
    fb = handle.potrf(handle.upload(Kb), allow_not_pd=True)
    assert not fb.is_positive_definite() and fb.info() == n - 3


# ---- the default look-ahead schedule against the reference itself ------------------------------------

def test_gp_n8192_default_schedule_vs_reference_fixture(handle, golden):
    """N = 8192 is the smallest size at which ab_potrf takes its default schedule (look-ahead on two streams,
    TMA-fed DSYRK); information / NLL / predictions / leave-one-out against outputs of the compiled reference
    (tests/golden/make_golden_r2.py `big`, ~5 min of single-core Eigen LDLT) at the north_star 1e-9."""
    _, ref = golden
    n = 8192
    x = np.random.default_rng(8192).uniform(0.0, 10.0, size=(n, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10.0 * x[:, 0])
    assert np.array_equal(np.array([x.sum(), x[4321, 1], y.sum()]), ref["big_x_checksum"]), "RNG stream changed"
    t = ref["big_test"]
    ops, pp = prog(6)
    f, info = handle.gp_fit(ops, pp, x, y)
    assert_close(info, ref["big_information"], RTOL, "information N=8192")
    nll = handle.gp_nll(ops, pp, x, y)
    assert abs(nll - float(ref["big_nll"])) <= RTOL * abs(float(ref["big_nll"])), (nll, float(ref["big_nll"]))
    mean, var, _ = handle.gp_predict(f, ops, pp, x, info, t, MARGINAL)
    assert_close(mean, ref["big_mean"], RTOL, "mean N=8192")
    assert np.max(np.abs(var - ref["big_var"])) <= RTOL * 1.01, "marginal variance (prior scale 1.01)"
    _, _, cov = handle.gp_predict(f, ops, pp, x, info, t, JOINT)
    assert np.max(np.abs(cov - ref["big_cov"])) <= RTOL * 1.01, "joint covariance"
    _, offsets, indices = capi.group_indexers(np.arange(n))
    m, v, _, s = handle.gp_cv(f, y, info, offsets, indices, MARGINAL, want_score=True)
    assert_close(m, ref["big_loo_mean"], RTOL, "LOO mean N=8192")
    assert_close(v, ref["big_loo_var"], RTOL, "LOO variance N=8192")
    assert abs(s - float(ref["big_loo_score"])) <= RTOL * abs(float(ref["big_loo_score"]))
    f.free()


def test_singular_psd_policy_next_to_the_reference(handle, golden):
    """Duplicate points without a noise term: K is positive SEMI-definite (rank 5 of 8).
    Reference (fixture from the compiled reference): the diagonally pivoted LDLT finishes with exact zeros in D
    (LDLT.h:316-338), is_positive_definite() is false, log_determinant() is -inf and the log-likelihood NaN;
    solve() goes through the pseudo-inverse of D (:568-585) and, for a right-hand side in the range of K,
    returns an exact solution.
    Device (policy): the unpivoted factorisation stops at the first non-positive pivot and
    reports it (AB_ERR_NOT_PD, ab_factor_info); the factor is not usable for solves.  The trait layer turns
    that into NaN outputs and is_positive_definite() == false, so likelihood-driven tuning sees what it sees
    with the reference (NaN -> +inf objective, tune.hpp:164-166); only solve() on a consistent singular
    system differs (NaN instead of a pseudo-inverse solution)."""
    _, ref = golden
    A, rhs = ref["psd_A"], ref["psd_rhs"]
    # what the reference does
    assert int(ref["psd_is_pd"]) == 0 and np.isneginf(float(ref["psd_logdet"])) and np.isnan(float(ref["psd_nll"]))
    assert np.sum(ref["psd_D"] == 0.0) == 3
    assert np.max(np.abs(A @ ref["psd_solve"] - rhs)) <= 1e-15
    # what the device does
    f = handle.potrf(handle.upload(A), allow_not_pd=True)
    assert not f.is_positive_definite()
    # x[3] duplicates x[2]: exact cancellation leaves a pivot of +-1e-17 or 0 there; the factorisation accepts
    # a pivot only above 64 eps times the original diagonal entry, so the singularity is caught AT pivot 3
    # whatever the sign of the rounding noise (a bare "> 0" test would let +1e-17 through and return garbage)
    assert f.info() == 3
    with pytest.raises(capi.AbError) as err:
        f.solve(rhs)
    assert err.value.status == 4    # AB_ERR_NOT_PD
    with pytest.raises(capi.AbError) as err:
        handle.gp_nll([capi.SE], [1.0, 1.0], ref["psd_x"], rhs)
    assert err.value.status == 4
    # IndependentNoise is value equality (noise.hpp:37-43): duplicated features share it, K stays singular
    ops, pp = prog(6)
    with pytest.raises(capi.AbError):
        handle.gp_nll(ops, pp, ref["psd_x"], rhs)
    # measurement variance on the targets (gp.hpp:65) is per observation: the fit is positive definite on both sides
    yvar = np.full(len(rhs), 0.01)
    f2, info = handle.gp_fit(ops, pp, ref["psd_x"], rhs, yvar=yvar)
    assert f2.is_positive_definite()
    assert_close(info, Restate.gp_fit(ops, pp, ref["psd_x"], rhs, yvar=yvar)["information"], RTOL, "fit with yvar")


# ---- incremental update (tests/test_gp.cc:182-219: a partial fit followed by update == a full fit) ----

@pytest.mark.parametrize("n,p", [(640, 128), (701, 150), (65, 1), (1500, 777)])
def test_update_equals_full_fit(handle, n, p):
    """ab_gp_update (gp.hpp:386-414 + BlockSymmetric, block_symmetric.hpp:46-133, as an extension of the
    Cholesky factor) against the oracle's full fit of the concatenated data."""
    ops, pp = prog(8)
    x = features(n + p, 3, 11 * n + p)
    y = targets(x)
    yvar = 0.01 + 0.02 * (np.arange(n + p) % 5)
    t = features(19, 3, 5)
    f0, info0 = handle.gp_fit(ops, pp, x[:n], y[:n], yvar=yvar[:n])
    f1, info1 = handle.gp_update(f0, ops, pp, x[:n], info0, x[n:], y[n:], yvar_new=yvar[n:])
    assert f1.n == n + p and f1.is_positive_definite()
    want = Restate.gp_fit(ops, pp, x, y, yvar=yvar)["information"]
    assert_close(info1, want, RTOL, "updated information")
    mean, var, _ = handle.gp_predict(f1, ops, pp, x, info1, t, MARGINAL)
    wm, wv, _ = Restate.gp_predict(ops, pp, x, y, t, 1, yvar=yvar)
    assert_close(mean, wm, RTOL, "mean after update")
    assert np.max(np.abs(var - wv)) <= RTOL * np.max(wv + 1.0)
    # the updated factor is an ordinary factor: the whole surface works on it
    full, _ = handle.gp_fit(ops, pp, x, y, yvar=yvar)
    assert abs(f1.log_determinant() - full.log_determinant()) <= RTOL * abs(full.log_determinant())
    rhs = np.random.default_rng(n).standard_normal((n + p, 3))
    assert_close(f1.solve(rhs), full.solve(rhs), 1e-10, "solve on the updated factor")
    # the old fit is untouched, and updates chain
    assert_close(f0.solve(y[:n]), info0, 1e-12, "old factor unchanged")
    half = p // 2
    if half > 0:
        fa, ia = handle.gp_update(f0, ops, pp, x[:n], info0, x[n:n + half], y[n:n + half], yvar_new=yvar[n:n + half])
        fb, ib = handle.gp_update(fa, ops, pp, x[:n + half], ia, x[n + half:], y[n + half:], yvar_new=yvar[n + half:])
        assert_close(ib, want, RTOL, "two chained updates")


# ---- on-disk interoperability: the packed LDLT the reference's cereal archives hold (§8f-3) ---------------

@pytest.mark.parametrize("n", [1, 64, 129, 500])
def test_import_packed_roundtrip_and_pivoted_reference_factor(handle, n):
    """ab_factor_import_packed: (i) export -> import of a device factor is bit-identical in every use;
    (ii) the PIVOTED packed factor + transpositions the reference's LDLT produces (what
    src/cereal/serializable_ldlt.hpp:18-32 writes) loads onto the device and solves like the reference."""
    A = spd(n, seed=31 * n + 7)
    scale = np.random.default_rng(n + 1).uniform(0.5, 2.0, size=n)   # unequal diagonal: the reference pivots
    A = A * scale[:, None] * scale[None, :]
    rhs = np.random.default_rng(n).standard_normal((n, 2))
    f = handle.potrf(handle.upload(A))
    LD, tr = f.export_packed()
    g = handle.import_packed(LD, tr)
    assert g.is_positive_definite() and g.n == n
    assert_close(g.solve(rhs), f.solve(rhs), 1e-13, "export -> import solve")
    assert abs(g.log_determinant() - f.log_determinant()) <= 1e-12 * max(1.0, abs(f.log_determinant()))
    # the reference's own factor: pivoted, with non-trivial transpositions
    ref = Restate.ldlt(A, rhs=rhs)
    h2 = handle.import_packed(ref["ldlt"], ref["transpositions"])
    assert_close(h2.solve(rhs), ref["solve"], RTOL, "solve from the reference's packed factor")
    assert abs(h2.log_determinant() - ref["logdet"]) <= RTOL * max(1.0, abs(ref["logdet"]))
    if n > 1:
        assert np.any(ref["transpositions"] != np.arange(n)) or n < 3  # the pivoted path was exercised
    # a non-positive D is rejected
    bad = LD.copy()
    bad[n // 2, n // 2] = -1.0
    with pytest.raises(capi.AbError) as err:
        handle.import_packed(bad, tr)
    assert err.value.status == 4
