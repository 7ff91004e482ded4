"""CPU tests: the oracle restatement is pinned against the reference's literal golden vectors
(tests/test_radial.cc, tests/test_evaluate.cc), against the committed outputs of the compiled
reference (tests/golden/ref_outputs.npz) and, when oracle/_ref is present, against the reference
itself.  These mirror the reference's own unit tests."""
import numpy as np
import pytest

from oracle.oracle import M32, M52, NOISE, PROD, SE, SUM, CONST, EXP, Ref, Restate, group_keys
from tests.helpers import GP_COVS, PARAMS, assert_close, features, prog, targets

needs_ref = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built here")


# ---- golden vectors of the reference's own tests ------------------------------------------------

def test_matern_gpytorch_tables(golden):
    """tests/test_radial.cc:339-351,482-489: |k - gpytorch| < 1e-15 on the 15x15 grid."""
    t, _ = golden
    x = np.array(t["x"])
    for op, key in ((M52, "matern52"), (M32, "matern32")):
        want = np.array(t[key]).reshape(15, 15)
        got = Restate.gram_cross([op], [t["length_scale"], t["sigma"]], x, x)
        assert np.max(np.abs(got - want)) < 1e-15
        for i in range(15):
            for j in range(15):
                v = Restate.cov_eval([op], [t["length_scale"], t["sigma"]], x[i], x[j])
                assert abs(v - want[i, j]) < 1e-15


def test_nll_known_answer(golden):
    """tests/test_evaluate.cc:34-63: scipy value 6.0946974293510134."""
    t, _ = golden
    k = t["nll_known_answer"]
    got = Restate.nll_dense(np.array(k["x"]), np.array(k["cov"]))
    assert abs(got - k["value"]) < 1e-6
    assert abs(got - k["value"]) < 1e-12


@pytest.mark.parametrize("op", [SE, EXP, M32, M52])
def test_radial_edge_cases(op):
    """tests/test_radial.cc:52-66."""
    sigma = 1.7
    p = [3.3, sigma]
    assert Restate.cov_eval([op], p, np.pi, np.pi) == sigma * sigma
    assert abs(Restate.cov_eval([op], p, np.pi, np.pi + 1e-16) - sigma * sigma) < 1e-8
    assert Restate.cov_eval([op], p, 0.0, 1e32) == 0.0
    # length_scale <= 0 -> 0 (radial.hpp:28-30)
    assert Restate.cov_eval([op], [0.0, sigma], 1.0, 2.0) == 0.0
    assert Restate.cov_eval([op], [-1.0, sigma], 1.0, 1.0) == 0.0


def test_sum_product_noise_semantics():
    """tests/test_covariance_functions.cc:33-93 (sum, product, value-equality noise)."""
    x, y = np.array([1.0, 2.0, 3.0]), np.array([1.0, 2.5, 3.0])
    se = Restate.cov_eval([SE], [2.0, 1.5], x, y)
    m52 = Restate.cov_eval([M52], [3.0, 0.7], x, y)
    assert Restate.cov_eval([SE, M52, SUM], [2.0, 1.5, 3.0, 0.7, 0, 0], x, y) == se + m52
    assert Restate.cov_eval([SE, M52, PROD], [2.0, 1.5, 3.0, 0.7, 0, 0], x, y) == se * m52
    assert Restate.cov_eval([NOISE], [0.3, 0], x, y) == 0.0
    assert Restate.cov_eval([NOISE], [0.3, 0], x, x.copy()) == 0.3 * 0.3
    # product short-circuit: 0 * inf stays 0 (covariance_function.hpp:362-366)
    assert Restate.cov_eval([NOISE, CONST, PROD], [0.3, 0, np.inf, 0, 0, 0], x, y) == 0.0
    assert np.isnan(Restate.cov_eval([CONST, NOISE, PROD], [np.inf, 0, 0.3, 0, 0, 0], x, y))


# ---- restatement vs the committed outputs of the compiled reference -----------------------------

def test_gram_vs_reference_fixture(golden):
    _, ref = golden
    for cid in PARAMS:
        ops, pp = prog(cid)
        assert_close(Restate.gram_sym(ops, pp, ref["x1"]), ref[f"gram1_{cid}"], 2e-15, f"gram1 {cid}")
        assert_close(Restate.gram_sym(ops, pp, ref["x3"]), ref[f"gram3_{cid}"], 2e-15, f"gram3 {cid}")
        assert_close(Restate.gram_cross(ops, pp, ref["x3"][:30], ref["x3"][25:70]),
                     ref[f"cross3_{cid}"], 2e-15, f"cross3 {cid}")


@pytest.mark.parametrize("cid", GP_COVS)
@pytest.mark.parametrize("tag", ["1", "3"])
def test_gp_vs_reference_fixture(golden, cid, tag):
    _, ref = golden
    ops, pp = prog(cid)
    x, y, t = ref[f"gp_x{tag}"], ref[f"gp_y{tag}"], ref[f"gp_test{tag}"]
    assert_close(Restate.gp_fit(ops, pp, x, y)["information"], ref[f"info{tag}_{cid}"], 1e-10)
    assert abs(Restate.gp_nll(ops, pp, x, y) - float(ref[f"nll{tag}_{cid}"])) <= 1e-10 * abs(
        float(ref[f"nll{tag}_{cid}"]))
    assert_close(Restate.gp_predict(ops, pp, x, y, t, 0)[0], ref[f"mean{tag}_{cid}"], 1e-10)
    assert_close(Restate.gp_predict(ops, pp, x, y, t, 1)[1], ref[f"var{tag}_{cid}"], 1e-9)
    assert_close(Restate.gp_predict(ops, pp, x, y, t, 2)[2], ref[f"cov{tag}_{cid}"], 1e-9)
    m, v, _, s = Restate.gp_cv(ops, pp, x, y, group_keys(x, 0), what=1, want_score=True)
    assert_close(m, ref[f"loo_mean{tag}_{cid}"], 1e-10)
    assert_close(v, ref[f"loo_var{tag}_{cid}"], 1e-10)
    assert abs(s - float(ref[f"loo_score{tag}_{cid}"])) <= 1e-10 * abs(float(ref[f"loo_score{tag}_{cid}"]))
    keys = group_keys(x, 1, 8)
    m, v, _, s = Restate.gp_cv(ops, pp, x, y, keys, what=1, want_score=True)
    assert_close(m, ref[f"logo_mean{tag}_{cid}"], 1e-10)
    assert_close(v, ref[f"logo_var{tag}_{cid}"], 1e-10)
    assert abs(s - float(ref[f"logo_score{tag}_{cid}"])) <= 1e-10 * abs(float(ref[f"logo_score{tag}_{cid}"]))
    _, _, j, _ = Restate.gp_cv(ops, pp, x, y, keys, what=2)
    assert_close(j, ref[f"logo_joint{tag}_{cid}"], 1e-9)


def test_fit_adds_target_variance_fixture(golden):
    """gp.hpp:65 vs gp.hpp:447-448."""
    _, ref = golden
    ops, pp = prog(6)
    x, y, yvar = ref["gp_x1"], ref["gp_y1"], ref["gp_yvar"]
    assert_close(Restate.gp_fit(ops, pp, x, y, yvar=yvar)["information"], ref["info1_6_yvar"], 1e-10)
    assert_close(Restate.gp_predict(ops, pp, x, y, ref["gp_test1"], 1, yvar=yvar)[1],
                 ref["var1_6_yvar"], 1e-9)


def test_ldlt_wrapper_fixture(golden):
    """tests/test_serializable_ldlt.cc:34-85 identities, against the compiled reference."""
    _, ref = golden
    A, rhs = ref["ldlt_A"], ref["ldlt_rhs"]
    l = Restate.ldlt(A, rhs=rhs, want_inverse_diagonal=True)
    assert_close(l["solve"], ref["ldlt_solve"], 1e-11)
    assert_close(l["sqrt_solve"], ref["ldlt_sqrt_solve"], 1e-11)
    assert abs(l["logdet"] - float(ref["ldlt_logdet"])) < 1e-10 * abs(float(ref["ldlt_logdet"]))
    assert_close(l["inverse_diagonal"], ref["ldlt_inverse_diagonal"], 1e-10)
    groups = [[0, 5, 9], [1], [100, 101, 102, 127], [64, 63]]
    got = np.concatenate([b.ravel(order="F") for b in Restate.inverse_blocks(A, groups)])
    assert_close(got, ref["ldlt_inverse_blocks"], 1e-10)
    # identities: inverse_diagonal == diag(inv(A)); sqrt_solve^T sqrt_solve == rhs^T A^-1 rhs
    assert_close(l["inverse_diagonal"], np.diag(np.linalg.inv(A)), 1e-8)
    assert_close(l["sqrt_solve"].T @ l["sqrt_solve"], rhs.T @ np.linalg.solve(A, rhs), 1e-10)


def test_integer_contract_fixture(golden):
    """Bit-exact group/fold indices."""
    _, ref = golden
    x = ref["gp_x1"]
    for tag, gk, ga in (("grp", 1, 8), ("grp2", 2, 3.7)):
        keys, offsets, indices = Restate.group_indexers(group_keys(x, gk, ga))
        assert np.array_equal(keys, ref[f"{tag}_keys"])
        assert np.array_equal(offsets, ref[f"{tag}_offsets"])
        assert np.array_equal(indices, ref[f"{tag}_indices"])
    for n, k in ((100, 7), (1000, 8), (32768, 32), (5, 8)):
        assert np.array_equal(Restate.partition_triangular(n, k), ref[f"ptri_{n}_{k}"])
    assert np.array_equal(Restate.indices_complement([3, 1, 7, 7, 12], 15), ref["complement"])


def test_partition_triangular_properties():
    """tests/test_indexing.cc:341-369: contiguous, covering, area-balanced."""
    for n, k in ((1000, 8), (4096, 32), (17, 3)):
        b = Restate.partition_triangular(n, k)
        assert b[0, 0] == 0 and b[-1, 1] == n
        assert np.all(b[1:, 0] == b[:-1, 1])


def test_sparse_gp_fixture(golden):
    _, ref = golden
    ops, pp = prog(6)
    x, y, u, t = ref["sp_x"], ref["sp_y"], ref["sp_u"], ref["sp_test"]
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        keys = group_keys(x, gk, ga)
        r = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=2, want_ll=True)
        assert_close(r["mean"], ref[f"sp_{tag}_mean"], 1e-9)
        assert_close(r["cov"], ref[f"sp_{tag}_cov"], 1e-8)
        assert abs(r["ll"] - float(ref[f"sp_{tag}_ll"])) < 1e-8 * abs(float(ref[f"sp_{tag}_ll"]))
        r = Restate.sparse_gp(ops, pp, x, y, u, keys, test=t, what=1)
        assert_close(r["var"], ref[f"sp_{tag}_var"], 1e-8)


def test_sparse_ll_close_to_dense():
    """tests/test_sparse_gp.cc:172-221: sparse log-likelihood approximates the dense one.  The
    reference test wraps its noise in measurement_only(); with a plain IndependentNoise the noise
    also lands on K_uu, so the approximation is looser here (10% instead of 1%)."""
    ops, pp = prog(6)
    x = features(300, 1, 3).ravel()
    y = targets(x)
    u = Restate.linspace(x.min(), x.max(), 60)
    sparse = Restate.sparse_gp(ops, pp, x, y, u, group_keys(x, 0), want_ll=True)["ll"]
    dense = -Restate.gp_nll(ops, pp, x, y)
    assert abs(sparse - dense) < 1e-1 * abs(dense)


# ---- restatement vs the live compiled reference (build container only) --------------------------

@needs_ref
def test_live_reference_gram_threaded_equals_serial():
    """tests/test_callers.cc:225-267: threaded Gram == serial.  The reference asserts bit equality
    under its own build flags; with -O3 -march=x86-64-v3 (FMA contraction) gcc emits different code
    for the two loops and they agree to 2 ulp, which is what is checked here."""
    x = Ref.random_features(257, 3, 9)
    serial = Ref.gram_sym(7, PARAMS[7], x, nthreads=1)
    for t in (2, 5, 8):
        threaded = Ref.gram_sym(7, PARAMS[7], x, nthreads=t)
        assert np.array_equal(threaded, threaded.T)
        assert_close(threaded, serial, 2e-15)
    ops, pp = prog(7)
    assert_close(Restate.gram_sym(ops, pp, x), serial, 2e-15)


@needs_ref
def test_live_reference_measurement_wrapping_is_transparent():
    """gp.hpp:288 wraps training features in Measurement<>; in-scope leaves ignore the wrapper."""
    x = Ref.random_features(64, 1, 2)
    x[5] = x[9]  # duplicate feature -> off-diagonal noise (noise.hpp:37-43)
    a = Ref.gram_sym(6, PARAMS[6], x, as_meas=True)
    b = Ref.gram_sym(6, PARAMS[6], x, as_meas=False)
    assert np.array_equal(a, b)
    assert a[5, 9] == a[5, 5]
    ops, pp = prog(6)
    assert_close(Restate.gram_sym(ops, pp, x), a, 2e-15)


@needs_ref
def test_live_reference_pivots_and_factor():
    x = Ref.random_features(200, 3, 4)
    y = Ref.random_targets(x)
    ops, pp = prog(8)
    a = Restate.gp_fit(ops, pp, x, y, want_factor=True)
    b = Ref.gp_fit(8, PARAMS[8], x, y, want_factor=True)
    assert np.array_equal(a["transpositions"], b["transpositions"])
    assert_close(np.tril(a["ldlt"]), np.tril(b["ldlt"]), 1e-11)
    assert_close(a["information"], b["information"], 1e-11)


@needs_ref
def test_live_reference_generators_match_numpy_port():
    """std::mt19937 + uniform_real_distribution as used by bench_utils.h, re-derived in numpy."""
    n, seed = 50, 7
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2 ** 32, size=2 * n, dtype=np.uint64).astype(np.float64)
    u = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    want = Ref.random_features(n, 1, seed).ravel()
    assert_close(u * 10.0, want, 1e-15)


# ---- Polynomial<order> (polynomials.hpp:63-90): the sinc example's covariance, menu entry 11 ---------------

P11 = [3.0, 0.7, 3.5, 5.7, 0.4]


def test_polynomial_fixture(golden):
    """Polynomial<1> + SE + measurement_only(noise) (examples/sinc_example.cc:84-87) — the restatement against
    the outputs of the compiled reference: both pairings of the Gram matrix, the exact GP and the sparse GP."""
    from oracle.oracle import menu_program, menu_program_plain
    _, ref = golden
    x, y, t, u = ref["poly_x"], ref["poly_y"], ref["poly_test"], ref["poly_u"]
    meas, plain = menu_program(11, P11), menu_program_plain(11, P11)
    assert_close(Restate.gram_sym(*meas, x[:40]), ref["poly_gram_meas"], 1e-13)
    assert_close(Restate.gram_sym(*plain, x[:40]), ref["poly_gram_plain"], 1e-13)
    assert_close(Restate.gram_cross(*plain, x[:40], t), ref["poly_gram_cross"], 1e-13)
    # sigma_0^2 + sigma_1^2 x y on top of the radial part
    i, j = 3, 17
    want = P11[0] ** 2 + P11[1] ** 2 * x[i] * x[j] + Restate.cov_eval([SE], P11[2:4], x[i], x[j])
    assert abs(ref["poly_gram_plain"][i, j] - want) <= 1e-13 * abs(want)
    assert_close(Restate.gp_fit(*meas, x, y)["information"], ref["poly_information"], 1e-8)
    nll = Restate.gp_nll(*meas, x, y)
    assert abs(nll - float(ref["poly_nll"])) <= 1e-10 * abs(float(ref["poly_nll"]))
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        r = Restate.sparse_gp(*meas, x, y, u, group_keys(x, gk, ga), test=t, what=2, want_ll=True,
                              fu=plain, uu=plain)
        assert_close(r["mean"], ref[f"poly_sp_{tag}_mean"], 1e-7)
        want = float(ref[f"poly_sp_{tag}_ll"])
        assert abs(r["ll"] - want) <= 1e-8 * abs(want)


@needs_ref
def test_live_reference_polynomial():
    from oracle.oracle import menu_program, menu_program_plain
    x = features(60, 1, 77).ravel() - 4.0
    got = Restate.gram_sym(*menu_program(11, P11), x)
    assert_close(got, Ref.gram_sym(11, P11, x, as_meas=True), 1e-14)
    assert_close(Restate.gram_sym(*menu_program_plain(11, P11), x), Ref.gram_sym(11, P11, x), 1e-14)
