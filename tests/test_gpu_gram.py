"""GPU parity: Gram construction through the C ABI vs the oracle (restatement, committed reference
fixtures and the reference's literal golden vectors).  fp64 tolerance: 1e-9 relative is the
north-star bar; the Gram entries themselves are held to a few ulp (4e-15 relative to the matrix
scale) and the gpytorch tables to the reference test's own 1e-15 absolute."""
import numpy as np
import pytest

from albatross_b200 import capi
from albatross_b200.capi import CONST, EXP, M32, M52, NOISE, PROD, SE, SUM
from oracle.oracle import Restate
from tests.helpers import PARAMS, assert_close, features, prog

pytestmark = pytest.mark.gpu
ULP_TOL = 4e-15


@pytest.mark.parametrize("cid", sorted(PARAMS))
@pytest.mark.parametrize("dim", [1, 2, 3, 5, 8])
def test_gram_sym_matches_oracle(handle, cid, dim):
    ops, pp = prog(cid)
    for n in (1, 63, 64, 65, 200):
        x = features(n, dim, seed=n + dim)
        got = handle.gram_sym(ops, pp, x).download()
        want = Restate.gram_sym(ops, pp, x)
        assert_close(got, want, ULP_TOL, f"cov {cid} dim {dim} n {n}")
        assert np.array_equal(got, got.T), "Gram must be exactly symmetric"


@pytest.mark.parametrize("cid", sorted(PARAMS))
def test_gram_cross_and_diag_match_oracle(handle, cid):
    ops, pp = prog(cid)
    for dim in (1, 3):
        x, y = features(130, dim, 1), features(77, dim, 2)
        y[:5] = x[10:15]  # shared points: value-equality noise must fire off the diagonal
        assert_close(handle.gram_cross(ops, pp, x, y).download(), Restate.gram_cross(ops, pp, x, y),
                     ULP_TOL, f"cross cov {cid} dim {dim}")
        assert_close(handle.gram_diag(ops, pp, x), Restate.gram_diag(ops, pp, x), ULP_TOL)


def test_gram_vs_reference_fixture(handle, golden):
    _, ref = golden
    for cid in PARAMS:
        ops, pp = prog(cid)
        assert_close(handle.gram_sym(ops, pp, ref["x1"]).download(), ref[f"gram1_{cid}"], ULP_TOL)
        assert_close(handle.gram_sym(ops, pp, ref["x3"]).download(), ref[f"gram3_{cid}"], ULP_TOL)
        assert_close(handle.gram_cross(ops, pp, ref["x3"][:30], ref["x3"][25:70]).download(),
                     ref[f"cross3_{cid}"], ULP_TOL)


def test_matern_gpytorch_tables(handle, golden):
    """The reference's own golden vectors (tests/test_radial.cc:212-489), 1e-15 absolute."""
    t, _ = golden
    x = np.array(t["x"])
    for op, key in ((M52, "matern52"), (M32, "matern32")):
        want = np.array(t[key]).reshape(15, 15)
        p = [t["length_scale"], t["sigma"]]
        assert np.max(np.abs(handle.gram_cross([op], p, x, x).download() - want)) < 1e-15
        assert np.max(np.abs(handle.gram_sym([op], p, x).download() - want)) < 1e-15


@pytest.mark.parametrize("op", [SE, EXP, M32, M52])
def test_radial_edge_cases(handle, op):
    """tests/test_radial.cc:52-66 on the device."""
    sigma = 1.7
    p = [3.3, sigma]
    x = np.array([np.pi, np.pi + 1e-16, 0.0, 1e32])
    k = handle.gram_sym([op], p, x).download()
    assert k[0, 0] == sigma * sigma
    assert abs(k[0, 1] - sigma * sigma) < 1e-8
    assert k[2, 3] == 0.0
    assert np.all(handle.gram_sym([op], [0.0, sigma], x).download() == 0.0)
    assert np.all(handle.gram_sym([op], [-2.0, sigma], x).download() == 0.0)


@pytest.mark.parametrize("op", [SE, EXP, M32, M52])
def test_nan_length_scale_propagates(handle, op):
    """radial.hpp:28-30 tests `length_scale <= 0.`, which is false for NaN: a corrupt hyper-parameter must
    surface as NaN covariances (and then as a not-positive-definite fit), never as a silent zero."""
    x = features(130, 3, 12)
    k = handle.gram_sym([op], [np.nan, 1.3], x).download()
    assert np.all(np.isnan(k))
    want = Restate.gram_sym([op], [np.nan, 1.3], x)
    assert np.all(np.isnan(want))
    k = handle.gram_sym([op, NOISE, SUM], [np.nan, 1.3, 0.1, 0, 0, 0], x[:, :1]).download()
    assert np.all(np.isnan(k))


def test_noise_is_value_equality(handle):
    """noise.hpp:37-43: duplicated features get off-diagonal noise; all coordinates must match."""
    x = features(70, 3, 5)
    x[7] = x[3]
    x[9, :2] = x[4, :2]  # only two of three coordinates equal -> not equal
    k = handle.gram_sym([NOISE], [0.3, 0], x).download()
    want = Restate.gram_sym([NOISE], [0.3, 0], x)
    assert np.array_equal(k, want)
    assert k[7, 3] == 0.3 * 0.3 and k[9, 4] == 0.0


def test_product_short_circuit_and_nesting(handle):
    """covariance_function.hpp:362-366 and programs that need the generic stack evaluator."""
    x = features(66, 2, 8)
    inf = np.inf
    # 0 * inf stays 0; inf * 0 is NaN (lhs evaluated first)
    k = handle.gram_sym([NOISE, CONST, PROD], [0.3, 0, inf, 0, 0, 0], x).download()
    assert k[0, 1] == 0.0 and np.isinf(k[0, 0])
    k = handle.gram_sym([CONST, NOISE, PROD], [inf, 0, 0.3, 0, 0, 0], x).download()
    assert np.isnan(k[0, 1])
    # (SE + M52) * EXP   and   SE * (M32 + CONST): nested sums inside products
    for ops, pp in (([SE, M52, SUM, EXP, PROD], [2, 1.5, 3, .7, 0, 0, 1.1, .9, 0, 0]),
                    ([SE, M32, CONST, SUM, PROD], [2, 1.5, 3, .7, 1.2, 0, 0, 0, 0, 0]),
                    ([SE, M32, EXP, CONST, NOISE, SUM, SUM, SUM, SUM],
                     [2, 1.5, 3, .7, 1.1, .9, 1.2, 0, .3, 0, 0, 0, 0, 0, 0, 0, 0, 0])):
        assert_close(handle.gram_sym(ops, pp, x).download(), Restate.gram_sym(ops, pp, x), ULP_TOL)
        assert_close(handle.gram_cross(ops, pp, x, x[:9]).download(),
                     Restate.gram_cross(ops, pp, x, x[:9]), ULP_TOL)


def test_lower_only_flag_and_empty(handle):
    ops, pp = prog(7)
    x = features(150, 3, 4)
    full = handle.gram_sym(ops, pp, x).download()
    low = handle.gram_sym(ops, pp, x, flags=capi.GRAM_LOWER_ONLY).download()
    assert np.array_equal(np.tril(low), np.tril(full))
    assert handle.gram_sym(ops, pp, np.zeros((0, 3))).shape == (0, 0)
    assert handle.gram_cross(ops, pp, x, np.zeros((0, 3))).shape == (150, 0)


def test_malformed_programs_are_rejected(handle):
    x = features(4, 1, 0)
    for ops, pp in (([SUM], [0, 0]), ([SE, SE], [1, 1, 1, 1]), ([SE, SUM], [1, 1, 0, 0]),
                    ([42], [1, 1])):
        with pytest.raises(capi.AbError) as e:
            handle.gram_sym(ops, pp, x)
        assert e.value.status == 1
    with pytest.raises(capi.AbError) as e:
        handle.gram_sym([SE], [1, 1], features(4, 9, 0))
    assert e.value.status in (1, 6)


def test_gram_device_resident_large_sampled(handle):
    """BASELINE config 2 shape at reduced n: full compare at 2048, sampled rows at 8192."""
    ops, pp = prog(7)
    x = features(2048, 3, 0)
    assert_close(handle.gram_sym(ops, pp, x).download(), Restate.gram_sym(ops, pp, x), ULP_TOL)
    x = features(8192, 3, 1)
    fd = handle.upload_features(x)
    K = handle.gram_sym_d(ops, pp, fd)
    rows = np.array([0, 1, 63, 64, 4095, 4096, 8191])
    for r in rows:
        got = K.download_block(int(r), 0, 1, 8192).ravel()
        want = Restate.gram_cross(ops, pp, x[r:r + 1], x).ravel()
        assert_close(got, want, ULP_TOL, f"row {r}")
        col = K.download_block(0, int(r), 8192, 1).ravel()
        assert np.array_equal(col, got), "mirror must be bit-identical"
