"""GPU parity for Polynomial<order> covariance terms (polynomials.hpp:63-90): the covariance of the reference's
sinc example, Polynomial<1> + SE + measurement_only(IndependentNoise) (examples/sinc_example.cc:84-87; menu
entry 11 of the oracle), through the C ABI — Gram matrices in both pairings, exact GP, LOO and the sparse GP —
against the outputs of the compiled reference (tests/golden/make_golden_r2.py, section `poly`) and against
the restatement on seeded inputs.  1e-9 relative."""
import numpy as np
import pytest

from albatross_b200 import capi
from albatross_b200.capi import GRAM_FULL, GRAM_LOWER_ONLY, JOINT, MARGINAL, POLY, PROD, SE, SUM
from oracle.oracle import Ref, Restate, group_keys, menu_program, menu_program_plain
from tests.helpers import assert_close, features

pytestmark = pytest.mark.gpu
P11 = [3.0, 0.7, 3.5, 5.7, 0.4]
MEAS = menu_program(11, P11)
PLAIN = menu_program_plain(11, P11)


def test_gram_fixture(handle, golden):
    _, ref = golden
    x, t = ref["poly_x"][:40], ref["poly_test"]
    assert_close(handle.gram_sym(*MEAS, x).download(), ref["poly_gram_meas"], 1e-13)
    assert_close(handle.gram_sym(*PLAIN, x).download(), ref["poly_gram_plain"], 1e-13)
    assert_close(handle.gram_cross(*PLAIN, x, t).download(), ref["poly_gram_cross"], 1e-13)
    assert_close(handle.gram_diag(*MEAS, x), np.diag(ref["poly_gram_meas"]), 1e-13)
    assert_close(handle.gram_diag(*PLAIN, x), np.diag(ref["poly_gram_plain"]), 1e-13)


@pytest.mark.parametrize("n,m", [(1, 1), (63, 5), (129, 200), (1000, 333), (2500, 64)])
def test_gram_vs_oracle(handle, n, m):
    """Tile edges and both outputs of the symmetric build (mirrored / lower only)."""
    x = features(n, 1, 5 * n).ravel() - 5.0
    y = features(m, 1, 7 * m + 1).ravel() - 5.0
    want = Restate.gram_sym(*MEAS, x)
    got = handle.gram_sym(*MEAS, x, GRAM_FULL).download()
    assert_close(got, want, 1e-13, f"sym n={n}")
    assert np.array_equal(got, got.T)
    low = handle.gram_sym(*MEAS, x, GRAM_LOWER_ONLY).download()
    assert np.array_equal(np.tril(low), np.tril(got))
    assert_close(handle.gram_cross(*PLAIN, x, y).download(), Restate.gram_cross(*PLAIN, x, y), 1e-13, "cross")
    assert_close(handle.gram_diag(*MEAS, x), np.diag(want), 1e-13, "diag")


def test_higher_orders_and_pure_polynomial(handle):
    """Polynomial<3> alone (no stationary part: the device program is 0 + the four terms), degrees 0..3."""
    x = np.linspace(-2.0, 2.0, 150)
    sig = [1.5, 0.9, 0.4, 0.2]
    ops = [POLY, POLY, SUM, POLY, SUM, POLY, SUM]
    pp = [sig[0], 0.0, sig[1], 1.0, 0, 0, sig[2], 2.0, 0, 0, sig[3], 3.0, 0, 0]
    want = sum(s * s * np.outer(x ** p, x ** p) for p, s in enumerate(sig))
    assert_close(handle.gram_sym(ops, pp, x).download(), want, 1e-13)
    assert_close(Restate.gram_sym(ops, pp, x), want, 1e-13)
    assert_close(handle.gram_cross(ops, pp, x, x[:7]).download(), want[:, :7], 1e-13)


def test_unsupported_forms_fail_loudly(handle):
    x = np.linspace(0.0, 1.0, 10)
    with pytest.raises(capi.AbError) as e:   # inside a product: no device form, no fallback
        handle.gram_sym([POLY, SE, PROD], [1.0, 1.0, 1.0, 1.0, 0, 0], x)
    assert "UNSUPPORTED" in str(e.value)
    with pytest.raises(capi.AbError) as e:   # Polynomial is defined between doubles only (polynomials.hpp:79)
        handle.gram_sym(*PLAIN, np.zeros((10, 3)))
    assert "UNSUPPORTED" in str(e.value)
    with pytest.raises(capi.AbError):        # non-integer degree
        handle.gram_sym([POLY], [1.0, 1.5], x)


def test_exact_gp_fixture(handle, golden):
    _, ref = golden
    x, y, t = ref["poly_x"], ref["poly_y"], ref["poly_test"]
    f, info = handle.gp_fit(*MEAS, x, y)
    assert_close(info, ref["poly_information"], 1e-9, "information")
    nll = handle.gp_nll(*MEAS, x, y)
    assert abs(nll - float(ref["poly_nll"])) <= 1e-9 * abs(float(ref["poly_nll"]))
    scale = np.max(np.abs(ref["poly_cov"]))
    mean, var, _ = handle.gp_predict(f, *PLAIN, x, info, t, MARGINAL)
    assert_close(mean, ref["poly_mean"], 1e-9, "mean")
    assert np.max(np.abs(var - ref["poly_var"])) <= 1e-9 * scale
    _, _, cov = handle.gp_predict(f, *PLAIN, x, info, t, JOINT)
    assert np.max(np.abs(cov - ref["poly_cov"])) <= 1e-9 * scale
    # leave-one-out, pure LOO route (cross_validation_utils.hpp:151-177)
    _, offsets, indices = capi.group_indexers(np.arange(len(x)))
    m, v, _, s = handle.gp_cv(f, y, info, offsets, indices, MARGINAL, want_score=True)
    assert_close(m, ref["poly_loo_mean"], 1e-9, "loo mean")
    assert_close(v, ref["poly_loo_var"], 1e-9, "loo var")
    assert abs(s - float(ref["poly_loo_score"])) <= 1e-9 * abs(float(ref["poly_loo_score"]))
    f.free()


def test_sparse_gp_fixture(handle, golden):
    _, ref = golden
    x, y, t, u = ref["poly_x"], ref["poly_y"], ref["poly_test"], ref["poly_u"]
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        _, offsets, indices = capi.group_indexers(group_keys(x, gk, ga))
        f, info, ll = handle.sparse_fit(*MEAS, x, y, u, offsets, indices, fu=PLAIN, uu=PLAIN)
        mean, _, cov = f.predict(*PLAIN, t, JOINT)
        want = float(ref[f"poly_sp_{tag}_ll"])
        assert abs(ll - want) <= 1e-9 * abs(want), (tag, ll, want)
        assert_close(mean, ref[f"poly_sp_{tag}_mean"], 1e-9, f"{tag} mean")
        # prior - Q** + S**: the error scales with the PRIOR covariance at the test points
        prior = np.max(np.abs(Restate.gram_sym(*PLAIN, t)))
        assert np.max(np.abs(cov - ref[f"poly_sp_{tag}_cov"])) <= 1e-9 * prior, tag
        f.free()


def test_exact_gp_vs_live_reference(handle):
    if not Ref.available():
        pytest.skip("oracle/_ref not shipped")
    n = 1200
    x = features(n, 1, 99).ravel() - 2.0
    y = 1.0 + 0.3 * x + 4.0 * np.sinc(0.25 * x)
    f, info = handle.gp_fit(*MEAS, x, y)
    assert_close(info, Ref.gp_fit(11, P11, x, y)["information"], 1e-9)
    t = np.linspace(-3.0, 9.0, 17)
    mean, var, _ = Ref.gp_predict(11, P11, x, y, t, 1)
    got_mean, got_var, _ = handle.gp_predict(f, *PLAIN, x, info, t, MARGINAL)
    assert_close(got_mean, mean, 1e-9)
    prior = np.max(np.abs(Restate.gram_sym(*PLAIN, t)))
    assert np.max(np.abs(got_var - var)) <= 1e-9 * prior
    f.free()
