"""Round-2 golden fixtures, generated from the reference itself like make_golden.py (run in the BUILD
container only: needs /root/reference and oracle/_ref/libref_oracle.so):

    python tests/golden/make_golden_r2.py [section ...]

Writes tests/golden/ref_outputs_r2.npz (merged into the `golden` fixture by tests/conftest.py).  Sections:

  sparse_mo   sparse GP whose covariance holds a MeasurementOnly term — the reference's standard sparse
              configuration (tests/lib/albatross/test/test_models.h:26-30, tests/test_sparse_gp.cc:118-377):
              menu entry 10 on the bench-shaped data of make_golden.py, FITC and PITC, with and without
              measurement variance, plus make_simple_covariance_function() itself (SE(100, 100) +
              measurement_only(IndependentNoise(0.1))) on make_toy_linear_data() with 25 uniformly spaced
              inducing points (MakeSparseGaussianProcess, test_models.h:44-57).
  ldlt        the rest of the SerializableLDLT surface (src/eigen/serializable_ldlt.hpp:58-126) on the
              bench-style PSD matrix of make_golden.py, and what the reference's pivoted LDLT returns for a
              singular PSD matrix (duplicate points without noise; LDLT.h:316-338, :568-585).
  big         exact GP at N = 8192 (3-D, SE + IndependentNoise; the default look-ahead schedule of the device
              factorisation starts at this size): information, NLL, predictions, LOO — about 3 minutes of
              single-core Eigen LDLT.

  poly        Polynomial<1> + SE + measurement_only(IndependentNoise) — the covariance of the reference's sinc
              example (examples/sinc_example.cc:84-87), menu entry 11: Gram matrices in both pairings, exact GP
              and sparse GP on sinc-shaped data with a linear trend.

Sections not named on the command line keep their previous arrays.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import Ref  # noqa: E402

OUT = os.path.join(HERE, "ref_outputs_r2.npz")
P10 = [1.0, 1.0, 0.1]


def toy_linear_data(n=10, a=5.0, b=1.0, sigma=0.1):
    """make_toy_linear_data, tests/lib/albatross/test/test_utils.h:41-59: x = 0..n-1,
    y = a + b x + N(0, sigma) from std::mt19937(3) (libstdc++'s normal_distribution scales the same
    N(0, 1) stream)."""
    x = np.arange(n, dtype=np.float64)
    return x, a + b * x + sigma * Ref.random_normal(n, 3)


def section_sparse_mo(out):
    xs = Ref.random_features(400, 1, 5).ravel()
    ys = Ref.random_targets(xs)
    u = Ref.uniform_inducing_points(xs, 20)
    ts = np.linspace(0.2, 9.8, 13)
    yvar = 0.01 + 0.02 * (np.arange(400) % 5)
    out["spmo_x"], out["spmo_y"], out["spmo_u"], out["spmo_test"], out["spmo_yvar"] = xs, ys, u, ts, yvar
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        for vtag, yv in (("", None), ("_yvar", yvar)):
            r = Ref.sparse_gp(10, P10, xs, ys, u, gk, ga, test=ts, what=2, want_ll=True, yvar=yv)
            k = f"spmo_{tag}{vtag}"
            out[f"{k}_mean"], out[f"{k}_cov"], out[f"{k}_ll"] = r["mean"], r["cov"], np.array(r["ll"])
            out[f"{k}_var"] = Ref.sparse_gp(10, P10, xs, ys, u, gk, ga, test=ts, what=1, yvar=yv)["var"]
    # the reference's own sparse test model on its own toy data
    x, y = toy_linear_data()
    u = Ref.uniform_inducing_points(x, 25)
    t = np.linspace(-1.0, 10.5, 24)
    p = [100.0, 100.0, 0.1]
    out["sptoy_x"], out["sptoy_y"], out["sptoy_u"], out["sptoy_test"] = x, y, u, t
    r = Ref.sparse_gp(10, p, x, y, u, 0, 0.0, test=t, what=2, want_ll=True)
    out["sptoy_mean"], out["sptoy_cov"], out["sptoy_ll"] = r["mean"], r["cov"], np.array(r["ll"])
    out["sptoy_var"] = Ref.sparse_gp(10, p, x, y, u, 0, 0.0, test=t, what=1)["var"]


def section_ldlt(out):
    """What the reference's diagonally pivoted LDLT does with a singular positive SEMI-definite matrix:
    duplicate points without a noise term (LDLT.h:316-338 pivots, :568-585 pseudo-inverse of D)."""
    x = np.array([0.0, 1.0, 2.0, 2.0, 3.0, 1.0, 4.5, 0.0])
    A = Ref.gram_sym(0, [1.0, 1.0], x)           # SE(1, 1), rank 5 of 8
    rhs = np.sin(x)                               # in the range of A (equal rows carry equal values)
    l = Ref.ldlt(A, rhs=rhs)
    out["psd_x"], out["psd_A"], out["psd_rhs"] = x, A, rhs
    out["psd_D"], out["psd_transpositions"] = l["D"], l["transpositions"]
    out["psd_solve"] = l["solve"].ravel()
    out["psd_logdet"] = np.array(l["logdet"])
    out["psd_is_pd"] = np.array(int(l["is_pd"]))
    out["psd_nll"] = np.array(Ref.gp_nll(6, [1.0, 1.0, 0.0], x, rhs)[0])  # SE + IndependentNoise(0)


def section_sparse_nested(out):
    """The nested sub-problem of BASELINE configs[4] the reference can still run (about 90 s per fit on one
    core): N = 131 072 bench-shaped observations, M = 512 uniformly spaced inducing points, FITC and PITC
    with 1024-point groups.  Inputs come from numpy's PCG64 (seed 0): only checksums are stored."""
    n, m = 131072, 512
    x = np.random.default_rng(0).uniform(0.0, 10.0, size=n)
    y = np.sin(x) + 0.1 * np.cos(10.0 * x)
    u = np.linspace(x.min(), x.max(), m)
    t = np.linspace(0.0, 10.0, 64)
    out["spn_n_m"] = np.array([n, m])
    out["spn_x_checksum"] = np.array([x.sum(), x[12345], y.sum()])
    out["spn_test"] = t
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, n / 10.0 / 1024.0)):
        r = Ref.sparse_gp(6, [1.0, 1.0, 0.1], x, y, u, gk, ga, test=t, what=1, want_ll=True)
        out[f"spn_{tag}_mean"], out[f"spn_{tag}_var"] = r["mean"], r["var"]
        out[f"spn_{tag}_ll"] = np.array(r["ll"])
        print("  sparse_nested", tag, "ll", r["ll"], flush=True)


def section_big(out):
    """Exact GP at N = 8192 (3-D, SE(1, 1) + IndependentNoise(0.1)): the smallest size at which the device
    factorisation takes its default look-ahead schedule.  Each reference pass is ~45 s of single-core LDLT."""
    n = 8192
    x = np.random.default_rng(8192).uniform(0.0, 10.0, size=(n, 3))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10.0 * x[:, 0])
    t = np.random.default_rng(1).uniform(0.0, 10.0, size=(48, 3))
    p = [1.0, 1.0, 0.1]
    out["big_x_checksum"] = np.array([x.sum(), x[4321, 1], y.sum()])
    out["big_test"] = t
    out["big_information"] = Ref.gp_fit(6, p, x, y)["information"]
    print("  big: fit done", flush=True)
    out["big_nll"] = np.array(Ref.gp_nll(6, p, x, y)[0])
    mean, _, cov = Ref.gp_predict(6, p, x, y, t, 2)
    out["big_mean"], out["big_cov"] = mean, cov
    out["big_var"] = Ref.gp_predict(6, p, x, y, t, 1)[1]
    print("  big: predictions done", flush=True)
    m, v, _, s = Ref.gp_cv(6, p, x, y, 0, 0.0, what=1, want_score=True)
    out["big_loo_mean"], out["big_loo_var"], out["big_loo_score"] = m, v, np.array(s)


P11 = [3.0, 0.7, 3.5, 5.7, 0.4]   # sigma_polynomial_0, _1, SE length scale, SE sigma, noise sigma


def section_poly(out):
    x = Ref.random_features(300, 1, 21).ravel() - 3.0      # (-3, 7): both signs, as in the sinc example
    y = 2.0 + 0.6 * x + 5.0 * np.sinc(0.3 * (x - 1.0)) + 0.4 * Ref.random_normal(300, 22)
    t = np.concatenate([x[:3], np.linspace(-4.0, 8.0, 9)])
    out["poly_x"], out["poly_y"], out["poly_test"] = x, y, t
    out["poly_gram_meas"] = Ref.gram_sym(11, P11, x[:40], as_meas=True)
    out["poly_gram_plain"] = Ref.gram_sym(11, P11, x[:40], as_meas=False)
    out["poly_gram_cross"] = Ref.gram_cross(11, P11, x[:40], t)
    out["poly_information"] = Ref.gp_fit(11, P11, x, y)["information"]
    out["poly_nll"] = np.array(Ref.gp_nll(11, P11, x, y)[0])
    mean, _, cov = Ref.gp_predict(11, P11, x, y, t, 2)
    out["poly_mean"], out["poly_cov"] = mean, cov
    out["poly_var"] = Ref.gp_predict(11, P11, x, y, t, 1)[1]
    out["poly_noisy_var"] = Ref.gp_predict(11, P11, x, y, t, 5)[1]
    m, v, _, s = Ref.gp_cv(11, P11, x, y, 0, 0.0, what=1, want_score=True)
    out["poly_loo_mean"], out["poly_loo_var"], out["poly_loo_score"] = m, v, np.array(s)
    u = Ref.uniform_inducing_points(x, 16)
    out["poly_u"] = u
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        r = Ref.sparse_gp(11, P11, x, y, u, gk, ga, test=t, what=2, want_ll=True)
        out[f"poly_sp_{tag}_mean"], out[f"poly_sp_{tag}_cov"] = r["mean"], r["cov"]
        out[f"poly_sp_{tag}_ll"] = np.array(r["ll"])


SECTIONS = {"poly": section_poly, "sparse_mo": section_sparse_mo, "ldlt": section_ldlt, "sparse_nested": section_sparse_nested,
            "big": section_big}


def main():
    want = sys.argv[1:] or list(SECTIONS)
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in want:
        SECTIONS[name](out)
        print("section", name, "done")
    np.savez_compressed(OUT, **out)
    print("wrote", len(out), "arrays to", OUT)


if __name__ == "__main__":
    main()
