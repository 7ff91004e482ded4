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


SECTIONS = {"sparse_mo": section_sparse_mo}


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
