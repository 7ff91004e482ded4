"""CPU checks of the C++ trait layer (include/albatross_b200/): it builds against the in-tree C-ABI
library with and without Eigen, its host-side logic (names, parameters, priors, programs, indexing)
behaves like the reference's, and types without a device form are COMPILE-TIME errors
(north_star: "no CPU fallback"; reference precedent ALBATROSS_FAIL, src/details/error_handling.hpp:49-52).
No device call is made here."""
import os
import shutil
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
LIB = os.path.join(ROOT, "albatross_b200", "csrc", "libalbatross_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        pytest.fail("libalbatross_b200.so missing: run __graft_entry__.build()")
    subprocess.check_call(["make", "-s", "-C", CPP])
    return CPP


def test_host_checks_standin_types(built):
    out = subprocess.run([os.path.join(built, "trait_layer_check"), "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_host_checks_eigen_types(built):
    exe = os.path.join(built, "trait_layer_check_eigen")
    if not os.path.exists(exe):
        pytest.skip("the reference's vendored Eigen is not on this machine")
    out = subprocess.run([exe, "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def _compile(snippet, tmp_path):
    src = tmp_path / "t.cc"
    src.write_text("#include <albatross_b200/albatross.hpp>\n#include <string>\n"
                   "namespace ab = albatross_b200;\n" + textwrap.dedent(snippet))
    return subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-DALBATROSS_B200_NO_EIGEN", "-I",
                           os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)


def test_supported_program_compiles(tmp_path):
    ok = _compile("""
        void f() {
          auto cov = ab::SquaredExponential<ab::EuclideanDistance>(1., 1.) + ab::IndependentNoise<double>(0.1);
          auto model = ab::gp_from_covariance(cov);
          ab::RegressionDataset<double> d({1., 2.}, ab::VectorXd({1., 2.}));
          auto fit = model.fit(d);
          (void)fit.predict(std::vector<double>{1.5}).marginal();
          // the sinc example's covariance: a Polynomial as a summand, next to a measurement-only term
          auto sinc = ab::Polynomial<1>(100.) + ab::SquaredExponential<ab::EuclideanDistance>(3.5, 5.7) +
                      ab::measurement_only(ab::IndependentNoise<double>(1.0));
          auto fit2 = ab::gp_from_covariance(sinc).fit(d);
          (void)fit2.predict(std::vector<double>{1.5}).joint();
        }""", tmp_path)
    assert ok.returncode == 0, ok.stderr


@pytest.mark.parametrize("snippet, message", [
    # a feature type without a device form
    ("""void f() {
          ab::Constant c(1.);
          std::vector<std::string> xs = {"a"};
          (void)c(xs);
        }""", "no device form"),
    # a covariance the device does not know (user-defined host functor) inside a model
    ("""struct Mine { double operator()(double, double) const { return 1.; } };
        void f() { ab::GaussianProcessRegression<Mine> m; (void)m; }""", "no device form"),
    # a covariance that is not defined for the feature type (IndependentNoise<double> on 3-D points)
    ("""void f() {
          ab::IndependentNoise<double> n(0.1);
          std::vector<std::array<double, 3>> xs(2);
          (void)n(xs);
        }""", "not defined for these feature types"),
    # a Polynomial as a factor of a product (the device adds polynomial terms to the finished matrix: summands only)
    ("""void f() {
          auto cov = ab::Polynomial<1>(1.) * ab::SquaredExponential<ab::EuclideanDistance>(1., 1.);
          (void)cov;
        }""", "a Polynomial inside a product has no device form"),
    # a distance metric without a device form
    ("""struct AngularDistance { std::string get_name() const { return "angular"; } };
        void f() { ab::Exponential<AngularDistance> e; (void)e; }""", "only EuclideanDistance has a device form"),
])
def test_types_without_device_form_fail_to_compile(tmp_path, snippet, message):
    res = _compile(snippet, tmp_path)
    assert res.returncode != 0
    assert message in res.stderr, res.stderr[-2000:]
