/*
 * TEST INFRASTRUCTURE ONLY.  Shared helpers for the compiled-reference oracle
 * (oracle/_ref/libref_oracle.so).  This translation unit family includes the REAL reference headers
 * from /root/reference in place (never copied into this repo) and exposes them through a small
 * extern "C" surface so tests and bench.py's cpu_baseline / --impl reference leg can call the
 * reference's own Eigen path.  Nothing in the product (albatross_b200/, include/) may include or
 * link this.
 *
 * Covariance "menu": the reference composes covariances at compile time, so the oracle instantiates a
 * fixed list of compositions, selected by cov_id; params are passed in constructor order.
 *
 *   id  covariance                                        params
 *   0   SquaredExponential<EuclideanDistance>             [l, s]
 *   1   Exponential<EuclideanDistance>                    [l, s]
 *   2   Matern32<EuclideanDistance>                       [l, s]
 *   3   Matern52<EuclideanDistance>                       [l, s]
 *   4   Constant                                          [s]
 *   5   IndependentNoise<X>                               [s]
 *   6   SE + IndependentNoise   (bench_covariance)        [l, s, sn]
 *   7   SE + Matern52           (BASELINE config 2)       [l1, s1, l2, s2]
 *   8   SE + Matern52 + IndependentNoise                  [l1, s1, l2, s2, sn]
 *   9   SE*Matern32 + Exponential*Constant + Noise        [l1, s1, l2, s2, l3, s3, sc, sn]
 *   10  SE + measurement_only(IndependentNoise<X>)        [l, s, sn]   (examples/sinc_example.cc:79-80)
 *   11  Polynomial<1> + SE + measurement_only(noise)      [s0, s1, l, s, sn]  (examples/sinc_example.cc:84-87;
 *                                                          scalar features only)
 *
 * Feature types: dim==1 -> double, dim==3 -> Eigen::Vector3d, otherwise Eigen::VectorXd
 * (AoS doubles, point i at feats[i*dim .. i*dim+dim)).
 */
#ifndef AB_ORACLE_REF_COMMON_H
#define AB_ORACLE_REF_COMMON_H

#include <albatross/GP>
#include <albatross/Evaluation>

#include <cstdint>
#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace refshim {

using albatross::Constant;
using albatross::EuclideanDistance;
using albatross::Exponential;
using albatross::IndependentNoise;
using albatross::Matern32;
using albatross::Matern52;
using albatross::SquaredExponential;

using SE = SquaredExponential<EuclideanDistance>;
using EXP = Exponential<EuclideanDistance>;
using M32 = Matern32<EuclideanDistance>;
using M52 = Matern52<EuclideanDistance>;

/* Polynomial<1> with its two sigmas set (polynomials.hpp:63-90). */
inline albatross::Polynomial<1> linear_polynomial(double s0, double s1) {
  albatross::Polynomial<1> p;
  p.set_param_value("sigma_polynomial_0", s0);
  p.set_param_value("sigma_polynomial_1", s1);
  return p;
}

/* Entry 11 exists for scalar features only (Polynomial is defined on double). */
template <typename X, typename F> struct Entry11 {
  static int call(const double *, F &&) { return -1; }
};
template <typename F> struct Entry11<double, F> {
  static int call(const double *p, F &&f) {
    f(linear_polynomial(p[0], p[1]) + SE(p[2], p[3]) + albatross::measurement_only(IndependentNoise<double>(p[4])));
    return 0;
  }
};

template <typename X> struct FeatureIO;

template <> struct FeatureIO<double> {
  static std::vector<double> load(const double *f, int64_t n, int) {
    return std::vector<double>(f, f + n);
  }
  static double first_coord(const double &x) { return x; }
};

template <> struct FeatureIO<Eigen::Vector3d> {
  static std::vector<Eigen::Vector3d> load(const double *f, int64_t n, int) {
    std::vector<Eigen::Vector3d> out(static_cast<std::size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
      out[static_cast<std::size_t>(i)] = Eigen::Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    }
    return out;
  }
  static double first_coord(const Eigen::Vector3d &x) { return x[0]; }
};

template <> struct FeatureIO<Eigen::VectorXd> {
  static std::vector<Eigen::VectorXd> load(const double *f, int64_t n, int dim) {
    std::vector<Eigen::VectorXd> out(static_cast<std::size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
      Eigen::VectorXd v(dim);
      for (int d = 0; d < dim; ++d) {
        v[d] = f[i * dim + d];
      }
      out[static_cast<std::size_t>(i)] = v;
    }
    return out;
  }
  static double first_coord(const Eigen::VectorXd &x) { return x[0]; }
};

/* Builds the cov_id-th composition for feature type X and hands it to `f`. */
template <typename X, typename F>
inline int with_cov(int cov_id, const double *p, F &&f) {
  switch (cov_id) {
  case 0:
    f(SE(p[0], p[1]));
    return 0;
  case 1:
    f(EXP(p[0], p[1]));
    return 0;
  case 2:
    f(M32(p[0], p[1]));
    return 0;
  case 3:
    f(M52(p[0], p[1]));
    return 0;
  case 4:
    f(Constant(p[0]));
    return 0;
  case 5:
    f(IndependentNoise<X>(p[0]));
    return 0;
  case 6:
    f(SE(p[0], p[1]) + IndependentNoise<X>(p[2]));
    return 0;
  case 7:
    f(SE(p[0], p[1]) + M52(p[2], p[3]));
    return 0;
  case 8:
    f(SE(p[0], p[1]) + M52(p[2], p[3]) + IndependentNoise<X>(p[4]));
    return 0;
  case 9:
    f(SE(p[0], p[1]) * M32(p[2], p[3]) + EXP(p[4], p[5]) * Constant(p[6]) +
      IndependentNoise<X>(p[7]));
    return 0;
  case 10:
    f(SE(p[0], p[1]) + albatross::measurement_only(IndependentNoise<X>(p[2])));
    return 0;
  case 11:
    return Entry11<X, F>::call(p, std::forward<F>(f));
  default:
    return -1;
  }
}

/* Only the compositions that make sense as GP priors (they carry a noise term). */
template <typename X, typename F>
inline int with_gp_cov(int cov_id, const double *p, F &&f) {
  switch (cov_id) {
  case 6:
    f(SE(p[0], p[1]) + IndependentNoise<X>(p[2]));
    return 0;
  case 8:
    f(SE(p[0], p[1]) + M52(p[2], p[3]) + IndependentNoise<X>(p[4]));
    return 0;
  case 9:
    f(SE(p[0], p[1]) * M32(p[2], p[3]) + EXP(p[4], p[5]) * Constant(p[6]) +
      IndependentNoise<X>(p[7]));
    return 0;
  case 10:
    f(SE(p[0], p[1]) + albatross::measurement_only(IndependentNoise<X>(p[2])));
    return 0;
  case 11:
    return Entry11<X, F>::call(p, std::forward<F>(f));
  default:
    return -1;
  }
}

/* dim -> feature type dispatch. */
template <typename F> inline int with_feature_type(int dim, F &&f) {
  if (dim == 1) {
    return f(static_cast<double *>(nullptr));
  } else if (dim == 3) {
    return f(static_cast<Eigen::Vector3d *>(nullptr));
  }
  return f(static_cast<Eigen::VectorXd *>(nullptr));
}

inline void copy_out(const Eigen::MatrixXd &m, double *out) {
  if (out != nullptr) {
    std::copy(m.data(), m.data() + m.size(), out);
  }
}

inline void copy_out(const Eigen::VectorXd &v, double *out) {
  if (out != nullptr) {
    std::copy(v.data(), v.data() + v.size(), out);
  }
}

/*
 * Groupers usable from C: kind 0 = LeaveOneOutGrouper, 1 = int(x0) % k (bench_loo_cv.cc:95-105),
 * 2 = int(floor(x0 * scale)).  x0 = first coordinate of the feature.
 */
template <typename X> struct ShimGrouper {
  int kind;
  double arg;
  long operator()(const X &x) const {
    const double v = FeatureIO<X>::first_coord(x);
    if (kind == 1) {
      return static_cast<long>(static_cast<int>(v) % static_cast<int>(arg));
    }
    return static_cast<long>(std::floor(v * arg));
  }
};

} // namespace refshim

#endif
