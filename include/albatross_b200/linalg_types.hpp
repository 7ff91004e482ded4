// albatross_b200 C++ trait layer — value types the user API returns.
//
// The reference returns Eigen::MatrixXd / VectorXd / DiagonalMatrix by value from every matrix call
// (src/covariance_functions/covariance_function.hpp:131,145; src/core/distribution.hpp).  When Eigen
// is on the include path (it is for every user of the reference) those very types are used.  When it
// is not (the GPU test box of this repo ships no Eigen) a minimal column-major stand-in with the
// same member subset is used, so that the layer, its tests and examples still build.
// Only the common subset is used inside the layer: rows(), cols(), size(), data(), (i, j), [i],
// diagonal().
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#if !defined(ALBATROSS_B200_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#define ALBATROSS_B200_HAVE_EIGEN 1
#endif
#endif

#ifdef ALBATROSS_B200_HAVE_EIGEN
#include <Eigen/Dense>
#endif

namespace albatross_b200 {

#ifdef ALBATROSS_B200_HAVE_EIGEN

using Index = Eigen::Index;
using MatrixXd = Eigen::MatrixXd;
using VectorXd = Eigen::VectorXd;
using DiagonalMatrixXd = Eigen::DiagonalMatrix<double, Eigen::Dynamic>;

#else

using Index = std::ptrdiff_t;

class VectorXd {
public:
  VectorXd() = default;
  explicit VectorXd(Index n) : v_(static_cast<std::size_t>(n), 0.) {}
  VectorXd(std::initializer_list<double> il) : v_(il) {}
  static VectorXd Zero(Index n) { return VectorXd(n); }
  static VectorXd Constant(Index n, double c) {
    VectorXd out(n);
    for (auto &x : out.v_) {
      x = c;
    }
    return out;
  }
  Index size() const { return static_cast<Index>(v_.size()); }
  Index rows() const { return size(); }
  Index cols() const { return 1; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator[](Index i) { return v_[static_cast<std::size_t>(i)]; }
  double operator[](Index i) const { return v_[static_cast<std::size_t>(i)]; }
  double &operator()(Index i) { return (*this)[i]; }
  double operator()(Index i) const { return (*this)[i]; }
  double sum() const {
    double s = 0.;
    for (double x : v_) {
      s += x;
    }
    return s;
  }
  double mean() const { return sum() / static_cast<double>(v_.size()); }
  bool operator==(const VectorXd &o) const { return v_ == o.v_; }

private:
  std::vector<double> v_;
};

class MatrixXd {
public:
  MatrixXd() = default;
  MatrixXd(Index r, Index c) : r_(r), c_(c), v_(static_cast<std::size_t>(r * c), 0.) {}
  static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c); }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator()(Index i, Index j) { return v_[static_cast<std::size_t>(i + j * r_)]; }
  double operator()(Index i, Index j) const { return v_[static_cast<std::size_t>(i + j * r_)]; }
  bool operator==(const MatrixXd &o) const { return r_ == o.r_ && c_ == o.c_ && v_ == o.v_; }

private:
  Index r_ = 0, c_ = 0;
  std::vector<double> v_; // column-major, as Eigen::MatrixXd
};

class DiagonalMatrixXd {
public:
  DiagonalMatrixXd() = default;
  explicit DiagonalMatrixXd(Index n) : d_(n) {}
  explicit DiagonalMatrixXd(const VectorXd &d) : d_(d) {}
  VectorXd &diagonal() { return d_; }
  const VectorXd &diagonal() const { return d_; }
  Index rows() const { return d_.size(); }
  Index cols() const { return d_.size(); }
  Index size() const { return d_.size(); }

private:
  VectorXd d_;
};

#endif

inline VectorXd make_vector(const double *p, Index n) {
  VectorXd v(n);
  for (Index i = 0; i < n; ++i) {
    v[i] = p[i];
  }
  return v;
}

inline VectorXd make_vector(const std::vector<double> &x) {
  return make_vector(x.data(), static_cast<Index>(x.size()));
}

} // namespace albatross_b200
