// albatross_b200 C++ trait layer — the covariance-function concept.
//
// Same user surface as the reference (src/covariance_functions/covariance_function.hpp:63-437,
// radial.hpp, noise.hpp, polynomials.hpp, measurement.hpp): CRTP leaves with named Parameter
// members, composition by operator+ / operator* returning Sum/ProductOfCovarianceFunctions<L, R>
// by value, matrix calls returning MatrixXd by value.  The difference is what a call does: the
// compile-time tree is flattened to a postfix ab_op program carrying the LIVE parameter values and
// handed to the device (ab_gram_sym / ab_gram_cross / ab_gram_diag).  There is no host evaluation:
// a feature or covariance type without a device form is a compile-time error (static_assert), which
// is the reference's own failure mode for undefined (covariance, feature) pairs
// (ALBATROSS_FAIL, src/details/error_handling.hpp:49-52; src/models/gp.hpp:296-303).
#pragma once

#include <array>
#include <string>
#include <type_traits>
#include <vector>

#include "device.hpp"
#include "parameters.hpp"

namespace albatross_b200 {

struct ThreadPool; // accepted and ignored wherever the reference takes a ThreadPool*

// ------------------------------------------------------------------------------------------------
// features
// ------------------------------------------------------------------------------------------------

// Measurement<X> tag, src/covariance_functions/measurement.hpp:18-29.
template <typename X> struct Measurement {
  Measurement() : value() {}
  Measurement(const X &x) : value(x) {}
  X value;
};

template <typename X> inline Measurement<X> as_measurement(const X &f) { return Measurement<X>(f); }

template <typename X>
inline std::vector<Measurement<X>> as_measurements(const std::vector<X> &features) {
  return std::vector<Measurement<X>>(features.begin(), features.end());
}

template <typename X> struct is_measurement : std::false_type {};
template <typename X> struct is_measurement<Measurement<X>> : std::true_type {};
template <typename X> struct unwrap_measurement { using type = X; };
template <typename X> struct unwrap_measurement<Measurement<X>> { using type = X; };
template <typename X> using unwrap_measurement_t = typename unwrap_measurement<X>::type;

/*
 * device_feature<X>: how a feature type lands in HBM (AoS doubles, point i at [i*dim, (i+1)*dim)).
 * Device forms exist for double, std::array<double, D> and Eigen column vectors (fixed or dynamic
 * size) — the feature types of every scoped config.  Specialise it to give another
 * plain-coordinate type a device form.
 */
template <typename X, typename Enable = void> struct device_feature {
  static constexpr bool value = false;
};

template <> struct device_feature<double> {
  static constexpr bool value = true;
  static int dim(const double &) { return 1; }
  static void pack(const double &x, double *out) { out[0] = x; }
};

template <std::size_t D> struct device_feature<std::array<double, D>> {
  static_assert(D >= 1 && D <= AB_MAX_DIM, "feature dimension outside the device range");
  static constexpr bool value = true;
  static int dim(const std::array<double, D> &) { return static_cast<int>(D); }
  static void pack(const std::array<double, D> &x, double *out) {
    for (std::size_t d = 0; d < D; ++d) {
      out[d] = x[d];
    }
  }
};

#ifdef ALBATROSS_B200_HAVE_EIGEN
template <int Rows> struct device_feature<Eigen::Matrix<double, Rows, 1>> {
  static_assert(Rows == Eigen::Dynamic || (Rows >= 1 && Rows <= AB_MAX_DIM),
                "feature dimension outside the device range");
  static constexpr bool value = true;
  static int dim(const Eigen::Matrix<double, Rows, 1> &x) { return static_cast<int>(x.size()); }
  static void pack(const Eigen::Matrix<double, Rows, 1> &x, double *out) {
    for (Eigen::Index d = 0; d < x.size(); ++d) {
      out[d] = x[d];
    }
  }
};
#endif

template <typename X> struct device_feature<Measurement<X>> {
  static constexpr bool value = device_feature<X>::value;
  static int dim(const Measurement<X> &x) { return device_feature<X>::dim(x.value); }
  static void pack(const Measurement<X> &x, double *out) { device_feature<X>::pack(x.value, out); }
};

template <typename X> struct is_device_feature : std::integral_constant<bool, device_feature<X>::value> {};

struct PackedFeatures {
  std::vector<double> data;
  int64_t n = 0;
  int dim = 1;
};

template <typename X> inline PackedFeatures pack_features(const std::vector<X> &xs) {
  static_assert(is_device_feature<X>::value,
                "albatross_b200: this feature type has no device form (and there is no CPU fallback); "
                "specialise albatross_b200::device_feature<X>");
  PackedFeatures out;
  out.n = static_cast<int64_t>(xs.size());
  out.dim = xs.empty() ? 1 : device_feature<X>::dim(xs[0]);
  if (out.dim < 1 || out.dim > AB_MAX_DIM) {
    check_status(AB_ERR_UNSUPPORTED, "feature dimension outside [1, AB_MAX_DIM]");
  }
  out.data.resize(xs.size() * static_cast<std::size_t>(out.dim));
  for (std::size_t i = 0; i < xs.size(); ++i) {
    if (device_feature<X>::dim(xs[i]) != out.dim) {
      check_status(AB_ERR_INVALID, "features of differing dimension");
    }
    device_feature<X>::pack(xs[i], out.data.data() + i * static_cast<std::size_t>(out.dim));
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// programs
// ------------------------------------------------------------------------------------------------

using Program = std::vector<ab_op>;

inline void push_op(Program *prog, ab_opcode op, double p0 = 0., double p1 = 0.) {
  ab_op o;
  o.op = static_cast<int32_t>(op);
  o.reserved = 0;
  o.p0 = p0;
  o.p1 = p1;
  prog->push_back(o);
}

// Distance metrics, src/covariance_functions/distance_metrics.hpp:29-44.  Only the Euclidean metric
// has a device form in this round; the struct is the tag the radial templates take.
struct EuclideanDistance {
  std::string get_name() const { return "euclidean_distance"; }
  template <typename X> static constexpr bool defined_for() { return device_feature<X>::value; }
};

template <typename Derived> class CovarianceFunction;
template <typename LHS, typename RHS> class SumOfCovarianceFunctions;
template <typename LHS, typename RHS> class ProductOfCovarianceFunctions;

// is_device_covariance<K>: K derives from this layer's CovarianceFunction<K>.
template <typename K>
struct is_device_covariance : std::is_base_of<CovarianceFunction<K>, K> {};

/*
 * CRTP base.  A derived class provides
 *   std::string name() const;
 *   ParameterStore get_params() const;  void set_param(name, Parameter);
 *   template <X, Y> static constexpr bool defined();      // is k(X, Y) defined (unwrapped types)
 *   template <X, Y> void emit(Program*, bool both_measurements) const;   // postfix ops
 */
template <typename Derived>
class CovarianceFunction : public ParameterHandling<Derived> {
public:
  std::string get_name() const { return derived().name(); }

  template <typename X, typename Y> static constexpr bool is_defined_for() {
    return Derived::template defined<unwrap_measurement_t<X>, unwrap_measurement_t<Y>>();
  }

  // The program of k(X-typed, Y-typed) with the parameter values of this instant.
  template <typename X, typename Y> Program program() const {
    static_assert(is_device_feature<X>::value && is_device_feature<Y>::value,
                  "albatross_b200: feature type has no device form (no CPU fallback)");
    static_assert(is_defined_for<X, Y>(),
                  "albatross_b200: covariance function is not defined for these feature types");
    Program prog;
    derived().template emit<unwrap_measurement_t<X>, unwrap_measurement_t<Y>>(
        &prog, is_measurement<X>::value && is_measurement<Y>::value);
    if (prog.size() > AB_MAX_OPS) {
      check_status(AB_ERR_UNSUPPORTED, "covariance program longer than AB_MAX_OPS");
    }
    return prog;
  }

  // covariance_function.hpp:128-137 — symmetric Gram, returned to the host like the reference does.
  template <typename X>
  MatrixXd operator()(const std::vector<X> &xs, ThreadPool * = nullptr) const {
    return device_gram(xs, AB_GRAM_FULL).to_host();
  }

  // covariance_function.hpp:142-151 — cross Gram.
  template <typename X, typename Y>
  MatrixXd operator()(const std::vector<X> &xs, const std::vector<Y> &ys, ThreadPool * = nullptr) const {
    return device_cross(xs, ys).to_host();
  }

  // covariance_function.hpp:108-123 — one pair (a 1 x 1 cross Gram on the device).
  template <typename X, typename Y,
            typename std::enable_if<is_device_feature<X>::value && is_device_feature<Y>::value, int>::type = 0>
  double operator()(const X &x, const Y &y) const {
    return (*this)(std::vector<X>{x}, std::vector<Y>{y})(0, 0);
  }

  // covariance_function.hpp:156-168.
  template <typename X> VectorXd diagonal(const std::vector<X> &xs) const {
    const Program prog = program<X, X>();
    const PackedFeatures f = pack_features(xs);
    VectorXd out(static_cast<Index>(f.n));
    ALBATROSS_B200_CHECK(ab_gram_diag(Device::default_device()->get(), prog.data(),
                                      static_cast<int>(prog.size()), f.data.data(), f.n, f.dim,
                                      out.data()));
    return out;
  }

  // Device-resident results: what the models use, so that K never visits the host.
  template <typename X>
  DeviceMatrix device_gram(const std::vector<X> &xs, uint32_t flags = AB_GRAM_FULL,
                           std::shared_ptr<Device> dev = Device::default_device()) const {
    const Program prog = program<X, X>();
    const PackedFeatures f = pack_features(xs);
    ab_matrix m = nullptr;
    ALBATROSS_B200_CHECK(ab_gram_sym(dev->get(), prog.data(), static_cast<int>(prog.size()),
                                     f.data.data(), f.n, f.dim, flags, &m));
    return DeviceMatrix(std::move(dev), m);
  }

  template <typename X, typename Y>
  DeviceMatrix device_cross(const std::vector<X> &xs, const std::vector<Y> &ys,
                            std::shared_ptr<Device> dev = Device::default_device()) const {
    const Program prog = program<X, Y>();
    const PackedFeatures fx = pack_features(xs);
    const PackedFeatures fy = pack_features(ys);
    if (fx.n > 0 && fy.n > 0 && fx.dim != fy.dim) {
      check_status(AB_ERR_INVALID, "cross covariance between features of differing dimension");
    }
    ab_matrix m = nullptr;
    ALBATROSS_B200_CHECK(ab_gram_cross(dev->get(), prog.data(), static_cast<int>(prog.size()),
                                       fx.data.data(), fx.n, fy.data.data(), fy.n,
                                       fx.n > 0 ? fx.dim : fy.dim, &m));
    return DeviceMatrix(std::move(dev), m);
  }

  template <typename Other>
  const SumOfCovarianceFunctions<Derived, Other> operator+(const CovarianceFunction<Other> &other) const {
    return SumOfCovarianceFunctions<Derived, Other>(derived(), other.derived());
  }

  template <typename Other>
  const ProductOfCovarianceFunctions<Derived, Other> operator*(const CovarianceFunction<Other> &other) const {
    return ProductOfCovarianceFunctions<Derived, Other>(derived(), other.derived());
  }

  Derived &derived() { return *static_cast<Derived *>(this); }
  const Derived &derived() const { return *static_cast<const Derived *>(this); }
};

// Declares get_params / set_param for a list of Parameter members, the job of
// ALBATROSS_DECLARE_PARAMS (src/core/parameter_macros.hpp:157-160).
#define ALBATROSS_B200_PARAMS_1(a)                                                                  \
  ::albatross_b200::ParameterStore get_params() const { return {{#a, a}}; }                         \
  void set_param(const ::albatross_b200::ParameterKey &name_, const ::albatross_b200::Parameter &p_) { \
    if (name_ == #a) {                                                                              \
      a = p_;                                                                                       \
    } else {                                                                                        \
      assert(false && "unknown parameter");                                                         \
    }                                                                                               \
  }                                                                                                 \
  bool has_param(const ::albatross_b200::ParameterKey &name_) const { return name_ == #a; }

#define ALBATROSS_B200_PARAMS_2(a, b)                                                               \
  ::albatross_b200::ParameterStore get_params() const { return {{#a, a}, {#b, b}}; }                \
  void set_param(const ::albatross_b200::ParameterKey &name_, const ::albatross_b200::Parameter &p_) { \
    if (name_ == #a) {                                                                              \
      a = p_;                                                                                       \
    } else if (name_ == #b) {                                                                       \
      b = p_;                                                                                       \
    } else {                                                                                        \
      assert(false && "unknown parameter");                                                         \
    }                                                                                               \
  }                                                                                                 \
  bool has_param(const ::albatross_b200::ParameterKey &name_) const {                               \
    return name_ == #a || name_ == #b;                                                              \
  }

constexpr double default_length_scale = 100000.; // radial.hpp:18-19
constexpr double default_radial_sigma = 10.;

// Radial leaves: k defined iff the distance metric is defined for X == Y (radial.hpp:175-186).
template <typename Derived, typename DistanceMetricType, ab_opcode OP>
class RadialCovariance : public CovarianceFunction<Derived> {
  static_assert(std::is_same<DistanceMetricType, EuclideanDistance>::value,
                "albatross_b200: only EuclideanDistance has a device form (no CPU fallback)");

public:
  template <typename X, typename Y> static constexpr bool defined() {
    return std::is_same<X, Y>::value && DistanceMetricType::template defined_for<X>();
  }
  DistanceMetricType distance_metric_;
};

// sigma^2 exp(-(d / length_scale)^2), radial.hpp:25-33,132-189.
template <class DistanceMetricType>
class SquaredExponential
    : public RadialCovariance<SquaredExponential<DistanceMetricType>, DistanceMetricType,
                              AB_OP_SQUARED_EXPONENTIAL> {
public:
  SquaredExponential(double length_scale_ = default_length_scale,
                     double sigma_squared_exponential_ = default_radial_sigma) {
    squared_exponential_length_scale = {length_scale_, PositivePrior()};
    sigma_squared_exponential = {sigma_squared_exponential_, NonNegativePrior()};
  }
  std::string name() const {
    return "squared_exponential[" + this->distance_metric_.get_name() + "]";
  }
  ALBATROSS_B200_PARAMS_2(squared_exponential_length_scale, sigma_squared_exponential)
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_SQUARED_EXPONENTIAL, squared_exponential_length_scale.value,
            sigma_squared_exponential.value);
  }
  Parameter squared_exponential_length_scale;
  Parameter sigma_squared_exponential;
};

// sigma^2 exp(-|d / length_scale|), radial.hpp:191-198,240-287.
template <class DistanceMetricType>
class Exponential
    : public RadialCovariance<Exponential<DistanceMetricType>, DistanceMetricType, AB_OP_EXPONENTIAL> {
public:
  Exponential(double length_scale_ = default_length_scale,
              double sigma_exponential_ = default_radial_sigma) {
    exponential_length_scale = {length_scale_, PositivePrior()};
    sigma_exponential = {sigma_exponential_, NonNegativePrior()};
  }
  std::string name() const { return "exponential[" + this->distance_metric_.get_name() + "]"; }
  ALBATROSS_B200_PARAMS_2(exponential_length_scale, sigma_exponential)
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_EXPONENTIAL, exponential_length_scale.value, sigma_exponential.value);
  }
  Parameter exponential_length_scale;
  Parameter sigma_exponential;
};

// radial.hpp:289-297,422-459.
template <class DistanceMetricType>
class Matern32
    : public RadialCovariance<Matern32<DistanceMetricType>, DistanceMetricType, AB_OP_MATERN32> {
public:
  Matern32(double length_scale_ = default_length_scale, double sigma_matern_32_ = default_radial_sigma) {
    matern_32_length_scale = {length_scale_, PositivePrior()};
    sigma_matern_32 = {sigma_matern_32_, NonNegativePrior()};
  }
  std::string name() const { return "matern_32[" + this->distance_metric_.get_name() + "]"; }
  ALBATROSS_B200_PARAMS_2(matern_32_length_scale, sigma_matern_32)
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_MATERN32, matern_32_length_scale.value, sigma_matern_32.value);
  }
  Parameter matern_32_length_scale;
  Parameter sigma_matern_32;
};

// radial.hpp:461-470,492-529.
template <class DistanceMetricType>
class Matern52
    : public RadialCovariance<Matern52<DistanceMetricType>, DistanceMetricType, AB_OP_MATERN52> {
public:
  Matern52(double length_scale_ = default_length_scale, double sigma_matern_52_ = default_radial_sigma) {
    matern_52_length_scale = {length_scale_, PositivePrior()};
    sigma_matern_52 = {sigma_matern_52_, NonNegativePrior()};
  }
  std::string name() const { return "matern_52[" + this->distance_metric_.get_name() + "]"; }
  ALBATROSS_B200_PARAMS_2(matern_52_length_scale, sigma_matern_52)
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_MATERN52, matern_52_length_scale.value, sigma_matern_52.value);
  }
  Parameter matern_52_length_scale;
  Parameter sigma_matern_52;
};

// sigma^2 for every pair of anything, polynomials.hpp:33-61.
class Constant : public CovarianceFunction<Constant> {
public:
  static constexpr double default_sigma = 100.; // polynomials.hpp:18
  Constant(double sigma_constant_ = default_sigma) {
    sigma_constant = {sigma_constant_, NonNegativePrior()};
  }
  std::string name() const { return "constant"; }
  ALBATROSS_B200_PARAMS_1(sigma_constant)
  template <typename X, typename Y> static constexpr bool defined() { return true; }
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_CONSTANT, sigma_constant.value);
  }
  Parameter sigma_constant;
};

/*
 * Polynomial<order>, polynomials.hpp:63-90: sum_p sigma_p^2 x^p y^p between two doubles, one parameter
 * "sigma_polynomial_<p>" per degree.  On the device every degree is one AB_OP_POLYNOMIAL_TERM leaf; the
 * Gram builder accepts them as summands of the top-level sum only (a product with a Polynomial has no
 * device form: AB_ERR_UNSUPPORTED from the library, there is no CPU fallback).
 */
template <int order> class Polynomial : public CovarianceFunction<Polynomial<order>> {
  static_assert(order >= 0 && order <= 7, "albatross_b200: Polynomial<order> carries at most 8 device terms");

public:
  Polynomial(double sigma = Constant::default_sigma) {
    for (int i = 0; i <= order; ++i) {
      sigmas_[static_cast<std::size_t>(i)] = {sigma, NonNegativePrior()};
    }
  }
  std::string name() const { return "polynomial_" + std::to_string(order); }
  static std::string param_name(int i) { return "sigma_polynomial_" + std::to_string(i); }
  ParameterStore get_params() const {
    ParameterStore out;
    for (int i = 0; i <= order; ++i) {
      out[param_name(i)] = sigmas_[static_cast<std::size_t>(i)];
    }
    return out;
  }
  bool has_param(const ParameterKey &name_) const {
    for (int i = 0; i <= order; ++i) {
      if (name_ == param_name(i)) {
        return true;
      }
    }
    return false;
  }
  void set_param(const ParameterKey &name_, const Parameter &p_) {
    for (int i = 0; i <= order; ++i) {
      if (name_ == param_name(i)) {
        sigmas_[static_cast<std::size_t>(i)] = p_;
        return;
      }
    }
    assert(false && "unknown parameter");
  }
  template <typename X, typename Y> static constexpr bool defined() {
    return std::is_same<X, double>::value && std::is_same<Y, double>::value;
  }
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    for (int i = 0; i <= order; ++i) {
      push_op(prog, AB_OP_POLYNOMIAL_TERM, sigmas_[static_cast<std::size_t>(i)].value, static_cast<double>(i));
      if (i > 0) {
        push_op(prog, AB_OP_SUM);
      }
    }
  }

private:
  std::array<Parameter, static_cast<std::size_t>(order + 1)> sigmas_;
};

// sigma^2 iff x == y, defined only between two `Observed`s, noise.hpp:20-45.
template <typename Observed> class IndependentNoise : public CovarianceFunction<IndependentNoise<Observed>> {
public:
  IndependentNoise(double sigma_noise = 0.1) { sigma_independent_noise = {sigma_noise, PositivePrior()}; }
  std::string name() const { return "independent_noise"; }
  ALBATROSS_B200_PARAMS_1(sigma_independent_noise)
  template <typename X, typename Y> static constexpr bool defined() {
    return std::is_same<X, Observed>::value && std::is_same<Y, Observed>::value &&
           device_feature<Observed>::value;
  }
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_INDEPENDENT_NOISE, sigma_independent_noise.value);
  }
  Parameter sigma_independent_noise;
};

// nugget_sigma^2 iff x == y, for any feature type with a device form, nugget.hpp:32-49 (fixed prior, 1e-8 by
// default).  On the device it is the value-equality leaf of IndependentNoise.
class Nugget : public CovarianceFunction<Nugget> {
public:
  static constexpr double default_nugget_noise = 1e-8; // nugget.hpp:16
  Nugget() { nugget_sigma = {default_nugget_noise, FixedPrior()}; }
  std::string name() const { return "nugget"; }
  ALBATROSS_B200_PARAMS_1(nugget_sigma)
  template <typename X, typename Y> static constexpr bool defined() {
    return std::is_same<X, Y>::value && device_feature<X>::value;
  }
  template <typename X, typename Y> void emit(Program *prog, bool) const {
    push_op(prog, AB_OP_INDEPENDENT_NOISE, nugget_sigma.value);
  }
  Parameter nugget_sigma;
};

/*
 * MeasurementOnly<Sub>, measurement.hpp:70-114: Sub between two Measurement<>s, exactly 0 otherwise.
 * The zero is emitted as a Constant with sigma 0 (0 * 0 = +0), so sums and products that contain it
 * reproduce the reference's arithmetic (lhs + 0, `lhs != 0 ? lhs * 0 : lhs`).
 */
template <typename SubCovariance>
class MeasurementOnly : public CovarianceFunction<MeasurementOnly<SubCovariance>> {
  static_assert(is_device_covariance<SubCovariance>::value,
                "albatross_b200: covariance type has no device form (no CPU fallback)");

public:
  MeasurementOnly() : sub_cov_() {}
  MeasurementOnly(const SubCovariance &sub_cov) : sub_cov_(sub_cov) {}
  std::string name() const { return "measurement[" + sub_cov_.get_name() + "]"; }
  ParameterStore get_params() const { return sub_cov_.get_params(); }
  void set_param(const ParameterKey &name, const Parameter &param) { sub_cov_.set_param(name, param); }
  bool has_param(const ParameterKey &name) const { return sub_cov_.has_param(name); }
  template <typename X, typename Y> static constexpr bool defined() {
    return SubCovariance::template defined<X, Y>();
  }
  template <typename X, typename Y> void emit(Program *prog, bool both_measurements) const {
    if (both_measurements) {
      sub_cov_.template emit<X, Y>(prog, false); // Sub sees the unwrapped values (:105-107)
    } else {
      push_op(prog, AB_OP_CONSTANT, 0.);
    }
  }

private:
  SubCovariance sub_cov_;
};

template <typename SubCovariance>
inline MeasurementOnly<SubCovariance> measurement_only(const SubCovariance &cov) {
  return MeasurementOnly<SubCovariance>(cov);
}

// contains_polynomial<K>: K holds a Polynomial<order> somewhere in its tree.  The device adds polynomial terms in a
// pass over the finished matrix, which works for summands only: a product with one is rejected at compile time.
template <typename K> struct contains_polynomial : std::false_type {};
template <int order> struct contains_polynomial<Polynomial<order>> : std::true_type {};
template <typename Sub> struct contains_polynomial<MeasurementOnly<Sub>> : contains_polynomial<Sub> {};
template <class LHS, class RHS>
struct contains_polynomial<SumOfCovarianceFunctions<LHS, RHS>>
    : std::integral_constant<bool, contains_polynomial<LHS>::value || contains_polynomial<RHS>::value> {};
template <class LHS, class RHS>
struct contains_polynomial<ProductOfCovarianceFunctions<LHS, RHS>>
    : std::integral_constant<bool, contains_polynomial<LHS>::value || contains_polynomial<RHS>::value> {};

// Shared by Sum and Product: parameter plumbing and the one-sided fallbacks of
// covariance_function.hpp:266-294 / :357-388 (when only one operand is defined for (X, Y) the
// result is that operand alone).
template <typename Derived, typename LHS, typename RHS, ab_opcode OP>
class BinaryCovariance : public CovarianceFunction<Derived> {
  static_assert(is_device_covariance<LHS>::value && is_device_covariance<RHS>::value,
                "albatross_b200: covariance type has no device form (no CPU fallback)");

public:
  BinaryCovariance() : lhs_(), rhs_() {}
  BinaryCovariance(const LHS &lhs, const RHS &rhs) : lhs_(lhs), rhs_(rhs) {}

  ParameterStore get_params() const { return map_join(lhs_.get_params(), rhs_.get_params()); }
  bool has_param(const ParameterKey &name) const { return lhs_.has_param(name) || rhs_.has_param(name); }
  void set_param(const ParameterKey &name, const Parameter &param) {
    // set_param_if_exists_in_any: every operand that knows the name takes it
    bool found = false;
    if (lhs_.has_param(name)) {
      lhs_.set_param(name, param);
      found = true;
    }
    if (rhs_.has_param(name)) {
      rhs_.set_param(name, param);
      found = true;
    }
    assert(found && "unknown parameter");
    (void)found;
  }

  template <typename X, typename Y> static constexpr bool defined() {
    return LHS::template defined<X, Y>() || RHS::template defined<X, Y>();
  }

  template <typename X, typename Y> void emit(Program *prog, bool both_measurements) const {
    constexpr bool l = LHS::template defined<X, Y>();
    constexpr bool r = RHS::template defined<X, Y>();
    if constexpr (l) {
      lhs_.template emit<X, Y>(prog, both_measurements);
    }
    if constexpr (r) {
      rhs_.template emit<X, Y>(prog, both_measurements);
    }
    if constexpr (l && r) {
      push_op(prog, OP);
    }
  }

  const LHS &lhs() const { return lhs_; }
  const RHS &rhs() const { return rhs_; }

protected:
  LHS lhs_;
  RHS rhs_;
};

template <class LHS, class RHS>
class SumOfCovarianceFunctions
    : public BinaryCovariance<SumOfCovarianceFunctions<LHS, RHS>, LHS, RHS, AB_OP_SUM> {
public:
  using BinaryCovariance<SumOfCovarianceFunctions<LHS, RHS>, LHS, RHS, AB_OP_SUM>::BinaryCovariance;
  std::string name() const { return "(" + this->lhs_.get_name() + "+" + this->rhs_.get_name() + ")"; }
};

template <class LHS, class RHS>
class ProductOfCovarianceFunctions
    : public BinaryCovariance<ProductOfCovarianceFunctions<LHS, RHS>, LHS, RHS, AB_OP_PRODUCT> {
  static_assert(!contains_polynomial<LHS>::value && !contains_polynomial<RHS>::value,
                "albatross_b200: a Polynomial inside a product has no device form (no CPU fallback)");

public:
  using BinaryCovariance<ProductOfCovarianceFunctions<LHS, RHS>, LHS, RHS, AB_OP_PRODUCT>::BinaryCovariance;
  std::string name() const { return "(" + this->lhs_.get_name() + "*" + this->rhs_.get_name() + ")"; }
};

} // namespace albatross_b200
