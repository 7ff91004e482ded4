// albatross_b200 C++ trait layer — parameters and priors (host scalars; the device never sees them
// except as the live values flattened into an ab_op program at every call).
//
// Mirrors src/core/priors.hpp:29-257, src/core/parameters.hpp:25-309 and
// src/core/parameter_handling_mixin.hpp:18-299: same names, same log_pdf formulas (including the
// reference's own form of the Gaussian normaliser), same get/set surface.
#pragma once

#include <cassert>
#include <cmath>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace albatross_b200 {

using ParameterValue = double;
using ParameterKey = std::string;

constexpr double LARGE_VAL = HUGE_VAL; // src/core/declarations.hpp
constexpr double PARAMETER_EPSILON = std::numeric_limits<ParameterValue>::epsilon();
constexpr double PARAMETER_MAX = std::numeric_limits<ParameterValue>::max();

// The reference holds one of a closed set of prior classes in a variant (priors.hpp:192-246); here
// the closed set is an enum + two doubles, which is all any of them carries.
class PriorContainer {
public:
  enum Kind { UNINFORMATIVE, FIXED, POSITIVE, NON_NEGATIVE, UNIFORM, LOG_SCALE_UNIFORM, GAUSSIAN,
              POSITIVE_GAUSSIAN, LOG_NORMAL };

  PriorContainer() = default;
  PriorContainer(Kind kind, double a = 0., double b = 0.) : kind_(kind), a_(a), b_(b) {}

  double log_pdf(double x) const {
    constexpr double LOG_2PI = 1.8378770664093453;
    constexpr double LOG_2 = 0.6931471805599453;
    switch (kind_) {
    case UNINFORMATIVE:
    case FIXED:
      return 0.;
    case POSITIVE:
      return x > 0. ? 0. : -LARGE_VAL; // priors.hpp:60
    case NON_NEGATIVE:
      return x >= 0. ? 0. : -LARGE_VAL; // priors.hpp:70
    case UNIFORM:
    case LOG_SCALE_UNIFORM:
      return (x >= a_ && x <= b_) ? -std::log(b_ - a_) : -LARGE_VAL; // priors.hpp:92-98
    case GAUSSIAN: {
      const double deviation = (x - a_) / b_;
      return -0.5 * (LOG_2PI * 2 * std::log(b_) + deviation * deviation); // priors.hpp:135-138 (verbatim form)
    }
    case POSITIVE_GAUSSIAN: {
      const double deviation = (x - a_) / b_;
      return -0.5 * (LOG_2PI * 2 * std::log(b_) + deviation * deviation) + LOG_2; // priors.hpp:163-166
    }
    case LOG_NORMAL: {
      const double deviation = (std::log(x) - a_) / b_;
      return -0.5 * LOG_2PI - std::log(b_) - std::log(x) - deviation * deviation; // priors.hpp:186-189
    }
    }
    return 0.;
  }

  double lower_bound() const {
    switch (kind_) {
    case POSITIVE:
      return std::numeric_limits<double>::epsilon();
    case NON_NEGATIVE:
    case POSITIVE_GAUSSIAN:
      return 0.;
    case UNIFORM:
    case LOG_SCALE_UNIFORM:
      return a_;
    default:
      return -LARGE_VAL;
    }
  }

  double upper_bound() const {
    switch (kind_) {
    case UNIFORM:
    case LOG_SCALE_UNIFORM:
      return b_;
    case POSITIVE_GAUSSIAN:
      return 10. * b_; // priors.hpp:155
    default:
      return LARGE_VAL;
    }
  }

  bool is_log_scale() const { return kind_ == LOG_SCALE_UNIFORM; }
  bool is_fixed() const { return kind_ == FIXED; }

  std::string get_name() const {
    std::ostringstream oss;
    switch (kind_) {
    case UNINFORMATIVE:
      return "uninformative";
    case FIXED:
      return "fixed";
    case POSITIVE:
      return "positive";
    case NON_NEGATIVE:
      return "non_negative";
    case UNIFORM:
      oss << "uniform[" << a_ << "," << b_ << "]";
      break;
    case LOG_SCALE_UNIFORM:
      oss << "log_scale_uniform[" << a_ << "," << b_ << "]";
      break;
    case GAUSSIAN:
      oss << "gaussian[" << a_ << "," << b_ << "]";
      break;
    case POSITIVE_GAUSSIAN:
      oss << "positive_gaussian[" << a_ << "," << b_ << "]";
      break;
    case LOG_NORMAL:
      oss << "log_normal[" << a_ << "," << b_ << "]";
      break;
    }
    return oss.str();
  }

  bool operator==(const PriorContainer &o) const {
    return kind_ == o.kind_ && a_ == o.a_ && b_ == o.b_;
  }
  bool operator!=(const PriorContainer &o) const { return !(*this == o); }

private:
  Kind kind_ = UNINFORMATIVE;
  double a_ = 0., b_ = 0.;
};

// Same spelling as the reference's prior classes, usable as `{value, PositivePrior()}`.
inline PriorContainer UninformativePrior() { return PriorContainer(PriorContainer::UNINFORMATIVE); }
inline PriorContainer FixedPrior() { return PriorContainer(PriorContainer::FIXED); }
inline PriorContainer PositivePrior() { return PriorContainer(PriorContainer::POSITIVE); }
inline PriorContainer NonNegativePrior() { return PriorContainer(PriorContainer::NON_NEGATIVE); }
inline PriorContainer UniformPrior(double lower = 0., double upper = 1.) {
  assert(upper > lower);
  return PriorContainer(PriorContainer::UNIFORM, lower, upper);
}
inline PriorContainer LogScaleUniformPrior(double lower = 1e-12, double upper = 1.e12) {
  return PriorContainer(PriorContainer::LOG_SCALE_UNIFORM, lower, upper);
}
inline PriorContainer GaussianPrior(double mu = 0., double sigma = 1.) {
  return PriorContainer(PriorContainer::GAUSSIAN, mu, sigma);
}
inline PriorContainer PositiveGaussianPrior(double mu = 0., double sigma = 1.) {
  return PriorContainer(PriorContainer::POSITIVE_GAUSSIAN, mu, sigma);
}
inline PriorContainer LogNormalPrior(double mu = 0., double sigma = 1.) {
  return PriorContainer(PriorContainer::LOG_NORMAL, mu, sigma);
}

struct Parameter { // parameters.hpp:25-60
  ParameterValue value = 0.;
  PriorContainer prior;

  Parameter() = default;
  Parameter(ParameterValue value_) : value(value_) {}
  Parameter(ParameterValue value_, const PriorContainer &prior_) : value(value_), prior(prior_) {}

  bool operator==(const Parameter &o) const { return value == o.value && prior == o.prior; }
  bool operator!=(const Parameter &o) const { return !(*this == o); }
  bool within_bounds() const { return value >= prior.lower_bound() && value <= prior.upper_bound(); }
  bool is_valid() const { return within_bounds(); }
  bool is_fixed() const { return prior.is_fixed(); }
  double prior_log_likelihood() const { return prior.log_pdf(value); }
};

using ParameterStore = std::map<ParameterKey, Parameter>;

struct TunableParameters { // parameters.hpp:18-23
  std::vector<std::string> names;
  std::vector<double> values;
  std::vector<double> lower_bounds;
  std::vector<double> upper_bounds;
};

inline ParameterStore map_join(const ParameterStore &a, const ParameterStore &b) {
  ParameterStore out(a);
  for (const auto &pair : b) {
    // the reference asserts on duplicate keys (src/utils/map_utils.hpp)
    assert(out.find(pair.first) == out.end() && "duplicate parameter name");
    out[pair.first] = pair.second;
  }
  return out;
}

inline std::string pretty_params(const ParameterStore &params) { // parameters.hpp:66-78
  std::ostringstream ss;
  ss.precision(12);
  ss << std::scientific << "{" << std::endl;
  for (const auto &pair : params) {
    ss << "    {\"" << pair.first << "\", " << pair.second.value << "}," << std::endl;
  }
  ss << "};" << std::endl;
  return ss.str();
}

inline double parameter_prior_log_likelihood(const ParameterStore &params) {
  double sum = 0.;
  for (const auto &pair : params) {
    sum += pair.second.prior_log_likelihood();
  }
  return sum;
}

// get_tunable_parameters / set_tunable_params_values, parameters.hpp:115-243: fixed parameters are
// skipped; log-scale priors are tuned in log space.
inline TunableParameters get_tunable_parameters(const ParameterStore &params) {
  TunableParameters out;
  for (const auto &pair : params) {
    if (pair.second.is_fixed()) {
      continue;
    }
    double v = pair.second.value;
    double lb = pair.second.prior.lower_bound();
    double ub = pair.second.prior.upper_bound();
    if (pair.second.prior.is_log_scale()) {
      v = std::log(v);
      lb = std::log(lb);
      ub = std::log(ub);
    }
    out.names.push_back(pair.first);
    out.values.push_back(v);
    out.lower_bounds.push_back(lb);
    out.upper_bounds.push_back(ub);
  }
  return out;
}

inline ParameterStore set_tunable_params_values(const ParameterStore &params,
                                                const std::vector<double> &x,
                                                bool force_bounds = true) {
  ParameterStore out(params);
  std::size_t i = 0;
  for (auto &pair : out) {
    if (pair.second.is_fixed()) {
      continue;
    }
    double v = x.at(i++);
    if (pair.second.prior.is_log_scale()) {
      v = std::exp(v);
    }
    if (force_bounds) {
      const double lb = pair.second.prior.lower_bound();
      const double ub = pair.second.prior.upper_bound();
      v = v < lb ? lb : (v > ub ? ub : v);
    }
    pair.second.value = v;
  }
  assert(i == x.size());
  return out;
}

// CRTP-free mixin: a class provides get_params() and set_param(name, Parameter); the rest of the
// reference's ParameterHandlingMixin surface (parameter_handling_mixin.hpp:33-118) is derived.
template <typename Derived> class ParameterHandling {
public:
  void set_params(const ParameterStore &params) {
    for (const auto &pair : params) {
      self().set_param(pair.first, pair.second);
    }
  }
  void set_param_values(const std::map<ParameterKey, ParameterValue> &values) {
    for (const auto &pair : values) {
      set_param_value(pair.first, pair.second);
    }
  }
  void set_param_value(const ParameterKey &name, ParameterValue value) {
    Parameter p = cself().get_params().at(name);
    p.value = value;
    self().set_param(name, p);
  }
  void set_prior(const ParameterKey &name, const PriorContainer &prior) {
    Parameter p = cself().get_params().at(name);
    p.prior = prior;
    self().set_param(name, p);
  }
  ParameterValue get_param_value(const ParameterKey &name) const {
    return cself().get_params().at(name).value;
  }
  double prior_log_likelihood() const { return parameter_prior_log_likelihood(cself().get_params()); }
  bool params_are_valid() const {
    for (const auto &pair : cself().get_params()) {
      if (!pair.second.is_valid()) {
        return false;
      }
    }
    return true;
  }
  std::vector<ParameterValue> get_params_as_vector() const {
    std::vector<ParameterValue> out;
    for (const auto &pair : cself().get_params()) {
      out.push_back(pair.second.value);
    }
    return out;
  }
  void set_params_from_vector(const std::vector<ParameterValue> &x) {
    ParameterStore params = cself().get_params();
    assert(x.size() == params.size());
    std::size_t i = 0;
    for (auto &pair : params) {
      pair.second.value = x[i++];
    }
    set_params(params);
  }
  TunableParameters get_tunable_parameters() const {
    return albatross_b200::get_tunable_parameters(cself().get_params());
  }
  void set_tunable_params_values(const std::vector<double> &x, bool force_bounds = true) {
    set_params(albatross_b200::set_tunable_params_values(cself().get_params(), x, force_bounds));
  }
  std::string pretty_string() const { return pretty_params(cself().get_params()); }

private:
  Derived &self() { return *static_cast<Derived *>(this); }
  const Derived &cself() const { return *static_cast<const Derived *>(this); }
};

} // namespace albatross_b200
