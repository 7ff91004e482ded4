// albatross_b200 C++ trait layer — the tuner loop, the caller that multiplies the hot
// path's cost: ModelTuner::tune() evaluates the objective 10^2 – 10^3 times, every evaluation a Gram build +
// factorisation (+ LOO algebra) on the device.
//
// User surface of src/tune/tune.hpp:110-330 and src/tune/finite_difference.hpp:18-81:
//     auto tuner = get_tuner(model, LeaveOneOutLikelihood<>(), dataset);
//     ParameterStore tuned = tuner.tune();
//     GenericTuner(params).tune(objective);     compute_gradient(f, params, f0);
// What differs, by design:
//   * the objective runs on the device (model.log_likelihood -> ab_gp_nll, the CV metrics -> ab_gp_cv_scores);
//     the handle recycles its N x N workspace between evaluations, so nothing but the hyper-parameters moves;
//   * finite-difference perturbations (one objective evaluation per tunable parameter,
//     finite_difference.hpp:25-31 hands them to a ThreadPool) are dealt over SEVERAL GPUs: one host thread
//     and one handle per device, perturbation i on device i % ndevices (set_devices);
//   * the optimiser: the reference drives nlopt (LN_SBPLX by default, tune.hpp:71-88), a third-party library
//     that is not part of the hot path.  This layer ships a bounded Nelder–Mead simplex (the building block of
//     SBPLX) with the reference's stopping rules (ftol_abs 1e-8, ftol_rel 1e-6) so that tune() works stand
//     alone; an application that links nlopt keeps using it with the objective below unchanged.
// NaN objectives are mapped to +inf exactly as tune.hpp:164-166 does.
#pragma once

#include <algorithm>
#include <cmath>
#include <functional>
#include <iostream>
#include <limits>
#include <thread>

#include "gp.hpp"

namespace albatross_b200 {

// ---- finite differences (finite_difference.hpp:18-81) ---------------------------------------------------

// Runs fn(i) for i in [0, count), dealing the indices over `lanes` host threads (lane = i % lanes); fn gets
// the lane so that it can pick its device.  lanes <= 1: serial, in order.
template <typename Fn> inline void deal_over_lanes(std::size_t count, std::size_t lanes, Fn fn) {
  if (lanes <= 1 || count <= 1) {
    for (std::size_t i = 0; i < count; ++i) {
      fn(i, std::size_t(0));
    }
    return;
  }
  std::vector<std::thread> workers;
  for (std::size_t lane = 0; lane < std::min(lanes, count); ++lane) {
    workers.emplace_back([=]() {
      for (std::size_t i = lane; i < count; i += lanes) {
        fn(i, lane);
      }
    });
  }
  for (auto &w : workers) {
    w.join();
  }
}

// f(std::vector<double>) form, :18-32: forward difference with epsilon = 1e-6.
template <typename Function>
inline std::vector<double> compute_gradient(Function f, const std::vector<double> &params, double f_val,
                                            std::size_t lanes = 1) {
  const double epsilon = 1e-6;
  std::vector<double> grad(params.size());
  deal_over_lanes(params.size(), lanes, [&](std::size_t i, std::size_t) {
    std::vector<double> perturbed(params);
    perturbed[i] += epsilon;
    grad[i] = (f(perturbed) - f_val) / epsilon;
  });
  return grad;
}

// f(ParameterStore, lane) form, :34-79: epsilon = 1e-8 * (upper - lower) when the range is finite, else 1e-6;
// the step flips sign when it leaves the valid region; gradients that point out of an active bound are zero.
template <typename Function>
inline std::vector<double> compute_gradient_on_lanes(Function f, const ParameterStore &params, double f_val,
                                                     std::size_t lanes) {
  const TunableParameters tunable = get_tunable_parameters(params);
  std::vector<double> grad(tunable.values.size());
  deal_over_lanes(tunable.values.size(), lanes, [&](std::size_t i, std::size_t lane) {
    auto perturbed_params = [&](double eps) {
      std::vector<double> x(tunable.values);
      x[i] += eps;
      return set_tunable_params_values(params, x);
    };
    double epsilon = 1e-6;
    const double range = tunable.upper_bounds[i] - tunable.lower_bounds[i];
    if (std::isfinite(range)) {
      epsilon = 1e-8 * range;
    }
    ParameterStore p = perturbed_params(epsilon);
    bool valid = true;
    for (const auto &pair : p) {
      valid = valid && pair.second.is_valid();
    }
    if (!valid) {
      epsilon *= -1;
      p = perturbed_params(epsilon);
    }
    double g = (f(p, lane) - f_val) / epsilon;
    if (tunable.values[i] >= tunable.upper_bounds[i] && g < 0) {
      g = 0;
    }
    if (tunable.values[i] <= tunable.lower_bounds[i] && g > 0) {
      g = 0;
    }
    grad[i] = g;
  });
  return grad;
}
template <typename Function>
inline std::vector<double> compute_gradient(Function f, const ParameterStore &params, double f_val) {
  return compute_gradient_on_lanes([&](const ParameterStore &p, std::size_t) { return f(p); }, params, f_val, 1);
}

// ---- the optimiser: bounded Nelder–Mead ----------------------------------------------------------------

struct SimplexOptions {
  double ftol_abs = 1e-8; // default_optimizer, tune.hpp:76-77
  double ftol_rel = 1e-6;
  std::size_t max_evaluations = 2000;
  double initial_step = 0.25; // of the bounded range (or of max(|x|, 1) when unbounded)
};

struct SimplexResult {
  std::vector<double> x;
  double f = std::numeric_limits<double>::infinity();
  std::size_t evaluations = 0;
  std::string termination;
};

template <typename Objective>
inline SimplexResult minimize_simplex(Objective objective, std::vector<double> x0, const std::vector<double> &lower,
                                      const std::vector<double> &upper, const SimplexOptions &opt = SimplexOptions()) {
  const std::size_t n = x0.size();
  SimplexResult res;
  auto clamp = [&](std::vector<double> x) {
    for (std::size_t i = 0; i < n; ++i) {
      x[i] = std::min(std::max(x[i], lower[i]), upper[i]);
    }
    return x;
  };
  auto eval = [&](const std::vector<double> &x) {
    ++res.evaluations;
    const double v = objective(x);
    return std::isnan(v) ? std::numeric_limits<double>::infinity() : v; // tune.hpp:164-166
  };
  if (n == 0) {
    res.x = x0;
    res.f = eval(x0);
    res.termination = "no tunable parameters";
    return res;
  }
  std::vector<std::vector<double>> pts(n + 1, clamp(x0));
  std::vector<double> vals(n + 1);
  for (std::size_t i = 0; i < n; ++i) {
    const double range = upper[i] - lower[i];
    double step = std::isfinite(range) ? opt.initial_step * range : opt.initial_step * std::max(std::fabs(x0[i]), 1.);
    if (pts[i + 1][i] + step > upper[i]) {
      step = -step;
    }
    pts[i + 1][i] += step;
    pts[i + 1] = clamp(pts[i + 1]);
  }
  for (std::size_t k = 0; k <= n; ++k) {
    vals[k] = eval(pts[k]);
  }
  res.termination = "maxeval reached";
  while (res.evaluations < opt.max_evaluations) {
    std::vector<std::size_t> order(n + 1);
    for (std::size_t k = 0; k <= n; ++k) {
      order[k] = k;
    }
    std::sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return vals[a] < vals[b]; });
    const std::size_t best = order[0], worst = order[n], second = order[n - 1];
    const double spread = std::fabs(vals[worst] - vals[best]);
    if (std::isfinite(vals[worst]) &&
        (spread <= opt.ftol_abs || spread <= opt.ftol_rel * std::fabs(vals[best]))) {
      res.termination = spread <= opt.ftol_abs ? "ftol_abs reached" : "ftol_rel reached";
      break;
    }
    std::vector<double> centroid(n, 0.);
    for (std::size_t k = 0; k <= n; ++k) {
      if (k != worst) {
        for (std::size_t i = 0; i < n; ++i) {
          centroid[i] += pts[k][i] / static_cast<double>(n);
        }
      }
    }
    auto along = [&](double t) {
      std::vector<double> x(n);
      for (std::size_t i = 0; i < n; ++i) {
        x[i] = centroid[i] + t * (pts[worst][i] - centroid[i]);
      }
      return clamp(x);
    };
    const std::vector<double> xr = along(-1.);
    const double fr = eval(xr);
    if (fr < vals[best]) {
      const std::vector<double> xe = along(-2.);
      const double fe = eval(xe);
      if (fe < fr) {
        pts[worst] = xe;
        vals[worst] = fe;
      } else {
        pts[worst] = xr;
        vals[worst] = fr;
      }
    } else if (fr < vals[second]) {
      pts[worst] = xr;
      vals[worst] = fr;
    } else {
      const std::vector<double> xc = fr < vals[worst] ? along(-0.5) : along(0.5);
      const double fc = eval(xc);
      if (fc < std::min(fr, vals[worst])) {
        pts[worst] = xc;
        vals[worst] = fc;
      } else { // shrink towards the best vertex
        for (std::size_t k = 0; k <= n; ++k) {
          if (k != best) {
            for (std::size_t i = 0; i < n; ++i) {
              pts[k][i] = pts[best][i] + 0.5 * (pts[k][i] - pts[best][i]);
            }
            vals[k] = eval(pts[k]);
          }
        }
      }
    }
  }
  const std::size_t best = static_cast<std::size_t>(std::min_element(vals.begin(), vals.end()) - vals.begin());
  res.x = pts[best];
  res.f = vals[best];
  return res;
}

// ---- GenericTuner / ModelTuner (tune.hpp:110-330) ----------------------------------------------------------

struct GenericTuner {
  ParameterStore initial_params;
  SimplexOptions options;
  std::ostream *output_stream;

  explicit GenericTuner(const ParameterStore &initial_params_, std::ostream &output_stream_ = std::cout)
      : initial_params(initial_params_), output_stream(&output_stream_) {}

  // objective: double f(const ParameterStore &)
  template <typename ObjectiveFunction> ParameterStore tune(ObjectiveFunction &objective) {
    const TunableParameters tunable = get_tunable_parameters(initial_params);
    auto wrapped = [&](const std::vector<double> &x) {
      const ParameterStore params = set_tunable_params_values(initial_params, x);
      double metric = objective(params);
      if (std::isnan(metric)) {
        metric = std::numeric_limits<double>::infinity();
      }
      (*output_stream) << "-------------------" << std::endl
                       << pretty_params(params) << "objective: " << metric << std::endl
                       << "-------------------" << std::endl;
      return metric;
    };
    last_result = minimize_simplex(wrapped, tunable.values, tunable.lower_bounds, tunable.upper_bounds, options);
    const ParameterStore output = set_tunable_params_values(initial_params, last_result.x);
    (*output_stream) << "==================" << std::endl
                     << "TUNED PARAMS" << std::endl
                     << "minimum: " << last_result.f << std::endl
                     << "termination: " << last_result.termination << " after " << last_result.evaluations
                     << " evaluations" << std::endl
                     << "==================" << std::endl
                     << pretty_params(output) << std::endl;
    return output;
  }

  SimplexResult last_result;
};

enum class TuningMetricAggregator { Sum, Mean }; // tune.hpp:232-243 (mean_aggregator is the default)

template <typename ModelType, typename MetricType, class FeatureType> struct ModelTuner {
  ModelType model;
  MetricType metric;
  std::vector<RegressionDataset<FeatureType>> datasets;
  TuningMetricAggregator aggregator = TuningMetricAggregator::Mean;
  std::ostream *output_stream;
  SimplexOptions options;
  std::vector<std::shared_ptr<Device>> devices; // finite differences / batches are dealt over these

  ModelTuner(const ModelType &model_, const MetricType &metric_,
             const std::vector<RegressionDataset<FeatureType>> &datasets_, std::ostream &output_stream_ = std::cout)
      : model(model_), metric(metric_), datasets(datasets_), output_stream(&output_stream_) {
    devices.push_back(model.device());
  }

  // One handle per GPU of this process: device ordinals 0 .. count-1.
  void set_devices(int count) {
    devices.clear();
    for (int d = 0; d < count; ++d) {
      devices.push_back(std::make_shared<Device>(d));
    }
  }

  // The objective of tune.hpp:277-286 on device `lane % devices.size()`: a copy of the model with the
  // candidate parameters, the metric on every dataset, aggregated.
  double objective(const ParameterStore &params, std::size_t lane = 0) const {
    ModelType m(model);
    m.set_params(params);
    m.set_device(devices[lane % devices.size()]);
    double total = 0.;
    for (const auto &d : datasets) {
      total += metric(d, m);
    }
    return aggregator == TuningMetricAggregator::Mean ? total / static_cast<double>(datasets.size()) : total;
  }

  // Objective values of a batch of candidates, candidate i on device i % devices.size() (one host thread
  // per device): what a population-based or finite-difference caller hands out per iteration.
  std::vector<double> evaluate(const std::vector<ParameterStore> &candidates) const {
    std::vector<double> out(candidates.size());
    deal_over_lanes(candidates.size(), devices.size(),
                    [&](std::size_t i, std::size_t lane) { out[i] = objective(candidates[i], lane); });
    return out;
  }

  std::vector<double> gradient(const ParameterStore &params, double f_val) const {
    return compute_gradient_on_lanes([&](const ParameterStore &p, std::size_t lane) { return objective(p, lane); },
                                     params, f_val, devices.size());
  }

  ParameterStore tune() {
    auto obj = [&](const ParameterStore &params) { return objective(params, 0); };
    GenericTuner generic(model.get_params(), *output_stream);
    generic.options = options;
    const ParameterStore out = generic.tune(obj);
    last_result = generic.last_result;
    return out;
  }

  SimplexResult last_result;
};

template <typename ModelType, typename MetricType, typename FeatureType>
auto get_tuner(const ModelType &model, const MetricType &metric,
               const std::vector<RegressionDataset<FeatureType>> &datasets, std::ostream &output_stream = std::cout) {
  return ModelTuner<ModelType, MetricType, FeatureType>(model, metric, datasets, output_stream);
}
template <typename ModelType, typename MetricType, typename FeatureType>
auto get_tuner(const ModelType &model, const MetricType &metric, const RegressionDataset<FeatureType> &dataset,
               std::ostream &output_stream = std::cout) {
  return get_tuner(model, metric, std::vector<RegressionDataset<FeatureType>>{dataset}, output_stream);
}

} // namespace albatross_b200
