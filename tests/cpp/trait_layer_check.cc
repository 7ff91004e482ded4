// Exercises the C++ trait layer (include/albatross_b200/) the way a reference user writes code
// (examples/sinc_example.cc, tests/test_models.cc, tests/test_cross_validation.cc, tests/test_sparse_gp.cc)
// and dumps inputs + outputs so that tests/test_gpu_cpp_layer.py can compare them with the oracle.
//
//   trait_layer_check host            host-only checks (parameters, names, indexing, programs); no device
//   trait_layer_check gpu <outfile>   the device scenarios; writes "<key> <count>\n<values...>\n" records
//
// Built by tests/cpp/Makefile twice: with the stand-in matrix types (ships to the GPU box) and,
// where the reference's vendored Eigen exists, with Eigen types (compile check + host mode).
#define ALBATROSS_B200_EXCEPTIONS 1
#include <albatross_b200/albatross.hpp>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <random>
#include <sstream>

namespace ab = albatross_b200;
using ab::Index;
using ab::MatrixXd;
using ab::VectorXd;

static int failures = 0;
#define EXPECT(cond)                                                                               \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      std::fprintf(stderr, "EXPECT failed %s:%d: %s\n", __FILE__, __LINE__, #cond);                \
      ++failures;                                                                                  \
    }                                                                                              \
  } while (0)

static FILE *out = nullptr;

static void dump(const std::string &key, const double *p, std::size_t n) {
  std::fprintf(out, "%s %zu\n", key.c_str(), n);
  for (std::size_t i = 0; i < n; ++i) {
    std::fprintf(out, "%.17g\n", p[i]);
  }
}
static void dump(const std::string &key, const VectorXd &v) { dump(key, v.data(), static_cast<std::size_t>(v.size())); }
static void dump(const std::string &key, const MatrixXd &m) { dump(key, m.data(), static_cast<std::size_t>(m.size())); }
static void dump(const std::string &key, const std::vector<double> &v) { dump(key, v.data(), v.size()); }
static void dump(const std::string &key, double v) { dump(key, &v, 1); }
static void dump(const std::string &key, const ab::MarginalDistribution &m) {
  dump(key + ".mean", m.mean);
  dump(key + ".var", VectorXd(m.covariance.diagonal()));
}
static void dump(const std::string &key, const ab::JointDistribution &j) {
  dump(key + ".mean", j.mean);
  dump(key + ".cov", j.covariance);
}

using SE = ab::SquaredExponential<ab::EuclideanDistance>;
using EXPO = ab::Exponential<ab::EuclideanDistance>;
using M32 = ab::Matern32<ab::EuclideanDistance>;
using M52 = ab::Matern52<ab::EuclideanDistance>;
using Vec3 = std::array<double, 3>;

static ab::RegressionDataset<double> make_1d(std::size_t n, unsigned seed, double lo, double hi) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> u(lo, hi);
  std::vector<double> xs(n);
  VectorXd y(static_cast<Index>(n));
  for (std::size_t i = 0; i < n; ++i) {
    xs[i] = u(gen);
    y[static_cast<Index>(i)] = std::sin(xs[i]) + 0.1 * std::cos(10. * xs[i]);
  }
  return ab::RegressionDataset<double>(xs, y);
}

static int kfold_like_grouper(const double &x) { return static_cast<int>(x) % 8; } // bench_loo_cv.cc:95-105
// GroupFunction<double>: same partition as the int grouper of scenario_sinc (string keys order differently)
static std::string string_grouper(const double &x) {
  return std::to_string(static_cast<int>(x) % 8);
}

// ------------------------------------------------------------------------------------------------
// host-only checks
// ------------------------------------------------------------------------------------------------

static void host_checks() {
  // names and parameter plumbing, tests/test_covariance_functions.cc / test_parameter_handling_mixin.cc
  SE se(3.5, 5.7);
  ab::IndependentNoise<double> noise(1.0);
  auto cov = se + ab::measurement_only(noise);
  EXPECT(cov.get_name() ==
         "(squared_exponential[euclidean_distance]+measurement[independent_noise])");
  auto params = cov.get_params();
  EXPECT(params.size() == 3);
  EXPECT(params.at("squared_exponential_length_scale").value == 3.5);
  EXPECT(params.at("sigma_squared_exponential").value == 5.7);
  EXPECT(params.at("sigma_independent_noise").value == 1.0);
  cov.set_param_value("sigma_independent_noise", 0.25);
  EXPECT(cov.get_param_value("sigma_independent_noise") == 0.25);
  EXPECT(cov.prior_log_likelihood() == 0.);
  cov.set_param_value("squared_exponential_length_scale", -1.);
  EXPECT(cov.prior_log_likelihood() == -HUGE_VAL); // PositivePrior
  cov.set_param_value("squared_exponential_length_scale", 3.5);

  auto prod = (se * M32(2., 1.)) + (EXPO(1., 1.) * ab::Constant(3.));
  EXPECT(prod.get_name() == "((squared_exponential[euclidean_distance]*matern_32[euclidean_distance])+"
                            "(exponential[euclidean_distance]*constant))");

  // programs: postfix order, live parameters, measurement-only switching, one-sided sums
  using M = ab::Measurement<double>;
  auto p_mm = cov.program<M, M>();
  EXPECT(p_mm.size() == 3 && p_mm[0].op == AB_OP_SQUARED_EXPONENTIAL && p_mm[1].op == AB_OP_INDEPENDENT_NOISE &&
         p_mm[2].op == AB_OP_SUM && p_mm[1].p0 == 0.25);
  auto p_xm = cov.program<double, M>();
  EXPECT(p_xm.size() == 3 && p_xm[1].op == AB_OP_CONSTANT && p_xm[1].p0 == 0.);
  auto p_xx = cov.program<double, double>();
  EXPECT(p_xx.size() == 3 && p_xx[1].op == AB_OP_CONSTANT);
  // IndependentNoise<double> is not defined for 3-D features: the sum falls back to its defined side
  auto p_3d = cov.program<Vec3, Vec3>();
  EXPECT(p_3d.size() == 1 && p_3d[0].op == AB_OP_SQUARED_EXPONENTIAL);
  auto p_prod = prod.program<double, double>();
  EXPECT(p_prod.size() == 7 && p_prod[2].op == AB_OP_PRODUCT && p_prod[5].op == AB_OP_PRODUCT &&
         p_prod[6].op == AB_OP_SUM);
  static_assert(!decltype(noise)::is_defined_for<Vec3, Vec3>(), "noise<double> undefined for Vec3");
  static_assert(ab::is_device_feature<double>::value && ab::is_device_feature<Vec3>::value &&
                    ab::is_device_feature<M>::value && !ab::is_device_feature<std::string>::value,
                "device feature trait");

  // Polynomial<order> (polynomials.hpp:63-90): one parameter and one device term per degree, doubles only
  ab::Polynomial<1> linear(100.);
  EXPECT(linear.get_name() == "polynomial_1");
  EXPECT(linear.get_params().size() == 2 && linear.get_param_value("sigma_polynomial_1") == 100.);
  linear.set_param_value("sigma_polynomial_0", 3.);
  auto sinc_cov = linear + se + ab::measurement_only(noise);
  EXPECT(sinc_cov.get_params().size() == 5);
  auto p_poly = sinc_cov.program<M, M>();
  EXPECT(p_poly.size() == 7 && p_poly[0].op == AB_OP_POLYNOMIAL_TERM && p_poly[0].p0 == 3. && p_poly[0].p1 == 0. &&
         p_poly[1].op == AB_OP_POLYNOMIAL_TERM && p_poly[1].p0 == 100. && p_poly[1].p1 == 1. &&
         p_poly[2].op == AB_OP_SUM && p_poly[3].op == AB_OP_SQUARED_EXPONENTIAL && p_poly[6].op == AB_OP_SUM);
  static_assert(!ab::Polynomial<2>::is_defined_for<Vec3, Vec3>(), "Polynomial is defined between doubles");
  auto p_poly3 = (ab::Polynomial<2>(1.) + se).program<Vec3, Vec3>();
  EXPECT(p_poly3.size() == 1 && p_poly3[0].op == AB_OP_SQUARED_EXPONENTIAL);

  // model parameters and tunable view
  auto model = ab::gp_from_covariance(cov, "check");
  EXPECT(model.get_name() == "check");
  EXPECT(model.get_params().size() == 3);
  model.set_param_value("sigma_squared_exponential", 2.);
  EXPECT(model.get_covariance().get_param_value("sigma_squared_exponential") == 2.);
  ab::Parameter p{1., ab::LogScaleUniformPrior(1e-3, 1e2)};
  model.set_param("sigma_independent_noise", p);
  auto tunable = model.get_tunable_parameters();
  EXPECT(tunable.names.size() == 3);
  for (std::size_t i = 0; i < tunable.names.size(); ++i) {
    if (tunable.names[i] == "sigma_independent_noise") {
      EXPECT(tunable.values[i] == 0. && std::fabs(tunable.lower_bounds[i] - std::log(1e-3)) < 1e-15);
    }
  }
  EXPECT(std::fabs(model.prior_log_likelihood() - (-std::log(1e2 - 1e-3))) < 1e-15);
  auto copy = model; // models are copied liberally by the reference (fit_model.hpp:112)
  copy.set_param_value("sigma_squared_exponential", 9.);
  EXPECT(model.get_param_value("sigma_squared_exponential") == 2.);

  // priors (src/core/priors.hpp)
  EXPECT(ab::GaussianPrior(1., 2.).log_pdf(1.) == -0.5 * (1.8378770664093453 * 2 * std::log(2.)));
  EXPECT(ab::UniformPrior(0., 4.).log_pdf(5.) == -HUGE_VAL);
  EXPECT(ab::PositiveGaussianPrior(0., 3.).upper_bound() == 30.);
  EXPECT(ab::FixedPrior().is_fixed());

  // the integer contract, src/indexing/group_by.hpp:349-435 and subset.hpp
  auto ds = make_1d(100, 27, 0., 10.);
  auto loo = ab::build_indexer(ab::LeaveOneOutGrouper(), ds.features);
  EXPECT(loo.size() == 100 && loo.at(7) == ab::GroupIndices{7});
  auto kf = ab::build_indexer(ab::KFoldGrouper(3), ds.features);
  EXPECT(kf.size() == 3 && kf.at(1)[0] == 1 && kf.at(1)[1] == 4 && kf.at(2).size() == 33);
  auto by = ds.group_by(kfold_like_grouper).indexers();
  std::size_t total = 0;
  int last_key = -1;
  for (const auto &pair : by) {
    EXPECT(pair.first > last_key);
    last_key = pair.first;
    for (std::size_t k = 1; k < pair.second.size(); ++k) {
      EXPECT(pair.second[k] > pair.second[k - 1]); // encounter order == ascending
    }
    for (std::size_t i : pair.second) {
      EXPECT(static_cast<int>(ds.features[i]) % 8 == pair.first);
    }
    total += pair.second.size();
  }
  EXPECT(total == 100);
  ab::GroupCSR csr = ab::to_csr(by);
  EXPECT(csr.ngroups() == static_cast<int64_t>(by.size()) && csr.offsets.back() == 100);
  EXPECT(ab::indices_complement({1, 3}, 5) == (ab::GroupIndices{0, 2, 4}));
  auto sub = ds.subset(std::vector<std::size_t>{5, 2});
  EXPECT(sub.size() == 2 && sub.features[0] == ds.features[5] && sub.targets.mean[1] == ds.targets.mean[2]);
  auto lin = ab::linspace(0., 1., 5);
  EXPECT(lin.size() == 5 && lin[0] == 0. && lin[2] == 0.5);
  ab::MarginalDistribution md(ds.targets.mean);
  EXPECT(!md.has_covariance() && md.size() == 100);
}

// ------------------------------------------------------------------------------------------------
// device scenarios
// ------------------------------------------------------------------------------------------------

static void scenario_sinc() {
  // BASELINE configs[0]: 1-D SE(3.5, 5.7) + IndependentNoise(1.0), N = 1000, x ~ U[-10, 23]
  const std::size_t n = 1000;
  std::mt19937 gen(3);
  std::uniform_real_distribution<double> u(-10., 23.);
  std::normal_distribution<double> noise(0., 1.);
  std::vector<double> xs(n);
  VectorXd y(static_cast<Index>(n));
  for (std::size_t i = 0; i < n; ++i) {
    xs[i] = u(gen);
    const double t = xs[i] - 3.;
    const double sinc = t == 0. ? 1. : std::sin(t) / t;
    y[static_cast<Index>(i)] = std::sqrt(2.) * xs[i] + 3.14159 + 10. * sinc + noise(gen);
  }
  ab::RegressionDataset<double> data(xs, y);
  dump("sinc.x", xs);
  dump("sinc.y", y);

  auto cov = SE(3.5, 5.7) + ab::IndependentNoise<double>(1.0);
  auto model = ab::gp_from_covariance(cov, "sinc");
  const auto fit_model = model.fit(data);
  dump("sinc.information", fit_model.get_fit().information);
  EXPECT(fit_model.get_fit().train_covariance.is_positive_definite());
  EXPECT(fit_model.get_fit().train_covariance.rows() == static_cast<Index>(n));

  std::vector<double> grid = ab::linspace(-20., 33., 161); // example_utils.h:154-158
  dump("sinc.grid", grid);
  dump("sinc.predict.mean", fit_model.predict(grid).mean());
  dump("sinc.predict.marginal", fit_model.predict(grid).marginal());
  std::vector<double> few = {-5., 0.5, 3., 9.5, 21.};
  dump("sinc.few", few);
  dump("sinc.predict.joint", fit_model.predict(few).joint());
  dump("sinc.nll", -(model.log_likelihood(data) - model.prior_log_likelihood()));

  // leave-one-out and grouped cross validation (tests/test_cross_validation.cc)
  const auto loo = model.cross_validate().predict(data, ab::LeaveOneOutGrouper());
  dump("sinc.loo.marginal", loo.marginal());
  dump("sinc.loo.mean", loo.mean());
  dump("sinc.loo.likelihood", ab::LeaveOneOutLikelihood<>()(data, model));
  dump("sinc.loo.likelihood_marginal", ab::LeaveOneOutLikelihood<ab::MarginalDistribution>()(data, model));
  dump("sinc.loo.rmse", ab::LeaveOneOutRMSE()(data, model));

  auto grouper = [](const double &x) { return static_cast<int>(x) % 8; }; // bench_loo_cv.cc:95-105
  const auto indexer = data.group_by(grouper).indexers();
  const auto cv = model.cross_validate().predict(data, indexer);
  dump("sinc.cv.marginal", cv.marginal());
  std::vector<double> keys, sizes, joint_flat, means_flat;
  for (const auto &pair : cv.joints()) {
    keys.push_back(pair.first);
    sizes.push_back(static_cast<double>(pair.second.size()));
    for (Index e = 0; e < pair.second.covariance.size(); ++e) {
      joint_flat.push_back(pair.second.covariance.data()[e]);
    }
    for (Index e = 0; e < pair.second.mean.size(); ++e) {
      means_flat.push_back(pair.second.mean[e]);
    }
  }
  dump("sinc.cv.keys", keys);
  dump("sinc.cv.sizes", sizes);
  dump("sinc.cv.joint_blocks", joint_flat);
  dump("sinc.cv.group_means", means_flat);
  std::vector<double> idx_flat;
  for (const auto &pair : indexer) {
    for (std::size_t i : pair.second) {
      idx_flat.push_back(static_cast<double>(i));
    }
  }
  dump("sinc.cv.indices", idx_flat);
  ab::NegativeLogLikelihood<ab::JointDistribution> nll;
  dump("sinc.cv.scores", model.cross_validate().scores(nll, data, indexer));
  dump("sinc.cv.logo_likelihood", ab::LeaveOneGroupOutLikelihood<double>(string_grouper)(data, model));

  // targets with measurement variance: fit adds it (gp.hpp:65), scores add it (prediction_metrics.hpp:112-119)
  VectorXd yvar(static_cast<Index>(n));
  for (Index i = 0; i < yvar.size(); ++i) {
    yvar[i] = 0.05 + 0.001 * static_cast<double>(i % 7);
  }
  ab::RegressionDataset<double> noisy(xs, ab::MarginalDistribution(y, yvar));
  dump("sinc.yvar", yvar);
  dump("sinc.noisy.information", model.fit(noisy).get_fit().information);
  dump("sinc.noisy.scores", model.cross_validate().scores(nll, noisy, indexer));

  // live parameters: set_param_value changes the next call (SURVEY.md §5 "Config")
  auto tuned = model;
  tuned.set_param_value("squared_exponential_length_scale", 2.0);
  tuned.set_param_value("sigma_independent_noise", 0.5);
  dump("sinc.tuned.nll", -(tuned.log_likelihood(data) - tuned.prior_log_likelihood()));
}

static void scenario_measurement_only() {
  // examples/sinc_example.cc:76-81 ("radial_only"): SE + measurement_only(noise)
  auto data = make_1d(300, 11, 0., 10.);
  dump("meas.x", data.features);
  dump("meas.y", data.targets.mean);
  auto cov = SE(1.5, 2.0) + ab::measurement_only(ab::IndependentNoise<double>(0.3));
  auto model = ab::gp_from_covariance(cov);
  const auto fit_model = model.fit(data);
  dump("meas.information", fit_model.get_fit().information);
  // predict at some training locations (where a non-measurement-only noise term WOULD fire) and new ones
  std::vector<double> test = {data.features[0], data.features[17], 2.5, 7.75, 11.};
  dump("meas.test", test);
  dump("meas.predict.marginal", fit_model.predict(test).marginal());
  dump("meas.predict.joint", fit_model.predict(test).joint());
  dump("meas.predict_with_noise.marginal", fit_model.predict_with_measurement_noise(test).marginal());
  dump("meas.nll", -(model.log_likelihood(data) - model.prior_log_likelihood()));
}

static void scenario_polynomial() {
  // examples/sinc_example.cc:82-89 ("radial"): Polynomial<1> + SE + measurement_only(noise) on a sinc with a
  // linear trend (the example's truth, sinc_example_utils.hpp)
  const std::size_t n = 250;
  std::mt19937 gen(5);
  std::uniform_real_distribution<double> u(-3., 7.);
  std::vector<double> xs(n);
  VectorXd y(static_cast<Index>(n));
  for (std::size_t i = 0; i < n; ++i) {
    xs[i] = u(gen);
    const double z = 0.9 * (xs[i] - 1.);
    y[static_cast<Index>(i)] = 2. + 0.6 * xs[i] + 5. * (z == 0. ? 1. : std::sin(z) / z);
  }
  ab::RegressionDataset<double> data(xs, y);
  dump("poly.x", data.features);
  dump("poly.y", data.targets.mean);
  ab::Polynomial<1> linear(100.);
  linear.set_param_value("sigma_polynomial_0", 3.0);
  linear.set_param_value("sigma_polynomial_1", 0.7);
  auto cov = linear + SE(3.5, 5.7) + ab::measurement_only(ab::IndependentNoise<double>(0.4));
  auto model = ab::gp_from_covariance(cov);
  const auto fit_model = model.fit(data);
  dump("poly.information", fit_model.get_fit().information);
  std::vector<double> test = {data.features[0], data.features[9], -4., 0., 2.5, 8.};
  dump("poly.test", test);
  dump("poly.predict.marginal", fit_model.predict(test).marginal());
  dump("poly.predict.joint", fit_model.predict(test).joint());
  dump("poly.predict_with_noise.marginal", fit_model.predict_with_measurement_noise(test).marginal());
  dump("poly.nll", -(model.log_likelihood(data) - model.prior_log_likelihood()));
  dump("poly.gram", cov(std::vector<double>(xs.begin(), xs.begin() + 12)));
  dump("poly.loo.mean", model.cross_validate().predict(data, ab::LeaveOneOutGrouper()).mean());
}

static void scenario_3d() {
  // BASELINE configs[1] shape at a checkable size: 3-D features, SE + Matern52 (+ noise for the GP)
  const std::size_t n = 600;
  std::mt19937 gen(0);
  std::uniform_real_distribution<double> u(0., 10.);
  std::vector<Vec3> xs(n);
  std::vector<double> flat;
  VectorXd y(static_cast<Index>(n));
  for (std::size_t i = 0; i < n; ++i) {
    xs[i] = {u(gen), u(gen), u(gen)};
    flat.insert(flat.end(), xs[i].begin(), xs[i].end());
    y[static_cast<Index>(i)] = std::sin(xs[i][0]) + 0.1 * std::cos(10. * xs[i][0]);
  }
  dump("v3.x", flat);
  dump("v3.y", y);
  auto cov = SE(2.0, 1.5) + M52(3.0, 0.7);
  dump("v3.gram", cov(xs));
  std::vector<Vec3> ys(xs.begin(), xs.begin() + 50);
  dump("v3.cross", cov(xs, ys));
  dump("v3.diag", cov.diagonal(xs));
  dump("v3.scalar", cov(xs[0], xs[1]));

  auto gp_cov = cov + ab::IndependentNoise<Vec3>(0.1);
  auto model = ab::gp_from_covariance(gp_cov, "v3");
  ab::RegressionDataset<Vec3> data(xs, y);
  const auto fit_model = model.fit(data);
  dump("v3.information", fit_model.get_fit().information);
  std::vector<Vec3> test(xs.begin() + 100, xs.begin() + 110);
  for (auto &t : test) {
    t[1] += 0.125;
  }
  std::vector<double> tflat;
  for (const auto &t : test) {
    tflat.insert(tflat.end(), t.begin(), t.end());
  }
  dump("v3.test", tflat);
  dump("v3.predict.joint", fit_model.predict(test).joint());
  dump("v3.nll", -(model.log_likelihood(data) - model.prior_log_likelihood()));

  // sum-of-products composition (oracle menu 9) on the first coordinate
  std::vector<double> x1(n);
  for (std::size_t i = 0; i < n; ++i) {
    x1[i] = xs[i][0];
  }
  auto sop = SE(2., 1.5) * M32(3., 0.7) + EXPO(1.5, 0.9) * ab::Constant(1.1) + ab::IndependentNoise<double>(0.2);
  dump("sop.gram", sop(x1));

  // the CovarianceRepresentation on its own (tests/test_serializable_ldlt.cc)
  MatrixXd K = gp_cov(ab::as_measurements(xs));
  ab::DeviceLDLT ldlt(K);
  EXPECT(ldlt.is_positive_definite() && ldlt.rows() == static_cast<Index>(n));
  MatrixXd rhs(static_cast<Index>(n), 3);
  for (Index j = 0; j < 3; ++j) {
    for (Index i = 0; i < rhs.rows(); ++i) {
      rhs(i, j) = std::sin(0.37 * static_cast<double>(i + 1) * static_cast<double>(j + 1));
    }
  }
  dump("ldlt.K", K);
  dump("ldlt.rhs", rhs);
  dump("ldlt.solve", ldlt.solve(rhs));
  dump("ldlt.sqrt_solve", ldlt.sqrt_solve(rhs));
  dump("ldlt.logdet", ldlt.log_determinant());
  dump("ldlt.inverse_diagonal", ldlt.inverse_diagonal());
  std::vector<ab::GroupIndices> blocks = {{0, 5, 9}, {17}, {400, 2, 3, 599}};
  std::vector<double> blk;
  for (const auto &b : ldlt.inverse_blocks(blocks)) {
    blk.insert(blk.end(), b.data(), b.data() + b.size());
  }
  dump("ldlt.inverse_blocks", blk);
  MatrixXd LD;
  std::vector<int64_t> tr;
  ldlt.export_packed(&LD, &tr);
  dump("ldlt.packed", LD);
  bool identity = true;
  for (std::size_t i = 0; i < tr.size(); ++i) {
    identity = identity && tr[i] == static_cast<int64_t>(i);
  }
  EXPECT(identity);

  // a matrix that is not positive definite is reported, not UB (SURVEY.md §8b "Errors")
  MatrixXd bad(3, 3);
  bad(0, 0) = 1.; bad(1, 1) = -1.; bad(2, 2) = 1.;
  ab::DeviceLDLT bad_ldlt(bad);
  EXPECT(!bad_ldlt.is_positive_definite());
}

static void scenario_sparse() {
  // tests/test_sparse_gp.cc shape: 1-D, uniformly spaced inducing points, FITC and grouped (PITC)
  auto data = make_1d(3000, 5, 0., 10.);
  dump("sparse.x", data.features);
  dump("sparse.y", data.targets.mean);
  auto cov = SE(1., 1.) + ab::IndependentNoise<double>(0.1);
  std::vector<double> test = ab::linspace(0.5, 9.5, 19);
  dump("sparse.test", test);

  auto fitc = ab::sparse_gp_from_covariance(cov, ab::LeaveOneOutGrouper(), ab::UniformlySpacedInducingPoints(48), "fitc");
  EXPECT(fitc.get_params().count("measurement_nugget") == 1 && fitc.get_params().count("inducing_nugget") == 1);
  const auto fitc_fit = fitc.fit(data);
  EXPECT(fitc_fit.get_fit().train_features.size() == 48);
  dump("sparse.inducing", fitc_fit.get_fit().train_features);
  dump("sparse.fitc.marginal", fitc_fit.predict(test).marginal());
  dump("sparse.fitc.joint", fitc_fit.predict(test).joint());
  dump("sparse.fitc.ll", fitc.log_likelihood(data) - fitc.prior_log_likelihood());

  auto grouper = [](const double &x) { return static_cast<long>(std::floor(x * 2.)); }; // 20 groups
  auto pitc = ab::sparse_gp_from_covariance(cov, grouper, ab::UniformlySpacedInducingPoints(48), "pitc");
  const auto pitc_fit = pitc.fit(data);
  dump("sparse.pitc.marginal", pitc_fit.predict(test).marginal());
  dump("sparse.pitc.ll", pitc.log_likelihood(data) - pitc.prior_log_likelihood());
  const MatrixXd R = pitc_fit.get_fit().sigma_R();
  EXPECT(R.rows() == 48 && R(5, 2) == 0.); // upper triangular
}

static void scenario_sparse_measurement_only() {
  // the reference's standard sparse configuration (tests/lib/albatross/test/test_models.h:26-30,44-57):
  // the noise sits behind measurement_only, so K_ff, K_fu and K_uu are three different programs
  auto data = make_1d(2000, 9, 0., 10.);
  dump("spmo.x", data.features);
  dump("spmo.y", data.targets.mean);
  auto cov = SE(1., 1.) + ab::measurement_only(ab::IndependentNoise<double>(0.1));
  std::vector<double> test = ab::linspace(0.25, 9.75, 17);
  dump("spmo.test", test);
  auto fitc = ab::sparse_gp_from_covariance(cov, ab::LeaveOneOutGrouper(), ab::UniformlySpacedInducingPoints(40), "fitc");
  const auto fitc_fit = fitc.fit(data);
  dump("spmo.inducing", fitc_fit.get_fit().train_features);
  dump("spmo.fitc.marginal", fitc_fit.predict(test).marginal());
  dump("spmo.fitc.joint", fitc_fit.predict(test).joint());
  dump("spmo.fitc.ll", fitc.log_likelihood(data) - fitc.prior_log_likelihood());
  auto grouper = [](const double &x) { return static_cast<long>(std::floor(x * 2.)); };
  auto pitc = ab::sparse_gp_from_covariance(cov, grouper, ab::UniformlySpacedInducingPoints(40), "pitc");
  dump("spmo.pitc.marginal", pitc.fit(data).predict(test).marginal());
  dump("spmo.pitc.ll", pitc.log_likelihood(data) - pitc.prior_log_likelihood());
}

static void scenario_block_diagonal_and_qr() {
  // BlockDiagonal(LDLT) (linalg/block_diagonal.hpp) and the QR concept (sparse_gp.hpp:72-89, qr_utils.hpp)
  // exercised the way compute_internal_components / compute_sigma_qr use them (sparse_gp.hpp:368-375, :632-706)
  auto data = make_1d(300, 21, 0., 10.);
  auto cov = SE(1., 1.) + ab::IndependentNoise<double>(0.1);
  ab::BlockDiagonal K_ff;
  std::vector<std::size_t> sizes = {70, 1, 129, 100};
  std::size_t at = 0;
  for (std::size_t sz : sizes) {
    std::vector<double> sub(data.features.begin() + at, data.features.begin() + at + sz);
    K_ff.blocks.push_back(cov(ab::as_measurements(sub)));
    at += sz;
  }
  const ab::BlockDiagonalLDLT A_ldlt = K_ff.ldlt();
  EXPECT(A_ldlt.rows() == 300 && A_ldlt.is_positive_definite());
  MatrixXd rhs(300, 2);
  for (Index i = 0; i < 300; ++i) {
    rhs(i, 0) = data.targets.mean[i];
    rhs(i, 1) = std::cos(0.3 * static_cast<double>(i));
  }
  dump("bd.x", data.features);
  dump("bd.rhs", rhs);
  dump("bd.solve", A_ldlt.solve(rhs));
  dump("bd.sqrt_solve", A_ldlt.sqrt_solve(rhs));
  dump("bd.log_determinant", A_ldlt.log_determinant());
  dump("bd.dense", K_ff.toDense());
  // QR of a tall matrix: B = [A^-1/2 rhs-like columns ; I]
  MatrixXd B(340, 40);
  std::mt19937 gen(3);
  std::normal_distribution<double> nd(0., 1.);
  for (Index j = 0; j < 40; ++j) {
    for (Index i = 0; i < 340; ++i) {
      B(i, j) = nd(gen) + (i == j ? 3. : 0.);
    }
  }
  const auto qr = ab::DenseQRImplementation::compute(B, nullptr);
  EXPECT(qr->rank() == 40 && qr->rows() == 340 && qr->cols() == 40);
  const MatrixXd R = ab::get_R(*qr);
  EXPECT(R(5, 2) == 0. && R(2, 2) > 0.);
  dump("qr.B", B);
  dump("qr.R", R);
  MatrixXd r2(40, 2);
  for (Index i = 0; i < 40; ++i) {
    r2(i, 0) = 1. + static_cast<double>(i);
    r2(i, 1) = std::sin(static_cast<double>(i));
  }
  dump("qr.rhs", r2);
  dump("qr.sqrt_solve", ab::sqrt_solve(R, ab::get_P(*qr), r2));
}

static void scenario_tuner() {
  // examples/sinc_example.cc:35-38 shape: get_tuner(model, LeaveOneOutLikelihood<>(), dataset).tune()
  auto data = make_1d(250, 33, 0., 10.);
  dump("tune.x", data.features);
  dump("tune.y", data.targets.mean);
  auto cov = SE(2.5, 0.7) + ab::IndependentNoise<double>(0.3); // deliberately off: truth is ~(1, 1, <0.1)
  auto model = ab::gp_from_covariance(cov);
  std::ostringstream log;
  auto tuner = ab::get_tuner(model, ab::LeaveOneOutLikelihood<>(), data, log);
  // a batch of candidates (what a finite-difference gradient or a population step evaluates): objective
  // values are compared with the reference's objective (tune.hpp:277-286) on the host
  std::vector<ab::ParameterStore> candidates;
  const double ls[5] = {2.5, 1.0, 0.6, 1.7, 3.1}, sg[5] = {0.7, 1.0, 1.4, 0.9, 0.5}, sn[5] = {0.3, 0.1, 0.05, 0.2, 0.6};
  std::vector<double> flat;
  for (int c = 0; c < 5; ++c) {
    ab::ParameterStore p = model.get_params();
    p["squared_exponential_length_scale"].value = ls[c];
    p["sigma_squared_exponential"].value = sg[c];
    p["sigma_independent_noise"].value = sn[c];
    candidates.push_back(p);
    flat.insert(flat.end(), {ls[c], sg[c], sn[c]});
  }
  dump("tune.candidates", flat);
  const std::vector<double> values = tuner.evaluate(candidates);
  dump("tune.candidate_objectives", values);
  const std::vector<double> grad = tuner.gradient(candidates[1], values[1]);
  EXPECT(grad.size() == 3);
  dump("tune.gradient_at_1", grad);
  // NegativeLogMarginalLikelihood objective at the same candidates
  auto tuner_ml = ab::get_tuner(model, ab::NegativeLogMarginalLikelihood(), data, log);
  dump("tune.candidate_nll", tuner_ml.evaluate(candidates));
  // and the loop itself
  tuner.options.max_evaluations = 150;
  const double before = tuner.objective(model.get_params());
  const ab::ParameterStore tuned = tuner.tune();
  const double after = tuner.objective(tuned);
  EXPECT(after < before - 1.);
  EXPECT(tuner.last_result.evaluations <= 150 + 4);
  dump("tune.before_after", std::vector<double>{before, after, static_cast<double>(tuner.last_result.evaluations)});
  dump("tune.tuned", std::vector<double>{tuned.at("squared_exponential_length_scale").value,
                                         tuned.at("sigma_squared_exponential").value,
                                         tuned.at("sigma_independent_noise").value});
  EXPECT(log.str().find("TUNED PARAMS") != std::string::npos);
}

static void scenario_update() {
  // tests/test_gp.cc:182-219: fit on the first part, update with the second == fit on everything
  auto data = make_1d(900, 41, 0., 10.);
  auto cov = SE(1.2, 1.1) + ab::measurement_only(ab::IndependentNoise<double>(0.2));
  auto model = ab::gp_from_covariance(cov);
  std::vector<double> xa(data.features.begin(), data.features.begin() + 600), xb(data.features.begin() + 600, data.features.end());
  VectorXd ya(600), yb(300);
  for (Index i = 0; i < 600; ++i) {
    ya[i] = data.targets.mean[i];
  }
  for (Index i = 0; i < 300; ++i) {
    yb[i] = data.targets.mean[600 + i];
  }
  const auto first = model.fit(ab::RegressionDataset<double>(xa, ya));
  const auto updated = first.update(ab::RegressionDataset<double>(xb, yb));
  const auto full = model.fit(data);
  EXPECT(updated.get_fit().train_features.size() == 900);
  std::vector<double> test = ab::linspace(0.3, 9.7, 23);
  dump("update.x", data.features);
  dump("update.y", data.targets.mean);
  dump("update.test", test);
  dump("update.information", updated.get_fit().information);
  dump("update.full_information", full.get_fit().information);
  dump("update.marginal", updated.predict(test).marginal());
}

static void scenario_nugget_and_linear_mean() {
  // §8f-4: Nugget (nugget.hpp:32-49) and LinearMean (polynomials.hpp:93-108) in the layer
  ab::Nugget nugget;
  EXPECT(nugget.get_name() == "nugget" && nugget.get_params().at("nugget_sigma").value == 1e-8);
  EXPECT(nugget.get_params().at("nugget_sigma").is_fixed());
  auto data = make_1d(400, 17, 0., 10.);
  nugget.set_param_value("nugget_sigma", 0.3);
  ab::IndependentNoise<double> noise(0.3);
  auto cov_n = SE(1.4, 1.2) + nugget;
  auto cov_i = SE(1.4, 1.2) + noise;
  const MatrixXd Kn = cov_n(data.features), Ki = cov_i(data.features);
  EXPECT(Kn == Ki); // the same device leaf
  // a linear trend handled by the mean function: y' = y + 0.7 x + 2 with LinearMean(0.7, 2) == y with ZeroMean
  ab::LinearMean mean;
  mean.set_param_value("slope", 0.7);
  mean.set_param_value("offset", 2.);
  VectorXd y2(data.targets.mean);
  for (Index i = 0; i < y2.size(); ++i) {
    y2[i] += 0.7 * data.features[static_cast<std::size_t>(i)] + 2.;
  }
  auto model_z = ab::gp_from_covariance(cov_i);
  auto model_l = ab::gp_from_covariance_and_mean(cov_i, mean);
  EXPECT(model_l.get_params().count("slope") == 1 && model_l.get_params().count("offset") == 1);
  const auto fit_z = model_z.fit(data);
  const auto fit_l = model_l.fit(ab::RegressionDataset<double>(data.features, y2));
  double worst = 0.;
  for (Index i = 0; i < y2.size(); ++i) {
    worst = std::max(worst, std::fabs(fit_z.get_fit().information[i] - fit_l.get_fit().information[i]));
  }
  EXPECT(worst <= 1e-9);
  std::vector<double> test = ab::linspace(0.5, 9.5, 7);
  const VectorXd mz = fit_z.predict(test).mean(), ml = fit_l.predict(test).mean();
  for (Index i = 0; i < mz.size(); ++i) {
    EXPECT(std::fabs(ml[i] - (mz[i] + 0.7 * test[static_cast<std::size_t>(i)] + 2.)) <= 1e-9);
  }
  EXPECT(std::fabs(model_l.log_likelihood(ab::RegressionDataset<double>(data.features, y2)) -
                   model_l.prior_log_likelihood() - (model_z.log_likelihood(data) - model_z.prior_log_likelihood())) <= 1e-7);
}

static void scenario_not_positive_definite() {
  // Duplicate points without a noise term: K is singular.  The reference's pivoted LDLT proceeds and its
  // outputs are NaN / inf (GenericTuner maps a NaN objective to +inf, tune.hpp:164-166); the device reports
  // AB_ERR_NOT_PD and the layer turns it into the same observable result instead of aborting.
  std::vector<double> xs = {0., 1., 2., 2., 3., 1.};
  VectorXd y(6);
  for (Index i = 0; i < 6; ++i) {
    y[i] = std::sin(xs[static_cast<std::size_t>(i)]);
  }
  ab::RegressionDataset<double> data(xs, y);
  auto model = ab::gp_from_covariance(SE(1., 1.));
  const auto fit_model = model.fit(data); // must not abort, must not leak the factor
  EXPECT(!fit_model.get_fit().train_covariance.is_positive_definite());
  EXPECT(std::isnan(fit_model.get_fit().information[0]));
  const double ll = model.log_likelihood(data);
  EXPECT(std::isnan(ll));
  const auto pred = fit_model.predict(std::vector<double>{0.5, 1.5}).marginal();
  EXPECT(std::isnan(pred.mean[0]) && std::isnan(pred.covariance.diagonal()[1]));
  ab::LeaveOneOutLikelihood<> loo;
  EXPECT(std::isnan(loo(data, model)));
  dump("notpd.ll_is_nan", std::isnan(ll) ? 1. : 0.);
  // many failed objective evaluations in a row (a tuner wandering in a bad region) recycle the matrix
  for (int rep = 0; rep < 50; ++rep) {
    EXPECT(std::isnan(model.log_likelihood(data)));
  }
}

// host-only checks of the tuner front and the QR helpers (no device call)
static void host_checks_tune_and_linalg() {
  // bounded Nelder-Mead on a shifted quadratic with a minimum outside the box on one axis
  auto f = [](const std::vector<double> &x) { return (x[0] - 1.5) * (x[0] - 1.5) + 3. * (x[1] + 4.) * (x[1] + 4.) + 2.; };
  const ab::SimplexResult r = ab::minimize_simplex(f, {0., 0.}, {-1., -2.}, {5., 5.});
  EXPECT(std::fabs(r.x[0] - 1.5) < 1e-3 && std::fabs(r.x[1] + 2.) < 1e-9); // x1 clamped to its lower bound
  EXPECT(std::fabs(r.f - (2. + 3. * 4.)) < 1e-5 && r.evaluations < 2000);
  // NaN objectives are +inf (tune.hpp:164-166): the simplex walks away from them
  auto g = [](const std::vector<double> &x) { return x[0] < 0.5 ? std::nan("") : (x[0] - 2.) * (x[0] - 2.); };
  const ab::SimplexResult r2 = ab::minimize_simplex(g, {1.}, {0.}, {4.});
  EXPECT(std::fabs(r2.x[0] - 2.) < 1e-3);
  // finite differences, vector form (finite_difference.hpp:18-32), dealt over 3 lanes
  auto q = [](const std::vector<double> &x) { return x[0] * x[0] + 3. * x[1] - x[2] * x[2] * x[2]; };
  const std::vector<double> x0 = {1., 2., 0.5};
  const std::vector<double> grad = ab::compute_gradient(q, x0, q(x0), 3);
  EXPECT(std::fabs(grad[0] - 2.) < 1e-5 && std::fabs(grad[1] - 3.) < 1e-5 && std::fabs(grad[2] + 0.75) < 1e-5);
  // ParameterStore form: log-scale priors are stepped in log space, fixed parameters are skipped
  ab::ParameterStore ps;
  ps["a"] = {2., ab::PositivePrior()};
  ps["b"] = {3., ab::FixedPrior()};
  auto pf = [](const ab::ParameterStore &p) { return p.at("a").value * p.at("a").value + p.at("b").value; };
  const std::vector<double> pg = ab::compute_gradient(pf, ps, pf(ps));
  EXPECT(pg.size() == 1 && std::fabs(pg[0] - 4.) < 1e-4);
  // GenericTuner over a ParameterStore objective
  std::ostringstream log;
  ab::GenericTuner tuner(ps, log);
  auto obj = [](const ab::ParameterStore &p) { return (p.at("a").value - 0.7) * (p.at("a").value - 0.7); };
  const ab::ParameterStore tuned = tuner.tune(obj);
  EXPECT(std::fabs(tuned.at("a").value - 0.7) < 1e-3 && tuned.at("b").value == 3.);
  // sqrt_solve(R, P, rhs) = R^-T P^T rhs (qr_utils.hpp:35-45)
  MatrixXd R(3, 3);
  R(0, 0) = 2.; R(0, 1) = 1.; R(0, 2) = -1.;
  R(1, 0) = 0.; R(1, 1) = 3.; R(1, 2) = 0.5;
  R(2, 0) = 0.; R(2, 1) = 0.; R(2, 2) = 4.;
  MatrixXd rhs(3, 1);
  rhs(0, 0) = 4.; rhs(1, 0) = 5.; rhs(2, 0) = 6.;
  const MatrixXd z = ab::sqrt_solve(R, std::vector<Index>{2, 0, 1}, rhs); // P^T rhs = (6, 4, 5)
  EXPECT(std::fabs(z(0, 0) - 3.) < 1e-15 && std::fabs(z(1, 0) - (4. - 1. * 3.) / 3.) < 1e-15);
  EXPECT(std::fabs(z(2, 0) - (5. + 3. - 0.5 * (1. / 3.)) / 4.) < 1e-15);
}

int main(int argc, char **argv) {
  if (argc >= 2 && std::strcmp(argv[1], "host") == 0) {
    host_checks();
    host_checks_tune_and_linalg();
    std::printf("trait_layer_check host: %d failure(s)\n", failures);
    return failures == 0 ? 0 : 1;
  }
  if (argc >= 3 && std::strcmp(argv[1], "gpu") == 0) {
    out = std::fopen(argv[2], "w");
    if (!out) {
      std::perror("open");
      return 2;
    }
    try {
      host_checks();
      scenario_sinc();
      scenario_measurement_only();
      scenario_polynomial();
      scenario_3d();
      scenario_sparse();
      scenario_sparse_measurement_only();
      scenario_not_positive_definite();
      scenario_block_diagonal_and_qr();
      scenario_tuner();
      scenario_update();
      scenario_nugget_and_linear_mean();
      const ab_phase_times t = ab::Device::default_device()->timings();
      dump("kernel_launches", static_cast<double>(t.kernel_launches));
    } catch (const ab::device_error &e) {
      std::fprintf(stderr, "device_error: %s\n", e.what());
      return 3;
    }
    std::fclose(out);
    std::printf("trait_layer_check gpu: %d failure(s)\n", failures);
    return failures == 0 ? 0 : 1;
  }
  std::fprintf(stderr, "usage: %s host | gpu <outfile>\n", argv[0]);
  return 2;
}
