// INTEGRATION.md route 1b, executed: the REAL reference (its headers are included where they lie under
// /root/reference; nothing is copied) with albatross_b200::DeviceLDLT as the CovarianceRepresentation of
// its own Fit<GPFit<CovarianceRepresentation, FeatureType>> (src/models/gp.hpp:42-78).  The reference's
// constructor (:61-69) builds the fit — K += targets.covariance, CovarianceRepresentation(K), information =
// solve(targets.mean) — and the reference's own generic _predict_impl (:305-366) and gp_*_prediction helpers
// (:82-113) run on top of DeviceLDLT::solve.  K is still built by the reference on the host here: this is
// the minimal first hook of route 1b (the factorisation and every solve on the B200).
//
// The binary carries both implementations, so it compares them in-process against the stock model
// (SerializableLDLT) on the same inputs and fails if any output differs by more than 1e-9 (north_star).
//
//   route1b_check gpu      run on the device (exit 0 = parity)
//   route1b_check host     compile / link check only
#include <albatross/GP>

#define ALBATROSS_B200_EXCEPTIONS 1
#include <albatross_b200/device.hpp>

#include <cstdio>
#include <cstring>
#include <random>

namespace ab = albatross_b200;

// A model in the style of AdaptedGaussianProcess (tests/lib/albatross/test/test_models.h:186-229): only
// _fit_impl is overridden; everything else is the reference's GaussianProcessBase.
template <typename CovFunc>
class DeviceFactorGaussianProcess
    : public albatross::GaussianProcessBase<CovFunc, albatross::ZeroMean, DeviceFactorGaussianProcess<CovFunc>> {
public:
  using Base = albatross::GaussianProcessBase<CovFunc, albatross::ZeroMean, DeviceFactorGaussianProcess<CovFunc>>;
  DeviceFactorGaussianProcess(const CovFunc &cov) : Base(cov, "device_factor_gp") {}

  template <typename FeatureType>
  auto _fit_impl(const std::vector<FeatureType> &features, const albatross::MarginalDistribution &targets) const {
    const auto measurement_features = albatross::as_measurements(features);
    const Eigen::MatrixXd cov = this->covariance_function_(measurement_features, Base::threads_.get());
    // the reference's own constructor, instantiated with the device representation
    return albatross::Fit<albatross::GPFit<ab::DeviceLDLT, FeatureType>>(features, cov, targets);
  }
};

static double rel(const Eigen::MatrixXd &a, const Eigen::MatrixXd &b) {
  return (a - b).cwiseAbs().maxCoeff() / std::max(b.cwiseAbs().maxCoeff(), 1e-300);
}

int main(int argc, char **argv) {
  static_assert(albatross::has_solve<ab::DeviceLDLT, Eigen::MatrixXd>::value,
                "DeviceLDLT satisfies the reference's CovarianceRepresentation concept (gp.hpp:44-45)");
  if (argc < 2 || std::strcmp(argv[1], "gpu") != 0) {
    std::printf("route1b_check: built against the reference headers (host mode: nothing to run)\n");
    return 0;
  }
  const std::size_t n = 1500;
  std::mt19937 gen(7);
  std::uniform_real_distribution<double> u(0., 10.);
  std::vector<double> xs(n);
  Eigen::VectorXd y(static_cast<Eigen::Index>(n)), yvar(static_cast<Eigen::Index>(n));
  for (std::size_t i = 0; i < n; ++i) {
    xs[i] = u(gen);
    y[static_cast<Eigen::Index>(i)] = std::sin(xs[i]) + 0.1 * std::cos(10. * xs[i]);
    yvar[static_cast<Eigen::Index>(i)] = 0.01 + 0.02 * static_cast<double>(i % 5);
  }
  const albatross::RegressionDataset<double> dataset(xs, albatross::MarginalDistribution(y, yvar));
  const auto cov = albatross::SquaredExponential<albatross::EuclideanDistance>(1.3, 1.1) +
                   albatross::Matern52<albatross::EuclideanDistance>(3.0, 0.7) +
                   albatross::measurement_only(albatross::IndependentNoise<double>(0.1));
  std::vector<double> test;
  for (int i = 0; i < 37; ++i) {
    test.push_back(-0.5 + 0.3 * i);
  }

  const auto stock = albatross::gp_from_covariance(cov);
  const auto stock_fit = stock.fit(dataset);
  const DeviceFactorGaussianProcess<std::decay_t<decltype(cov)>> device_model(cov);
  int failures = 0;
  auto check = [&](const char *what, double err) {
    const bool ok = err <= 1e-9;
    std::printf("route1b %-28s rel err %.3e %s\n", what, err, ok ? "OK" : "FAIL");
    failures += ok ? 0 : 1;
  };
  try {
    const auto device_fit = device_model.fit(dataset);
    check("information", rel(device_fit.get_fit().information, stock_fit.get_fit().information));
    check("predict.mean", rel(device_fit.predict(test).mean(), stock_fit.predict(test).mean()));
    const albatross::MarginalDistribution dm = device_fit.predict(test).marginal(), sm = stock_fit.predict(test).marginal();
    check("predict.marginal mean", rel(dm.mean, sm.mean));
    check("predict.marginal variance", (Eigen::VectorXd(dm.covariance.diagonal()) - Eigen::VectorXd(sm.covariance.diagonal())).cwiseAbs().maxCoeff() / 1.7);
    const albatross::JointDistribution dj = device_fit.predict(test).joint(), sj = stock_fit.predict(test).joint();
    check("predict.joint covariance", (dj.covariance - sj.covariance).cwiseAbs().maxCoeff() / 1.7);
    // the rest of the representation's surface against SerializableLDLT on the same matrix
    const Eigen::MatrixXd K = cov(albatross::as_measurements(xs)) + Eigen::MatrixXd(yvar.asDiagonal());
    const Eigen::SerializableLDLT ref_ldlt(K);
    const ab::DeviceLDLT dev_ldlt(K);
    check("log_determinant", std::fabs(dev_ldlt.log_determinant() - ref_ldlt.log_determinant()) /
                                 std::fabs(ref_ldlt.log_determinant()));
    check("inverse_diagonal", rel(dev_ldlt.inverse_diagonal(), ref_ldlt.inverse_diagonal()));
    const Eigen::MatrixXd rhs = Eigen::MatrixXd::Random(static_cast<Eigen::Index>(n), 3);
    check("solve", rel(dev_ldlt.solve(rhs), ref_ldlt.solve(rhs)));
    const Eigen::MatrixXd s_dev = dev_ldlt.sqrt_solve(rhs), s_ref = ref_ldlt.sqrt_solve(rhs);
    check("sqrt_solve (Gram)", rel(s_dev.transpose() * s_dev, s_ref.transpose() * s_ref));
    // the square root itself is representation dependent (pivot order): S S^T rhs = K rhs is not
    check("sqrt_product (S S^T rhs)", rel(dev_ldlt.sqrt_product(dev_ldlt.sqrt_transpose() * rhs), K * rhs));
    check("sqrt_transpose_solve", rel(dev_ldlt.sqrt_transpose_solve(dev_ldlt.sqrt_solve(rhs)), ref_ldlt.solve(rhs)));
    check("negative_log_likelihood",
          std::fabs(dev_ldlt.negative_log_likelihood(y) - albatross::negative_log_likelihood(y, ref_ldlt)) /
              std::fabs(albatross::negative_log_likelihood(y, ref_ldlt)));
    if (!dev_ldlt.is_positive_definite() || !ref_ldlt.is_positive_definite()) {
      std::printf("route1b is_positive_definite FAIL\n");
      ++failures;
    }
  } catch (const ab::device_error &e) {
    std::fprintf(stderr, "device_error: %s\n", e.what());
    return 3;
  }
  std::printf("route1b_check gpu: %d failure(s)\n", failures);
  return failures == 0 ? 0 : 1;
}
