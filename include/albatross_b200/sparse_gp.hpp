// albatross_b200 C++ trait layer — the sparse (FITC / PITC) Gaussian process.
//
// User surface of src/models/sparse_gp.hpp:225-800: sparse_gp_from_covariance(cov, grouper, strategy,
// name), model.fit(dataset), predict(xs).mean()/marginal()/joint(), log_likelihood(dataset),
// measurement_nugget / inducing_nugget parameters, UniformlySpacedInducingPoints.  The reference
// builds K_fu, P and B = [A^-1/2 K_fu ; K_uu^T/2] on the host (three 32 GiB matrices at N = 2^20,
// M = 4096) and QR-factors B through a QRImplementation; here B is never formed on the host, so the
// model overrides fit wholesale and calls ab_sparse_fit (SURVEY.md §8b "QR concept").
#pragma once

#include <algorithm>

#include "gp.hpp"
#include "linalg.hpp"

namespace albatross_b200 {

namespace details {
constexpr double DEFAULT_NUGGET = 1e-8; // sparse_gp.hpp:20
inline std::string measurement_nugget_name() { return "measurement_nugget"; }
inline std::string inducing_nugget_name() { return "inducing_nugget"; }
} // namespace details

struct UniformlySpacedInducingPoints { // sparse_gp.hpp:34-47
  UniformlySpacedInducingPoints(std::size_t num_points_ = 10) : num_points(num_points_) {}
  template <typename CovarianceFunction>
  std::vector<double> operator()(const CovarianceFunction &, const std::vector<double> &features) const {
    const double min = *std::min_element(features.begin(), features.end());
    const double max = *std::max_element(features.begin(), features.end());
    return linspace(min, max, num_points);
  }
  std::size_t num_points;
};

// DenseQRImplementation / DeviceQR / BlockDiagonalLDLT: linalg.hpp.

template <typename InducingFeatureType> struct SparseGPFit {};

// Fit<SparseGPFit<U>>, sparse_gp.hpp:93-124: inducing features + device-resident (K_uu factor, R, v).
template <typename InducingFeatureType> struct Fit<SparseGPFit<InducingFeatureType>> {
  struct Holder {
    std::shared_ptr<Device> dev;
    ab_sparse f = nullptr;
    ~Holder() {
      if (f) {
        ab_sparse_free(dev->get(), f);
      }
    }
  };
  std::vector<InducingFeatureType> train_features;
  VectorXd information;
  std::shared_ptr<Holder> device_fit;

  Index numerical_rank() const { return static_cast<Index>(train_features.size()); }
  // sigma_R with identity permutation (R^T R = B^T B), linalg/qr_utils.hpp:18-27
  MatrixXd sigma_R() const {
    const Index m = static_cast<Index>(train_features.size());
    MatrixXd R(m, m);
    ALBATROSS_B200_CHECK(ab_sparse_export_R(device_fit->dev->get(), device_fit->f, R.data()));
    return R;
  }
  bool operator==(const Fit &o) const { return device_fit == o.device_fit; }
};

template <typename CovFunc, typename MeanFunc, typename GrouperFunction, typename InducingPointStrategy,
          typename QRImplementation = DenseQRImplementation>
class SparseGaussianProcessRegression
    : public GaussianProcessBase<CovFunc, MeanFunc,
                                 SparseGaussianProcessRegression<CovFunc, MeanFunc, GrouperFunction,
                                                                 InducingPointStrategy, QRImplementation>> {
public:
  using Base = GaussianProcessBase<CovFunc, MeanFunc,
                                   SparseGaussianProcessRegression<CovFunc, MeanFunc, GrouperFunction,
                                                                   InducingPointStrategy, QRImplementation>>;

  SparseGaussianProcessRegression() : Base() { initialize_params(); }
  SparseGaussianProcessRegression(const CovFunc &covariance_function, const MeanFunc &mean_function)
      : Base(covariance_function, mean_function) {
    initialize_params();
  }
  SparseGaussianProcessRegression(const CovFunc &covariance_function, const MeanFunc &mean_function,
                                  const GrouperFunction &independent_group_function,
                                  const InducingPointStrategy &inducing_point_strategy,
                                  const std::string &model_name)
      : Base(covariance_function, mean_function, model_name),
        inducing_point_strategy_(inducing_point_strategy),
        independent_group_function_(independent_group_function) {
    initialize_params();
  }

  void initialize_params() { // sparse_gp.hpp:279-287
    measurement_nugget_ = {details::DEFAULT_NUGGET, LogScaleUniformPrior(PARAMETER_EPSILON, PARAMETER_MAX)};
    inducing_nugget_ = {details::DEFAULT_NUGGET, LogScaleUniformPrior(PARAMETER_EPSILON, PARAMETER_MAX)};
  }

  ParameterStore get_params() const {
    ParameterStore params = Base::get_params();
    params[details::measurement_nugget_name()] = measurement_nugget_;
    params[details::inducing_nugget_name()] = inducing_nugget_;
    return params;
  }
  void set_param(const ParameterKey &name, const Parameter &param) {
    if (name == details::measurement_nugget_name()) {
      measurement_nugget_ = param;
    } else if (name == details::inducing_nugget_name()) {
      inducing_nugget_ = param;
    } else {
      Base::set_param(name, param);
    }
  }
  // ParameterHandling<Impl> lives in the base with Impl = this class, but name lookup for the
  // convenience setters must find this class's get_params / set_param: they do (CRTP downcast).

  // _fit_impl, sparse_gp.hpp:381-404.
  template <typename FeatureType>
  auto _fit_impl(const std::vector<FeatureType> &features, const MarginalDistribution &targets) const {
    const auto u = inducing_point_strategy_(this->covariance_function_, features);
    assert(u.size() > 0 && "Empty inducing points!");
    using U = typename std::decay<decltype(u[0])>::type;
    Fit<SparseGPFit<U>> fit;
    fit.train_features = u;
    fit.information = VectorXd(static_cast<Index>(u.size()));
    fit.device_fit = std::make_shared<typename Fit<SparseGPFit<U>>::Holder>();
    fit.device_fit->dev = this->device();
    sparse_call(u, features, targets, &fit.device_fit->f, fit.information.data(), nullptr);
    return fit;
  }

  // _predict_impl x3, sparse_gp.hpp:468-536.
  template <typename FeatureType, typename U>
  VectorXd _predict_impl(const std::vector<FeatureType> &features, const Fit<SparseGPFit<U>> &fit,
                         PredictTypeIdentity<VectorXd> &&) const {
    VectorXd mean(static_cast<Index>(features.size()));
    sparse_predict_call<U>(features, fit, AB_PREDICT_MEAN, mean.data(), nullptr, nullptr);
    add_mean(this->mean_function_, features, &mean);
    return mean;
  }
  template <typename FeatureType, typename U>
  MarginalDistribution _predict_impl(const std::vector<FeatureType> &features, const Fit<SparseGPFit<U>> &fit,
                                     PredictTypeIdentity<MarginalDistribution> &&) const {
    VectorXd mean(static_cast<Index>(features.size())), var(static_cast<Index>(features.size()));
    sparse_predict_call<U>(features, fit, AB_PREDICT_MARGINAL, mean.data(), var.data(), nullptr);
    add_mean(this->mean_function_, features, &mean);
    return MarginalDistribution(mean, var);
  }
  template <typename FeatureType, typename U>
  JointDistribution _predict_impl(const std::vector<FeatureType> &features, const Fit<SparseGPFit<U>> &fit,
                                  PredictTypeIdentity<JointDistribution> &&) const {
    const Index p = static_cast<Index>(features.size());
    VectorXd mean(p);
    MatrixXd cov(p, p);
    sparse_predict_call<U>(features, fit, AB_PREDICT_JOINT, mean.data(), nullptr, cov.data());
    add_mean(this->mean_function_, features, &mean);
    return JointDistribution(mean, cov);
  }

  // log_likelihood, sparse_gp.hpp:539-603.
  template <typename FeatureType> double log_likelihood(const RegressionDataset<FeatureType> &dataset) const {
    const auto u = inducing_point_strategy_(this->covariance_function_, dataset.features);
    double ll = 0.;
    sparse_call(u, dataset.features, dataset.targets, nullptr, nullptr, &ll);
    return ll + this->prior_log_likelihood();
  }

  InducingPointStrategy get_inducing_point_strategy() const { return inducing_point_strategy_; }
  GrouperFunction get_grouper_function() const { return independent_group_function_; }

private:
  template <typename U, typename FeatureType>
  void sparse_call(const std::vector<U> &u, const std::vector<FeatureType> &features,
                   const MarginalDistribution &targets, ab_sparse *fit_out, double *information,
                   double *log_likelihood) const {
    using M = Measurement<FeatureType>;
    // K_ff blocks are k(Measurement, Measurement), K_fu is k(Measurement, U), K_uu is k(U, U)
    // (sparse_gp.hpp:646-679): three programs, which differ when the tree holds a MeasurementOnly term
    // (the reference's standard sparse configuration, tests/lib/albatross/test/test_models.h:26-30).
    const Program p_ff = this->covariance_function_.template program<M, M>();
    const Program p_fu = this->covariance_function_.template program<M, U>();
    const Program p_uu = this->covariance_function_.template program<U, U>();
    const auto indexer = build_indexer(independent_group_function_, features);
    const GroupCSR csr = to_csr(indexer);
    const PackedFeatures f = pack_features(features);
    const PackedFeatures fu = pack_features(u);
    VectorXd y(targets.mean);
    // Appendix B.7 of SURVEY.md: the reference copies y BEFORE removing the mean (:665 precedes
    // :667-668), so y is not de-meaned; reproduced by not touching y here.
    const double *yvar = targets.has_covariance() ? targets.covariance.diagonal().data() : nullptr;
    const ab_handle h = this->device()->get();
    bool not_pd = false;
    if (fit_out != nullptr) {
      not_pd = ALBATROSS_B200_NOT_PD(ab_sparse_fit2(
          h, p_ff.data(), static_cast<int>(p_ff.size()), p_fu.data(), static_cast<int>(p_fu.size()), p_uu.data(),
          static_cast<int>(p_uu.size()), f.data.data(), f.n, f.dim, y.data(), yvar, fu.data.data(), fu.n,
          csr.indices.data(), csr.offsets.data(), csr.ngroups(), measurement_nugget_.value,
          inducing_nugget_.value, fit_out, information, log_likelihood));
    } else {
      not_pd = ALBATROSS_B200_NOT_PD(ab_sparse_log_likelihood2(
          h, p_ff.data(), static_cast<int>(p_ff.size()), p_fu.data(), static_cast<int>(p_fu.size()), p_uu.data(),
          static_cast<int>(p_uu.size()), f.data.data(), f.n, f.dim, y.data(), yvar, fu.data.data(), fu.n,
          csr.indices.data(), csr.offsets.data(), csr.ngroups(), measurement_nugget_.value,
          inducing_nugget_.value, log_likelihood));
    }
    if (not_pd) { // K_uu or a block of A was not positive definite: NaN, as the reference's LDLT would yield
      fill_nan(information, static_cast<std::size_t>(fu.n));
      fill_nan(log_likelihood, 1);
    }
  }

  template <typename U, typename FeatureType>
  void sparse_predict_call(const std::vector<FeatureType> &features, const Fit<SparseGPFit<U>> &fit,
                           int what, double *mean, double *var, double *cov) const {
    const Program cross = this->covariance_function_.template program<U, FeatureType>();
    const Program prior = this->covariance_function_.template program<FeatureType, FeatureType>();
    const PackedFeatures test = pack_features(features);
    if (fit.device_fit->f == nullptr) { // the fit met a matrix that was not positive definite
      const std::size_t p = static_cast<std::size_t>(test.n);
      fill_nan(mean, p);
      fill_nan(var, p);
      fill_nan(cov, p * p);
      return;
    }
    ALBATROSS_B200_CHECK(ab_sparse_predict2(fit.device_fit->dev->get(), fit.device_fit->f, cross.data(),
                                            static_cast<int>(cross.size()), prior.data(),
                                            static_cast<int>(prior.size()), test.data.data(), test.n, what, mean,
                                            var, cov));
  }

  Parameter measurement_nugget_;
  Parameter inducing_nugget_;
  InducingPointStrategy inducing_point_strategy_;
  GrouperFunction independent_group_function_;
};

// sparse_gp.hpp:733-798
template <typename CovFunc, typename MeanFunc, typename GrouperFunction, typename InducingPointStrategy,
          typename QRImplementation = DenseQRImplementation>
auto sparse_gp_from_covariance_and_mean(CovFunc &&covariance_function, MeanFunc &&mean_function,
                                        GrouperFunction &&grouper_function, InducingPointStrategy &&strategy,
                                        const std::string &model_name, QRImplementation = DenseQRImplementation{}) {
  return SparseGaussianProcessRegression<typename std::decay<CovFunc>::type, typename std::decay<MeanFunc>::type,
                                         typename std::decay<GrouperFunction>::type,
                                         typename std::decay<InducingPointStrategy>::type,
                                         typename std::decay<QRImplementation>::type>(
      covariance_function, mean_function, grouper_function, strategy, model_name);
}

template <typename CovFunc, typename GrouperFunction, typename InducingPointStrategy,
          typename QRImplementation = DenseQRImplementation>
auto sparse_gp_from_covariance(CovFunc &&covariance_function, GrouperFunction &&grouper_function,
                               InducingPointStrategy &&strategy, const std::string &model_name,
                               QRImplementation qr = DenseQRImplementation{}) {
  return sparse_gp_from_covariance_and_mean(std::forward<CovFunc>(covariance_function), ZeroMean(),
                                            std::forward<GrouperFunction>(grouper_function),
                                            std::forward<InducingPointStrategy>(strategy), model_name, qr);
}

} // namespace albatross_b200
