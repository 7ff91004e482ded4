// albatross_b200 C++ trait layer — the exact Gaussian-process model.
//
// User surface of the reference, unchanged (src/models/gp.hpp:115-560, src/core/model.hpp:22-170,
// src/core/fit_model.hpp:18-123, src/core/prediction.hpp:32-240, src/evaluation/cross_validation.hpp,
// src/evaluation/model_metrics.hpp):
//
//   auto model = gp_from_covariance(cov [, name]);
//   auto fit_model = model.fit(dataset);
//   fit_model.predict(xs).mean() / .marginal() / .joint();
//   model.log_likelihood(dataset);
//   model.cross_validate().predict(dataset, grouper).means() / .marginals() / .joints() / .mean() / .marginal();
//   model.cross_validate().scores(metric, dataset, grouper);   LeaveOneOutLikelihood<>()(dataset, model);
//   get_params / set_params / set_param_value / prior_log_likelihood / set_thread_pool (accepted, ignored).
//
// Every numerical step is one C-ABI call; K, its factor and all O(N^2) intermediates stay in HBM.
#pragma once

#include <cmath>
#include <memory>
#include <string>

#include "core.hpp"
#include "covariance.hpp"

namespace albatross_b200 {

template <typename T> struct PredictTypeIdentity { typedef T type; };

// Mean functions (src/covariance_functions/mean_function.hpp): an O(N) host step either side of the
// device path.  ZeroMean is the one every scoped config uses; any class with
// `double call(const X&) const`, get_params and set_param can take its place.
struct ZeroMean : public ParameterHandling<ZeroMean> {
  std::string get_name() const { return "zero_mean"; }
  ParameterStore get_params() const { return {}; }
  void set_param(const ParameterKey &, const Parameter &) { assert(false && "ZeroMean has no parameters"); }
  bool has_param(const ParameterKey &) const { return false; }
  template <typename X> double call(const X &) const { return 0.; }
};

// slope * x + offset on scalar features, polynomials.hpp:93-108 (Gaussian(0, 1000) priors).  Mean functions stay
// on the host: they are subtracted from the targets before the device fit and added to the predicted means.
struct LinearMean : public ParameterHandling<LinearMean> {
  LinearMean() {
    slope = {0., GaussianPrior(0., 1000.)};
    offset = {0., GaussianPrior(0., 1000.)};
  }
  std::string get_name() const { return "linear"; }
  ParameterStore get_params() const { return {{"slope", slope}, {"offset", offset}}; }
  void set_param(const ParameterKey &name, const Parameter &param) {
    if (name == "slope") {
      slope = param;
    } else if (name == "offset") {
      offset = param;
    } else {
      assert(false && "unknown parameter");
    }
  }
  bool has_param(const ParameterKey &name) const { return name == "slope" || name == "offset"; }
  double call(const double &x) const { return slope.value * x + offset.value; }
  template <typename X> double call(const Measurement<X> &x) const { return call(x.value); }
  Parameter slope, offset;
};

template <typename MeanFunc, typename X>
inline void remove_mean(const MeanFunc &mean_function, const std::vector<X> &features, VectorXd *y) {
  if (std::is_same<MeanFunc, ZeroMean>::value) {
    return;
  }
  for (std::size_t i = 0; i < features.size(); ++i) {
    (*y)[static_cast<Index>(i)] -= mean_function.call(features[i]);
  }
}

template <typename MeanFunc, typename X>
inline void add_mean(const MeanFunc &mean_function, const std::vector<X> &features, VectorXd *y) {
  if (std::is_same<MeanFunc, ZeroMean>::value) {
    return;
  }
  for (std::size_t i = 0; i < features.size(); ++i) {
    (*y)[static_cast<Index>(i)] += mean_function.call(features[i]);
  }
}

// Fit<GPFit<CovarianceRepresentation, FeatureType>>, gp.hpp:33-80, with the device factor as the
// covariance representation.
template <typename CovarianceRepresentation, typename FeatureType> struct GPFit {};
template <typename FitType> struct Fit {};

template <typename FeatureType> struct Fit<GPFit<DeviceLDLT, FeatureType>> {
  std::vector<FeatureType> train_features;
  DeviceLDLT train_covariance;
  VectorXd information;

  Fit() = default;
  Fit(const std::vector<FeatureType> &features, DeviceLDLT &&factor, VectorXd &&information_)
      : train_features(features), train_covariance(std::move(factor)),
        information(std::move(information_)) {}

  // gp.hpp:61-69: from a host covariance (cov += targets.covariance; factor; information = K^-1 y)
  Fit(const std::vector<FeatureType> &features, const MatrixXd &train_cov, const MarginalDistribution &targets)
      : train_features(features) {
    DeviceMatrix K(train_cov);
    K.add_diagonal(targets.covariance.diagonal());
    train_covariance = DeviceLDLT(std::move(K));
    information = train_covariance.solve(targets.mean);
  }

  bool operator==(const Fit &other) const {
    return train_covariance == other.train_covariance && information == other.information;
  }
};

template <typename FeatureType> using DeviceGPFit = Fit<GPFit<DeviceLDLT, FeatureType>>;

template <typename ModelType, typename FeatureType, typename FitType> class Prediction;
template <typename ModelType> class CrossValidation;

// FitModel, fit_model.hpp:18-123.
template <typename ModelType, typename FitType> class FitModel {
public:
  typedef ModelType model_type;
  typedef FitType fit_type;
  FitModel() = default;
  FitModel(const ModelType &model, FitType &&fit) : model_(model), fit_(std::move(fit)) {}

  template <typename PredictFeatureType>
  Prediction<ModelType, PredictFeatureType, FitType>
  predict(const std::vector<PredictFeatureType> &features) const {
    return Prediction<ModelType, PredictFeatureType, FitType>(model_, fit_, features);
  }

  template <typename PredictFeatureType>
  auto predict_with_measurement_noise(const std::vector<PredictFeatureType> &features) const {
    return predict(as_measurements(features));
  }

  // fit_model.update(dataset), fit_model.hpp:64-95: the fit of the concatenated data without a refit.
  template <typename FeatureType>
  auto update(const std::vector<FeatureType> &features, const MarginalDistribution &targets) const {
    auto updated = model_._update_impl(fit_, features, targets);
    return FitModel<ModelType, decltype(updated)>(model_, std::move(updated));
  }
  template <typename FeatureType> auto update(const RegressionDataset<FeatureType> &dataset) const {
    return update(dataset.features, dataset.targets);
  }
  template <typename FeatureType> void update_in_place(const RegressionDataset<FeatureType> &dataset) {
    fit_ = model_._update_impl(fit_, dataset.features, dataset.targets);
  }

  const FitType &get_fit() const { return fit_; }
  FitType &get_fit() { return fit_; }
  ModelType get_model() const { return model_; }

private:
  ModelType model_;
  FitType fit_; // cheap to copy: the device factor is shared
};

// Prediction, prediction.hpp:32-113: lazily evaluated .mean() / .marginal() / .joint().
template <typename ModelType, typename FeatureType, typename FitType> class Prediction {
public:
  Prediction(const ModelType &model, const FitType &fit, const std::vector<FeatureType> &features)
      : model_(model), fit_(fit), features_(features) {}

  VectorXd mean() const { return model_._predict_impl(features_, fit_, PredictTypeIdentity<VectorXd>()); }
  MarginalDistribution marginal() const {
    return model_._predict_impl(features_, fit_, PredictTypeIdentity<MarginalDistribution>());
  }
  JointDistribution joint() const {
    return model_._predict_impl(features_, fit_, PredictTypeIdentity<JointDistribution>());
  }
  template <typename PredictType> PredictType get(PredictTypeIdentity<PredictType> = PredictTypeIdentity<PredictType>()) const {
    return model_._predict_impl(features_, fit_, PredictTypeIdentity<PredictType>());
  }
  std::size_t size() const { return features_.size(); }

private:
  ModelType model_;
  FitType fit_;
  std::vector<FeatureType> features_;
};

// ---- prediction metrics (src/evaluation/prediction_metrics.hpp) --------------------------------

inline double negative_log_likelihood(double deviation, double variance) { // likelihood.hpp:21-24
  double ll = -deviation * deviation / (2 * variance);
  ll -= 0.5 * std::log(2 * M_PI * variance);
  return -ll;
}

// likelihood.hpp:53-67: 1 x 1 closed form, otherwise factor (on the device) and reduce.
inline double negative_log_likelihood(const VectorXd &deviation, const MatrixXd &covariance) {
  assert(deviation.size() == covariance.rows() && covariance.rows() == covariance.cols());
  if (deviation.size() == 1) {
    return negative_log_likelihood(deviation[0], covariance(0, 0));
  }
  return DeviceLDLT(covariance).negative_log_likelihood(deviation);
}

inline double negative_log_likelihood(const VectorXd &deviation, const DiagonalMatrixXd &covariance) {
  double nll = 0.; // likelihood.hpp:74-86
  for (Index i = 0; i < deviation.size(); ++i) {
    nll += negative_log_likelihood(deviation[i], covariance.diagonal()[i]);
  }
  return nll;
}

template <typename RequiredPredictType> struct PredictionMetric {
  typedef RequiredPredictType predict_type;
};

template <typename PredictType = JointDistribution> struct NegativeLogLikelihood;

template <> struct NegativeLogLikelihood<JointDistribution> : public PredictionMetric<JointDistribution> {
  double operator()(const JointDistribution &prediction, const MarginalDistribution &truth) const {
    VectorXd dev(prediction.mean.size()); // prediction_metrics.hpp:112-119
    MatrixXd cov(prediction.covariance);
    for (Index i = 0; i < dev.size(); ++i) {
      dev[i] = prediction.mean[i] - truth.mean[i];
      cov(i, i) += truth.covariance.diagonal()[i];
    }
    return negative_log_likelihood(dev, cov);
  }
};

template <> struct NegativeLogLikelihood<MarginalDistribution> : public PredictionMetric<MarginalDistribution> {
  double operator()(const MarginalDistribution &prediction, const MarginalDistribution &truth) const {
    VectorXd dev(prediction.mean.size()); // prediction_metrics.hpp:121-128
    VectorXd var(prediction.mean.size());
    for (Index i = 0; i < dev.size(); ++i) {
      dev[i] = prediction.mean[i] - truth.mean[i];
      var[i] = prediction.covariance.diagonal()[i] + truth.covariance.diagonal()[i];
    }
    return negative_log_likelihood(dev, DiagonalMatrixXd(var));
  }
};

struct RootMeanSquareError : public PredictionMetric<VectorXd> { // prediction_metrics.hpp:56-71
  double operator()(const VectorXd &prediction, const MarginalDistribution &truth) const {
    double sse = 0.;
    for (Index i = 0; i < prediction.size(); ++i) {
      const double e = prediction[i] - truth.mean[i];
      sse += e * e;
    }
    return std::sqrt(sse / static_cast<double>(prediction.size()));
  }
};

// ---- cross validation ----------------------------------------------------------------------------

// Prediction<CrossValidation<Model>, Feature, GroupIndexer<Key>>, cross_validation.hpp:29-247.  The
// model's held-out predictions come from ONE fit through ab_gp_cv (the reference's specialised path,
// gp.hpp:465-482 -> cross_validation_utils.hpp:199-232); the generic refit-per-fold path
// (folds.hpp:35-58, O(N^2) memory for leave-one-out) is deliberately not offered.
template <typename ModelType, typename FeatureType, typename GroupKey> class CVPrediction {
public:
  CVPrediction(const ModelType &model, const RegressionDataset<FeatureType> &dataset,
               const GroupIndexer<GroupKey> &indexer)
      : model_(model), dataset_(dataset), indexer_(indexer) {}

  Grouped<GroupKey, VectorXd> means() const {
    return Grouped<GroupKey, VectorXd>(
        model_.cross_validated_predictions(dataset_, indexer_, PredictTypeIdentity<VectorXd>()));
  }
  Grouped<GroupKey, MarginalDistribution> marginals() const {
    return Grouped<GroupKey, MarginalDistribution>(
        model_.cross_validated_predictions(dataset_, indexer_, PredictTypeIdentity<MarginalDistribution>()));
  }
  Grouped<GroupKey, JointDistribution> joints() const {
    return Grouped<GroupKey, JointDistribution>(
        model_.cross_validated_predictions(dataset_, indexer_, PredictTypeIdentity<JointDistribution>()));
  }
  // concatenate_{mean,marginal}_predictions, cross_validation_utils.hpp:59-100
  VectorXd mean() const {
    VectorXd pred(static_cast<Index>(dataset_.size()));
    for (const auto &pair : means()) {
      set_subset(pair.second, indexer_.at(pair.first), &pred);
    }
    return pred;
  }
  MarginalDistribution marginal() const {
    VectorXd m(static_cast<Index>(dataset_.size())), v(static_cast<Index>(dataset_.size()));
    for (const auto &pair : marginals()) {
      set_subset(pair.second.mean, indexer_.at(pair.first), &m);
      set_subset(VectorXd(pair.second.covariance.diagonal()), indexer_.at(pair.first), &v);
    }
    return MarginalDistribution(m, v);
  }
  JointDistribution joint() const = delete; // cross_validation.hpp:200-203

  Grouped<GroupKey, VectorXd> get(PredictTypeIdentity<VectorXd>) const { return means(); }
  Grouped<GroupKey, MarginalDistribution> get(PredictTypeIdentity<MarginalDistribution>) const { return marginals(); }
  Grouped<GroupKey, JointDistribution> get(PredictTypeIdentity<JointDistribution>) const { return joints(); }

private:
  ModelType model_;
  RegressionDataset<FeatureType> dataset_;
  GroupIndexer<GroupKey> indexer_;
};

template <typename T> struct is_group_indexer : std::false_type {};
template <typename K> struct is_group_indexer<std::map<K, GroupIndices>> : std::true_type {};

template <typename ModelType> class CrossValidation { // cross_validation.hpp:255-326
public:
  CrossValidation(const ModelType &model) : model_(model) {}

  template <typename FeatureType, typename GroupKey>
  CVPrediction<ModelType, FeatureType, GroupKey>
  predict(const RegressionDataset<FeatureType> &dataset, const GroupIndexer<GroupKey> &indexer) const {
    return CVPrediction<ModelType, FeatureType, GroupKey>(model_, dataset, indexer);
  }

  template <typename FeatureType, typename GrouperFunction,
            typename std::enable_if<!is_group_indexer<GrouperFunction>::value, int>::type = 0>
  auto predict(const RegressionDataset<FeatureType> &dataset, const GrouperFunction &grouper) const {
    return predict(dataset, build_indexer(grouper, dataset.features));
  }

  // One score per group, in key order (cross_validation.hpp:297-325 -> cross_validation_utils.hpp:102-130).
  template <typename MetricType, typename FeatureType, typename GroupKey>
  VectorXd scores(const MetricType &metric, const RegressionDataset<FeatureType> &dataset,
                  const GroupIndexer<GroupKey> &indexer) const {
    using PredictType = typename MetricType::predict_type;
    if (std::is_same<MetricType, NegativeLogLikelihood<JointDistribution>>::value &&
        !dataset.targets.has_covariance()) {
      return model_.cross_validated_joint_nll(dataset, indexer); // reduced on the device
    }
    const auto preds = predict(dataset, indexer).get(PredictTypeIdentity<PredictType>());
    VectorXd out(static_cast<Index>(indexer.size()));
    Index g = 0;
    for (const auto &pair : indexer) {
      out[g++] = metric(preds.at(pair.first), dataset.targets.subset(pair.second));
    }
    return out;
  }

  template <typename MetricType, typename FeatureType, typename GrouperFunction,
            typename std::enable_if<!is_group_indexer<GrouperFunction>::value, int>::type = 0>
  VectorXd scores(const MetricType &metric, const RegressionDataset<FeatureType> &dataset,
                  const GrouperFunction &grouper) const {
    return scores(metric, dataset, build_indexer(grouper, dataset.features));
  }

private:
  ModelType model_;
};

// ---- the model --------------------------------------------------------------------------------------

template <typename CovFunc, typename MeanFunc, typename ImplType>
class GaussianProcessBase : public ParameterHandling<ImplType> {
  static_assert(is_device_covariance<CovFunc>::value,
                "albatross_b200: covariance type has no device form (and there is no CPU fallback)");

public:
  GaussianProcessBase() : covariance_function_(), mean_function_(), model_name_(default_name()) {}
  GaussianProcessBase(const CovFunc &covariance_function)
      : covariance_function_(covariance_function), mean_function_(), model_name_(default_name()) {}
  GaussianProcessBase(const CovFunc &covariance_function, const std::string &model_name)
      : covariance_function_(covariance_function), mean_function_(), model_name_(model_name) {}
  GaussianProcessBase(const CovFunc &covariance_function, const MeanFunc &mean_function)
      : covariance_function_(covariance_function), mean_function_(mean_function),
        model_name_(default_name()) {}
  GaussianProcessBase(const CovFunc &covariance_function, const MeanFunc &mean_function,
                      const std::string &model_name)
      : covariance_function_(covariance_function), mean_function_(mean_function),
        model_name_(model_name) {}

  std::string get_name() const { return model_name_; }

  ParameterStore get_params() const { // gp.hpp:243-246
    return map_join(mean_function_.get_params(), covariance_function_.get_params());
  }
  void set_param(const ParameterKey &name, const Parameter &param) {
    bool found = false;
    if (covariance_function_.has_param(name)) {
      covariance_function_.set_param(name, param);
      found = true;
    }
    if (mean_function_.has_param(name)) {
      mean_function_.set_param(name, param);
      found = true;
    }
    assert(found && "unknown parameter");
    (void)found;
  }

  CovFunc get_covariance() const { return covariance_function_; }
  MeanFunc get_mean() const { return mean_function_; }
  template <typename Pool> void set_thread_pool(const std::shared_ptr<Pool> &) {} // accepted, ignored
  void set_device(std::shared_ptr<Device> dev) { device_ = std::move(dev); }
  std::shared_ptr<Device> device() const { return device_ ? device_ : Device::default_device(); }

  // model.hpp:118-130
  template <typename FeatureType>
  auto fit(const std::vector<FeatureType> &features, const MarginalDistribution &targets) const {
    auto f = impl()._fit_impl(features, targets);
    return FitModel<ImplType, decltype(f)>(impl(), std::move(f));
  }
  template <typename FeatureType> auto fit(const RegressionDataset<FeatureType> &dataset) const {
    return fit(dataset.features, dataset.targets);
  }

  CrossValidation<ImplType> cross_validate() const { return CrossValidation<ImplType>(impl()); }

  // _fit_impl, gp.hpp:285-294: training features are wrapped as Measurement<>s (:288); one fused
  // device call builds K (lower triangle), adds targets.covariance, factors in place and solves.
  template <typename FeatureType>
  DeviceGPFit<FeatureType> _fit_impl(const std::vector<FeatureType> &features,
                                     const MarginalDistribution &targets) const {
    using M = Measurement<FeatureType>;
    const Program prog = covariance_function_.template program<M, M>();
    const PackedFeatures f = pack_features(features);
    VectorXd y(targets.mean);
    remove_mean(mean_function_, features, &y);
    VectorXd information(y.size());
    const std::shared_ptr<Device> dev = device();
    ab_factor factor = nullptr;
    const double *yvar = targets.has_covariance() ? targets.covariance.diagonal().data() : nullptr;
    // not positive definite: the (unusable) factor is still owned by the fit, so nothing leaks and
    // is_positive_definite() reports it; information is NaN as the reference's would be
    if (ALBATROSS_B200_NOT_PD(ab_gp_fit(dev->get(), prog.data(), static_cast<int>(prog.size()), f.data.data(),
                                        f.n, f.dim, y.data(), yvar, &factor, information.data()))) {
      fill_nan(information.data(), static_cast<std::size_t>(information.size()));
    }
    return DeviceGPFit<FeatureType>(features, DeviceLDLT(dev, factor), std::move(information));
  }

  // _update_impl, gp.hpp:386-414 (+ BlockSymmetric, linalg/block_symmetric.hpp:46-133): the reference keeps the
  // old factor, A^-1 B and the Schur complement's factor side by side; the device extends the Cholesky factor
  // itself (ab_gp_update), so the updated fit is an ordinary DeviceGPFit of the concatenated features.
  template <typename FeatureType>
  DeviceGPFit<FeatureType> _update_impl(const DeviceGPFit<FeatureType> &fit_, const std::vector<FeatureType> &features,
                                        const MarginalDistribution &targets) const {
    using M = Measurement<FeatureType>;
    const Program prog = covariance_function_.template program<M, M>();
    std::vector<FeatureType> all(fit_.train_features);
    all.insert(all.end(), features.begin(), features.end());
    const PackedFeatures f_old = pack_features(fit_.train_features);
    const PackedFeatures f_new = pack_features(features);
    VectorXd y(targets.mean);
    remove_mean(mean_function_, features, &y);
    VectorXd information(static_cast<Index>(all.size()));
    const std::shared_ptr<Device> dev = fit_.train_covariance.device();
    ab_factor factor = nullptr;
    const double *yvar = targets.has_covariance() ? targets.covariance.diagonal().data() : nullptr;
    if (ALBATROSS_B200_NOT_PD(ab_gp_update(dev->get(), fit_.train_covariance.get(), prog.data(),
                                           static_cast<int>(prog.size()), f_old.data.data(), f_old.n, f_old.dim,
                                           fit_.information.data(), f_new.data.data(), f_new.n, y.data(), yvar, &factor,
                                           information.data()))) {
      fill_nan(information.data(), static_cast<std::size_t>(information.size()));
    }
    return DeviceGPFit<FeatureType>(all, DeviceLDLT(dev, factor), std::move(information));
  }

  // _predict_impl x3, gp.hpp:313-366.  The cross covariance is k(train as stored in the fit, test):
  // the fit keeps the UNWRAPPED features (gp.hpp:293), so measurement-only terms vanish unless the
  // caller predicts at Measurement<> features of a fit of Measurement<> features.
  template <typename FeatureType, typename FitFeatureType>
  VectorXd _predict_impl(const std::vector<FeatureType> &features, const DeviceGPFit<FitFeatureType> &gp_fit,
                         PredictTypeIdentity<VectorXd> &&) const {
    VectorXd mean(static_cast<Index>(features.size()));
    predict_call(features, gp_fit, AB_PREDICT_MEAN, mean.data(), nullptr, nullptr);
    add_mean(mean_function_, features, &mean);
    return mean;
  }
  template <typename FeatureType, typename FitFeatureType>
  MarginalDistribution _predict_impl(const std::vector<FeatureType> &features,
                                     const DeviceGPFit<FitFeatureType> &gp_fit,
                                     PredictTypeIdentity<MarginalDistribution> &&) const {
    VectorXd mean(static_cast<Index>(features.size())), var(static_cast<Index>(features.size()));
    predict_call(features, gp_fit, AB_PREDICT_MARGINAL, mean.data(), var.data(), nullptr);
    add_mean(mean_function_, features, &mean);
    return MarginalDistribution(mean, var);
  }
  template <typename FeatureType, typename FitFeatureType>
  JointDistribution _predict_impl(const std::vector<FeatureType> &features,
                                  const DeviceGPFit<FitFeatureType> &gp_fit,
                                  PredictTypeIdentity<JointDistribution> &&) const {
    const Index p = static_cast<Index>(features.size());
    VectorXd mean(p);
    MatrixXd cov(p, p);
    predict_call(features, gp_fit, AB_PREDICT_JOINT, mean.data(), nullptr, cov.data());
    add_mean(mean_function_, features, &mean);
    return JointDistribution(mean, cov);
  }

  // gp.hpp:417-423
  template <typename FeatureType> JointDistribution prior(const std::vector<FeatureType> &features) const {
    const auto m = as_measurements(features);
    VectorXd mean(static_cast<Index>(features.size()));
    for (Index i = 0; i < mean.size(); ++i) {
      mean[i] = 0.;
    }
    add_mean(mean_function_, features, &mean);
    return JointDistribution(mean, covariance_function_(m));
  }
  template <typename FeatureType> MatrixXd compute_covariance(const std::vector<FeatureType> &features) const {
    return covariance_function_(features);
  }

  // gp.hpp:442-451: a fresh Gram (no targets.covariance) + factorisation, reduced on the device.
  template <typename FeatureType> double log_likelihood(const RegressionDataset<FeatureType> &dataset) const {
    using M = Measurement<FeatureType>;
    const Program prog = covariance_function_.template program<M, M>();
    const PackedFeatures f = pack_features(dataset.features);
    VectorXd y(dataset.targets.mean);
    remove_mean(mean_function_, dataset.features, &y);
    double nll = 0.;
    if (ALBATROSS_B200_NOT_PD(ab_gp_nll(device()->get(), prog.data(), static_cast<int>(prog.size()),
                                        f.data.data(), f.n, f.dim, y.data(), &nll))) {
      return quiet_nan(); // the tuner maps NaN to an infinite objective (tune.hpp:164-166)
    }
    return -nll + this->prior_log_likelihood();
  }

  // gp_cross_validated_predictions, gp.hpp:465-482 (one fit, every group held out in turn).
  template <typename FeatureType, typename GroupKey>
  std::map<GroupKey, VectorXd>
  cross_validated_predictions(const RegressionDataset<FeatureType> &dataset,
                              const GroupIndexer<GroupKey> &indexer, PredictTypeIdentity<VectorXd>) const {
    const CVRaw raw = cv_call(dataset, indexer, AB_PREDICT_MEAN, false);
    std::map<GroupKey, VectorXd> out;
    for (const auto &pair : indexer) {
      out[pair.first] = subset(raw.mean, pair.second);
    }
    return out;
  }
  template <typename FeatureType, typename GroupKey>
  std::map<GroupKey, MarginalDistribution>
  cross_validated_predictions(const RegressionDataset<FeatureType> &dataset,
                              const GroupIndexer<GroupKey> &indexer,
                              PredictTypeIdentity<MarginalDistribution>) const {
    const CVRaw raw = cv_call(dataset, indexer, AB_PREDICT_MARGINAL, false);
    std::map<GroupKey, MarginalDistribution> out;
    for (const auto &pair : indexer) {
      out[pair.first] = MarginalDistribution(subset(raw.mean, pair.second), subset(raw.var, pair.second));
    }
    return out;
  }
  template <typename FeatureType, typename GroupKey>
  std::map<GroupKey, JointDistribution>
  cross_validated_predictions(const RegressionDataset<FeatureType> &dataset,
                              const GroupIndexer<GroupKey> &indexer,
                              PredictTypeIdentity<JointDistribution>) const {
    const CVRaw raw = cv_call(dataset, indexer, AB_PREDICT_JOINT, false);
    std::map<GroupKey, JointDistribution> out;
    std::size_t at = 0;
    for (const auto &pair : indexer) {
      const Index k = static_cast<Index>(pair.second.size());
      MatrixXd cov(k, k);
      for (Index e = 0; e < k * k; ++e) {
        cov.data()[e] = raw.joint[at + static_cast<std::size_t>(e)];
      }
      at += pair.second.size() * pair.second.size();
      out[pair.first] = JointDistribution(subset(raw.mean, pair.second), cov);
    }
    return out;
  }

  // Per-group NLL of the held-out joint predictions against the (noise-free) truth, reduced on the
  // device: what scores(NegativeLogLikelihood<JointDistribution>, ...) returns.
  template <typename FeatureType, typename GroupKey>
  VectorXd cross_validated_joint_nll(const RegressionDataset<FeatureType> &dataset,
                                     const GroupIndexer<GroupKey> &indexer) const {
    return cv_call(dataset, indexer, AB_PREDICT_MEAN, true).scores;
  }

protected:
  struct CVRaw {
    VectorXd mean, var, scores;
    std::vector<double> joint;
  };

  template <typename FeatureType, typename GroupKey>
  CVRaw cv_call(const RegressionDataset<FeatureType> &dataset, const GroupIndexer<GroupKey> &indexer,
                int what, bool want_scores) const {
    const auto fit_model = impl().fit(dataset);
    const auto &gp_fit = fit_model.get_fit();
    const GroupCSR csr = to_csr(indexer);
    const Index n = static_cast<Index>(dataset.size());
    CVRaw raw;
    raw.mean = VectorXd(n);
    if (what == AB_PREDICT_MARGINAL) {
      raw.var = VectorXd(n);
    }
    if (what == AB_PREDICT_JOINT) {
      std::size_t total = 0;
      for (const auto &pair : indexer) {
        total += pair.second.size() * pair.second.size();
      }
      raw.joint.assign(total, 0.);
    }
    if (want_scores) {
      raw.scores = VectorXd(static_cast<Index>(indexer.size()));
    }
    // Note (gp.hpp:472-476): the held-out algebra works on the raw targets; the information vector
    // already accounts for the mean function.
    if (ALBATROSS_B200_NOT_PD(ab_gp_cv_scores(
            device()->get(), gp_fit.train_covariance.get(), dataset.targets.mean.data(),
            gp_fit.information.data(), csr.indices.data(), csr.offsets.data(), csr.ngroups(), what,
            raw.mean.data(), what == AB_PREDICT_MARGINAL ? raw.var.data() : nullptr,
            what == AB_PREDICT_JOINT ? raw.joint.data() : nullptr, nullptr,
            want_scores ? raw.scores.data() : nullptr))) {
      fill_nan(raw.mean.data(), static_cast<std::size_t>(raw.mean.size()));
      fill_nan(raw.var.data(), static_cast<std::size_t>(raw.var.size()));
      fill_nan(raw.scores.data(), static_cast<std::size_t>(raw.scores.size()));
      fill_nan(raw.joint.data(), raw.joint.size());
    }
    return raw;
  }

  template <typename FeatureType, typename FitFeatureType>
  void predict_call(const std::vector<FeatureType> &features, const DeviceGPFit<FitFeatureType> &gp_fit,
                    int what, double *mean, double *var, double *cov) const {
    static_assert(CovFunc::template is_defined_for<FeatureType, FeatureType>() &&
                      CovFunc::template is_defined_for<FitFeatureType, FeatureType>(),
                  "albatross_b200: CovFunc is not defined for FeatureType and FitFeatureType");
    // cross = k(train, test) and prior = k(test, test) must flatten to the same program for one fused
    // call; they differ only when exactly one side is a Measurement<> and the tree holds a
    // measurement-only term, in which case both contributions of that term are zero anyway except
    // on the prior of Measurement<> test features.
    const Program cross = covariance_function_.template program<FitFeatureType, FeatureType>();
    const Program prior = covariance_function_.template program<FeatureType, FeatureType>();
    const bool same = cross.size() == prior.size() &&
                      std::equal(cross.begin(), cross.end(), prior.begin(), [](const ab_op &a, const ab_op &b) {
                        return a.op == b.op && a.p0 == b.p0 && a.p1 == b.p1;
                      });
    const PackedFeatures train = pack_features(gp_fit.train_features);
    const PackedFeatures test = pack_features(features);
    const ab_handle h = device()->get();
    const int status =
        same ? ab_gp_predict(h, gp_fit.train_covariance.get(), cross.data(), static_cast<int>(cross.size()),
                             train.data.data(), train.n, train.dim, gp_fit.information.data(), test.data.data(),
                             test.n, what, mean, var, cov)
             : ab_gp_predict2(h, gp_fit.train_covariance.get(), cross.data(), static_cast<int>(cross.size()),
                              prior.data(), static_cast<int>(prior.size()), train.data.data(), train.n, train.dim,
                              gp_fit.information.data(), test.data.data(), test.n, what, mean, var, cov);
    if (is_not_positive_definite(status, "ab_gp_predict")) { // a fit of a matrix that was not PD
      const std::size_t p = static_cast<std::size_t>(test.n);
      fill_nan(mean, p);
      fill_nan(var, p);
      fill_nan(cov, p * p);
    }
  }

  static std::string default_name() { return "gaussian_process_regression"; }
  ImplType &impl() { return *static_cast<ImplType *>(this); }
  const ImplType &impl() const { return *static_cast<const ImplType *>(this); }

  CovFunc covariance_function_;
  MeanFunc mean_function_;
  std::string model_name_;
  std::shared_ptr<Device> device_;
};

template <typename CovFunc, typename MeanFunc = ZeroMean>
class GaussianProcessRegression
    : public GaussianProcessBase<CovFunc, MeanFunc, GaussianProcessRegression<CovFunc, MeanFunc>> {
public:
  using Base = GaussianProcessBase<CovFunc, MeanFunc, GaussianProcessRegression<CovFunc, MeanFunc>>;
  using Base::Base;
};

template <typename CovFunc> auto gp_from_covariance(CovFunc &&covariance_function, const std::string &model_name) {
  return GaussianProcessRegression<typename std::decay<CovFunc>::type>(
      std::forward<CovFunc>(covariance_function), model_name);
}
template <typename CovFunc> auto gp_from_covariance(CovFunc &&covariance_function) {
  return GaussianProcessRegression<typename std::decay<CovFunc>::type>(std::forward<CovFunc>(covariance_function));
}
template <typename CovFunc, typename MeanFunc>
auto gp_from_covariance_and_mean(CovFunc &&covariance_function, MeanFunc &&mean_function,
                                 const std::string &model_name = "gaussian_process_regression") {
  return GaussianProcessRegression<typename std::decay<CovFunc>::type, typename std::decay<MeanFunc>::type>(
      std::forward<CovFunc>(covariance_function), std::forward<MeanFunc>(mean_function), model_name);
}

// ---- model metrics: the tune() objectives (src/evaluation/model_metrics.hpp:59-106) ------------

template <typename PredictType = JointDistribution> struct LeaveOneOutLikelihood {
  template <typename FeatureType, typename ModelType>
  double operator()(const RegressionDataset<FeatureType> &dataset, const ModelType &model) const {
    NegativeLogLikelihood<PredictType> nll;
    const VectorXd scores = model.cross_validate().scores(nll, dataset, LeaveOneOutGrouper());
    return scores.sum() - model.prior_log_likelihood();
  }
};

// GroupFunction<FeatureType>, src/core/declarations.hpp:149.
template <typename FeatureType> using GroupFunction = std::string (*)(const FeatureType &);

template <typename FeatureType, typename PredictType = JointDistribution> class LeaveOneGroupOutLikelihood {
public:
  explicit LeaveOneGroupOutLikelihood(const GroupFunction<FeatureType> &grouper) : grouper_(grouper) {}
  template <typename ModelType>
  double operator()(const RegressionDataset<FeatureType> &dataset, const ModelType &model) const {
    NegativeLogLikelihood<PredictType> nll;
    const VectorXd scores = model.cross_validate().scores(nll, dataset, grouper_);
    return scores.sum() - model.prior_log_likelihood();
  }

private:
  GroupFunction<FeatureType> grouper_;
};

struct LeaveOneOutRMSE {
  template <typename FeatureType, typename ModelType>
  double operator()(const RegressionDataset<FeatureType> &dataset, const ModelType &model) const {
    RootMeanSquareError rmse;
    return model.cross_validate().scores(rmse, dataset, LeaveOneOutGrouper()).mean();
  }
};

// The marginal-likelihood objective of tune (src/tune/tune.hpp): -log_likelihood(dataset).
struct NegativeLogMarginalLikelihood {
  template <typename FeatureType, typename ModelType>
  double operator()(const RegressionDataset<FeatureType> &dataset, const ModelType &model) const {
    return -model.log_likelihood(dataset);
  }
};

} // namespace albatross_b200
