/*
 * TEST INFRASTRUCTURE ONLY.  Exact-GP path of the reference (fit / predict / log_likelihood /
 * cross-validation), compiled from the reference headers in place.
 *
 * Reference entry points exercised:
 *   gp_from_covariance, GaussianProcessBase::{_fit_impl,_predict_impl,log_likelihood}
 *       include/albatross/src/models/gp.hpp:285-366,443-451
 *   Fit<GPFit<...>> ctor                     gp.hpp:61-69
 *   gp_cross_validated_predictions           gp.hpp:467-482
 *   held_out_predictions                     include/albatross/src/evaluation/cross_validation_utils.hpp:199-232
 *   LeaveOneOutLikelihood                    include/albatross/src/evaluation/model_metrics.hpp:59-71
 */
#include "ref_common.h"

using namespace refshim;

namespace {

template <typename X>
albatross::RegressionDataset<X> make_dataset(const double *feats, int64_t n, int dim,
                                             const double *y, const double *yvar) {
  const auto xs = FeatureIO<X>::load(feats, n, dim);
  const Eigen::Map<const Eigen::VectorXd> ymap(y, n);
  if (yvar != nullptr) {
    const Eigen::Map<const Eigen::VectorXd> vmap(yvar, n);
    const Eigen::VectorXd mean = ymap;
    const Eigen::VectorXd variance = vmap;
    const albatross::MarginalDistribution targets(mean, variance);
    return albatross::RegressionDataset<X>(xs, targets);
  }
  return albatross::RegressionDataset<X>(xs, Eigen::VectorXd(ymap));
}

} // namespace

/*
 * model.fit(dataset): information = (K + diag(yvar))^-1 y.  Optional outputs: the packed LDLT
 * (n*n), transpositions, vectorD.
 */
REF_API int ref_gp_fit(int cov_id, const double *params, const double *feats, int64_t n, int dim,
                       const double *y, const double *yvar, int nthreads, double *information,
                       double *ldlt_out, int64_t *transpositions, double *vector_d) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto dataset = make_dataset<X>(feats, n, dim, y, yvar);
    return with_gp_cov<X>(cov_id, params, [&](const auto &cov) {
      auto model = albatross::gp_from_covariance(cov, "oracle");
      if (nthreads > 1) {
        model.set_thread_pool(
            std::make_shared<ThreadPool>(static_cast<std::size_t>(nthreads)));
      }
      const auto fit_model = model.fit(dataset);
      const auto &fit = fit_model.get_fit();
      copy_out(fit.information, information);
      if (ldlt_out != nullptr) {
        copy_out(Eigen::MatrixXd(fit.train_covariance.matrixLDLT()), ldlt_out);
      }
      if (transpositions != nullptr) {
        for (int64_t i = 0; i < n; ++i) {
          transpositions[i] = fit.train_covariance.transpositionsP().indices()[i];
        }
      }
      if (vector_d != nullptr) {
        copy_out(Eigen::VectorXd(fit.train_covariance.vectorD()), vector_d);
      }
    });
  });
}

/*
 * model.fit(dataset).predict(test).{mean,marginal,joint}().  what: 0 mean, 1 marginal, 2 joint,
 * 5 (= 1 | 4) marginal of predict_with_measurement_noise(test).
 * mean_out: p ; var_out: p (marginal) ; cov_out: p*p (joint).
 */
REF_API int ref_gp_predict(int cov_id, const double *params, const double *feats, int64_t n,
                           int dim, const double *y, const double *yvar, const double *test,
                           int64_t p, int what, double *mean_out, double *var_out,
                           double *cov_out) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto dataset = make_dataset<X>(feats, n, dim, y, yvar);
    const auto test_features = FeatureIO<X>::load(test, p, dim);
    return with_gp_cov<X>(cov_id, params, [&](const auto &cov) {
      const auto model = albatross::gp_from_covariance(cov, "oracle");
      const auto fit_model = model.fit(dataset);
      if (what & 4) { // fit_model.predict_with_measurement_noise(test), fit_model.hpp:54-62
        const albatross::MarginalDistribution m =
            fit_model.predict_with_measurement_noise(test_features).marginal();
        copy_out(m.mean, mean_out);
        copy_out(Eigen::VectorXd(m.covariance.diagonal()), var_out);
        return;
      }
      const auto prediction = fit_model.predict(test_features);
      if (what == 0) {
        copy_out(prediction.mean(), mean_out);
      } else if (what == 1) {
        const albatross::MarginalDistribution m = prediction.marginal();
        copy_out(m.mean, mean_out);
        copy_out(Eigen::VectorXd(m.covariance.diagonal()), var_out);
      } else {
        const albatross::JointDistribution j = prediction.joint();
        copy_out(j.mean, mean_out);
        copy_out(j.covariance, cov_out);
      }
    });
  });
}

/*
 * nll_out = -model.log_likelihood(dataset) with the prior term removed, i.e. the data term
 * 0.5 (log|K| + y^T K^-1 y + n log 2 pi) of likelihood.hpp:38-47; prior_out = prior_log_likelihood().
 */
REF_API int ref_gp_nll(int cov_id, const double *params, const double *feats, int64_t n, int dim,
                       const double *y, double *nll_out, double *prior_out) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto dataset = make_dataset<X>(feats, n, dim, y, nullptr);
    return with_gp_cov<X>(cov_id, params, [&](const auto &cov) {
      const auto model = albatross::gp_from_covariance(cov, "oracle");
      const double prior = model.prior_log_likelihood();
      const double ll = model.log_likelihood(dataset);
      if (nll_out != nullptr) {
        *nll_out = -(ll - prior);
      }
      if (prior_out != nullptr) {
        *prior_out = prior;
      }
    });
  });
}

/*
 * model.cross_validate().predict(dataset, grouper).{means,marginals,joints}().
 * what 0: mean_out (n, original order, concatenate_mean_predictions)
 * what 1: mean_out + var_out (n, original order, concatenate_marginal_predictions)
 * what 2: mean_out (n, original order) + joint_out: the per-group covariance blocks back to back in
 *         std::map key order (col-major each).
 * loo_nll_out (optional): LeaveOneOutLikelihood / LeaveOneGroupOutLikelihood style score
 *         sum_g NLL(joint_g, truth_g) (model_metrics.hpp:59-90) without the prior term.
 */
REF_API int ref_gp_cv(int cov_id, const double *params, const double *feats, int64_t n, int dim,
                      const double *y, int grouper_kind, double grouper_arg, int nthreads,
                      int what, double *mean_out, double *var_out, double *joint_out,
                      double *loo_nll_out) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto dataset = make_dataset<X>(feats, n, dim, y, nullptr);
    return with_gp_cov<X>(cov_id, params, [&](const auto &cov) {
      auto model = albatross::gp_from_covariance(cov, "oracle");
      if (nthreads > 1) {
        model.set_thread_pool(
            std::make_shared<ThreadPool>(static_cast<std::size_t>(nthreads)));
      }
      auto run = [&](const auto &grouper) {
        const auto indexer = albatross::group_by(dataset, grouper).indexers();
        const auto cv = model.cross_validate().predict(dataset, indexer);
        if (what == 0) {
          copy_out(cv.mean(), mean_out);
        } else if (what == 1) {
          const albatross::MarginalDistribution m = cv.marginal();
          copy_out(m.mean, mean_out);
          copy_out(Eigen::VectorXd(m.covariance.diagonal()), var_out);
        } else {
          const auto joints = cv.joints();
          Eigen::VectorXd mean(n);
          double *cursor = joint_out;
          for (const auto &pair : indexer) {
            const auto &j = joints.at(pair.first);
            albatross::set_subset(j.mean, pair.second, &mean);
            if (cursor != nullptr) {
              std::copy(j.covariance.data(), j.covariance.data() + j.covariance.size(), cursor);
              cursor += j.covariance.size();
            }
          }
          copy_out(mean, mean_out);
        }
        if (loo_nll_out != nullptr) {
          const albatross::NegativeLogLikelihood<albatross::JointDistribution> nll;
          const Eigen::VectorXd scores = model.cross_validate().scores(nll, dataset, indexer);
          *loo_nll_out = scores.sum();
        }
      };
      if (grouper_kind == 0) {
        run(albatross::LeaveOneOutGrouper());
      } else {
        run(ShimGrouper<X>{grouper_kind, grouper_arg});
      }
    });
  });
}
