/*
 * TEST INFRASTRUCTURE ONLY.  Sparse (FITC/PITC) GP path of the reference with the default
 * DenseQRImplementation (Eigen::ColPivHouseholderQR), compiled from the reference headers in place.
 *
 * include/albatross/SparseGP:16-18 hard-includes SuiteSparse (absent here, and its SPQR variant is
 * out of scope), so this TU bypasses the umbrella: a dummy Eigen::SPQR / SPQR_create satisfies the
 * declarations that sparse_gp.hpp:72-79 needs, and the reference's own sparse_gp.hpp is included
 * directly.
 *
 * Reference entry points exercised:
 *   SparseGaussianProcessRegression::{_fit_impl, _predict_impl x3, log_likelihood}
 *       include/albatross/src/models/sparse_gp.hpp:381-404,468-536,539-603
 *   UniformlySpacedInducingPoints              sparse_gp.hpp:34-47
 */
#include "ref_common.h"

#include <Eigen/Sparse>

namespace Eigen {
template <class M> class SPQR {
public:
  SPQR() {}
};
} // namespace Eigen

namespace albatross {
inline std::unique_ptr<Eigen::SPQR<Eigen::SparseMatrix<double>>>
SPQR_create(const Eigen::SparseMatrix<double> &, const ThreadPool * = nullptr) {
  std::abort();
}
} // namespace albatross

#include <albatross/src/linalg/qr_utils.hpp>
#include <albatross/src/linalg/block_utils.hpp>
#include <albatross/src/utils/eigen_utils.hpp>
#include <albatross/src/models/sparse_gp.hpp>

using namespace refshim;

namespace {

/* Inducing points handed in from C (so any strategy can be reproduced by the caller). */
struct FixedInducingPoints {
  std::vector<double> u;
  template <typename CovarianceFunction>
  std::vector<double> operator()(const CovarianceFunction &, const std::vector<double> &) const {
    return u;
  }
};

template <typename Cov, typename Grouper>
auto make_sparse(const Cov &cov, const Grouper &grouper, const double *inducing, int64_t m,
                 double measurement_nugget, double inducing_nugget) {
  FixedInducingPoints strategy{std::vector<double>(inducing, inducing + m)};
  auto model = albatross::sparse_gp_from_covariance(cov, grouper, strategy, "oracle_sparse");
  if (measurement_nugget >= 0.) {
    model.set_param_value(albatross::details::measurement_nugget_name(), measurement_nugget);
  }
  if (inducing_nugget >= 0.) {
    model.set_param_value(albatross::details::inducing_nugget_name(), inducing_nugget);
  }
  return model;
}

} // namespace

/* linspace(min, max, m) as UniformlySpacedInducingPoints (sparse_gp.hpp:34-47) produces it. */
REF_API void ref_uniform_inducing_points(const double *feats, int64_t n, int64_t m, double *out) {
  const std::vector<double> xs(feats, feats + n);
  albatross::UniformlySpacedInducingPoints strategy(static_cast<std::size_t>(m));
  const auto u = strategy(0, xs);
  std::copy(u.begin(), u.end(), out);
}

/*
 * 1-D features only (the in-scope sparse configs are 1-D).
 * Nuggets < 0 keep the reference defaults (1e-8).
 * what: 0 mean, 1 marginal, 2 joint, -1 fit only.
 * Outputs (any may be null): information v (m), R (m*m), perm (m, P.indices()), rank,
 *   mean_out (p), var_out (p), cov_out (p*p), ll_out = model.log_likelihood(dataset) - prior.
 */
REF_API int ref_sparse_gp(int cov_id, const double *params, const double *feats, int64_t n,
                          const double *y, const double *yvar, const double *inducing, int64_t m,
                          int grouper_kind, double grouper_arg, double measurement_nugget,
                          double inducing_nugget, const double *test, int64_t p, int what,
                          double *information, double *R_out, int64_t *perm_out, int64_t *rank_out,
                          double *mean_out, double *var_out, double *cov_out, double *ll_out) {
  using X = double;
  const auto xs = FeatureIO<X>::load(feats, n, 1);
  const Eigen::Map<const Eigen::VectorXd> ymap(y, n);
  albatross::RegressionDataset<X> dataset =
      (yvar != nullptr)
          ? albatross::RegressionDataset<X>(
                xs, albatross::MarginalDistribution(
                        Eigen::VectorXd(ymap),
                        Eigen::VectorXd(Eigen::Map<const Eigen::VectorXd>(yvar, n))))
          : albatross::RegressionDataset<X>(xs, Eigen::VectorXd(ymap));
  return with_gp_cov<X>(cov_id, params, [&](const auto &cov) {
    auto run = [&](const auto &grouper) {
      const auto model =
          make_sparse(cov, grouper, inducing, m, measurement_nugget, inducing_nugget);
      if (what >= -1 && (information != nullptr || R_out != nullptr || what >= 0)) {
        const auto fit_model = model.fit(dataset);
        const auto &fit = fit_model.get_fit();
        copy_out(fit.information, information);
        copy_out(fit.R, R_out);
        if (perm_out != nullptr) {
          for (int64_t i = 0; i < m; ++i) {
            perm_out[i] = fit.P.indices()[i];
          }
        }
        if (rank_out != nullptr) {
          *rank_out = fit.numerical_rank;
        }
        if (what >= 0) {
          const std::vector<double> test_features(test, test + p);
          const auto prediction = fit_model.predict(test_features);
          if (what == 0) {
            copy_out(prediction.mean(), mean_out);
          } else if (what == 1) {
            const albatross::MarginalDistribution md = prediction.marginal();
            copy_out(md.mean, mean_out);
            copy_out(Eigen::VectorXd(md.covariance.diagonal()), var_out);
          } else {
            const albatross::JointDistribution jd = prediction.joint();
            copy_out(jd.mean, mean_out);
            copy_out(jd.covariance, cov_out);
          }
        }
      }
      if (ll_out != nullptr) {
        *ll_out = model.log_likelihood(dataset) - model.prior_log_likelihood();
      }
    };
    if (grouper_kind == 0) {
      run(albatross::LeaveOneOutGrouper());
    } else {
      run(ShimGrouper<X>{grouper_kind, grouper_arg});
    }
  });
}
