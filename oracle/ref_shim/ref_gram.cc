/*
 * TEST INFRASTRUCTURE ONLY.  Gram construction, LDLT wrapper, indexing and the benchmark data
 * generators of the reference, compiled from the reference headers in place.
 *
 * Reference entry points exercised:
 *   CovarianceFunction::operator()(vector<X>[, vector<Y>][, ThreadPool*])
 *       include/albatross/src/covariance_functions/covariance_function.hpp:128-151
 *   compute_covariance_matrix  include/albatross/src/covariance_functions/callers.hpp:38-166
 *   Eigen::SerializableLDLT    include/albatross/src/eigen/serializable_ldlt.hpp:19-199
 *   group_by / LeaveOneOutGrouper  include/albatross/src/indexing/group_by.hpp:349-403
 *   partition_triangular       include/albatross/src/indexing/block.hpp:25-44
 *   bench generators           benchmarks/bench_utils.h:25-85 (re-stated here: that header pulls in
 *                              google_benchmark-free code only, but lives outside include/)
 */
#include "ref_common.h"

#include <random>

using namespace refshim;

REF_API int ref_hardware_threads() {
  return static_cast<int>(std::thread::hardware_concurrency());
}

/* benchmarks/bench_utils.h:25-35 — U[0,10] from std::mt19937(seed), one draw per coordinate. */
REF_API void ref_random_features(int64_t n, int dim, uint32_t seed, double *out) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> dist(0., 10.);
  for (int64_t i = 0; i < n * dim; ++i) {
    out[i] = dist(gen);
  }
}

/* benchmarks/bench_utils.h:76-85 — y = sin x + 0.1 cos 10x (first coordinate). */
REF_API void ref_random_targets(const double *feats, int64_t n, int dim, double *out) {
  for (int64_t i = 0; i < n; ++i) {
    out[i] = std::sin(feats[i * dim]) + 0.1 * std::cos(10. * feats[i * dim]);
  }
}

/* benchmarks/bench_utils.h:37-45 — N(0,1) from std::mt19937(seed). */
REF_API void ref_random_normal(int64_t n, uint32_t seed, double *out) {
  std::mt19937 gen(seed);
  std::normal_distribution<double> dist(0., 1.);
  for (int64_t i = 0; i < n; ++i) {
    out[i] = dist(gen);
  }
}

REF_API int ref_gram_sym(int cov_id, const double *params, const double *feats, int64_t n, int dim,
                         int as_meas, int nthreads, double *out) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto xs = FeatureIO<X>::load(feats, n, dim);
    std::unique_ptr<ThreadPool> pool;
    if (nthreads > 1) {
      pool = std::make_unique<ThreadPool>(static_cast<std::size_t>(nthreads));
    }
    return with_cov<X>(cov_id, params, [&](const auto &cov) {
      Eigen::MatrixXd K;
      if (as_meas) {
        K = cov(albatross::as_measurements(xs), pool.get());
      } else {
        K = cov(xs, pool.get());
      }
      copy_out(K, out);
    });
  });
}

REF_API int ref_gram_cross(int cov_id, const double *params, const double *fx, int64_t n,
                           const double *fy, int64_t m, int dim, int nthreads, double *out) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto xs = FeatureIO<X>::load(fx, n, dim);
    const auto ys = FeatureIO<X>::load(fy, m, dim);
    std::unique_ptr<ThreadPool> pool;
    if (nthreads > 1) {
      pool = std::make_unique<ThreadPool>(static_cast<std::size_t>(nthreads));
    }
    return with_cov<X>(cov_id, params, [&](const auto &cov) {
      const Eigen::MatrixXd K = cov(xs, ys, pool.get());
      copy_out(K, out);
    });
  });
}

REF_API int ref_gram_diag(int cov_id, const double *params, const double *feats, int64_t n, int dim,
                          double *out) {
  return with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto xs = FeatureIO<X>::load(feats, n, dim);
    return with_cov<X>(cov_id, params, [&](const auto &cov) {
      for (std::size_t i = 0; i < xs.size(); ++i) {
        out[i] = cov(xs[i], xs[i]);
      }
    });
  });
}

/* Scalar evaluation cov(x, y) for 1-D features (golden tables, edge cases). */
REF_API double ref_cov_scalar(int cov_id, const double *params, double x, double y) {
  double value = NAN;
  with_cov<double>(cov_id, params, [&](const auto &cov) { value = cov(x, y); });
  return value;
}

/*
 * SerializableLDLT on a caller-provided symmetric matrix (col-major n*n).  Any output may be null.
 *   ldlt_out       packed factor (strict lower = L, diagonal = D), n*n
 *   transpositions n entries
 *   solve_rhs      n*nrhs  ->  solve_out = A^-1 rhs ; sqrt_solve_out = D^-1/2 L^-1 P rhs
 */
REF_API int ref_ldlt(const double *A, int64_t n, double *ldlt_out, int64_t *transpositions,
                     double *vector_d, double *logdet, int *is_pd, const double *rhs, int64_t nrhs,
                     double *solve_out, double *sqrt_solve_out, double *inverse_diagonal_out) {
  const Eigen::Map<const Eigen::MatrixXd> Amap(A, n, n);
  const Eigen::SerializableLDLT ldlt{Eigen::MatrixXd(Amap)};
  if (ldlt_out != nullptr) {
    copy_out(Eigen::MatrixXd(ldlt.matrixLDLT()), ldlt_out);
  }
  if (transpositions != nullptr) {
    for (int64_t i = 0; i < n; ++i) {
      transpositions[i] = ldlt.transpositionsP().indices()[i];
    }
  }
  if (vector_d != nullptr) {
    copy_out(Eigen::VectorXd(ldlt.vectorD()), vector_d);
  }
  if (logdet != nullptr) {
    *logdet = ldlt.log_determinant();
  }
  if (is_pd != nullptr) {
    *is_pd = ldlt.is_positive_definite() ? 1 : 0;
  }
  if (rhs != nullptr) {
    const Eigen::Map<const Eigen::MatrixXd> rhs_map(rhs, n, nrhs);
    const Eigen::MatrixXd rhs_mat(rhs_map);
    if (solve_out != nullptr) {
      copy_out(Eigen::MatrixXd(ldlt.solve(rhs_mat)), solve_out);
    }
    if (sqrt_solve_out != nullptr) {
      copy_out(ldlt.sqrt_solve(rhs_mat), sqrt_solve_out);
    }
  }
  if (inverse_diagonal_out != nullptr) {
    copy_out(ldlt.inverse_diagonal(), inverse_diagonal_out);
  }
  return 0;
}

/*
 * SerializableLDLT::inverse_blocks (serializable_ldlt.hpp:137-175).  Groups are given CSR-style:
 * indices[offsets[g] .. offsets[g+1]); blocks are written back to back (col-major each).
 */
REF_API int ref_ldlt_inverse_blocks(const double *A, int64_t n, const int64_t *indices,
                                    const int64_t *offsets, int64_t ngroups, int nthreads,
                                    double *out) {
  const Eigen::Map<const Eigen::MatrixXd> Amap(A, n, n);
  const Eigen::SerializableLDLT ldlt{Eigen::MatrixXd(Amap)};
  std::vector<std::vector<std::size_t>> blocks(static_cast<std::size_t>(ngroups));
  for (int64_t g = 0; g < ngroups; ++g) {
    for (int64_t k = offsets[g]; k < offsets[g + 1]; ++k) {
      blocks[static_cast<std::size_t>(g)].push_back(static_cast<std::size_t>(indices[k]));
    }
  }
  std::unique_ptr<ThreadPool> pool;
  if (nthreads > 1) {
    pool = std::make_unique<ThreadPool>(static_cast<std::size_t>(nthreads));
  }
  const auto inv = ldlt.inverse_blocks(blocks, pool.get());
  double *cursor = out;
  for (const auto &b : inv) {
    std::copy(b.data(), b.data() + b.size(), cursor);
    cursor += b.size();
  }
  return 0;
}

/*
 * group_by(features, grouper).indexers() (group_by.hpp:349-403) — the bit-exact integer contract.
 * Output CSR: keys[g] ascending (std::map order), indices of group g at
 * indices[offsets[g] .. offsets[g+1]).  Returns the number of groups.
 */
REF_API int64_t ref_group_indexers(const double *feats, int64_t n, int dim, int grouper_kind,
                                   double grouper_arg, int64_t *keys, int64_t *offsets,
                                   int64_t *indices) {
  int64_t ngroups = 0;
  with_feature_type(dim, [&](auto *tag) {
    using X = std::remove_pointer_t<decltype(tag)>;
    const auto xs = FeatureIO<X>::load(feats, n, dim);
    auto emit = [&](const auto &indexer) {
      int64_t g = 0;
      int64_t cursor = 0;
      offsets[0] = 0;
      for (const auto &pair : indexer) {
        keys[g] = static_cast<int64_t>(pair.first);
        for (const auto &i : pair.second) {
          indices[cursor++] = static_cast<int64_t>(i);
        }
        offsets[++g] = cursor;
      }
      ngroups = g;
    };
    if (grouper_kind == 0) {
      emit(albatross::group_by(xs, albatross::LeaveOneOutGrouper()).indexers());
    } else {
      emit(albatross::group_by(xs, ShimGrouper<X>{grouper_kind, grouper_arg}).indexers());
    }
    return 0;
  });
  return ngroups;
}

/* partition_triangular (indexing/block.hpp:25-44): writes 2*count [start, end) pairs. */
REF_API int64_t ref_partition_triangular(int64_t n, int64_t num_blocks, int64_t *out) {
  const auto blocks = albatross::detail::partition_triangular(n, num_blocks);
  int64_t k = 0;
  for (const auto &b : blocks) {
    out[2 * k] = b.first;
    out[2 * k + 1] = b.second;
    ++k;
  }
  return k;
}

/* indices_complement (indexing/subset.hpp) — fold train indices. */
REF_API int64_t ref_indices_complement(const int64_t *indices, int64_t count, int64_t n,
                                       int64_t *out) {
  std::vector<std::size_t> idx(indices, indices + count);
  const auto comp = albatross::indices_complement(idx, static_cast<std::size_t>(n));
  for (std::size_t i = 0; i < comp.size(); ++i) {
    out[i] = static_cast<int64_t>(comp[i]);
  }
  return static_cast<int64_t>(comp.size());
}

/* negative_log_likelihood(deviation, covariance) (evaluation/likelihood.hpp:53-67). */
REF_API double ref_nll_dense(const double *deviation, const double *cov, int64_t n) {
  const Eigen::Map<const Eigen::VectorXd> d(deviation, n);
  const Eigen::Map<const Eigen::MatrixXd> c(cov, n, n);
  return albatross::negative_log_likelihood(Eigen::VectorXd(d), Eigen::MatrixXd(c));
}
