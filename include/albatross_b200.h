/*
 * albatross_b200 — C ABI of the B200-native exact-GP hot path.
 *
 * This is the drop-in boundary.  The reference (swift-nav/albatross) has no FFI:
 * its "operator API" is a set of C++ template concepts.  The C++ trait layer shipped in
 * include/albatross_b200/ (umbrella header albatross.hpp) re-creates those concepts (same names and signatures) and calls
 * ONLY the functions below; INTEGRATION.md shows the binding a reference maintainer would add.
 * Each entry point cites the reference routine it replaces; paths are relative to the reference
 * root, `src/` = include/albatross/src/.
 *
 * Conventions
 *   - plain C types only; all sizes int64_t; matrices column-major fp64 (Eigen::MatrixXd layout);
 *   - features are AoS doubles, point i at feats[i*dim .. i*dim+dim) (std::vector<double> for dim 1,
 *     std::vector<Eigen::Matrix<double,dim,1>> otherwise) — equivalently a dim x n col-major matrix;
 *   - `const double *` / `double *` arguments are HOST pointers; device-resident data only ever
 *     appears behind the opaque ab_matrix / ab_factor handles;
 *   - every function returns an ab_status (0 = ok); ab_last_error() describes the last failure on the
 *     calling thread.  The reference itself has no error channel (ALBATROSS_ASSERT,
 *     src/details/error_handling.hpp:37-45); the C++ layer asserts on non-zero;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     AB_ERR_CUDA.
 *   - a handle is bound to one GPU and one stream (one process per GPU); calls on one handle are
 *     serialised by an internal mutex, so concurrent tuner threads may share it.
 */
#ifndef ALBATROSS_B200_H
#define ALBATROSS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB_VERSION 1

#if defined(__GNUC__)
#define AB_API __attribute__((visibility("default")))
#else
#define AB_API
#endif

typedef enum {
  AB_OK = 0,
  AB_ERR_INVALID = 1,      /* bad argument / malformed covariance program */
  AB_ERR_CUDA = 2,         /* CUDA runtime error or no device */
  AB_ERR_ALLOC = 3,        /* device or host allocation failed */
  AB_ERR_NOT_PD = 4,       /* non-positive pivot met in the factorisation (see ab_factor_info) */
  AB_ERR_NCCL = 5,         /* collective failed */
  AB_ERR_UNSUPPORTED = 6   /* feature/covariance type without a device form */
} ab_status;

/*
 * Covariance program: the compile-time Sum/Product tree of the reference
 * (src/covariance_functions/covariance_function.hpp:222-420) flattened to postfix by the C++ trait
 * layer at every call (hyper-parameters are re-read live, SURVEY.md §5 "Config").
 * Leaves push k(x,y); AB_OP_SUM / AB_OP_PRODUCT pop rhs then lhs and push the combination.
 */
typedef enum {
  AB_OP_SQUARED_EXPONENTIAL = 1, /* p0 = length_scale, p1 = sigma   src/covariance_functions/radial.hpp:25-33   */
  AB_OP_EXPONENTIAL = 2,         /* p0 = length_scale, p1 = sigma   radial.hpp:191-198 */
  AB_OP_MATERN32 = 3,            /* p0 = length_scale, p1 = sigma   radial.hpp:289-297 */
  AB_OP_MATERN52 = 4,            /* p0 = length_scale, p1 = sigma   radial.hpp:461-470 */
  AB_OP_CONSTANT = 5,            /* p0 = sigma                      src/covariance_functions/polynomials.hpp:56-60 */
  AB_OP_INDEPENDENT_NOISE = 6,   /* p0 = sigma; value equality      src/covariance_functions/noise.hpp:37-43 */
  AB_OP_SUM = 7,                 /* lhs + rhs                       covariance_function.hpp:270-272 */
  AB_OP_PRODUCT = 8,             /* lhs != 0 ? lhs * rhs : lhs      covariance_function.hpp:361-367 */
  AB_OP_POLYNOMIAL_TERM = 9      /* p0 = sigma, p1 = degree p: sigma^2 x^p y^p, one term of Polynomial<order>
                                    (polynomials.hpp:63-90; scalar features).  Only as a summand of the top-level
                                    sum (AB_ERR_UNSUPPORTED inside a product). */
} ab_opcode;

typedef struct {
  int32_t op; /* ab_opcode */
  int32_t reserved;
  double p0;
  double p1;
} ab_op;

#define AB_MAX_OPS 32
#define AB_MAX_DIM 8

typedef struct ab_handle_s *ab_handle;
typedef struct ab_matrix_s *ab_matrix; /* device-resident column-major fp64 matrix */
typedef struct ab_factor_s *ab_factor; /* device-resident factor K = L L^T (D = diag(L)^2, P = I) */

/* Per-phase device times of the most recent GP-level call, CUDA events, milliseconds. */
typedef struct {
  double h2d_ms;
  double gram_ms;
  double factor_ms;
  double solve_ms;
  double reduce_ms;
  double predict_ms;
  double d2h_ms;
  double total_ms;
  int64_t kernel_launches; /* kernels launched by this library since ab_create / last reset */
} ab_phase_times;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Binds a handle to CUDA device `device` with its own non-blocking stream. */
AB_API int ab_create(ab_handle *out, int device);
/* Same, but work is enqueued on the caller's stream (a cudaStream_t passed as void*). */
AB_API int ab_create_on_stream(ab_handle *out, int device, void *cuda_stream);
AB_API int ab_destroy(ab_handle h);
AB_API const char *ab_last_error(void);
AB_API int ab_version(void);
AB_API int ab_device_count(int *count);
AB_API int ab_synchronize(ab_handle h);
AB_API int ab_timings(ab_handle h, ab_phase_times *out);
AB_API int ab_reset_counters(ab_handle h);
/* Releases the handle's cached device workspace (buffers are otherwise recycled between calls). */
AB_API int ab_trim(ab_handle h);

/* ---- device matrices ----------------------------------------------------------------------- */

AB_API int ab_matrix_upload(ab_handle h, const double *host, int64_t rows, int64_t cols, ab_matrix *out);
AB_API int ab_matrix_alloc(ab_handle h, int64_t rows, int64_t cols, ab_matrix *out);
AB_API int ab_matrix_download(ab_handle h, ab_matrix m, double *host);
/* Copies the rows x cols block starting at (row0, col0) into `host` (column-major, ld = rows). */
AB_API int ab_matrix_download_block(ab_handle h, ab_matrix m, int64_t row0, int64_t col0, int64_t rows,
                             int64_t cols, double *host);
AB_API int ab_matrix_dims(ab_matrix m, int64_t *rows, int64_t *cols);
AB_API int ab_matrix_free(ab_handle h, ab_matrix m);
/* Raw device pointer + leading dimension, for callers that own CUDA code themselves. */
AB_API int ab_matrix_device_ptr(ab_matrix m, void **ptr, int64_t *ld);
/* K(i,i) += d[i]: `cov += targets.covariance` of src/models/gp.hpp:65. */
AB_API int ab_matrix_add_diag(ab_handle h, ab_matrix m, const double *d);

/* ---- Gram construction --------------------------------------------------------------------- */

#define AB_GRAM_FULL 0u        /* full symmetric n x n (what the reference returns) */
#define AB_GRAM_LOWER_ONLY 1u  /* only i >= j is written (enough for ab_potrf) */

/*
 * Symmetric Gram K_ij = k(x_i, x_j).
 * Replaces CovarianceFunction::operator()(const std::vector<X>&, ThreadPool*)
 * src/covariance_functions/covariance_function.hpp:128-137 -> compute_covariance_matrix
 * src/covariance_functions/callers.hpp:107-166.
 */
AB_API int ab_gram_sym(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
                uint32_t flags, ab_matrix *out);
/*
 * Cross Gram C_ij = k(x_i, y_j), n x m.
 * Replaces operator()(const std::vector<X>&, const std::vector<Y>&, ThreadPool*)
 * covariance_function.hpp:142-151 -> callers.hpp:38-102.
 */
AB_API int ab_gram_cross(ab_handle h, const ab_op *prog, int nops, const double *fx, int64_t n,
                  const double *fy, int64_t m, int dim, ab_matrix *out);
/*
 * k(x_i, x_i) -> out[n] (host).  Replaces CovarianceFunction::diagonal
 * covariance_function.hpp:156-168 and the prior-variance loop src/models/gp.hpp:339-343.
 */
AB_API int ab_gram_diag(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
                 double *out);
/* Device-resident variants: feats is a dim x n ab_matrix; used when inputs already live in HBM. */
AB_API int ab_gram_sym_d(ab_handle h, const ab_op *prog, int nops, ab_matrix feats, uint32_t flags,
                  ab_matrix *out);
AB_API int ab_gram_cross_d(ab_handle h, const ab_op *prog, int nops, ab_matrix fx, ab_matrix fy,
                    ab_matrix *out);

/* ---- factorisation (the CovarianceRepresentation concept, src/models/gp.hpp:42-45) ---------- */

/*
 * In-place blocked Cholesky of the lower triangle of `m`; `m` is consumed (owned by the factor).
 * Replaces Eigen::SerializableLDLT(const MatrixXd&) src/eigen/serializable_ldlt.hpp:27 ->
 * LDLT::compute third_party/eigen/Eigen/src/Cholesky/LDLT.h:488-521.  The device factor is
 * unpivoted (P = I, D = diag(L)^2); results agree with the pivoted reference to rounding.
 * Returns AB_ERR_NOT_PD if a pivot <= 0 (or NaN) is met; the factor is still returned so that
 * ab_factor_info can report the offending index, but must not be used for solves.
 */
AB_API int ab_potrf(ab_handle h, ab_matrix m, ab_factor *out);
AB_API int ab_factor_free(ab_handle h, ab_factor f);
AB_API int ab_factor_rows(ab_factor f, int64_t *n);
/* first_bad_pivot = -1 when the matrix was positive definite (is_positive_definite(), :36). */
AB_API int ab_factor_info(ab_factor f, int64_t *first_bad_pivot);
/* out = K^-1 rhs, n x nrhs.  LDLT::solve, LDLT.h:558-592. */
AB_API int ab_factor_solve(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out);
/* out = D^-1/2 L^-1 P rhs = L_chol^-1 rhs.  SerializableLDLT::sqrt_solve, serializable_ldlt.hpp:100-109. */
AB_API int ab_factor_sqrt_solve(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out);
/* sum log D_ii.  serializable_ldlt.hpp:128-135. */
AB_API int ab_factor_logdet(ab_handle h, ab_factor f, double *out);
/* 0.5 (log|K| + dev^T K^-1 dev + n log 2pi).  src/evaluation/likelihood.hpp:38-47. */
AB_API int ab_factor_nll(ab_handle h, ab_factor f, const double *deviation, double *out);
/* diag(K^-1).  serializable_ldlt.hpp:181-199. */
AB_API int ab_factor_inverse_diagonal(ab_handle h, ab_factor f, double *out);
/*
 * (K^-1)_gg for each group g = indices[offsets[g] .. offsets[g+1]); blocks written back to back,
 * column-major each.  serializable_ldlt.hpp:137-175.
 */
AB_API int ab_factor_inverse_blocks(ab_handle h, ab_factor f, const int64_t *indices,
                             const int64_t *offsets, int64_t ngroups, double *out);
/*
 * Materialises the factor in Eigen::SerializableLDLT's packed layout: strict lower = unit L,
 * diagonal = D, transpositions = identity (src/cereal/serializable_ldlt.hpp:18-32).  Streamed in column
 * panels through a bounded device buffer: works at any n the host can hold (n = 65 536 is 32 GiB).
 */
AB_API int ab_factor_export_packed(ab_handle h, ab_factor f, double *LD, int64_t *transpositions);

/*
 * The reverse: a device factor from Eigen::SerializableLDLT's packed form — what the reference's cereal
 * archives hold (src/cereal/serializable_ldlt.hpp:18-32: the lower triangle with unit L below the diagonal and
 * D on it, and the transpositions; src/cereal/gp.hpp:28-52 stores it inside Fit<GPFit>) — so that a fit
 * serialised by the reference can be loaded onto the GPU.  LD: n x n column-major (only i >= j is read).
 * Identity transpositions (every archive written from a device fit via ab_factor_export_packed): O(n^2).
 * Otherwise K = P^T L D L^T P is rebuilt on the device (one DSYRK) and factored without pivoting.
 * AB_ERR_NOT_PD when some D_ii <= 0.
 */
AB_API int ab_factor_import_packed(ab_handle h, const double *LD, const int64_t *transpositions, int64_t n,
                            ab_factor *out);
/* out = P^T L D^1/2 rhs = L_chol rhs.  SerializableLDLT::sqrt_product, serializable_ldlt.hpp:91-94. */
AB_API int ab_factor_sqrt_product(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out);
/* out = P^T L^-T D^-1/2 rhs = L_chol^-T rhs.  sqrt_transpose_solve, serializable_ldlt.hpp:123-126. */
AB_API int ab_factor_sqrt_transpose_solve(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs,
                                   double *out);
/* out (n x n, column-major) = D^1/2 (P^T L)^T = L_chol^T, upper triangular.  sqrt_transpose, :111-115. */
AB_API int ab_factor_sqrt_transpose(ab_handle h, ab_factor f, double *out);
/* out[n] = sqrt(D_ii) = diag(L_chol).  diagonal_sqrt :74-84 (and, inverted, diagonal_sqrt_inverse :58-69;
 * a usable factor has every D_ii > 0, so the reference's clamping of non-positive pivots never fires). */
AB_API int ab_factor_diagonal_sqrt(ab_handle h, ab_factor f, double *out);

/* ---- exact GP (src/models/gp.hpp) ---------------------------------------------------------- */

/*
 * model.fit(dataset): K = k(X,X) (+ diag(yvar) when yvar != NULL), factor, information = K^-1 y.
 * Replaces GaussianProcessBase::_fit_impl gp.hpp:285-294 + Fit<GPFit<...>> ctor gp.hpp:61-69.
 * `y` must already have the mean function removed by the caller (ZeroMean: unchanged).
 * information may be NULL.  A NaN in K yields AB_ERR_NOT_PD (ALBATROSS_ASSERT(!cov.hasNaN())).
 */
AB_API int ab_gp_fit(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
              const double *y, const double *yvar, ab_factor *factor, double *information);
/*
 * Data term of -model.log_likelihood(dataset): builds and factors K afresh, exactly like
 * gp.hpp:443-451 -> likelihood.hpp:53-67 (no targets.covariance).  The prior term is a host scalar
 * added by the caller.
 */
AB_API int ab_gp_nll(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
              const double *y, double *nll);
/*
 * fit + nll sharing ONE Gram build and ONE factorisation (valid when yvar == NULL, where the two
 * reference matrices coincide).  Offered for tuner loops; not what model.fit + log_likelihood do.
 */
AB_API int ab_gp_fit_nll(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n,
                  int dim, const double *y, ab_factor *factor, double *information, double *nll);

#define AB_PREDICT_MEAN 0
#define AB_PREDICT_MARGINAL 1
#define AB_PREDICT_JOINT 2
/*
 * fit_model.predict(test).{mean,marginal,joint}().  Replaces _predict_impl x3 gp.hpp:313-366 and
 * gp_{mean,marginal,joint}_prediction gp.hpp:82-113.  mean: p; var: p (marginal); cov: p*p (joint).
 */
AB_API int ab_gp_predict(ab_handle h, ab_factor factor, const ab_op *prog, int nops,
                  const double *train_feats, int64_t n, int dim, const double *information,
                  const double *test_feats, int64_t p, int what, double *mean, double *var,
                  double *cov);
/*
 * ab_gp_predict with separate programs for the cross covariance k(train, test) and the prior
 * k(test, test).  The two differ when exactly one side of the cross call is a Measurement<> and the
 * tree holds a MeasurementOnly term (src/covariance_functions/measurement.hpp:70-114), e.g.
 * fit_model.predict_with_measurement_noise(x) (src/core/fit_model.hpp:54-62).
 */
AB_API int ab_gp_predict2(ab_handle h, ab_factor factor, const ab_op *cross_prog, int cross_nops,
                   const ab_op *prior_prog, int prior_nops, const double *train_feats, int64_t n,
                   int dim, const double *information, const double *test_feats, int64_t p, int what,
                   double *mean, double *var, double *cov);
/*
 * Leave-one-group-out predictions from an existing fit.  Replaces
 * details::held_out_predictions src/evaluation/cross_validation_utils.hpp:199-232 ->
 * held_out_prediction :172-197, scattered back by index (concatenate_*_predictions :59-100).
 * groups: CSR (indices, offsets, ngroups) exactly as GroupIndexer iterates (std::map key order).
 * what = MEAN: mean[n]; MARGINAL: mean[n], var[n]; JOINT: mean[n], joint = A_g^-1 blocks back to
 * back.  score (optional) = sum_g NLL(joint_g, truth_g) (model_metrics.hpp:59-90, data term).
 */
AB_API int ab_gp_cv(ab_handle h, ab_factor factor, const double *y, const double *information,
             const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
             double *mean, double *var, double *joint, double *score);

/*
 * ab_gp_cv that also returns one score per group (group_scores[ngroups], key order; optional): what
 * CrossValidation::scores(NegativeLogLikelihood<JointDistribution>, dataset, indexer) returns
 * (src/evaluation/cross_validation.hpp:297-325 -> cross_validated_scores
 * cross_validation_utils.hpp:102-130) for targets without measurement covariance.
 */
AB_API int ab_gp_cv_scores(ab_handle h, ab_factor factor, const double *y, const double *information,
                    const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
                    double *mean, double *var, double *joint, double *score, double *group_scores);

/*
 * One shard of ab_gp_cv: only the groups g with g % nshards == shard (for pure leave-one-out: the
 * 2048-column chunks c of the inverse diagonal with c % nshards == shard) are processed, through
 * per-shard triangular solves instead of one explicit inverse factor; entries of other shards are
 * returned as zero, so that the element-wise SUM over shards is the full ab_gp_cv result.  This is
 * the unit of work ab_dist_gp_cv gives each rank; a single process driving several handles can use
 * it directly.
 */
AB_API int ab_gp_cv_shard(ab_handle h, ab_factor factor, const double *y, const double *information,
                   const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
                   int shard, int nshards, double *mean, double *var, double *score);

/*
 * fit_model.update(dataset): the fit of [train; new] from the fit of train without refactoring.  Replaces
 * GaussianProcessBase::_update_impl src/models/gp.hpp:386-414 and BlockSymmetric
 * src/linalg/block_symmetric.hpp:46-133 (Schur complement S = C - B^T A^-1 B): with a Cholesky factor the
 * same algebra yields the factor of the enlarged matrix itself, [L 0; (L^-1 B)^T chol(S)], so the result is
 * an ordinary ab_factor of size n + p (usable with every ab_factor_* / ab_gp_* call); O(n^2 p + p^3) work.
 *   old_factor   factor of k(train, train) + diag(yvar_train), left untouched
 *   prog         k(Measurement, Measurement) as in ab_gp_fit
 *   information_old  K^-1 y of the old fit (a fit does not keep its targets, gp.hpp:49-51: they are
 *                recovered on the device as y = L L^T information_old)
 *   y_new        p new targets, mean function already removed; yvar_new may be NULL
 * information (n + p doubles, optional) = K'^-1 [y; y_new].
 */
AB_API int ab_gp_update(ab_handle h, ab_factor old_factor, const ab_op *prog, int nops,
                 const double *train_feats, int64_t n, int dim, const double *information_old,
                 const double *new_feats, int64_t p, const double *y_new, const double *yvar_new,
                 ab_factor *factor, double *information);

/* Device-resident variants (inputs already in HBM; results stay in HBM unless a host ptr is given). */
AB_API int ab_gp_fit_d(ab_handle h, const ab_op *prog, int nops, ab_matrix feats, ab_matrix y,
                ab_matrix yvar, ab_factor *factor, ab_matrix *information);
AB_API int ab_gp_nll_d(ab_handle h, const ab_op *prog, int nops, ab_matrix feats, ab_matrix y,
                double *nll);

/* ---- sparse GP (src/models/sparse_gp.hpp) --------------------------------------------------- */

typedef struct ab_sparse_fit_s *ab_sparse; /* device-resident Fit<SparseGPFit>: u, K_uu factor, R, v */

/*
 * SparseGaussianProcessRegression::_fit_impl sparse_gp.hpp:381-404 (compute_internal_components
 * :632-706 + compute_sigma_qr :368-375) for the FITC / PITC approximation.
 *   feats/y/yvar     n observations (yvar may be NULL = zero measurement variance);
 *   inducing         m inducing features (the caller runs its InducingPointStrategy, e.g.
 *                    UniformlySpacedInducingPoints sparse_gp.hpp:34-47, on the host);
 *   indices/offsets  the GroupIndexer of group_by(features, grouper).indexers() as CSR in std::map
 *                    key order (ab_group_indexers): observations are reordered by it (:649-668).
 *                    LeaveOneOutGrouper (all groups singletons) = FITC;
 *   nuggets          measurement / inducing nugget (defaults 1e-8, sparse_gp.hpp:20-26, :281-287).
 * information (m doubles, optional) = v of :396-398; log_likelihood (optional) = the data term of
 * log_likelihood() :539-603 for the same inputs (the prior term is a host scalar of the caller).
 * QR of B is taken as CholQR2 of the better-conditioned C = B L_u^-T (see sparse.cu); R and the
 * column permutation are representation-internal; ab_sparse_export_R returns an R with R^T R = B^T B.
 * When the handle is part of a distributed group (ab_dist_init) each rank passes ITS shard of the
 * observations/groups; the m x m Gram products and the scalars are all-reduced, the resulting fit
 * is replicated on every rank.
 */
AB_API int ab_sparse_fit(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n,
                  int dim, const double *y, const double *yvar, const double *inducing, int64_t m,
                  const int64_t *indices, const int64_t *offsets, int64_t ngroups,
                  double measurement_nugget, double inducing_nugget, ab_sparse *out,
                  double *information, double *log_likelihood);
/*
 * ab_sparse_fit with the three covariance programs the reference evaluates (sparse_gp.hpp:646-679):
 *   prog_ff  k(Measurement<X>, Measurement<X>)  diagonal blocks of K_ff (:657)
 *   prog_fu  k(Measurement<X>, U)               K_fu (:670-671)
 *   prog_uu  k(U, U)                            K_uu (:673-674)
 * They differ when the tree holds a MeasurementOnly term (src/covariance_functions/measurement.hpp:
 * 70-114), the configuration of the reference's own sparse tests
 * (tests/lib/albatross/test/test_models.h:26-30).  ab_sparse_fit passes one program three times.
 */
AB_API int ab_sparse_fit2(ab_handle h, const ab_op *prog_ff, int nops_ff, const ab_op *prog_fu,
                   int nops_fu, const ab_op *prog_uu, int nops_uu, const double *feats, int64_t n,
                   int dim, const double *y, const double *yvar, const double *inducing, int64_t m,
                   const int64_t *indices, const int64_t *offsets, int64_t ngroups,
                   double measurement_nugget, double inducing_nugget, ab_sparse *out,
                   double *information, double *log_likelihood);
AB_API int ab_sparse_free(ab_handle h, ab_sparse f);
AB_API int ab_sparse_info(ab_sparse f, int64_t *m, double *log_likelihood);
/* model.log_likelihood(dataset) for the sparse model (:539-603): a fresh fit, only the scalar kept. */
AB_API int ab_sparse_log_likelihood(ab_handle h, const ab_op *prog, int nops, const double *feats,
                             int64_t n, int dim, const double *y, const double *yvar,
                             const double *inducing, int64_t m, const int64_t *indices,
                             const int64_t *offsets, int64_t ngroups, double measurement_nugget,
                             double inducing_nugget, double *log_likelihood);
AB_API int ab_sparse_log_likelihood2(ab_handle h, const ab_op *prog_ff, int nops_ff, const ab_op *prog_fu,
                              int nops_fu, const ab_op *prog_uu, int nops_uu, const double *feats,
                              int64_t n, int dim, const double *y, const double *yvar,
                              const double *inducing, int64_t m, const int64_t *indices,
                              const int64_t *offsets, int64_t ngroups, double measurement_nugget,
                              double inducing_nugget, double *log_likelihood);
/* _predict_impl x3 sparse_gp.hpp:468-536; outputs as ab_gp_predict. */
AB_API int ab_sparse_predict(ab_handle h, ab_sparse f, const ab_op *prog, int nops,
                      const double *test_feats, int64_t p, int what, double *mean, double *var,
                      double *cov);
/* ab_sparse_predict with separate programs for the cross covariance k(U, X*) (:470, :483, :509) and
 * the prior k(X*, X*) (:491-495, :516). */
AB_API int ab_sparse_predict2(ab_handle h, ab_sparse f, const ab_op *cross_prog, int cross_nops,
                       const ab_op *prior_prog, int prior_nops, const double *test_feats, int64_t p,
                       int what, double *mean, double *var, double *cov);
/*
 * R (cols x cols, upper triangular, column-major) of a thin QR of the dense host matrix B (rows x cols,
 * rows >= cols): what QRImplementation::compute + get_R (sparse_gp.hpp:72-89, linalg/qr_utils.hpp:18-27)
 * return, with P = I.  Computed as CholQR2 on the device (two passes of B^T B = L L^T, B <- B L^-T): valid
 * for cond(B) below ~1e7, AB_ERR_NOT_PD beyond (the sparse model itself factors a better-conditioned
 * matrix, see ab_sparse_fit).  R^T R = B^T B; the signs of R's diagonal are positive.
 */
AB_API int ab_qr_r(ab_handle h, const double *B, int64_t rows, int64_t cols, double *R);
/* sigma_R (m x m upper triangular, column-major, get_R linalg/qr_utils.hpp:18-27) with P = I. */
AB_API int ab_sparse_export_R(ab_handle h, ab_sparse f, double *R);

/* ---- one process per GPU: distributed group (NCCL over NVLink / NVSwitch) -------------------- */

#define AB_DIST_ID_BYTES 128
/*
 * Rank 0 creates an id (ncclGetUniqueId) and hands the bytes to the other ranks by whatever the
 * launcher offers (torch.distributed broadcast, MPI, a file); every rank then calls ab_dist_init.
 * libnccl.so.2 is loaded on first use (dlopen); single-GPU users never need it.
 */
AB_API int ab_dist_unique_id(void *id_out);
AB_API int ab_dist_init(ab_handle h, int rank, int world, const void *id);
AB_API int ab_dist_finalize(ab_handle h);
AB_API int ab_dist_info(ab_handle h, int *rank, int *world);

typedef struct ab_dist_factor_s *ab_dist_factor; /* block-column-cyclic factor, one shard per rank */

/*
 * Distributed model.fit(dataset) for matrices that outgrow one GPU (BASELINE configs[2],
 * N = 131 072): the Gram matrix is generated directly in block-column-cyclic layout (block size
 * `nb`, block column j on rank j % world; every rank holds all features), factorised by a
 * right-looking blocked Cholesky with one NCCL panel broadcast per block column (look-ahead 1,
 * overlapped with the trailing DSYRK/DGEMM update), and the information vector K^-1 y is computed by
 * pipelined block substitutions.  All ranks call with identical arguments; information (n doubles,
 * optional) and nll (optional; 0.5 (log|K| + y^T K^-1 y + n log 2 pi), likelihood.hpp:38-47) are
 * returned on every rank.  yvar as in ab_gp_fit.  nb = 0 picks the default.
 */
AB_API int ab_dist_gp_fit(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n,
                   int dim, const double *y, const double *yvar, int64_t nb, ab_dist_factor *factor,
                   double *information, double *nll);
AB_API int ab_dist_factor_free(ab_handle h, ab_dist_factor f);
/*
 * Where the most recent ab_dist_gp_fit on this rank spent its factorisation phase, from CUDA events on the
 * update stream: wait_ms = time the trailing updates stood waiting for a panel broadcast (summed over the
 * `steps` block columns), panel_ms = time of the panel chains this rank ran (diagonal factorisation + TRSM +
 * packing, overlapped with the updates).  The rest of ab_timings().factor_ms is DSYRK / DGEMM work.
 */
AB_API int ab_dist_fit_breakdown(ab_handle h, double *wait_ms, double *panel_ms, int64_t *steps);
/* Owner rank of block column j and its local index: the integer contract of the layout. */
AB_API int ab_dist_block_owner(int64_t block, int world, int *rank, int64_t *local_block);
/*
 * Gram row-block sharding: rank r builds rows [row0, row0 + rows) x all columns of
 * the symmetric Gram; the split balances rows.  Returns the local shard as an ab_matrix (rows x n).
 */
AB_API int ab_dist_gram_rows(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n,
                      int dim, int64_t *row0, int64_t *rows, ab_matrix *out);
/*
 * Replicates a single-GPU factor (ab_gp_fit / ab_potrf on rank `root`) onto every rank with one
 * ncclBroadcast of L over NVLink (at N = 32 768 the factor is 8.6 GB, a few tens of
 * milliseconds, against eight redundant 0.4 s factorisations).  The root passes its factor, which stays
 * valid; every other rank passes a null handle and receives a new factor that it frees itself.
 * ab_timings().h2d_ms reports the transfer.
 */
AB_API int ab_dist_factor_broadcast(ab_handle h, ab_factor *factor, int root);
/*
 * Leave-one-group-out CV with folds sharded over ranks: every rank holds the same
 * single-GPU factor (N <= 65 536 fits one GPU; fit once and ab_dist_factor_broadcast it); rank r processes
 * groups g with g % world == r and
 * the per-observation results are all-reduced.  Outputs as ab_gp_cv (joint blocks are returned only
 * for the caller's own groups, others zero-filled).
 */
AB_API int ab_dist_gp_cv(ab_handle h, ab_factor factor, const double *y, const double *information,
                  const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
                  double *mean, double *var, double *score);

/* ---- dense building block -------------------------------------------------------------------- */

#define AB_GEMM_TRANS_A 1u /* op(A) = A^T */
#define AB_GEMM_TRANS_B 2u /* op(B) = B^T */
#define AB_GEMM_LOWER 4u   /* C square: only the tiles touching the lower triangle are computed */
/*
 * C = alpha * op(A) * op(B) + beta * C on device-resident matrices (the DMMA kernel that performs
 * every O(n^3) step of the factorisation and the solves).  Exposed for callers that compose their
 * own block algebra (BlockSymmetric updates, src/linalg/block_symmetric.hpp:46-133) and for
 * benchmarking the kernel in isolation.
 */
AB_API int ab_gemm(ab_handle h, uint32_t flags, double alpha, ab_matrix A, ab_matrix B, double beta,
            ab_matrix C);

/* ---- integer contract helpers (host; bit-exact with src/indexing/) -------------------------- */

/*
 * group_by(...).indexers() given one integer key per item: keys ascending, member indices in
 * encounter order (IndexerBuilder::build src/indexing/group_by.hpp:349-376).  Returns ngroups
 * through *ngroups.  keys/offsets/indices must hold n, n+1, n entries.
 */
AB_API int ab_group_indexers(const int64_t *item_keys, int64_t n, int64_t *keys, int64_t *offsets,
                      int64_t *indices, int64_t *ngroups);

/*
 * detail::partition_triangular src/indexing/block.hpp:25-44: `count` row (lower-triangular) or column
 * (upper-triangular) blocks [start, end) of an n x n triangle with approximately equal areas — what the
 * reference's threaded Gram build hands to its pool (callers.hpp:134-166) and what a symmetric Gram build
 * sharded over `count` GPUs uses as row ranges.  bounds receives 2 * count entries: start_0, end_0, ...
 */
AB_API int ab_partition_triangular(int64_t n, int64_t count, int64_t *bounds);

#ifdef __cplusplus
}
#endif
#endif /* ALBATROSS_B200_H */
