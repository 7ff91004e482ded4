/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's exact-GP hot path.
 * See restate.h for the parity-pinning statement.  Citations are relative to /root/reference/.
 *
 * This file is a checker.  It is deliberately simple (scalar loops, the reference's own loop
 * structure) and is never linked into, called from, or used as a fallback by the product.
 */
#include "restate.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j, ld) ((size_t)(i) + (size_t)(j) * (size_t)(ld))

/* ------------------------------------------------------------------------------------------- */
/* Covariance leaves: include/albatross/src/covariance_functions/radial.hpp                      */
/* ------------------------------------------------------------------------------------------- */

/* EuclideanDistance, distance_metrics.hpp:36-44: fabs(x-y) for doubles, (x-y).norm() for vectors. */
static double euclidean_distance(const double *x, const double *y, int dim) {
  if (dim == 1) {
    return fabs(x[0] - y[0]);
  }
  double ss = 0.;
  for (int d = 0; d < dim; ++d) {
    const double diff = x[d] - y[d];
    ss += diff * diff;
  }
  return sqrt(ss);
}

/* radial.hpp:25-33 */
static double squared_exponential(double distance, double length_scale, double sigma) {
  if (length_scale <= 0.) {
    return 0.;
  }
  return sigma * sigma * exp(-pow(distance / length_scale, 2));
}

/* radial.hpp:191-198 */
static double exponential(double distance, double length_scale, double sigma) {
  if (length_scale <= 0.) {
    return 0.;
  }
  return sigma * sigma * exp(-fabs(distance / length_scale));
}

/* radial.hpp:289-297 */
static double matern_32(double distance, double length_scale, double sigma) {
  if (length_scale <= 0.) {
    return 0.;
  }
  const double sqrt_3_d = sqrt(3.) * distance / length_scale;
  return sigma * sigma * (1 + sqrt_3_d) * exp(-sqrt_3_d);
}

/* radial.hpp:461-470 */
static double matern_52(double distance, double length_scale, double sigma) {
  if (length_scale <= 0.) {
    return 0.;
  }
  const double sqrt_5_d = sqrt(5.) * distance / length_scale;
  return sigma * sigma * (1 + sqrt_5_d + sqrt_5_d * sqrt_5_d / 3.) * exp(-sqrt_5_d);
}

/* noise.hpp:37-43 — VALUE equality of the features (all coordinates). */
static int features_equal(const double *x, const double *y, int dim) {
  for (int d = 0; d < dim; ++d) {
    if (!(x[d] == y[d])) {
      return 0;
    }
  }
  return 1;
}

/*
 * Postfix evaluation of a composed covariance.  Sum: covariance_function.hpp:270-272.
 * Product: covariance_function.hpp:361-367 (`if (output != 0.) output *= rhs(x, y)`).
 * Measurement<X> wrappers are transparent to every in-scope leaf (callers.hpp:258-319), so the
 * same program serves K(meas, meas), K(train, test) and K(test, test).
 */
double rs_cov_eval(const rs_op *prog, int nops, const double *x, const double *y, int dim) {
  double stack[16];
  int sp = 0;
  double distance = -1.;
  for (int k = 0; k < nops; ++k) {
    const rs_op *o = &prog[k];
    switch (o->op) {
    case RS_OP_SQUARED_EXPONENTIAL:
    case RS_OP_EXPONENTIAL:
    case RS_OP_MATERN32:
    case RS_OP_MATERN52:
      if (distance < 0.) {
        distance = euclidean_distance(x, y, dim);
      }
      if (o->op == RS_OP_SQUARED_EXPONENTIAL) {
        stack[sp++] = squared_exponential(distance, o->p0, o->p1);
      } else if (o->op == RS_OP_EXPONENTIAL) {
        stack[sp++] = exponential(distance, o->p0, o->p1);
      } else if (o->op == RS_OP_MATERN32) {
        stack[sp++] = matern_32(distance, o->p0, o->p1);
      } else {
        stack[sp++] = matern_52(distance, o->p0, o->p1);
      }
      break;
    case RS_OP_CONSTANT: /* polynomials.hpp:56-60 */
      stack[sp++] = o->p0 * o->p0;
      break;
    case RS_OP_INDEPENDENT_NOISE:
      stack[sp++] = features_equal(x, y, dim) ? o->p0 * o->p0 : 0.;
      break;
    case RS_OP_POLYNOMIAL_TERM: /* polynomials.hpp:79-87: sigma^2 pow(x, p) pow(y, p), scalar features */
      stack[sp++] = o->p0 * o->p0 * pow(x[0], o->p1) * pow(y[0], o->p1);
      break;
    case RS_OP_SUM: {
      const double rhs = stack[--sp];
      const double lhs = stack[--sp];
      stack[sp++] = lhs + rhs;
      break;
    }
    case RS_OP_PRODUCT: {
      const double rhs = stack[--sp];
      double out = stack[--sp];
      if (out != 0.) {
        out *= rhs;
      }
      stack[sp++] = out;
      break;
    }
    default:
      return NAN;
    }
  }
  return sp == 1 ? stack[0] : NAN;
}

/* Serial symmetric Gram, callers.hpp:107-129: lower triangle evaluated, mirrored to the upper. */
void rs_gram_sym(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                 double *out) {
  for (int64_t i = 0; i < n; ++i) {
    for (int64_t j = 0; j <= i; ++j) {
      const double v = rs_cov_eval(prog, nops, feats + i * dim, feats + j * dim, dim);
      out[IDX(i, j, n)] = v;
      out[IDX(j, i, n)] = v;
    }
  }
}

/* Cross Gram, callers.hpp:38-60: C(i,j) = k(x_i, y_j), n x m. */
void rs_gram_cross(const rs_op *prog, int nops, const double *fx, int64_t n, const double *fy,
                   int64_t m, int dim, double *out) {
  for (int64_t j = 0; j < m; ++j) {
    for (int64_t i = 0; i < n; ++i) {
      out[IDX(i, j, n)] = rs_cov_eval(prog, nops, fx + i * dim, fy + j * dim, dim);
    }
  }
}

/* covariance_function.hpp:159-168 / gp.hpp:339-343 */
void rs_gram_diag(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                  double *out) {
  for (int64_t i = 0; i < n; ++i) {
    out[i] = rs_cov_eval(prog, nops, feats + i * dim, feats + i * dim, dim);
  }
}

/* ------------------------------------------------------------------------------------------- */
/* Diagonally pivoted LDL^T: third_party/eigen/Eigen/src/Cholesky/LDLT.h:294-394 (Lower)        */
/* ------------------------------------------------------------------------------------------- */

int rs_ldlt(double *A, int64_t n, int64_t *tr) {
  if (n <= 1) {
    if (n == 1) {
      tr[0] = 0;
    }
    return 0;
  }
  double *temp = (double *)malloc((size_t)n * sizeof(double));
  int found_zero_pivot = 0;
  int ret = 0;
  for (int64_t k = 0; k < n; ++k) {
    /* largest |diagonal| in the trailing corner (first maximum wins) */
    int64_t big = k;
    double best = fabs(A[IDX(k, k, n)]);
    for (int64_t i = k + 1; i < n; ++i) {
      const double v = fabs(A[IDX(i, i, n)]);
      if (v > best) {
        best = v;
        big = i;
      }
    }
    tr[k] = big;
    if (k != big) {
      /* symmetric swap touching only the lower triangle (LDLT.h:322-338) */
      for (int64_t j = 0; j < k; ++j) {
        const double t = A[IDX(k, j, n)];
        A[IDX(k, j, n)] = A[IDX(big, j, n)];
        A[IDX(big, j, n)] = t;
      }
      for (int64_t i = big + 1; i < n; ++i) {
        const double t = A[IDX(i, k, n)];
        A[IDX(i, k, n)] = A[IDX(i, big, n)];
        A[IDX(i, big, n)] = t;
      }
      {
        const double t = A[IDX(k, k, n)];
        A[IDX(k, k, n)] = A[IDX(big, big, n)];
        A[IDX(big, big, n)] = t;
      }
      for (int64_t i = k + 1; i < big; ++i) {
        const double t = A[IDX(i, k, n)];
        A[IDX(i, k, n)] = A[IDX(big, i, n)];
        A[IDX(big, i, n)] = t;
      }
    }
    const int64_t rs = n - k - 1;
    if (k > 0) {
      /* temp = D(0:k) .* A10^T ; A(k,k) -= A10 temp ; A21 -= A20 temp  (LDLT.h:349-355) */
      double dot = 0.;
      for (int64_t j = 0; j < k; ++j) {
        temp[j] = A[IDX(j, j, n)] * A[IDX(k, j, n)];
        dot += A[IDX(k, j, n)] * temp[j];
      }
      A[IDX(k, k, n)] -= dot;
      for (int64_t j = 0; j < k; ++j) {
        const double t = temp[j];
        const double *col = &A[IDX(k + 1, j, n)];
        double *dst = &A[IDX(k + 1, k, n)];
        for (int64_t i = 0; i < rs; ++i) {
          dst[i] -= col[i] * t;
        }
      }
    }
    const double akk = A[IDX(k, k, n)];
    const int pivot_is_valid = fabs(akk) > 0.;
    if (k == 0 && !pivot_is_valid) {
      for (int64_t j = 0; j < n; ++j) {
        tr[j] = j;
      }
      free(temp);
      return 0;
    }
    if (rs > 0 && pivot_is_valid) {
      for (int64_t i = 0; i < rs; ++i) {
        A[IDX(k + 1 + i, k, n)] /= akk;
      }
    }
    if (found_zero_pivot && pivot_is_valid) {
      ret = 1; /* NumericalIssue */
    } else if (!pivot_is_valid) {
      found_zero_pivot = 1;
    }
  }
  free(temp);
  return ret;
}

static void apply_transpositions(const int64_t *tr, int64_t n, double *B, int64_t k) {
  /* dst = P b : swaps applied in order (Transpositions * rhs). */
  for (int64_t c = 0; c < k; ++c) {
    double *b = B + (size_t)c * (size_t)n;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t j = tr[i];
      if (j != i) {
        const double t = b[i];
        b[i] = b[j];
        b[j] = t;
      }
    }
  }
}

static void apply_transpositions_T(const int64_t *tr, int64_t n, double *B, int64_t k) {
  for (int64_t c = 0; c < k; ++c) {
    double *b = B + (size_t)c * (size_t)n;
    for (int64_t i = n - 1; i >= 0; --i) {
      const int64_t j = tr[i];
      if (j != i) {
        const double t = b[i];
        b[i] = b[j];
        b[j] = t;
      }
    }
  }
}

static void unit_lower_solve(const double *LD, int64_t n, double *B, int64_t k) {
  for (int64_t c = 0; c < k; ++c) {
    double *b = B + (size_t)c * (size_t)n;
    for (int64_t j = 0; j < n; ++j) {
      const double bj = b[j];
      if (bj != 0.) {
        const double *col = &LD[IDX(0, j, n)];
        for (int64_t i = j + 1; i < n; ++i) {
          b[i] -= col[i] * bj;
        }
      }
    }
  }
}

static void unit_lower_T_solve(const double *LD, int64_t n, double *B, int64_t k) {
  for (int64_t c = 0; c < k; ++c) {
    double *b = B + (size_t)c * (size_t)n;
    for (int64_t j = n - 1; j >= 0; --j) {
      const double *col = &LD[IDX(0, j, n)];
      double s = b[j];
      for (int64_t i = j + 1; i < n; ++i) {
        s -= col[i] * b[i];
      }
      b[j] = s;
    }
  }
}

/* LDLT::_solve_impl, LDLT.h:558-592 (pseudo-inverse of D with tolerance 1/DBL_MAX). */
void rs_ldlt_solve(const double *LD, const int64_t *tr, int64_t n, double *B, int64_t k) {
  apply_transpositions(tr, n, B, k);
  unit_lower_solve(LD, n, B, k);
  const double tolerance = 1. / DBL_MAX;
  for (int64_t i = 0; i < n; ++i) {
    const double d = LD[IDX(i, i, n)];
    for (int64_t c = 0; c < k; ++c) {
      if (fabs(d) > tolerance) {
        B[IDX(i, c, n)] /= d;
      } else {
        B[IDX(i, c, n)] = 0.;
      }
    }
  }
  unit_lower_T_solve(LD, n, B, k);
  apply_transpositions_T(tr, n, B, k);
}

/* SerializableLDLT::sqrt_solve, serializable_ldlt.hpp:100-109: D^-1/2 L^-1 P rhs, D clamped at 0. */
void rs_ldlt_sqrt_solve(const double *LD, const int64_t *tr, int64_t n, double *B, int64_t k) {
  apply_transpositions(tr, n, B, k);
  unit_lower_solve(LD, n, B, k);
  for (int64_t i = 0; i < n; ++i) {
    const double d = LD[IDX(i, i, n)];
    const double s = d > 0. ? 1. / sqrt(d) : 0.;
    for (int64_t c = 0; c < k; ++c) {
      B[IDX(i, c, n)] *= s;
    }
  }
}

/* serializable_ldlt.hpp:128-135 / likelihood.hpp:26-32 */
double rs_ldlt_logdet(const double *LD, int64_t n) {
  double sum = 0.;
  for (int64_t i = 0; i < n; ++i) {
    sum += log(LD[IDX(i, i, n)]);
  }
  return sum;
}

/* serializable_ldlt.hpp:137-175: R^-1 = D^-1/2 L^-1 P materialised, then Q_g^T Q_g per group. */
void rs_ldlt_inverse_blocks(const double *LD, const int64_t *tr, int64_t n, const int64_t *indices,
                            const int64_t *offsets, int64_t ngroups, double *out) {
  double *inv = (double *)calloc((size_t)n * (size_t)n, sizeof(double));
  for (int64_t i = 0; i < n; ++i) {
    inv[IDX(i, i, n)] = 1.;
  }
  rs_ldlt_sqrt_solve(LD, tr, n, inv, n);
  double *cursor = out;
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t *idx = indices + offsets[g];
    const int64_t sz = offsets[g + 1] - offsets[g];
    for (int64_t b = 0; b < sz; ++b) {
      for (int64_t a = 0; a < sz; ++a) {
        const double *ca = &inv[IDX(0, idx[a], n)];
        const double *cb = &inv[IDX(0, idx[b], n)];
        double s = 0.;
        for (int64_t i = 0; i < n; ++i) {
          s += ca[i] * cb[i];
        }
        cursor[IDX(a, b, sz)] = s;
      }
    }
    cursor += sz * sz;
  }
  free(inv);
}

/* serializable_ldlt.hpp:181-199 */
void rs_ldlt_inverse_diagonal(const double *LD, const int64_t *tr, int64_t n, double *out) {
  int64_t *indices = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  int64_t *offsets = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) {
    indices[i] = i;
    offsets[i] = i;
  }
  offsets[n] = n;
  rs_ldlt_inverse_blocks(LD, tr, n, indices, offsets, n, out);
  free(indices);
  free(offsets);
}

/* stats/gaussian.hpp:19-23 via likelihood.hpp:21-24 */
static double univariate_nll(double deviation, double variance) {
  double ll = -deviation * deviation / (2 * variance);
  ll -= 0.5 * log(2 * M_PI * variance);
  return -ll;
}

/* likelihood.hpp:38-67 */
static double nll_from_ldlt(const double *deviation, const double *LD, const int64_t *tr,
                            int64_t n) {
  double *sol = (double *)malloc((size_t)n * sizeof(double));
  memcpy(sol, deviation, (size_t)n * sizeof(double));
  rs_ldlt_solve(LD, tr, n, sol, 1);
  double mahalanobis = 0.;
  for (int64_t i = 0; i < n; ++i) {
    mahalanobis += deviation[i] * sol[i];
  }
  free(sol);
  const double log_det = rs_ldlt_logdet(LD, n);
  return 0.5 * (log_det + mahalanobis + (double)n * log(2 * M_PI));
}

double rs_nll_dense(const double *deviation, const double *cov, int64_t n) {
  if (n == 1) {
    return univariate_nll(deviation[0], cov[0]);
  }
  double *A = (double *)malloc((size_t)n * (size_t)n * sizeof(double));
  int64_t *tr = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  memcpy(A, cov, (size_t)n * (size_t)n * sizeof(double));
  rs_ldlt(A, n, tr);
  const double out = nll_from_ldlt(deviation, A, tr, n);
  free(A);
  free(tr);
  return out;
}

/* ------------------------------------------------------------------------------------------- */
/* Exact GP: include/albatross/src/models/gp.hpp                                                */
/* ------------------------------------------------------------------------------------------- */

/* _fit_impl gp.hpp:285-294 + Fit ctor gp.hpp:61-69: K + diag(yvar) -> LDLT -> information. */
int rs_gp_fit(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
              const double *y, const double *yvar, double *information, double *LD_out,
              int64_t *tr_out) {
  double *A = LD_out ? LD_out : (double *)malloc((size_t)n * (size_t)n * sizeof(double));
  int64_t *tr = tr_out ? tr_out : (int64_t *)malloc((size_t)n * sizeof(int64_t));
  rs_gram_sym(prog, nops, feats, n, dim, A);
  if (yvar) {
    for (int64_t i = 0; i < n; ++i) {
      A[IDX(i, i, n)] += yvar[i];
    }
  }
  const int rc = rs_ldlt(A, n, tr);
  memcpy(information, y, (size_t)n * sizeof(double));
  rs_ldlt_solve(A, tr, n, information, 1);
  if (!LD_out) {
    free(A);
  }
  if (!tr_out) {
    free(tr);
  }
  return rc;
}

/*
 * _predict_impl gp.hpp:313-366 + gp_{mean,marginal,joint}_prediction gp.hpp:82-113.
 * what: 0 mean, 1 marginal (var), 2 joint (cov, p x p).
 */
int rs_gp_predict(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                  const double *y, const double *yvar, const double *test, int64_t p, int what,
                  double *mean, double *var, double *cov) {
  double *LD = (double *)malloc((size_t)n * (size_t)n * sizeof(double));
  int64_t *tr = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  double *information = (double *)malloc((size_t)n * sizeof(double));
  rs_gp_fit(prog, nops, feats, n, dim, y, yvar, information, LD, tr);
  double *cross = (double *)malloc((size_t)n * (size_t)p * sizeof(double));
  rs_gram_cross(prog, nops, feats, n, test, p, dim, cross);
  for (int64_t j = 0; j < p; ++j) {
    double s = 0.;
    for (int64_t i = 0; i < n; ++i) {
      s += cross[IDX(i, j, n)] * information[i];
    }
    mean[j] = s;
  }
  if (what >= 1) {
    double *explained = (double *)malloc((size_t)n * (size_t)p * sizeof(double));
    memcpy(explained, cross, (size_t)n * (size_t)p * sizeof(double));
    rs_ldlt_solve(LD, tr, n, explained, p);
    if (what == 1) {
      rs_gram_diag(prog, nops, test, p, dim, var);
      for (int64_t j = 0; j < p; ++j) {
        double s = 0.;
        for (int64_t i = 0; i < n; ++i) {
          s += explained[IDX(i, j, n)] * cross[IDX(i, j, n)];
        }
        var[j] -= s;
      }
    } else {
      rs_gram_sym(prog, nops, test, p, dim, cov);
      for (int64_t b = 0; b < p; ++b) {
        for (int64_t a = 0; a < p; ++a) {
          double s = 0.;
          for (int64_t i = 0; i < n; ++i) {
            s += cross[IDX(i, a, n)] * explained[IDX(i, b, n)];
          }
          cov[IDX(a, b, p)] -= s;
        }
      }
    }
    free(explained);
  }
  free(cross);
  free(information);
  free(tr);
  free(LD);
  return 0;
}

/* log_likelihood gp.hpp:443-451 (data term; no targets.covariance). */
double rs_gp_nll(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                 const double *y) {
  double *K = (double *)malloc((size_t)n * (size_t)n * sizeof(double));
  rs_gram_sym(prog, nops, feats, n, dim, K);
  const double out = rs_nll_dense(y, K, n);
  free(K);
  return out;
}

/* small dense helpers for held_out_prediction: LDLT-based solve / inverse of a |g| x |g| block */
static void small_inverse(const double *A, int64_t k, double *inv) {
  double *LD = (double *)malloc((size_t)k * (size_t)k * sizeof(double));
  int64_t *tr = (int64_t *)malloc((size_t)k * sizeof(int64_t));
  memcpy(LD, A, (size_t)k * (size_t)k * sizeof(double));
  rs_ldlt(LD, k, tr);
  memset(inv, 0, (size_t)k * (size_t)k * sizeof(double));
  for (int64_t i = 0; i < k; ++i) {
    inv[IDX(i, i, k)] = 1.;
  }
  rs_ldlt_solve(LD, tr, k, inv, k);
  free(LD);
  free(tr);
}

/*
 * gp_cross_validated_predictions gp.hpp:467-482 -> held_out_predictions
 * cross_validation_utils.hpp:199-232 -> held_out_prediction :172-197.
 * what 0: means; 1: marginals (mean + diag(A^-1)); 2: joints (mean + A^-1 blocks back to back).
 * mean/var are scattered back to the original order (concatenate_*_predictions :59-100).
 * score (optional) = sum_g negative_log_likelihood(truth_g - mean_g, cov_g) with the joint
 * covariance (prediction_metrics.hpp:112-134 for JointDistribution).
 */
int rs_gp_cv(const rs_op *prog, int nops, const double *feats, int64_t n, int dim, const double *y,
             const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
             double *mean, double *var, double *joint, double *score) {
  double *LD = (double *)malloc((size_t)n * (size_t)n * sizeof(double));
  int64_t *tr = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  double *information = (double *)malloc((size_t)n * sizeof(double));
  rs_gp_fit(prog, nops, feats, n, dim, y, NULL, information, LD, tr);
  size_t total = 0;
  int64_t maxg = 0;
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t sz = offsets[g + 1] - offsets[g];
    total += (size_t)sz * (size_t)sz;
    if (sz > maxg) {
      maxg = sz;
    }
  }
  double *blocks = (double *)malloc(total * sizeof(double));
  rs_ldlt_inverse_blocks(LD, tr, n, indices, offsets, ngroups, blocks);
  double *Ainv = (double *)malloc((size_t)maxg * (size_t)maxg * sizeof(double));
  double *dev = (double *)malloc((size_t)maxg * sizeof(double));
  const double *blk = blocks;
  double *jcur = joint;
  double total_score = 0.;
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t *idx = indices + offsets[g];
    const int64_t sz = offsets[g + 1] - offsets[g];
    small_inverse(blk, sz, Ainv);
    for (int64_t a = 0; a < sz; ++a) {
      double s = 0.;
      for (int64_t b = 0; b < sz; ++b) {
        s += Ainv[IDX(a, b, sz)] * information[idx[b]];
      }
      mean[idx[a]] = y[idx[a]] - s;
      dev[a] = s; /* truth - mean */
      if (what == 1 && var) {
        var[idx[a]] = Ainv[IDX(a, a, sz)];
      }
    }
    if (what == 2 && jcur) {
      memcpy(jcur, Ainv, (size_t)sz * (size_t)sz * sizeof(double));
      jcur += sz * sz;
    }
    if (score) {
      total_score += rs_nll_dense(dev, Ainv, sz);
    }
    blk += sz * sz;
  }
  if (score) {
    *score = total_score;
  }
  free(dev);
  free(Ainv);
  free(blocks);
  free(information);
  free(tr);
  free(LD);
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Integer contract: include/albatross/src/indexing/                                            */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
  int64_t key;
  int64_t index;
} key_index;

static int key_index_cmp(const void *a, const void *b) {
  const key_index *x = (const key_index *)a;
  const key_index *y = (const key_index *)b;
  if (x->key != y->key) {
    return x->key < y->key ? -1 : 1;
  }
  return x->index < y->index ? -1 : (x->index > y->index ? 1 : 0);
}

/*
 * IndexerBuilder::build group_by.hpp:349-376: std::map<key, indices> — keys ascending, indices of a
 * group in encounter (ascending) order.  group_keys[i] = grouper(feature_i) evaluated by the caller.
 */
int64_t rs_group_indexers(const int64_t *group_keys, int64_t n, int64_t *keys, int64_t *offsets,
                          int64_t *indices) {
  key_index *ki = (key_index *)malloc((size_t)n * sizeof(key_index));
  for (int64_t i = 0; i < n; ++i) {
    ki[i].key = group_keys[i];
    ki[i].index = i;
  }
  qsort(ki, (size_t)n, sizeof(key_index), key_index_cmp);
  int64_t g = 0;
  offsets[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (i == 0 || ki[i].key != ki[i - 1].key) {
      if (i > 0) {
        offsets[g] = i;
      }
      keys[g++] = ki[i].key;
    }
    indices[i] = ki[i].index;
  }
  offsets[g] = n;
  free(ki);
  return g;
}

/* indexing/block.hpp:25-44 */
int64_t rs_partition_triangular(int64_t n, int64_t count, int64_t *out) {
  double area = 0;
  int64_t start = 0;
  for (int64_t b = 0; b < count; ++b) {
    const double end_fraction = sqrt(1 / (double)count + area);
    area = end_fraction * end_fraction;
    const int64_t end = (int64_t)rint((double)n * end_fraction);
    out[2 * b] = start;
    out[2 * b + 1] = end;
    start = end;
  }
  if (count > 0 && out[2 * (count - 1) + 1] > n) {
    out[2 * (count - 1) + 1] = n;
  }
  return count;
}

/* indexing/subset.hpp:198-205 */
int64_t rs_indices_complement(const int64_t *indices, int64_t count, int64_t n, int64_t *out) {
  char *present = (char *)calloc((size_t)n, 1);
  for (int64_t i = 0; i < count; ++i) {
    if (indices[i] >= 0 && indices[i] < n) {
      present[indices[i]] = 1;
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (!present[i]) {
      out[k++] = i;
    }
  }
  free(present);
  return k;
}

/* utils/vector_utils.hpp:24-33 (accumulating form) */
void rs_linspace(double a, double b, int64_t n, double *out) {
  const double step = (b - a) / (double)(n - 1);
  double val = a;
  for (int64_t i = 0; i < n; ++i) {
    out[i] = val;
    val += step;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* Sparse GP: include/albatross/src/models/sparse_gp.hpp                                         */
/* ------------------------------------------------------------------------------------------- */

/*
 * Householder QR of B (rows x cols, rows >= cols), unpivoted, applied to one rhs.
 * The reference uses Eigen::ColPivHouseholderQR (sparse_gp.hpp:81-89); column pivoting only
 * permutes R's columns — v = argmin |B v - y_aug|, predictions and the log-likelihood are invariant
 * to it for full-column-rank B (which the K_uu^{T/2} block guarantees) — so the permutation and
 * R itself are representation-internal (SURVEY.md §7 hard part 3) and are not restated.
 * On exit B's upper triangle holds R and rhs holds Q^T rhs.
 */
static void householder_qr(double *B, int64_t rows, int64_t cols, double *rhs) {
  double *v = (double *)malloc((size_t)rows * sizeof(double));
  for (int64_t k = 0; k < cols; ++k) {
    double norm2 = 0.;
    for (int64_t i = k; i < rows; ++i) {
      norm2 += B[IDX(i, k, rows)] * B[IDX(i, k, rows)];
    }
    const double alpha = B[IDX(k, k, rows)];
    const double normx = sqrt(norm2);
    if (normx == 0.) {
      continue;
    }
    const double beta = alpha >= 0. ? -normx : normx;
    v[k] = alpha - beta;
    for (int64_t i = k + 1; i < rows; ++i) {
      v[i] = B[IDX(i, k, rows)];
    }
    double vtv = 0.;
    for (int64_t i = k; i < rows; ++i) {
      vtv += v[i] * v[i];
    }
    if (vtv == 0.) {
      continue;
    }
    const double tau = 2. / vtv;
    for (int64_t j = k; j < cols; ++j) {
      double s = 0.;
      for (int64_t i = k; i < rows; ++i) {
        s += v[i] * B[IDX(i, j, rows)];
      }
      s *= tau;
      for (int64_t i = k; i < rows; ++i) {
        B[IDX(i, j, rows)] -= s * v[i];
      }
    }
    if (rhs) {
      double s = 0.;
      for (int64_t i = k; i < rows; ++i) {
        s += v[i] * rhs[i];
      }
      s *= tau;
      for (int64_t i = k; i < rows; ++i) {
        rhs[i] -= s * v[i];
      }
    }
  }
  free(v);
}

static void upper_solve(const double *R, int64_t ld, int64_t m, double *b) {
  for (int64_t j = m - 1; j >= 0; --j) {
    b[j] /= R[IDX(j, j, ld)];
    for (int64_t i = 0; i < j; ++i) {
      b[i] -= R[IDX(i, j, ld)] * b[j];
    }
  }
}

/* R^-T rhs, qr_utils.hpp:37-45 (P = identity here). */
static void upper_T_solve(const double *R, int64_t ld, int64_t m, double *b) {
  for (int64_t j = 0; j < m; ++j) {
    double s = b[j];
    for (int64_t i = 0; i < j; ++i) {
      s -= R[IDX(i, j, ld)] * b[i];
    }
    b[j] = s / R[IDX(j, j, ld)];
  }
}

/*
 * _fit_impl sparse_gp.hpp:381-404, compute_internal_components :632-706, compute_sigma_qr :368-375,
 * _predict_impl :468-536, log_likelihood :539-603.  1-D or dim-D features are both fine here but the
 * signature is 1-D (the scoped configs).  Groups come in as a CSR indexer (keys ascending).
 * what: -1 fit only, 0 mean, 1 marginal, 2 joint.  ll (optional) = log_likelihood without prior.
 */
int rs_sparse_gp(const rs_op *prog, int nops, const double *feats, int64_t n, const double *y,
                 const double *yvar, const double *inducing, int64_t m, const int64_t *indices,
                 const int64_t *offsets, int64_t ngroups, double measurement_nugget,
                 double inducing_nugget, const double *test, int64_t p, int what,
                 double *information, double *mean, double *var, double *cov, double *ll) {
  return rs_sparse_gp2(prog, nops, prog, nops, prog, nops, feats, n, y, yvar, inducing, m, indices,
                       offsets, ngroups, measurement_nugget, inducing_nugget, test, p, what,
                       information, mean, var, cov, ll);
}

/*
 * The same with the three programs the reference evaluates (sparse_gp.hpp:646-679): prog = k(Measurement,
 * Measurement) for the K_ff blocks, prog_fu = k(Measurement, U) for K_fu, prog_uu = k(U, U) for K_uu and for
 * the predictions' cross / prior covariances (inducing and test features are both unwrapped there,
 * :470-536).  They differ when the tree holds a MeasurementOnly term (measurement.hpp:70-114).
 */
int rs_sparse_gp2(const rs_op *prog, int nops, const rs_op *prog_fu, int nops_fu, const rs_op *prog_uu,
                  int nops_uu, const double *feats, int64_t n, const double *y, const double *yvar,
                  const double *inducing, int64_t m, const int64_t *indices, const int64_t *offsets,
                  int64_t ngroups, double measurement_nugget, double inducing_nugget,
                  const double *test, int64_t p, int what, double *information, double *mean,
                  double *var, double *cov, double *ll) {
  const int dim = 1;
  /* reorder by group (:649-668) */
  double *xf = (double *)malloc((size_t)n * sizeof(double));
  double *yr = (double *)malloc((size_t)n * sizeof(double));
  double *vr = (double *)calloc((size_t)n, sizeof(double));
  for (int64_t i = 0; i < n; ++i) {
    xf[i] = feats[indices[i]];
    yr[i] = y[indices[i]];
    if (yvar) {
      vr[i] = yvar[indices[i]];
    }
  }
  /* K_fu (:670-671), K_uu + nugget (:673-679) */
  double *K_fu = (double *)malloc((size_t)n * (size_t)m * sizeof(double));
  rs_gram_cross(prog_fu, nops_fu, xf, n, inducing, m, dim, K_fu);
  double *Kuu = (double *)malloc((size_t)m * (size_t)m * sizeof(double));
  int64_t *tru = (int64_t *)malloc((size_t)m * sizeof(int64_t));
  rs_gram_sym(prog_uu, nops_uu, inducing, m, dim, Kuu);
  for (int64_t i = 0; i < m; ++i) {
    Kuu[IDX(i, i, m)] += inducing_nugget;
  }
  rs_ldlt(Kuu, m, tru);
  /* P = K_uu^-1/2 K_uf  (m x n) (:684) */
  double *P = (double *)malloc((size_t)m * (size_t)n * sizeof(double));
  for (int64_t j = 0; j < n; ++j) {
    for (int64_t i = 0; i < m; ++i) {
      P[IDX(i, j, m)] = K_fu[IDX(j, i, n)];
    }
  }
  rs_ldlt_sqrt_solve(Kuu, tru, m, P, n);
  /* per-group A_g = K_ff,g + diag(var) - P_g^T P_g + nugget I, LDLT (:652-705);
     B top = A^-1/2 K_fu, y_a = A^-1/2 y (block_diagonal.hpp:96-123,180-218) */
  const int64_t rows = n + m;
  double *B = (double *)calloc((size_t)rows * (size_t)m, sizeof(double));
  double *y_aug = (double *)calloc((size_t)rows, sizeof(double));
  double *y_a = (double *)malloc((size_t)n * sizeof(double)); /* A^-1 y */
  double log_det_a = 0.;
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t o = offsets[g];
    const int64_t sz = offsets[g + 1] - o;
    double *A = (double *)malloc((size_t)sz * (size_t)sz * sizeof(double));
    int64_t *tra = (int64_t *)malloc((size_t)sz * sizeof(int64_t));
    rs_gram_sym(prog, nops, xf + o, sz, dim, A);
    for (int64_t b = 0; b < sz; ++b) {
      for (int64_t a = 0; a < sz; ++a) {
        double s = 0.;
        for (int64_t i = 0; i < m; ++i) {
          s += P[IDX(i, o + a, m)] * P[IDX(i, o + b, m)];
        }
        A[IDX(a, b, sz)] -= s;
      }
      A[IDX(b, b, sz)] += vr[o + b];
    }
    for (int64_t b = 0; b < sz; ++b) {
      A[IDX(b, b, sz)] += measurement_nugget;
    }
    rs_ldlt(A, sz, tra);
    log_det_a += rs_ldlt_logdet(A, sz);
    /* block rows of B */
    double *blk = (double *)malloc((size_t)sz * (size_t)(m + 1) * sizeof(double));
    for (int64_t j = 0; j < m; ++j) {
      for (int64_t a = 0; a < sz; ++a) {
        blk[IDX(a, j, sz)] = K_fu[IDX(o + a, j, n)];
      }
    }
    for (int64_t a = 0; a < sz; ++a) {
      blk[IDX(a, m, sz)] = yr[o + a];
    }
    rs_ldlt_sqrt_solve(A, tra, sz, blk, m + 1);
    for (int64_t j = 0; j < m; ++j) {
      for (int64_t a = 0; a < sz; ++a) {
        B[IDX(o + a, j, rows)] = blk[IDX(a, j, sz)];
      }
    }
    for (int64_t a = 0; a < sz; ++a) {
      y_aug[o + a] = blk[IDX(a, m, sz)];
      y_a[o + a] = yr[o + a];
    }
    rs_ldlt_solve(A, tra, sz, y_a + o, 1);
    free(blk);
    free(tra);
    free(A);
  }
  /* bottom rows: K_uu^{T/2} = D^1/2 (P^T L)^T  (serializable_ldlt.hpp:111-115) */
  {
    double *PtL = (double *)calloc((size_t)m * (size_t)m, sizeof(double));
    for (int64_t j = 0; j < m; ++j) {
      PtL[IDX(j, j, m)] = 1.;
      for (int64_t i = j + 1; i < m; ++i) {
        PtL[IDX(i, j, m)] = Kuu[IDX(i, j, m)];
      }
    }
    apply_transpositions_T(tru, m, PtL, m);
    for (int64_t i = 0; i < m; ++i) {
      const double d = Kuu[IDX(i, i, m)];
      const double s = d > 0. ? sqrt(d) : 0.;
      for (int64_t j = 0; j < m; ++j) {
        B[IDX(n + i, j, rows)] = s * PtL[IDX(j, i, m)];
      }
    }
    free(PtL);
  }
  householder_qr(B, rows, m, y_aug);
  /* v = R^-1 (Q^T y_aug)(0:m) (:396-398) */
  double *v = (double *)malloc((size_t)m * sizeof(double));
  memcpy(v, y_aug, (size_t)m * sizeof(double));
  upper_solve(B, rows, m, v);
  if (information) {
    memcpy(information, v, (size_t)m * sizeof(double));
  }
  if (ll) {
    double log_det_r = 0.;
    for (int64_t i = 0; i < m; ++i) {
      log_det_r += log(fabs(B[IDX(i, i, rows)]));
    }
    const double log_det = log_det_a + 2 * log_det_r - rs_ldlt_logdet(Kuu, m);
    double *y_b = (double *)calloc((size_t)m, sizeof(double));
    double quad = 0.;
    for (int64_t i = 0; i < n; ++i) {
      quad += yr[i] * y_a[i];
    }
    for (int64_t j = 0; j < m; ++j) {
      double s = 0.;
      for (int64_t i = 0; i < n; ++i) {
        s += K_fu[IDX(i, j, n)] * y_a[i];
      }
      y_b[j] = s;
    }
    upper_T_solve(B, rows, m, y_b);
    for (int64_t j = 0; j < m; ++j) {
      quad -= y_b[j] * y_b[j];
    }
    *ll = -0.5 * (log_det + quad + (double)n * log(2 * M_PI));
    free(y_b);
  }
  if (what >= 0) {
    double *cross = (double *)malloc((size_t)m * (size_t)p * sizeof(double));
    rs_gram_cross(prog_uu, nops_uu, inducing, m, test, p, dim, cross);
    for (int64_t j = 0; j < p; ++j) {
      double s = 0.;
      for (int64_t i = 0; i < m; ++i) {
        s += cross[IDX(i, j, m)] * v[i];
      }
      mean[j] = s;
    }
    if (what >= 1) {
      double *Q = (double *)malloc((size_t)m * (size_t)p * sizeof(double));
      double *S = (double *)malloc((size_t)m * (size_t)p * sizeof(double));
      memcpy(Q, cross, (size_t)m * (size_t)p * sizeof(double));
      memcpy(S, cross, (size_t)m * (size_t)p * sizeof(double));
      rs_ldlt_sqrt_solve(Kuu, tru, m, Q, p);
      for (int64_t j = 0; j < p; ++j) {
        upper_T_solve(B, rows, m, S + (size_t)j * (size_t)m);
      }
      if (what == 1) {
        rs_gram_diag(prog_uu, nops_uu, test, p, dim, var);
        for (int64_t j = 0; j < p; ++j) {
          double q = 0., s = 0.;
          for (int64_t i = 0; i < m; ++i) {
            q += Q[IDX(i, j, m)] * Q[IDX(i, j, m)];
            s += S[IDX(i, j, m)] * S[IDX(i, j, m)];
          }
          var[j] = var[j] - q + s;
        }
      } else {
        rs_gram_sym(prog_uu, nops_uu, test, p, dim, cov);
        for (int64_t b = 0; b < p; ++b) {
          for (int64_t a = 0; a < p; ++a) {
            double q = 0., s = 0.;
            for (int64_t i = 0; i < m; ++i) {
              q += Q[IDX(i, a, m)] * Q[IDX(i, b, m)];
              s += S[IDX(i, a, m)] * S[IDX(i, b, m)];
            }
            cov[IDX(a, b, p)] = cov[IDX(a, b, p)] - q + s;
          }
        }
      }
      free(Q);
      free(S);
    }
    free(cross);
  }
  free(v);
  free(y_a);
  free(y_aug);
  free(B);
  free(P);
  free(tru);
  free(Kuu);
  free(K_fu);
  free(vr);
  free(yr);
  free(xf);
  return 0;
}
