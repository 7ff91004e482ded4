/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's exact-GP hot path.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle.py) against
 *   (1) the reference's literal golden vectors: the gpytorch Matern-5/2 and -3/2 15x15 tables
 *       (tests/test_radial.cc:212-489), the scipy NLL known answer 6.0946974293510134
 *       (tests/test_evaluate.cc:34-63), the radial edge cases (tests/test_radial.cc:52-66);
 *   (2) outputs of the reference itself, compiled in place into oracle/_ref/libref_oracle.so
 *       (oracle/ref_shim/), on seeded inputs, and the fixtures under tests/golden/ generated from it
 *       by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this library.  All reference citations are relative to /root/reference/.
 *
 * Conventions: matrices column-major fp64 (Eigen::MatrixXd), sizes int64_t, features AoS
 * (point i at feats[i*dim .. i*dim+dim)), covariance given as a postfix program of rs_op.
 */
#ifndef AB_ORACLE_RESTATE_H
#define AB_ORACLE_RESTATE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  RS_OP_SQUARED_EXPONENTIAL = 1, /* p0 = length_scale, p1 = sigma */
  RS_OP_EXPONENTIAL = 2,
  RS_OP_MATERN32 = 3,
  RS_OP_MATERN52 = 4,
  RS_OP_CONSTANT = 5,          /* p0 = sigma */
  RS_OP_INDEPENDENT_NOISE = 6, /* p0 = sigma */
  RS_OP_SUM = 7,               /* pops rhs, lhs; pushes lhs + rhs */
  RS_OP_PRODUCT = 8,
  RS_OP_POLYNOMIAL_TERM = 9 /* p0 = sigma, p1 = degree: one term of Polynomial<order>, polynomials.hpp:79-87 */            /* pops rhs, lhs; pushes lhs != 0 ? lhs * rhs : lhs */
};

typedef struct {
  int32_t op;
  int32_t reserved;
  double p0;
  double p1;
} rs_op;

double rs_cov_eval(const rs_op *prog, int nops, const double *x, const double *y, int dim);
void rs_gram_sym(const rs_op *prog, int nops, const double *feats, int64_t n, int dim, double *out);
void rs_gram_cross(const rs_op *prog, int nops, const double *fx, int64_t n, const double *fy,
                   int64_t m, int dim, double *out);
void rs_gram_diag(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                  double *out);

int rs_ldlt(double *A, int64_t n, int64_t *transpositions);
void rs_ldlt_solve(const double *LD, const int64_t *tr, int64_t n, double *B, int64_t k);
void rs_ldlt_sqrt_solve(const double *LD, const int64_t *tr, int64_t n, double *B, int64_t k);
double rs_ldlt_logdet(const double *LD, int64_t n);
void rs_ldlt_inverse_blocks(const double *LD, const int64_t *tr, int64_t n, const int64_t *indices,
                            const int64_t *offsets, int64_t ngroups, double *out);
void rs_ldlt_inverse_diagonal(const double *LD, const int64_t *tr, int64_t n, double *out);
double rs_nll_dense(const double *deviation, const double *cov, int64_t n);

int rs_gp_fit(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
              const double *y, const double *yvar, double *information, double *LD_out,
              int64_t *tr_out);
int rs_gp_predict(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                  const double *y, const double *yvar, const double *test, int64_t p, int what,
                  double *mean, double *var, double *cov);
double rs_gp_nll(const rs_op *prog, int nops, const double *feats, int64_t n, int dim,
                 const double *y);
int rs_gp_cv(const rs_op *prog, int nops, const double *feats, int64_t n, int dim, const double *y,
             const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
             double *mean, double *var, double *joint, double *score);

int64_t rs_group_indexers(const int64_t *group_keys, int64_t n, int64_t *keys, int64_t *offsets,
                          int64_t *indices);
int64_t rs_partition_triangular(int64_t n, int64_t count, int64_t *out);
int64_t rs_indices_complement(const int64_t *indices, int64_t count, int64_t n, int64_t *out);
void rs_linspace(double a, double b, int64_t n, double *out);

int rs_sparse_gp(const rs_op *prog, int nops, const double *feats, int64_t n, const double *y,
                 const double *yvar, const double *inducing, int64_t m, const int64_t *indices,
                 const int64_t *offsets, int64_t ngroups, double measurement_nugget,
                 double inducing_nugget, const double *test, int64_t p, int what,
                 double *information, double *mean, double *var, double *cov, double *ll);

int rs_sparse_gp2(const rs_op *prog, int nops, const rs_op *prog_fu, int nops_fu, const rs_op *prog_uu,
                  int nops_uu, const double *feats, int64_t n, const double *y, const double *yvar,
                  const double *inducing, int64_t m, const int64_t *indices, const int64_t *offsets,
                  int64_t ngroups, double measurement_nugget, double inducing_nugget,
                  const double *test, int64_t p, int what, double *information, double *mean,
                  double *var, double *cov, double *ll);

#ifdef __cplusplus
}
#endif
#endif
