// Dense fp64 building blocks on the device: DMMA GEMM family, blocked Cholesky, triangular solves,
// reductions.  All matrices column-major.
#pragma once

#include "common.cuh"

namespace ab {

constexpr int LEAF = 64; // diagonal-block size of the recursive factorisation / solves

struct MatView {
  double *p;
  int64_t ld;
  __host__ __device__ MatView sub(int64_t r, int64_t c) const { return MatView{p + r + c * ld, ld}; }
};

enum GemmFlags : unsigned {
  GEMM_TRANS_A = 1u, // op(A) = A^T, A stored k x m
  GEMM_TRANS_B = 2u, // op(B) = B^T, B stored n x k
  GEMM_LOWER = 4u    // C is square (m == n): only tiles touching the lower triangle are computed
};

// C[m x n] = alpha * op(A) * op(B) + beta * C.   beta == 0 never reads C.
// In-place use (C aliasing A or B) is allowed only when one CTA owns the whole aliased extent:
// C == A with n <= 64 (= the CTA tile's BN) and k == n (right-multiplication by a small matrix), or
// C == B with m <= 128 (= BM) and k == m (left-multiplication).  The callers are the LEAF = 64 wide
// triangular-solve leaves.
int gemm(ab_handle_s *h, unsigned flags, int64_t m, int64_t n, int64_t k, double alpha, MatView A,
         MatView B, double beta, MatView C);

// TMA-fed kernel of the C = alpha A B^T + beta C product (gemm_tma.cu; on unless AB_GEMM_TMA=0).
// AB_ERR_UNSUPPORTED = shape / alignment it does not cover: use the cp.async kernel.
bool gemm_tma_enabled();
// Block-cyclic B operand (the one-launch trailing update of the distributed factorisation, dist.cu): column
// block q (width blk, a multiple of 64) of C multiplies rows row0 + q * stride + [0, blk) of B, which has
// `rows` rows in all; with lower == true a tile is skipped when it lies entirely above that stretched
// diagonal (its last row < row index q * stride + c of its first column).
struct CyclicB {
  int64_t blk = 0;
  int64_t stride = 0;
  int64_t row0 = 0;
  int64_t rows = 0;
};
int gemm_nt_tma(ab_handle_s *h, bool lower, int64_t m, int64_t n, int64_t k, double alpha, MatView A,
                MatView B, double beta, MatView C, const CyclicB *cyc = nullptr);

// Blocked Cholesky of the lower triangle of the n x n matrix A, in place.  dinv receives the
// explicit inverses of the LEAF x LEAF diagonal blocks of L (block j at dinv + j*LEAF*LEAF,
// column-major LEAF x LEAF, upper triangle zero).  d_bad: device int, atomicMin'ed with the global
// index of the first non-positive pivot (initialise to INT_MAX).
// d_floor (optional, n doubles on the device): pivot i is accepted only if it exceeds d_floor[i]; default:
// 64 eps times the diagonal of A as passed in (pivot_floor).  Callers that factor an already-updated block
// pass the floor of its ORIGINAL diagonal.
int potrf(ab_handle_s *h, MatView A, int64_t n, double *dinv, int *d_bad, const double *d_floor = nullptr);
int pivot_floor(ab_handle_s *h, MatView A, int64_t n, double *d_floor);

// Creates (once) the handle's high-priority panel stream used by the look-ahead factorisations.
int ensure_panel_stream(ab_handle_s *h);

// X <- L^-1 X  (n x p), X <- L^-T X, X <- X L^-T (m x n, L n x n), using dinv for the leaves.
int trsm_left_lower(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView X, int64_t p);
int trsm_left_lower_T(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView X,
                      int64_t p);
int trsm_right_lower_T(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView X,
                       int64_t m);

// Y <- L X (trans == false) or Y <- L^T X (trans == true), n x p; only the lower triangle of L is read
// (the strict upper triangle of an in-place factor holds stale data).  Y must not alias X.
int trmm_left_lower(ab_handle_s *h, MatView L, int64_t n, bool trans, MatView X, MatView Y, int64_t p);

// ---- one right-hand side (trsv.cu) --------------------------------------------------------------------
// y[m] = beta y + alpha A[m x k] x[k] and y[k] = beta y + alpha A[m x k]^T x[m]; 16-byte loads: callers check
// gemv_fast_ok (A 16-byte aligned with an even leading dimension, x 16-byte aligned).
bool gemv_fast_ok(MatView A, const double *x);
int gemv_n(ab_handle_s *h, int64_t m, int64_t k, double alpha, MatView A, const double *x, double beta,
           double *y);
int gemv_t(ab_handle_s *h, int64_t m, int64_t k, double alpha, MatView A, const double *x, double beta,
           double *y);
// One diagonal block of at most 1024 rows solved by one CTA: x <- L^-1 x or L^-T x.
int trsv_block(ab_handle_s *h, bool trans, MatView L, const double *dinv, int64_t nb, double *x);
// x <- L^-1 x, x <- L^-T x for a contiguous vector; L 16-byte aligned with an even leading dimension.
int trsv_lower(ab_handle_s *h, MatView L, const double *dinv, int64_t n, double *x);
int trsv_lower_T(ab_handle_s *h, MatView L, const double *dinv, int64_t n, double *x);

// out[0] = sum_i 2 log(L_ii)
int logdet_chol(ab_handle_s *h, MatView L, int64_t n, double *d_out);
// out[0] = sum_i a_i * b_i
int dot(ab_handle_s *h, const double *a, const double *b, int64_t n, double *d_out);
// out[j] = sum_i A(i,j) * B(i,j), j < cols
int column_dots(ab_handle_s *h, MatView A, MatView B, int64_t rows, int64_t cols, double *d_out);
int add_diag(ab_handle_s *h, MatView A, int64_t n, const double *d_diag);
int fill(ab_handle_s *h, MatView A, int64_t rows, int64_t cols, double value);
int set_identity(ab_handle_s *h, MatView A, int64_t n);
// NaN scan of the lower triangle (gp.hpp:66): sets *d_flag = 1 if any NaN.
int has_nan_lower(ab_handle_s *h, MatView A, int64_t n, int *d_flag);

} // namespace ab
