// Sparse (FITC / PITC) Gaussian process on the device.
//
// Replaces SparseGaussianProcessRegression::{compute_internal_components, compute_sigma_qr,
// _fit_impl, _predict_impl x3, log_likelihood} (reference include/albatross/src/models/
// sparse_gp.hpp:368-404, 468-603, 632-706), BlockDiagonalLDLT (linalg/block_diagonal.hpp:96-218)
// and the QR helpers (linalg/qr_utils.hpp:18-53).
//
// The reference solves  min_v | B v - [A^-1/2 y ; 0] |  with  B = [A^-1/2 K_fu ; K_uu^{T/2}]  by a
// column-pivoted Householder QR of the (n+m) x m matrix B (level-2 BLAS, 2(n+m)m^2 flops).  Here:
//
//   K_uu = L_u L_u^T,   P^T = K_fu L_u^-T (n x m),   A_g = K_ff,g + diag(var) - P_g^T P_g + nugget I
//   B = C L_u^T  with  C = [A^-1/2 P^T ; I]          (because K_fu = P^T L_u^T)
//
// C is far better conditioned than B (C^T C = I + P A^-1 P^T, no K_uu factor), so its QR is taken
// by CholQR2 — two passes of  S = C^T C (DSYRK on the FP64 tensor pipe), S = L L^T,  C <- C L^-T —
// which is as stable as Householder for cond(C) < 1e8 and is pure level-3 work:
//
//   C = Q (L1 L2)^T,  R_c = (L1 L2)^T,  R = R_c L_u^T  (the reference's sigma_R up to its column
//   permutation: R^T R = B^T B),  Q^T y_aug = L2^-1 (C^T y_aug),  v = L_u^-T R_c^-1 Q^T y_aug.
//
// y_aug rides along as column m of C so that every row operation (A^-1/2, the Gram products) is
// applied to it for free.  C is (n+m) x (m+1) fp64 and stays resident in HBM for the second pass:
// 34.4 GB at n = 2^20, m = 4096 — sized for a 180 GB part; K_fu, P and B are never materialised
// separately (the reference holds all three plus copies).
//
//   log|K|  = log|A| + 2 log|R| - log|K_uu| = log|A| + 2 log|R_c|
//   y^T K^-1 y = |A^-1/2 y|^2 - |Q^T y_aug|^2                               (sparse_gp.hpp:563-603)
#include "internal.cuh"

#include <cmath>
#include <cstring>
#include <numeric>

struct ab_sparse_fit_s {
  int64_t n = 0; // observations seen by this rank
  int64_t m = 0; // inducing points
  int dim = 1;
  ab_matrix_s *u = nullptr;  // dim x m inducing features (packed AoS)
  ab_factor_s *Ku = nullptr; // chol(K_uu + inducing_nugget I)
  ab_factor_s *L1 = nullptr; // CholQR pass 1
  ab_factor_s *L2 = nullptr; // CholQR pass 2
  ab_matrix_s *v = nullptr;  // information, m x 1
  double log_likelihood = 0.;
};

namespace ab {
namespace {

constexpr int RED_THREADS = 1024;

__device__ __forceinline__ double block_reduce_sum(double v) {
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  double total = 0.;
  if (threadIdx.x < 32) {
    total = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.;
    for (int o = 16; o > 0; o >>= 1) {
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
  }
  __syncthreads();
  return total; // valid in warp 0
}

// dst[i*dim + d] = src[idx[i]*dim + d]
__global__ void gather_features_kernel(const double *src, int dim, const int64_t *idx, int64_t n,
                                       double *dst) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t < n * dim) {
    const int64_t i = t / dim;
    const int d = static_cast<int>(t - i * dim);
    dst[t] = src[idx[i] * dim + d];
  }
}

// out[i] = src != nullptr ? src[idx[i]] : 0
__global__ void gather_or_zero_kernel(const double *src, const int64_t *idx, int64_t n,
                                      double *out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    out[i] = src != nullptr ? src[idx[i]] : 0.;
  }
}

// Bottom block of C: rows [n, n+m) = [I_m | 0].
__global__ void bottom_block_kernel(double *C, int64_t ld, int64_t n, int64_t m) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < m) {
    C[n + i + j * ld] = (i == j) ? 1. : 0.;
  }
}

__global__ void add_diag_scalar_kernel(double *A, int64_t ld, int64_t n, double value) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    A[i + i * ld] += value;
  }
}

// FITC: partial row sums of squares.  part[c * n + i] = sum_{j in chunk c} X(i, j)^2; thread = row
// (coalesced along the rows of the column-major panel), blockIdx.y = column chunk.
constexpr int ROW_CHUNK = 128;
__global__ void __launch_bounds__(256)
row_sumsq_partial_kernel(const double *X, int64_t ld, int64_t n, int64_t m, double *part) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= n) {
    return;
  }
  const int64_t j0 = c * ROW_CHUNK;
  const int64_t j1 = j0 + ROW_CHUNK < m ? j0 + ROW_CHUNK : m;
  double acc = 0.;
  for (int64_t j = j0; j < j1; ++j) {
    const double x = X[i + j * ld];
    acc = fma(x, x, acc);
  }
  part[c * n + i] = acc;
}

// FITC: a_i = ((k(x_i,x_i) + var_i) - |P_i|^2) + nugget  (sparse_gp.hpp:652-705 with 1x1 blocks);
// scale_i = a_i^-1/2 (sqrt_solve of a 1x1 LDLT), logs_i = log a_i.
__global__ void fitc_finalize_kernel(const double *kd, const double *var, const double *part,
                                     int64_t nchunks, int64_t n, double nugget, double *scale,
                                     double *logs, int *d_bad) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) {
    return;
  }
  double q = 0.;
  for (int64_t c = 0; c < nchunks; ++c) {
    q += part[c * n + i];
  }
  const double a = ((kd[i] + var[i]) - q) + nugget;
  if (!(a > 0.)) {
    atomicMin(d_bad, static_cast<int>(i < INT_MAX ? i : INT_MAX - 1));
  }
  scale[i] = 1. / sqrt(a);
  logs[i] = log(a);
}

// X(i, j) *= scale[i]
__global__ void row_scale_kernel(double *X, int64_t ld, int64_t n, const double *scale) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < n) {
    X[i + j * ld] *= scale[i];
  }
}

__global__ void __launch_bounds__(RED_THREADS) sum_kernel(const double *x, int64_t n,
                                                          double *out) {
  double acc = 0.;
  for (int64_t i = threadIdx.x; i < n; i += RED_THREADS) {
    acc += x[i];
  }
  const double total = block_reduce_sum(acc);
  if (threadIdx.x == 0) {
    out[0] = total;
  }
}

// dst (clean lower triangle, zeros above) = tril(src)
__global__ void tril_copy_kernel(const double *src, int64_t lds, int64_t n, double *dst,
                                 int64_t ldd) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < n) {
    dst[i + j * ldd] = i >= j ? src[i + j * lds] : 0.;
  }
}

// out(i, j) = src(j, i)
__global__ void transpose_kernel(const double *src, int64_t lds, int64_t n, double *dst,
                                 int64_t ldd) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < n) {
    dst[i + j * ldd] = src[j + i * lds];
  }
}

dim3 grid2(int64_t rows, int64_t cols) {
  return dim3(static_cast<unsigned>((rows + 255) / 256), static_cast<unsigned>(cols));
}

void free_sparse(ab_handle_s *h, ab_sparse_fit_s *f) {
  if (f == nullptr) {
    return;
  }
  matrix_delete(h, f->u);
  matrix_delete(h, f->v);
  delete_factor(h, f->Ku);
  delete_factor(h, f->L1);
  delete_factor(h, f->L2);
  delete f;
}

// One CholQR pass over the first `rows` rows of C (rows x (m+1), last column = y_aug):
//   S = C^T C (all-reduced over ranks);  g = S[m, 0:m] = C[:, :m]^T y_aug;  yy = S[m, m];
//   leading m x m of S -> L (factor).
int cholqr_pass(ab_handle_s *h, Scope &sc, MatView C, int64_t rows, int64_t m, ab_factor_s **L,
                double *d_g, double *d_yy) {
  ab_matrix_s *S = nullptr;
  AB_TRY(matrix_new(h, m + 1, m + 1, &S));
  // GEMM_LOWER writes only the lower triangle: the strict upper triangle and the ld padding are
  // recycled pool memory and must not ride through the all-reduce (or stay in the stored factors)
  int s = AB_OK;
  if (cudaMemsetAsync(S->d, 0, static_cast<size_t>(S->ld) * (m + 1) * sizeof(double), h->stream) !=
      cudaSuccess) {
    set_error("cholqr memset failed");
    s = AB_ERR_CUDA;
  }
  if (s == AB_OK) {
    s = gemm(h, GEMM_TRANS_A | GEMM_LOWER, m + 1, m + 1, rows, 1., C, C, 0., view(S));
  }
  if (s == AB_OK) {
    s = dist_allreduce_sum(h, S->d, S->ld * (m + 1));
  }
  if (s != AB_OK) {
    matrix_delete(h, S);
    return s;
  }
  // row m of the lower triangle -> contiguous vector; S[m, m] -> scalar
  cudaError_t e = cudaMemcpy2DAsync(d_g, sizeof(double), S->d + m, S->ld * sizeof(double),
                                    sizeof(double), static_cast<size_t>(m),
                                    cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess) {
    e = cudaMemcpyAsync(d_yy, S->d + m + m * S->ld, sizeof(double), cudaMemcpyDeviceToDevice,
                        h->stream);
  }
  if (e != cudaSuccess) {
    matrix_delete(h, S);
    set_error("cholqr copy failed: %s", cudaGetErrorString(e));
    return AB_ERR_CUDA;
  }
  (void)sc;
  S->rows = S->cols = m; // factor the leading m x m block in place (allocation size unchanged)
  return factorize(h, S, L);
}

// P_ff: k(Measurement, Measurement) for the diagonal blocks of K_ff, P_fu: k(Measurement, U), P_uu:
// k(U, U) (sparse_gp.hpp:646-679); the three differ when the tree holds a MeasurementOnly term.
int sparse_fit_impl(ab_handle_s *h, const DevProg &P_ff, const DevProg &P_fu, const DevProg &P_uu,
                    const double *feats, int64_t n, int dim,
                    const double *y, const double *yvar, const double *inducing, int64_t m,
                    const int64_t *indices, const int64_t *offsets, int64_t ngroups,
                    double measurement_nugget, double inducing_nugget, ab_sparse_fit_s *fit) {
  Scope sc(h);
  fit->n = n;
  fit->m = m;
  fit->dim = dim;
  bool all_singletons = true;
  int64_t maxg = 0;
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t sz = offsets[g + 1] - offsets[g];
    AB_REQUIRE(sz >= 0, "offsets must be non-decreasing");
    all_singletons = all_singletons && sz == 1;
    maxg = std::max(maxg, sz);
  }

  // ---- uploads and the group reordering (sparse_gp.hpp:649-668) ------------------------------
  phase_begin(h, PH_H2D);
  ab_matrix_s *F = nullptr, *XF = nullptr;
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  AB_TRY(upload_features(h, inducing, m, dim, &fit->u));
  void *d_idx = nullptr, *d_y = nullptr, *d_yvar = nullptr, *d_var = nullptr;
  const size_t nb = static_cast<size_t>(n < 1 ? 1 : n) * sizeof(double);
  AB_TRY(upload_bytes(h, sc, indices, static_cast<size_t>(n) * sizeof(int64_t), &d_idx));
  AB_TRY(upload_bytes(h, sc, y, static_cast<size_t>(n) * sizeof(double), &d_y));
  if (yvar != nullptr) {
    AB_TRY(upload_bytes(h, sc, yvar, static_cast<size_t>(n) * sizeof(double), &d_yvar));
  }
  phase_end(h, PH_H2D);
  AB_TRY(sc.alloc(nb, &d_var));
  {
    auto *xf = new ab_matrix_s();
    xf->rows = dim;
    xf->cols = n;
    xf->ld = dim;
    xf->bytes = static_cast<size_t>(dim) * static_cast<size_t>(n < 1 ? 1 : n) * sizeof(double);
    void *p = nullptr;
    int s = dev_alloc(h, xf->bytes, &p);
    if (s != AB_OK) {
      delete xf;
      return s;
    }
    xf->d = static_cast<double *>(p);
    XF = sc.own(xf);
  }
  const int64_t rows = n + m;
  ab_matrix_s *C = nullptr;
  AB_TRY(matrix_new(h, rows, m + 1, &C));
  sc.own(C);
  const MatView Cv = view(C);
  if (n > 0) {
    const unsigned gb = static_cast<unsigned>((n * dim + 255) / 256);
    gather_features_kernel<<<gb, 256, 0, h->stream>>>(F->d, dim, static_cast<int64_t *>(d_idx), n,
                                                      XF->d);
    AB_LAUNCHED(h);
    const unsigned nbk = static_cast<unsigned>((n + 255) / 256);
    gather_or_zero_kernel<<<nbk, 256, 0, h->stream>>>(static_cast<double *>(d_yvar),
                                                      static_cast<int64_t *>(d_idx), n,
                                                      static_cast<double *>(d_var));
    AB_LAUNCHED(h);
    // y_aug column (before A^-1/2 is applied): C[0:n, m] = y[indices]
    gather_or_zero_kernel<<<nbk, 256, 0, h->stream>>>(static_cast<double *>(d_y),
                                                      static_cast<int64_t *>(d_idx), n,
                                                      Cv.sub(0, m).p);
    AB_LAUNCHED(h);
  }
  bottom_block_kernel<<<grid2(m, m + 1), 256, 0, h->stream>>>(C->d, C->ld, n, m);
  AB_LAUNCHED(h);

  // ---- K_uu = L_u L_u^T (sparse_gp.hpp:673-679) -----------------------------------------------
  phase_begin(h, PH_GRAM);
  ab_matrix_s *Kuu = nullptr;
  AB_TRY(gram_sym_device(h, P_uu, fit->u, AB_GRAM_LOWER_ONLY, &Kuu));
  add_diag_scalar_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, h->stream>>>(
      Kuu->d, Kuu->ld, m, inducing_nugget);
  AB_LAUNCHED(h);
  // ---- K_fu straight into the top block of C (sparse_gp.hpp:670-671) ---------------------------
  if (n > 0) {
    int s = gram_into(h, P_fu, dim, false, XF->d, XF->ld, n, fit->u->d, fit->u->ld, m, C->d, C->ld,
                      0u);
    if (s != AB_OK) {
      matrix_delete(h, Kuu);
      return s;
    }
  }
  phase_end(h, PH_GRAM);
  phase_begin(h, PH_FACTOR);
  {
    int s = factorize(h, Kuu, &fit->Ku);
    if (s != AB_OK) {
      if (s == AB_ERR_NOT_PD) {
        set_error("K_uu + inducing_nugget I is not positive definite (pivot %lld)",
                  static_cast<long long>(fit->Ku->bad_pivot));
      }
      return s;
    }
  }
  // ---- P^T = K_fu L_u^-T  (sqrt_solve, sparse_gp.hpp:684), in place ----------------------------
  AB_TRY(trsm_right_lower_T(h, view(fit->Ku->m), fit->Ku->dinv, m, Cv, n));

  // ---- A = K_ff - Q_ff + noise (block diagonal), C_top = A^-1/2 [P^T | y]  (:652-705, :368-375)
  void *d_logs = nullptr;
  AB_TRY(sc.alloc(nb, &d_logs));
  AB_CUDA(cudaMemsetAsync(d_logs, 0, nb, h->stream));
  h->h_flags[0] = INT_MAX;
  AB_CUDA(cudaMemcpyAsync(h->d_flags, h->h_flags, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (n > 0 && all_singletons) {
    void *d_kd = nullptr, *d_part = nullptr, *d_scale = nullptr;
    const int64_t nchunks = (m + ROW_CHUNK - 1) / ROW_CHUNK;
    AB_TRY(sc.alloc(nb, &d_kd));
    AB_TRY(sc.alloc(nb, &d_scale));
    AB_TRY(sc.alloc(nb * static_cast<size_t>(nchunks), &d_part));
    AB_TRY(gram_diag_into(h, P_ff, dim, XF->d, XF->ld, n, static_cast<double *>(d_kd)));
    row_sumsq_partial_kernel<<<grid2(n, nchunks), 256, 0, h->stream>>>(
        C->d, C->ld, n, m, static_cast<double *>(d_part));
    AB_LAUNCHED(h);
    fitc_finalize_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(
        static_cast<double *>(d_kd), static_cast<double *>(d_var), static_cast<double *>(d_part),
        nchunks, n, measurement_nugget, static_cast<double *>(d_scale),
        static_cast<double *>(d_logs), h->d_flags);
    AB_LAUNCHED(h);
    row_scale_kernel<<<grid2(n, m + 1), 256, 0, h->stream>>>(C->d, C->ld, n,
                                                             static_cast<double *>(d_scale));
    AB_LAUNCHED(h);
  } else if (n > 0) {
    ab_matrix_s *A = nullptr;
    AB_TRY(matrix_new(h, maxg, maxg, &A));
    sc.own(A);
    void *d_ainv = nullptr;
    const int64_t nleaf = (maxg + LEAF - 1) / LEAF;
    AB_TRY(sc.alloc(static_cast<size_t>(nleaf) * LEAF * LEAF * sizeof(double), &d_ainv));
    for (int64_t g = 0; g < ngroups; ++g) {
      const int64_t o = offsets[g];
      const int64_t sz = offsets[g + 1] - o;
      if (sz == 0) {
        continue;
      }
      const MatView Cg = Cv.sub(o, 0);
      AB_TRY(gram_into(h, P_ff, dim, true, XF->d + o * dim, XF->ld, sz, nullptr, 0, 0, A->d, A->ld,
                       AB_GRAM_LOWER_ONLY));
      AB_TRY(add_diag(h, view(A), sz, static_cast<double *>(d_var) + o));
      AB_TRY(gemm(h, GEMM_TRANS_B | GEMM_LOWER, sz, sz, m, -1., Cg, Cg, 1., view(A)));
      add_diag_scalar_kernel<<<static_cast<unsigned>((sz + 255) / 256), 256, 0, h->stream>>>(
          A->d, A->ld, sz, measurement_nugget);
      AB_LAUNCHED(h);
      AB_TRY(potrf(h, view(A), sz, static_cast<double *>(d_ainv), h->d_flags));
      AB_TRY(trsm_left_lower(h, view(A), static_cast<double *>(d_ainv), sz, Cg, m + 1));
      AB_TRY(logdet_chol(h, view(A), sz, static_cast<double *>(d_logs) + o));
    }
  }
  AB_TRY(download_bytes(h, h->d_flags, sizeof(int), h->h_flags));
  // every rank must take the same sequence of collectives: agree on failure before the first one
  const bool local_bad = h->h_flags[0] != INT_MAX;
  if (dist_total(h, local_bad ? 1 : 0) > 0) {
    set_error(local_bad ? "a block of A = K_ff - Q_ff + noise is not positive definite"
                        : "a block of A = K_ff - Q_ff + noise is not positive definite on another rank");
    return AB_ERR_NOT_PD;
  }

  // ---- QR of C by CholQR2 (compute_sigma_qr, sparse_gp.hpp:368-375) ----------------------------
  // Only rank 0 carries the [I | 0] bottom block when the observations are sharded over ranks.
  const int64_t rows_used = n + (dist_rank(h) == 0 ? m : 0);
  void *d_g = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(m) * sizeof(double), &d_g));
  double *d_sc = h->d_scalars; // [0] yy  [1] zz  [2] log|A|  [3] log|L1|^2  [4] log|L2|^2
  AB_TRY(cholqr_pass(h, sc, Cv, rows_used, m, &fit->L1, static_cast<double *>(d_g), d_sc));
  AB_TRY(trsm_right_lower_T(h, view(fit->L1->m), fit->L1->dinv, m, Cv, rows_used));
  AB_TRY(cholqr_pass(h, sc, Cv, rows_used, m, &fit->L2, static_cast<double *>(d_g), d_sc));
  phase_end(h, PH_FACTOR);

  // ---- v = L_u^-T L1^-T L2^-T z,  z = Q^T y_aug = L2^-1 g  (sparse_gp.hpp:396-398) --------------
  phase_begin(h, PH_SOLVE);
  AB_TRY(matrix_new(h, m, 1, &fit->v));
  AB_CUDA(cudaMemcpyAsync(fit->v->d, d_g, static_cast<size_t>(m) * sizeof(double),
                          cudaMemcpyDeviceToDevice, h->stream));
  const MatView vv = view(fit->v);
  AB_TRY(trsm_left_lower(h, view(fit->L2->m), fit->L2->dinv, m, vv, 1));
  AB_TRY(dot(h, fit->v->d, fit->v->d, m, d_sc + 1));
  AB_TRY(trsm_left_lower_T(h, view(fit->L2->m), fit->L2->dinv, m, vv, 1));
  AB_TRY(trsm_left_lower_T(h, view(fit->L1->m), fit->L1->dinv, m, vv, 1));
  AB_TRY(trsm_left_lower_T(h, view(fit->Ku->m), fit->Ku->dinv, m, vv, 1));
  phase_end(h, PH_SOLVE);

  // ---- log-likelihood pieces (sparse_gp.hpp:539-603) -------------------------------------------
  phase_begin(h, PH_REDUCE);
  sum_kernel<<<1, RED_THREADS, 0, h->stream>>>(static_cast<double *>(d_logs), n, d_sc + 2);
  AB_LAUNCHED(h);
  AB_TRY(dist_allreduce_sum(h, d_sc + 2, 1));
  AB_TRY(logdet_chol(h, view(fit->L1->m), m, d_sc + 3));
  AB_TRY(logdet_chol(h, view(fit->L2->m), m, d_sc + 4));
  phase_end(h, PH_REDUCE);
  AB_TRY(download_bytes(h, d_sc, 5 * sizeof(double), h->h_scalars));
  const double yy = h->h_scalars[0], zz = h->h_scalars[1], log_det_a = h->h_scalars[2];
  const double log_det = log_det_a + h->h_scalars[3] + h->h_scalars[4];
  const double total_n = static_cast<double>(dist_total(h, n));
  fit->log_likelihood = -0.5 * (log_det + (yy - zz) + total_n * std::log(2 * M_PI));
  return AB_OK;
}

} // namespace
} // namespace ab

using namespace ab;

extern "C" {

int ab_sparse_fit2(ab_handle h, const ab_op *prog_ff, int nops_ff, const ab_op *prog_fu, int nops_fu,
                   const ab_op *prog_uu, int nops_uu, const double *feats, int64_t n, int dim,
                   const double *y, const double *yvar, const double *inducing, int64_t m,
                   const int64_t *indices, const int64_t *offsets, int64_t ngroups,
                   double measurement_nugget, double inducing_nugget, ab_sparse *out,
                   double *information, double *log_likelihood) {
  AB_REQUIRE(h != nullptr && out != nullptr && n >= 0 && m >= 1 && ngroups >= 0, "null / sizes");
  AB_REQUIRE(inducing != nullptr && offsets != nullptr && (n == 0 || (feats && y && indices)),
             "null inputs");
  AB_REQUIRE(offsets[0] == 0 && offsets[ngroups] == n, "group offsets must cover all observations");
  AB_REQUIRE(dim >= 1 && dim <= AB_MAX_DIM, "feature dimension");
  Lock lock(h);
  DevProg P_ff, P_fu, P_uu;
  AB_TRY(compile_program(prog_ff, nops_ff, &P_ff));
  AB_TRY(compile_program(prog_fu, nops_fu, &P_fu));
  AB_TRY(compile_program(prog_uu, nops_uu, &P_uu));
  timings_reset(h);
  auto *fit = new ab_sparse_fit_s();
  int s = sparse_fit_impl(h, P_ff, P_fu, P_uu, feats, n, dim, y, yvar, inducing, m, indices, offsets,
                          ngroups, measurement_nugget, inducing_nugget, fit);
  cudaEventRecord(h->ev_total_end, h->stream);
  if (s != AB_OK) {
    cudaStreamSynchronize(h->stream);
    free_sparse(h, fit);
    *out = nullptr;
    return s;
  }
  if (information != nullptr) {
    s = download(h, fit->v, 0, 0, m, 1, information);
  }
  if (log_likelihood != nullptr) {
    *log_likelihood = fit->log_likelihood;
  }
  *out = fit;
  return s;
}

int ab_sparse_fit(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
                  const double *y, const double *yvar, const double *inducing, int64_t m,
                  const int64_t *indices, const int64_t *offsets, int64_t ngroups,
                  double measurement_nugget, double inducing_nugget, ab_sparse *out,
                  double *information, double *log_likelihood) {
  return ab_sparse_fit2(h, prog, nops, prog, nops, prog, nops, feats, n, dim, y, yvar, inducing, m,
                        indices, offsets, ngroups, measurement_nugget, inducing_nugget, out,
                        information, log_likelihood);
}

int ab_sparse_free(ab_handle h, ab_sparse f) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  free_sparse(h, f);
  return AB_OK;
}

int ab_sparse_info(ab_sparse f, int64_t *m, double *log_likelihood) {
  AB_REQUIRE(f != nullptr, "null");
  if (m != nullptr) {
    *m = f->m;
  }
  if (log_likelihood != nullptr) {
    *log_likelihood = f->log_likelihood;
  }
  return AB_OK;
}

int ab_sparse_log_likelihood(ab_handle h, const ab_op *prog, int nops, const double *feats,
                             int64_t n, int dim, const double *y, const double *yvar,
                             const double *inducing, int64_t m, const int64_t *indices,
                             const int64_t *offsets, int64_t ngroups, double measurement_nugget,
                             double inducing_nugget, double *log_likelihood) {
  AB_REQUIRE(log_likelihood != nullptr, "null");
  ab_sparse f = nullptr;
  AB_TRY(ab_sparse_fit(h, prog, nops, feats, n, dim, y, yvar, inducing, m, indices, offsets,
                       ngroups, measurement_nugget, inducing_nugget, &f, nullptr, log_likelihood));
  return ab_sparse_free(h, f);
}

int ab_sparse_log_likelihood2(ab_handle h, const ab_op *prog_ff, int nops_ff, const ab_op *prog_fu,
                              int nops_fu, const ab_op *prog_uu, int nops_uu, const double *feats,
                              int64_t n, int dim, const double *y, const double *yvar,
                              const double *inducing, int64_t m, const int64_t *indices,
                              const int64_t *offsets, int64_t ngroups, double measurement_nugget,
                              double inducing_nugget, double *log_likelihood) {
  AB_REQUIRE(log_likelihood != nullptr, "null");
  ab_sparse f = nullptr;
  AB_TRY(ab_sparse_fit2(h, prog_ff, nops_ff, prog_fu, nops_fu, prog_uu, nops_uu, feats, n, dim, y,
                        yvar, inducing, m, indices, offsets, ngroups, measurement_nugget,
                        inducing_nugget, &f, nullptr, log_likelihood));
  return ab_sparse_free(h, f);
}

int ab_sparse_predict2(ab_handle h, ab_sparse f, const ab_op *prog, int nops,
                       const ab_op *prior_prog, int prior_nops, const double *test_feats, int64_t p,
                       int what, double *mean, double *var, double *cov) {
  AB_REQUIRE(h != nullptr && f != nullptr && p >= 0 && (p == 0 || (test_feats && mean)), "null");
  AB_REQUIRE(what == AB_PREDICT_MEAN || (what == AB_PREDICT_MARGINAL && var != nullptr) ||
                 (what == AB_PREDICT_JOINT && cov != nullptr),
             "prediction kind / outputs");
  if (p == 0) {
    return AB_OK;
  }
  Lock lock(h);
  DevProg P, PP; // cross k(u, test) and prior k(test, test) (sparse_gp.hpp:470, 491-495, 516)
  AB_TRY(compile_program(prog, nops, &P));
  AB_TRY(compile_program(prior_prog, prior_nops, &PP));
  Scope sc(h);
  timings_reset(h);
  const int64_t m = f->m;
  ab_matrix_s *T = nullptr, *cross = nullptr, *mu = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload_features(h, test_feats, p, f->dim, &T));
  sc.own(T);
  phase_end(h, PH_H2D);
  phase_begin(h, PH_PREDICT);
  // cross = K(u, test), m x p; mean = cross^T v  (sparse_gp.hpp:468-478)
  AB_TRY(gram_cross_device(h, P, f->u, T, &cross));
  sc.own(cross);
  AB_TRY(matrix_new(h, p, 1, &mu));
  sc.own(mu);
  AB_TRY(gemm(h, GEMM_TRANS_A, p, 1, m, 1., view(cross), view(f->v), 0., view(mu)));
  if (what != AB_PREDICT_MEAN) {
    // Q* = K_uu^-1/2 K_u*  ;  S* = R^-T P^T K_u* = L2^-1 L1^-1 Q*   (sparse_gp.hpp:481-536)
    ab_matrix_s *S = nullptr;
    AB_TRY(trsm_left_lower(h, view(f->Ku->m), f->Ku->dinv, m, view(cross), p));
    AB_TRY(matrix_new(h, m, p, &S));
    sc.own(S);
    AB_CUDA(cudaMemcpy2DAsync(S->d, S->ld * sizeof(double), cross->d, cross->ld * sizeof(double),
                              static_cast<size_t>(m) * sizeof(double), static_cast<size_t>(p),
                              cudaMemcpyDeviceToDevice, h->stream));
    AB_TRY(trsm_left_lower(h, view(f->L1->m), f->L1->dinv, m, view(S), p));
    AB_TRY(trsm_left_lower(h, view(f->L2->m), f->L2->dinv, m, view(S), p));
    if (what == AB_PREDICT_MARGINAL) {
      void *d_prior = nullptr, *d_q = nullptr, *d_s = nullptr;
      const size_t pb = static_cast<size_t>(p) * sizeof(double);
      AB_TRY(sc.alloc(pb, &d_prior));
      AB_TRY(sc.alloc(pb, &d_q));
      AB_TRY(sc.alloc(pb, &d_s));
      AB_TRY(gram_diag_device(h, PP, T, static_cast<double *>(d_prior)));
      AB_TRY(column_dots(h, view(cross), view(cross), m, p, static_cast<double *>(d_q)));
      AB_TRY(column_dots(h, view(S), view(S), m, p, static_cast<double *>(d_s)));
      phase_end(h, PH_PREDICT);
      std::vector<double> prior(p), q(p), s2(p);
      AB_TRY(download_bytes(h, d_prior, pb, prior.data()));
      AB_TRY(download_bytes(h, d_q, pb, q.data()));
      AB_TRY(download_bytes(h, d_s, pb, s2.data()));
      for (int64_t i = 0; i < p; ++i) {
        var[i] = prior[i] - q[i] + s2[i];
      }
    } else {
      ab_matrix_s *prior = nullptr;
      AB_TRY(gram_sym_device(h, PP, T, AB_GRAM_FULL, &prior));
      sc.own(prior);
      AB_TRY(gemm(h, GEMM_TRANS_A, p, p, m, -1., view(cross), view(cross), 1., view(prior)));
      AB_TRY(gemm(h, GEMM_TRANS_A, p, p, m, 1., view(S), view(S), 1., view(prior)));
      phase_end(h, PH_PREDICT);
      AB_TRY(download(h, prior, 0, 0, p, p, cov));
    }
  } else {
    phase_end(h, PH_PREDICT);
  }
  cudaEventRecord(h->ev_total_end, h->stream);
  return download(h, mu, 0, 0, p, 1, mean);
}

int ab_sparse_predict(ab_handle h, ab_sparse f, const ab_op *prog, int nops,
                      const double *test_feats, int64_t p, int what, double *mean, double *var,
                      double *cov) {
  return ab_sparse_predict2(h, f, prog, nops, prog, nops, test_feats, p, what, mean, var, cov);
}

// R of a thin QR of a dense host matrix (the QR concept of sparse_gp.hpp:72-89 for callers that build their
// own B): CholQR2 on the device — S = B^T B, S = L1 L1^T, B <- B L1^-T, once more, R = (L1 L2)^T.
int ab_qr_r(ab_handle h, const double *B, int64_t rows, int64_t cols, double *R) {
  AB_REQUIRE(h != nullptr && B != nullptr && R != nullptr && rows >= cols && cols >= 1, "null / shape");
  Lock lock(h);
  Scope sc(h);
  timings_reset(h);
  ab_matrix_s *Bd = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload(h, B, rows, cols, &Bd));
  sc.own(Bd);
  phase_end(h, PH_H2D);
  phase_begin(h, PH_FACTOR);
  ab_factor_s *L[2] = {nullptr, nullptr};
  int status = AB_OK;
  for (int pass = 0; pass < 2 && status == AB_OK; ++pass) {
    ab_matrix_s *S = nullptr;
    status = matrix_new(h, cols, cols, &S);
    if (status != AB_OK) {
      break;
    }
    status = gemm(h, GEMM_TRANS_A | GEMM_LOWER, cols, cols, rows, 1., view(Bd), view(Bd), 0., view(S));
    if (status != AB_OK) {
      matrix_delete(h, S);
      break;
    }
    status = factorize(h, S, &L[pass]); // consumes S
    if (status == AB_OK && pass == 0) {
      status = trsm_right_lower_T(h, view(L[0]->m), L[0]->dinv, cols, view(Bd), rows);
    }
  }
  phase_end(h, PH_FACTOR);
  if (status == AB_OK) {
    // R = (L1 L2)^T
    ab_matrix_s *a = nullptr, *b = nullptr, *c = nullptr;
    if ((status = matrix_new(h, cols, cols, &a)) == AB_OK) {
      sc.own(a);
    }
    if (status == AB_OK && (status = matrix_new(h, cols, cols, &b)) == AB_OK) {
      sc.own(b);
    }
    if (status == AB_OK && (status = matrix_new(h, cols, cols, &c)) == AB_OK) {
      sc.own(c);
    }
    if (status == AB_OK) {
      const dim3 g = grid2(cols, cols);
      tril_copy_kernel<<<g, 256, 0, h->stream>>>(L[0]->m->d, L[0]->m->ld, cols, a->d, a->ld);
      tril_copy_kernel<<<g, 256, 0, h->stream>>>(L[1]->m->d, L[1]->m->ld, cols, b->d, b->ld);
      h->launches += 2;
      status = gemm(h, 0u, cols, cols, cols, 1., view(a), view(b), 0., view(c));
      if (status == AB_OK) {
        transpose_kernel<<<g, 256, 0, h->stream>>>(c->d, c->ld, cols, a->d, a->ld);
        h->launches++;
        cudaEventRecord(h->ev_total_end, h->stream);
        status = download(h, a, 0, 0, cols, cols, R);
      }
    }
  } else if (status == AB_ERR_NOT_PD) {
    set_error("ab_qr_r: B^T B is numerically singular (CholQR2 needs cond(B) below ~1e7)");
  }
  delete_factor(h, L[0]);
  delete_factor(h, L[1]);
  return status;
}

int ab_sparse_export_R(ab_handle h, ab_sparse f, double *R) {
  AB_REQUIRE(h != nullptr && f != nullptr && R != nullptr, "null");
  Lock lock(h);
  Scope sc(h);
  const int64_t m = f->m;
  // R = (L_u L1 L2)^T
  ab_matrix_s *a = nullptr, *b = nullptr, *c = nullptr;
  AB_TRY(matrix_new(h, m, m, &a));
  sc.own(a);
  AB_TRY(matrix_new(h, m, m, &b));
  sc.own(b);
  AB_TRY(matrix_new(h, m, m, &c));
  sc.own(c);
  const dim3 g = grid2(m, m);
  tril_copy_kernel<<<g, 256, 0, h->stream>>>(f->Ku->m->d, f->Ku->m->ld, m, a->d, a->ld);
  AB_LAUNCHED(h);
  tril_copy_kernel<<<g, 256, 0, h->stream>>>(f->L1->m->d, f->L1->m->ld, m, b->d, b->ld);
  AB_LAUNCHED(h);
  AB_TRY(gemm(h, 0u, m, m, m, 1., view(a), view(b), 0., view(c)));
  tril_copy_kernel<<<g, 256, 0, h->stream>>>(f->L2->m->d, f->L2->m->ld, m, b->d, b->ld);
  AB_LAUNCHED(h);
  AB_TRY(gemm(h, 0u, m, m, m, 1., view(c), view(b), 0., view(a)));
  transpose_kernel<<<g, 256, 0, h->stream>>>(a->d, a->ld, m, c->d, c->ld);
  AB_LAUNCHED(h);
  return download(h, c, 0, 0, m, m, R);
}

} // extern "C"
