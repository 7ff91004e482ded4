// GP-level orchestration and the extern "C" surface declared in include/albatross_b200.h.
//
// Each entry point cites the reference routine it replaces in the header; this file only sequences
// device kernels (gram.cu, gemm.cu, linalg.cu) on the handle's stream and moves results to the host.
#include "internal.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

namespace ab {
namespace {

// W <- L^-1 (lower triangular, W's strict upper triangle must already be zero).
// T: workspace of at least ceil(n/2) x ceil(n/2) (+LEAF slack) doubles with leading dimension ldt.
int trtri_rec(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView W, MatView T) {
  if (n <= LEAF) {
    AB_CUDA(cudaMemcpy2DAsync(W.p, W.ld * sizeof(double), dinv, LEAF * sizeof(double),
                              n * sizeof(double), n, cudaMemcpyDeviceToDevice, h->stream));
    return AB_OK;
  }
  int64_t n1 = round_up((n + 1) / 2, LEAF);
  if (n1 >= n) {
    n1 = round_up(n, LEAF) - LEAF;
  }
  const int64_t n2 = n - n1;
  AB_TRY(trtri_rec(h, L, dinv, n1, W, T));
  AB_TRY(trtri_rec(h, L.sub(n1, n1), dinv + (n1 / LEAF) * LEAF * LEAF, n2, W.sub(n1, n1), T));
  // W21 = -W22 * (L21 * W11)
  AB_TRY(gemm(h, 0u, n2, n1, n1, 1., L.sub(n1, 0), W, 0., T));
  return gemm(h, 0u, n2, n1, n2, -1., W.sub(n1, n1), T, 0., W.sub(n1, 0));
}

int inverse_factor(ab_handle_s *h, Scope &sc, const ab_factor_s *f, ab_matrix_s **Wout) {
  const int64_t n = f->n;
  ab_matrix_s *W = nullptr;
  AB_TRY(matrix_new(h, n, n, &W));
  sc.own(W);
  AB_TRY(fill(h, view(W), n, n, 0.));
  const int64_t half = round_up((n + 1) / 2, LEAF) + LEAF;
  ab_matrix_s *T = nullptr;
  AB_TRY(matrix_new(h, half, half, &T));
  sc.own(T);
  AB_TRY(trtri_rec(h, view(f->m), f->dinv, n, view(W), view(T)));
  *Wout = W;
  return AB_OK;
}

__global__ void gather_cols_kernel(const double *W, int64_t ldw, int64_t row0, int64_t rows,
                                   const int64_t *idx, double *G, int64_t ldg) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (r < rows) {
    G[r + c * ldg] = W[row0 + r + idx[c] * ldw];
  }
}

__global__ void set_diag_one_kernel(double *G, int64_t ldg, int64_t k) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c < k) {
    G[c + c * ldg] = 1.;
  }
}

__global__ void gather_vec_kernel(const double *v, const int64_t *idx, int64_t k, double *out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < k) {
    out[i] = v[idx[i]];
  }
}

// mean[idx[i]] = y[idx[i]] - x[i]; optional var[idx[i]] = diag[i]
__global__ void heldout_scatter_kernel(const double *y, const double *x, const double *diag,
                                       const int64_t *idx, int64_t k, double *mean, double *var) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < k) {
    const int64_t t = idx[i];
    mean[t] = y[t] - x[i];
    if (var != nullptr) {
      var[t] = diag[i];
    }
  }
}

// Leave-one-out: a_i = (K^-1)_ii; variance 1/a_i, mean y_i - information_i / a_i
// (cross_validation_utils.hpp:172-197 with 1x1 blocks).  per-point score terms written to `terms`.
__global__ void loo_kernel(const double *a, const double *y, const double *info, int64_t n,
                           double *mean, double *var, double *terms) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const double variance = 1. / a[i];
    const double x = info[i] * variance;
    mean[i] = y[i] - x;
    var[i] = variance;
    // univariate NLL of deviation x under `variance` (stats/gaussian.hpp:19-23)
    terms[i] = x * x / (2. * variance) + 0.5 * log(2. * M_PI * variance);
  }
}

} // namespace
} // namespace ab

using namespace ab;

extern "C" {

// ---- Gram -------------------------------------------------------------------------------------

int ab_gram_sym_d(ab_handle h, const ab_op *prog, int nops, ab_matrix feats, uint32_t flags,
                  ab_matrix *out) {
  AB_REQUIRE(h != nullptr && feats != nullptr && out != nullptr, "null");
  Lock lock(h);
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  timings_reset(h);
  phase_begin(h, PH_GRAM);
  int s = gram_sym_device(h, P, feats, flags, out);
  phase_end(h, PH_GRAM);
  cudaEventRecord(h->ev_total_end, h->stream);
  return s;
}

int ab_gram_cross_d(ab_handle h, const ab_op *prog, int nops, ab_matrix fx, ab_matrix fy,
                    ab_matrix *out) {
  AB_REQUIRE(h != nullptr && fx != nullptr && fy != nullptr && out != nullptr, "null");
  Lock lock(h);
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  timings_reset(h);
  phase_begin(h, PH_GRAM);
  int s = gram_cross_device(h, P, fx, fy, out);
  phase_end(h, PH_GRAM);
  cudaEventRecord(h->ev_total_end, h->stream);
  return s;
}

int ab_gram_sym(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
                uint32_t flags, ab_matrix *out) {
  AB_REQUIRE(h != nullptr && out != nullptr && (feats != nullptr || n == 0) && n >= 0, "null");
  Lock lock(h);
  Scope sc(h);
  ab_matrix_s *F = nullptr;
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  int s = ab_gram_sym_d(h, prog, nops, F, flags, out);
  cudaStreamSynchronize(h->stream); // `feats` may be pageable: finish the copy before returning
  return s;
}

int ab_gram_cross(ab_handle h, const ab_op *prog, int nops, const double *fx, int64_t n,
                  const double *fy, int64_t m, int dim, ab_matrix *out) {
  AB_REQUIRE(h != nullptr && out != nullptr && n >= 0 && m >= 0, "null");
  Lock lock(h);
  Scope sc(h);
  ab_matrix_s *FX = nullptr, *FY = nullptr;
  AB_TRY(upload_features(h, fx, n, dim, &FX));
  sc.own(FX);
  AB_TRY(upload_features(h, fy, m, dim, &FY));
  sc.own(FY);
  int s = ab_gram_cross_d(h, prog, nops, FX, FY, out);
  cudaStreamSynchronize(h->stream);
  return s;
}

int ab_gram_diag(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
                 double *out) {
  AB_REQUIRE(h != nullptr && (out != nullptr || n == 0) && n >= 0, "null");
  Lock lock(h);
  Scope sc(h);
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  ab_matrix_s *F = nullptr;
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  void *d = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(n < 1 ? 1 : n) * sizeof(double), &d));
  AB_TRY(gram_diag_device(h, P, F, static_cast<double *>(d)));
  return download_bytes(h, d, static_cast<size_t>(n) * sizeof(double), out);
}

int ab_matrix_add_diag(ab_handle h, ab_matrix m, const double *d) {
  AB_REQUIRE(h != nullptr && m != nullptr && d != nullptr && m->rows == m->cols, "null/shape");
  Lock lock(h);
  Scope sc(h);
  void *dd = nullptr;
  AB_TRY(upload_bytes(h, sc, d, static_cast<size_t>(m->rows) * sizeof(double), &dd));
  AB_TRY(add_diag(h, view(m), m->rows, static_cast<double *>(dd)));
  AB_CUDA(cudaStreamSynchronize(h->stream));
  return AB_OK;
}

// ---- factor -----------------------------------------------------------------------------------

int ab_potrf(ab_handle h, ab_matrix m, ab_factor *out) {
  AB_REQUIRE(h != nullptr && m != nullptr && out != nullptr, "null");
  Lock lock(h);
  timings_reset(h);
  phase_begin(h, PH_FACTOR);
  int s = factorize(h, m, out);
  phase_end(h, PH_FACTOR);
  cudaEventRecord(h->ev_total_end, h->stream);
  return s;
}

int ab_factor_free(ab_handle h, ab_factor f) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  delete_factor(h, f);
  return AB_OK;
}

int ab_factor_rows(ab_factor f, int64_t *n) {
  AB_REQUIRE(f != nullptr && n != nullptr, "null");
  *n = f->n;
  return AB_OK;
}

int ab_factor_info(ab_factor f, int64_t *first_bad_pivot) {
  AB_REQUIRE(f != nullptr && first_bad_pivot != nullptr, "null");
  *first_bad_pivot = f->bad_pivot;
  return AB_OK;
}

static int solve_impl(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out,
                      bool sqrt_only) {
  AB_REQUIRE(h != nullptr && (rhs != nullptr || nrhs == 0) && (out != nullptr || nrhs == 0) &&
                 nrhs >= 0,
             "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  if (nrhs == 0 || f->n == 0) {
    return AB_OK;
  }
  Scope sc(h);
  timings_reset(h);
  ab_matrix_s *X = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload(h, rhs, f->n, nrhs, &X));
  phase_end(h, PH_H2D);
  sc.own(X);
  phase_begin(h, PH_SOLVE);
  AB_TRY(trsm_left_lower(h, view(f->m), f->dinv, f->n, view(X), nrhs));
  if (!sqrt_only) {
    AB_TRY(trsm_left_lower_T(h, view(f->m), f->dinv, f->n, view(X), nrhs));
  }
  phase_end(h, PH_SOLVE);
  cudaEventRecord(h->ev_total_end, h->stream);
  return download(h, X, 0, 0, f->n, nrhs, out);
}

int ab_factor_solve(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out) {
  return solve_impl(h, f, rhs, nrhs, out, false);
}

int ab_factor_sqrt_solve(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out) {
  return solve_impl(h, f, rhs, nrhs, out, true);
}

int ab_factor_logdet(ab_handle h, ab_factor f, double *out) {
  AB_REQUIRE(h != nullptr && out != nullptr, "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  if (f->n == 0) {
    *out = 0.;
    return AB_OK;
  }
  AB_TRY(logdet_chol(h, view(f->m), f->n, h->d_scalars));
  AB_TRY(download_bytes(h, h->d_scalars, sizeof(double), h->h_scalars));
  *out = h->h_scalars[0];
  return AB_OK;
}

// nll from an existing factor and a device-resident deviation vector (n x 1 matrix).
static int nll_device(ab_handle h, ab_factor f, ab_matrix_s *dev_copy, double *out) {
  // quad = |L^-1 dev|^2 = dev^T K^-1 dev ; logdet = sum 2 log L_ii
  phase_begin(h, PH_SOLVE);
  AB_TRY(trsm_left_lower(h, view(f->m), f->dinv, f->n, view(dev_copy), 1));
  phase_end(h, PH_SOLVE);
  phase_begin(h, PH_REDUCE);
  AB_TRY(dot(h, dev_copy->d, dev_copy->d, f->n, h->d_scalars + 1));
  AB_TRY(logdet_chol(h, view(f->m), f->n, h->d_scalars));
  phase_end(h, PH_REDUCE);
  AB_TRY(download_bytes(h, h->d_scalars, 2 * sizeof(double), h->h_scalars));
  const double log_det = h->h_scalars[0];
  const double mahalanobis = h->h_scalars[1];
  *out = 0.5 * (log_det + mahalanobis + static_cast<double>(f->n) * std::log(2 * M_PI));
  return AB_OK;
}

int ab_factor_nll(ab_handle h, ab_factor f, const double *deviation, double *out) {
  AB_REQUIRE(h != nullptr && deviation != nullptr && out != nullptr, "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  Scope sc(h);
  timings_reset(h);
  ab_matrix_s *d = nullptr;
  AB_TRY(upload(h, deviation, f->n, 1, &d));
  sc.own(d);
  int s = nll_device(h, f, d, out);
  cudaEventRecord(h->ev_total_end, h->stream);
  return s;
}

int ab_factor_inverse_diagonal(ab_handle h, ab_factor f, double *out) {
  AB_REQUIRE(h != nullptr && out != nullptr, "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  if (f->n == 0) {
    return AB_OK;
  }
  Scope sc(h);
  ab_matrix_s *W = nullptr;
  AB_TRY(inverse_factor(h, sc, f, &W));
  void *d = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(f->n) * sizeof(double), &d));
  AB_TRY(column_dots(h, view(W), view(W), f->n, f->n, static_cast<double *>(d)));
  return download_bytes(h, d, static_cast<size_t>(f->n) * sizeof(double), out);
}

__global__ void unit_columns_kernel(double *G, int64_t ldg, int64_t row0, const int64_t *idx,
                                    int64_t k) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c < k) {
    G[idx[c] - row0 + c * ldg] = 1.;
  }
}

// (K^-1)_gg = W[:, g]^T W[:, g] written to A (k x k); W = L^-1.  With W == nullptr the k columns
// of L^-1 are produced on the fly by a triangular solve on unit vectors (rows above the smallest
// index are structurally zero and skipped): the form used when folds are sharded over GPUs and no
// rank wants the whole inverse factor.
static int inverse_block_device(ab_handle h, const ab_factor_s *f, const ab_matrix_s *W,
                                const int64_t *d_idx, const int64_t *h_idx, int64_t k,
                                ab_matrix_s *G, MatView A) {
  const int64_t n = f->n;
  int64_t row0 = n;
  for (int64_t i = 0; i < k; ++i) {
    AB_REQUIRE(h_idx[i] >= 0 && h_idx[i] < n, "group index out of range");
    row0 = std::min(row0, h_idx[i]);
  }
  if (W != nullptr) {
    row0 = row0 / 2 * 2; // keep 16-byte alignment of the gathered panel rows
    const int64_t rows = n - row0;
    const dim3 grid(static_cast<unsigned>((rows + 255) / 256), static_cast<unsigned>(k));
    gather_cols_kernel<<<grid, 256, 0, h->stream>>>(W->d, W->ld, row0, rows, d_idx, G->d, G->ld);
    AB_LAUNCHED(h);
    return gemm(h, GEMM_TRANS_A, k, k, rows, 1., view(G), view(G), 0., A);
  }
  row0 = row0 / LEAF * LEAF; // start the solve on a leaf boundary so that dinv blocks line up
  const int64_t rows = n - row0;
  AB_TRY(fill(h, view(G), rows, k, 0.));
  unit_columns_kernel<<<static_cast<unsigned>((k + 255) / 256), 256, 0, h->stream>>>(
      G->d, G->ld, row0, d_idx, k);
  AB_LAUNCHED(h);
  AB_TRY(trsm_left_lower(h, view(f->m).sub(row0, row0), f->dinv + (row0 / LEAF) * LEAF * LEAF,
                         rows, view(G), k));
  return gemm(h, GEMM_TRANS_A, k, k, rows, 1., view(G), view(G), 0., A);
}

int ab_factor_inverse_blocks(ab_handle h, ab_factor f, const int64_t *indices,
                             const int64_t *offsets, int64_t ngroups, double *out) {
  AB_REQUIRE(h != nullptr && indices != nullptr && offsets != nullptr && out != nullptr &&
                 ngroups >= 0,
             "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  if (ngroups == 0) {
    return AB_OK;
  }
  Scope sc(h);
  const int64_t total = offsets[ngroups];
  int64_t maxg = 0;
  for (int64_t g = 0; g < ngroups; ++g) {
    maxg = std::max(maxg, offsets[g + 1] - offsets[g]);
  }
  ab_matrix_s *W = nullptr;
  AB_TRY(inverse_factor(h, sc, f, &W));
  void *d_idx = nullptr;
  AB_TRY(upload_bytes(h, sc, indices, static_cast<size_t>(total) * sizeof(int64_t), &d_idx));
  ab_matrix_s *G = nullptr, *A = nullptr;
  AB_TRY(matrix_new(h, f->n, maxg, &G));
  sc.own(G);
  AB_TRY(matrix_new(h, maxg, maxg, &A));
  sc.own(A);
  double *cursor = out;
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t k = offsets[g + 1] - offsets[g];
    AB_TRY(inverse_block_device(h, f, W, static_cast<int64_t *>(d_idx) + offsets[g],
                                indices + offsets[g], k, G, view(A)));
    AB_TRY(download(h, A, 0, 0, k, k, cursor));
    cursor += k * k;
  }
  return AB_OK;
}

// Columns [c0, c0 + w) of a dense n x n host view of the factor, written to a packed n x w panel.
//   mode 0  Eigen's packed LDLT (serializable_ldlt.hpp / cereal layout): strict lower = unit L, diagonal
//           = D, strict upper = the transpose (Eigen's matrixLDLT is read through triangular views)
//   mode 1  sqrt_transpose(): D^1/2 (P^T L)^T = L_chol^T, upper triangular, zeros below
__global__ void export_panel_kernel(const double *L, int64_t ld, int64_t n, int64_t c0, int64_t w,
                                    int mode, double *out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t jl = blockIdx.y;
  const int64_t j = c0 + jl;
  if (i >= n || jl >= w) {
    return;
  }
  double v;
  if (mode == 0) {
    const double ljj = L[j + j * ld];
    if (i > j) {
      v = L[i + j * ld] / ljj;
    } else if (i == j) {
      v = ljj * ljj;
    } else {
      v = L[j + i * ld] / L[i + i * ld];
    }
  } else {
    v = i <= j ? L[j + i * ld] : 0.;
  }
  out[i + jl * n] = v;
}

// Streams the n x n host matrix in column panels through a bounded device buffer (no n^2 scratch: the
// headline size n = 65 536 is 34 GB of factor already).
static int export_dense(ab_handle h, ab_factor f, int mode, double *host) {
  const int64_t n = f->n;
  int64_t w = std::max<int64_t>(1, (int64_t(256) << 20) / (n * static_cast<int64_t>(sizeof(double))));
  if (const char *e = std::getenv("AB_EXPORT_PANEL_COLS")) { // test hook: force several panels
    w = std::max<int64_t>(1, std::atoll(e));
  }
  w = std::min<int64_t>(std::min<int64_t>(w, n), 65535);
  Scope sc(h);
  void *d = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(n) * w * sizeof(double), &d));
  for (int64_t c0 = 0; c0 < n; c0 += w) {
    const int64_t wc = std::min(w, n - c0);
    const dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(wc));
    export_panel_kernel<<<grid, 256, 0, h->stream>>>(f->m->d, f->m->ld, n, c0, wc, mode,
                                                     static_cast<double *>(d));
    AB_LAUNCHED(h);
    AB_TRY(download_bytes(h, d, static_cast<size_t>(n) * wc * sizeof(double), host + c0 * n));
  }
  return AB_OK;
}

int ab_factor_export_packed(ab_handle h, ab_factor f, double *LD, int64_t *transpositions) {
  AB_REQUIRE(h != nullptr && LD != nullptr, "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  const int64_t n = f->n;
  if (n == 0) {
    return AB_OK;
  }
  AB_TRY(export_dense(h, f, 0, LD));
  if (transpositions != nullptr) {
    std::iota(transpositions, transpositions + n, int64_t(0));
  }
  return AB_OK;
}

static __global__ void leaf_inverse_kernel(const double *L, int64_t ld, int64_t n, double *dinv);

// M(i, j) = unit-L(i, j) * sqrt(D_j) (lower triangle, zeros above) from Eigen's packed LDLT; flags a
// non-positive D_j.
static __global__ void packed_to_chol_kernel(const double *LD, int64_t n, double *M, int64_t ld, int *d_bad) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= n) {
    return;
  }
  const double d = LD[j + j * n];
  if (i == j && !(d > 0.)) {
    atomicMin(d_bad, static_cast<int>(j));
  }
  const double s = sqrt(d);
  M[i + j * ld] = i > j ? LD[i + j * n] * s : (i == j ? s : 0.);
}

// K(a, b) = S(max(i, j), min(i, j)) with i = inv[a], j = inv[b], for a >= b: K = P^T S P from the lower
// triangle of S, where (P x)_i = x_{perm(i)} and inv = perm^-1.
static __global__ void unpermute_kernel(const double *S, int64_t lds, const int64_t *inv, int64_t n, double *K,
                                        int64_t ldk) {
  const int64_t a = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t b = blockIdx.y;
  if (a >= n || a < b) {
    return;
  }
  const int64_t i = inv[a], j = inv[b];
  K[a + b * ldk] = i >= j ? S[i + j * lds] : S[j + i * lds];
}

/*
 * The reverse of ab_factor_export_packed: a device factor from Eigen::SerializableLDLT's packed form (what
 * cereal writes, src/cereal/serializable_ldlt.hpp:18-32: lower triangle = unit L below the diagonal and D
 * on it, plus the transpositions) so that fits serialised by the reference can be loaded onto the GPU.
 * Identity transpositions (every file written from a device fit): L_chol = L D^1/2 directly, O(n^2).
 * Otherwise K = P^T L D L^T P is rebuilt on the device (one DSYRK) and factored without pivoting.
 */
int ab_factor_import_packed(ab_handle h, const double *LD, const int64_t *transpositions, int64_t n,
                            ab_factor *out) {
  AB_REQUIRE(h != nullptr && LD != nullptr && out != nullptr && n >= 1, "null / size");
  AB_REQUIRE(n <= 65535, "ab_factor_import_packed takes a dense host matrix: n <= 65535");
  Lock lock(h);
  Scope sc(h);
  timings_reset(h);
  *out = nullptr;
  // the permutation the transpositions compose to, and whether it is the identity
  std::vector<int64_t> perm(static_cast<size_t>(n));
  std::iota(perm.begin(), perm.end(), int64_t(0));
  bool identity = true;
  if (transpositions != nullptr) {
    for (int64_t i = 0; i < n; ++i) {
      const int64_t j = transpositions[i];
      AB_REQUIRE(j >= 0 && j < n, "transposition out of range");
      if (j != i) {
        std::swap(perm[static_cast<size_t>(i)], perm[static_cast<size_t>(j)]);
        identity = false;
      }
    }
  }
  ab_matrix_s *M = nullptr;
  void *d_ld = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload_bytes(h, sc, LD, static_cast<size_t>(n) * n * sizeof(double), &d_ld));
  phase_end(h, PH_H2D);
  AB_TRY(matrix_new(h, n, n, &M));
  h->h_flags[0] = INT_MAX;
  AB_CUDA(cudaMemcpyAsync(h->d_flags, h->h_flags, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  phase_begin(h, PH_FACTOR);
  {
    const dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(n));
    packed_to_chol_kernel<<<grid, 256, 0, h->stream>>>(static_cast<double *>(d_ld), n, M->d, M->ld, h->d_flags);
    AB_LAUNCHED(h);
  }
  int status = download_bytes(h, h->d_flags, sizeof(int), h->h_flags);
  if (status == AB_OK && h->h_flags[0] != INT_MAX) {
    set_error("imported LDLT has a non-positive pivot D[%d]: not positive definite", h->h_flags[0]);
    status = AB_ERR_NOT_PD;
  }
  if (status != AB_OK) {
    matrix_delete(h, M);
    return status;
  }
  ab_factor_s *f = nullptr;
  if (identity) {
    status = new_factor(h, M, &f);
    if (status != AB_OK) {
      matrix_delete(h, M);
      return status;
    }
    leaf_inverse_kernel<<<static_cast<unsigned>((n + LEAF - 1) / LEAF), LEAF, 0, h->stream>>>(M->d, M->ld, n, f->dinv);
    h->launches++;
    f->bad_pivot = -1;
  } else {
    sc.own(M);
    ab_matrix_s *S = nullptr, *K = nullptr;
    AB_TRY(matrix_new(h, n, n, &S));
    sc.own(S);
    AB_TRY(gemm(h, GEMM_TRANS_B | GEMM_LOWER, n, n, n, 1., view(M), view(M), 0., view(S)));
    std::vector<int64_t> inv(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
      inv[static_cast<size_t>(perm[static_cast<size_t>(i)])] = i;
    }
    void *d_inv = nullptr;
    AB_TRY(upload_bytes(h, sc, inv.data(), inv.size() * sizeof(int64_t), &d_inv));
    AB_TRY(matrix_new(h, n, n, &K));
    const dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(n));
    unpermute_kernel<<<grid, 256, 0, h->stream>>>(S->d, S->ld, static_cast<int64_t *>(d_inv), n, K->d, K->ld);
    h->launches++;
    AB_CUDA(cudaStreamSynchronize(h->stream)); // `inv` is a stack temporary
    status = factorize(h, K, &f);             // consumes K
    if (status != AB_OK) {
      delete_factor(h, f);
      return status;
    }
  }
  phase_end(h, PH_FACTOR);
  cudaEventRecord(h->ev_total_end, h->stream);
  AB_CUDA(cudaStreamSynchronize(h->stream));
  *out = f;
  return AB_OK;
}

int ab_factor_sqrt_transpose(ab_handle h, ab_factor f, double *out) {
  AB_REQUIRE(h != nullptr && out != nullptr, "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  return f->n == 0 ? AB_OK : export_dense(h, f, 1, out);
}

__global__ void diag_kernel(const double *L, int64_t ld, int64_t n, double *out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    out[i] = L[i + i * ld];
  }
}

int ab_factor_diagonal_sqrt(ab_handle h, ab_factor f, double *out) {
  AB_REQUIRE(h != nullptr && out != nullptr, "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  const int64_t n = f->n;
  if (n == 0) {
    return AB_OK;
  }
  Scope sc(h);
  void *d = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(n) * sizeof(double), &d));
  diag_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(
      f->m->d, f->m->ld, n, static_cast<double *>(d));
  AB_LAUNCHED(h);
  return download_bytes(h, d, static_cast<size_t>(n) * sizeof(double), out);
}

int ab_factor_sqrt_product(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs, double *out) {
  AB_REQUIRE(h != nullptr && (rhs != nullptr || nrhs == 0) && (out != nullptr || nrhs == 0) &&
                 nrhs >= 0,
             "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  if (nrhs == 0 || f->n == 0) {
    return AB_OK;
  }
  Scope sc(h);
  timings_reset(h);
  ab_matrix_s *X = nullptr, *Y = nullptr;
  AB_TRY(upload(h, rhs, f->n, nrhs, &X));
  sc.own(X);
  AB_TRY(matrix_new(h, f->n, nrhs, &Y));
  sc.own(Y);
  phase_begin(h, PH_SOLVE);
  AB_TRY(trmm_left_lower(h, view(f->m), f->n, false, view(X), view(Y), nrhs));
  phase_end(h, PH_SOLVE);
  cudaEventRecord(h->ev_total_end, h->stream);
  return download(h, Y, 0, 0, f->n, nrhs, out);
}

int ab_factor_sqrt_transpose_solve(ab_handle h, ab_factor f, const double *rhs, int64_t nrhs,
                                   double *out) {
  AB_REQUIRE(h != nullptr && (rhs != nullptr || nrhs == 0) && (out != nullptr || nrhs == 0) &&
                 nrhs >= 0,
             "null");
  Lock lock(h);
  AB_TRY(require_usable(f));
  if (nrhs == 0 || f->n == 0) {
    return AB_OK;
  }
  Scope sc(h);
  timings_reset(h);
  ab_matrix_s *X = nullptr;
  AB_TRY(upload(h, rhs, f->n, nrhs, &X));
  sc.own(X);
  phase_begin(h, PH_SOLVE);
  AB_TRY(trsm_left_lower_T(h, view(f->m), f->dinv, f->n, view(X), nrhs));
  phase_end(h, PH_SOLVE);
  cudaEventRecord(h->ev_total_end, h->stream);
  return download(h, X, 0, 0, f->n, nrhs, out);
}

// ---- exact GP ---------------------------------------------------------------------------------

// K (+diag) -> factor.  feats: dim x n device matrix.  Consumes nothing; returns a new factor.
static int build_and_factor(ab_handle h, const DevProg &P, const ab_matrix_s *feats,
                            const double *d_yvar, ab_factor_s **out) {
  ab_matrix_s *K = nullptr;
  phase_begin(h, PH_GRAM);
  AB_TRY(gram_sym_device(h, P, feats, AB_GRAM_LOWER_ONLY, &K));
  if (d_yvar != nullptr) {
    int s = add_diag(h, view(K), K->rows, d_yvar);
    if (s != AB_OK) {
      matrix_delete(h, K);
      return s;
    }
  }
  phase_end(h, PH_GRAM);
  phase_begin(h, PH_FACTOR);
  int s = factorize(h, K, out); // a NaN anywhere in K surfaces as a NaN pivot -> AB_ERR_NOT_PD
  phase_end(h, PH_FACTOR);
  return s;
}

int ab_gp_fit_d(ab_handle h, const ab_op *prog, int nops, ab_matrix feats, ab_matrix y,
                ab_matrix yvar, ab_factor *factor, ab_matrix *information) {
  AB_REQUIRE(h != nullptr && feats != nullptr && y != nullptr && factor != nullptr, "null");
  AB_REQUIRE(y->rows == feats->cols && y->cols == 1, "targets shape");
  AB_REQUIRE(yvar == nullptr || (yvar->rows == feats->cols && yvar->cols == 1), "variance shape");
  Lock lock(h);
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  timings_reset(h);
  ab_factor_s *f = nullptr;
  int s = build_and_factor(h, P, feats, yvar != nullptr ? yvar->d : nullptr, &f);
  if (s != AB_OK) {
    *factor = f; // non-null only for AB_ERR_NOT_PD
    cudaEventRecord(h->ev_total_end, h->stream);
    return s;
  }
  *factor = f;
  if (information != nullptr) {
    ab_matrix_s *x = nullptr;
    AB_TRY(matrix_new(h, f->n, 1, &x));
    phase_begin(h, PH_SOLVE);
    cudaMemcpyAsync(x->d, y->d, static_cast<size_t>(f->n) * sizeof(double),
                    cudaMemcpyDeviceToDevice, h->stream);
    s = trsm_left_lower(h, view(f->m), f->dinv, f->n, view(x), 1);
    if (s == AB_OK) {
      s = trsm_left_lower_T(h, view(f->m), f->dinv, f->n, view(x), 1);
    }
    phase_end(h, PH_SOLVE);
    if (s != AB_OK) {
      matrix_delete(h, x);
      return s;
    }
    *information = x;
  }
  cudaEventRecord(h->ev_total_end, h->stream);
  return AB_OK;
}

int ab_gp_fit(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
              const double *y, const double *yvar, ab_factor *factor, double *information) {
  AB_REQUIRE(h != nullptr && factor != nullptr && n >= 0 && (n == 0 || (feats && y)), "null");
  Lock lock(h);
  Scope sc(h);
  *factor = nullptr;
  ab_matrix_s *F = nullptr, *Y = nullptr, *V = nullptr, *info = nullptr;
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  AB_TRY(upload(h, y, n, 1, &Y));
  sc.own(Y);
  if (yvar != nullptr) {
    AB_TRY(upload(h, yvar, n, 1, &V));
    sc.own(V);
  }
  int s = ab_gp_fit_d(h, prog, nops, F, Y, V, factor, information != nullptr ? &info : nullptr);
  if (s != AB_OK) {
    cudaStreamSynchronize(h->stream);
    return s;
  }
  if (info != nullptr) {
    sc.own(info);
    AB_TRY(download(h, info, 0, 0, n, 1, information));
  }
  AB_CUDA(cudaStreamSynchronize(h->stream));
  return AB_OK;
}

int ab_gp_nll_d(ab_handle h, const ab_op *prog, int nops, ab_matrix feats, ab_matrix y,
                double *nll) {
  AB_REQUIRE(h != nullptr && feats != nullptr && y != nullptr && nll != nullptr, "null");
  AB_REQUIRE(y->rows == feats->cols && y->cols == 1, "targets shape");
  Lock lock(h);
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  Scope sc(h);
  timings_reset(h);
  const int64_t n = feats->cols;
  if (n == 0) {
    *nll = 0.;
    return AB_OK;
  }
  ab_factor_s *f = nullptr;
  int s = build_and_factor(h, P, feats, nullptr, &f);
  if (s != AB_OK) {
    delete_factor(h, f);
    return s;
  }
  ab_matrix_s *d = nullptr;
  s = matrix_new(h, n, 1, &d);
  if (s == AB_OK) {
    sc.own(d);
    cudaMemcpyAsync(d->d, y->d, static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToDevice,
                    h->stream);
    s = nll_device(h, f, d, nll);
  }
  cudaEventRecord(h->ev_total_end, h->stream);
  delete_factor(h, f);
  return s;
}

int ab_gp_nll(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
              const double *y, double *nll) {
  AB_REQUIRE(h != nullptr && nll != nullptr && n >= 0 && (n == 0 || (feats && y)), "null");
  Lock lock(h);
  Scope sc(h);
  ab_matrix_s *F = nullptr, *Y = nullptr;
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  AB_TRY(upload(h, y, n, 1, &Y));
  sc.own(Y);
  int s = ab_gp_nll_d(h, prog, nops, F, Y, nll);
  cudaStreamSynchronize(h->stream);
  return s;
}

int ab_gp_fit_nll(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n,
                  int dim, const double *y, ab_factor *factor, double *information, double *nll) {
  AB_REQUIRE(h != nullptr && factor != nullptr && nll != nullptr, "null");
  Lock lock(h);
  AB_TRY(ab_gp_fit(h, prog, nops, feats, n, dim, y, nullptr, factor, information));
  // dev^T K^-1 dev = y . information when information is available; use the factor directly.
  return ab_factor_nll(h, *factor, y, nll);
}

static int gp_predict_impl(ab_handle h, ab_factor f, const ab_op *prog, int nops,
                           const ab_op *prior_prog, int prior_nops, const double *train_feats,
                           int64_t n, int dim, const double *information, const double *test_feats,
                           int64_t p, int what, double *mean, double *var, double *cov) {
  AB_REQUIRE(h != nullptr && train_feats != nullptr && information != nullptr && p >= 0 &&
                 (p == 0 || (test_feats != nullptr && mean != nullptr)),
             "null");
  AB_REQUIRE(what == AB_PREDICT_MEAN || (what == AB_PREDICT_MARGINAL && var != nullptr) ||
                 (what == AB_PREDICT_JOINT && cov != nullptr),
             "prediction kind / outputs");
  Lock lock(h);
  AB_TRY(require_usable(f));
  AB_REQUIRE(f->n == n, "factor size differs from training set");
  if (p == 0) {
    return AB_OK;
  }
  DevProg P, PP;
  AB_TRY(compile_program(prog, nops, &P));
  AB_TRY(compile_program(prior_prog, prior_nops, &PP));
  Scope sc(h);
  timings_reset(h);
  ab_matrix_s *FX = nullptr, *FT = nullptr, *info = nullptr, *cross = nullptr, *m = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload_features(h, train_feats, n, dim, &FX));
  sc.own(FX);
  AB_TRY(upload_features(h, test_feats, p, dim, &FT));
  sc.own(FT);
  AB_TRY(upload(h, information, n, 1, &info));
  sc.own(info);
  phase_end(h, PH_H2D);
  phase_begin(h, PH_PREDICT);
  // cross_cov = K(train, test), n x p  (gp.hpp:317-318)
  AB_TRY(gram_cross_device(h, P, FX, FT, &cross));
  sc.own(cross);
  // mean = cross^T information  (gp.hpp:82-85)
  AB_TRY(matrix_new(h, p, 1, &m));
  sc.own(m);
  AB_TRY(gemm(h, GEMM_TRANS_A, p, 1, n, 1., view(cross), view(info), 0., view(m)));
  if (what == AB_PREDICT_MARGINAL) {
    // var = k(x*,x*) - colsum((L^-1 cross)^2)  (gp.hpp:87-101)
    void *d_prior = nullptr, *d_expl = nullptr;
    AB_TRY(sc.alloc(static_cast<size_t>(p) * sizeof(double), &d_prior));
    AB_TRY(sc.alloc(static_cast<size_t>(p) * sizeof(double), &d_expl));
    AB_TRY(gram_diag_device(h, PP, FT, static_cast<double *>(d_prior)));
    AB_TRY(trsm_left_lower(h, view(f->m), f->dinv, n, view(cross), p));
    AB_TRY(column_dots(h, view(cross), view(cross), n, p, static_cast<double *>(d_expl)));
    phase_end(h, PH_PREDICT);
    std::vector<double> prior(p), expl(p);
    AB_TRY(download_bytes(h, d_prior, static_cast<size_t>(p) * sizeof(double), prior.data()));
    AB_TRY(download_bytes(h, d_expl, static_cast<size_t>(p) * sizeof(double), expl.data()));
    for (int64_t i = 0; i < p; ++i) {
      var[i] = prior[i] - expl[i];
    }
  } else if (what == AB_PREDICT_JOINT) {
    // cov = K(test,test) - (L^-1 cross)^T (L^-1 cross)  (gp.hpp:103-113)
    ab_matrix_s *prior = nullptr;
    AB_TRY(gram_sym_device(h, PP, FT, AB_GRAM_FULL, &prior));
    sc.own(prior);
    AB_TRY(trsm_left_lower(h, view(f->m), f->dinv, n, view(cross), p));
    AB_TRY(gemm(h, GEMM_TRANS_A, p, p, n, -1., view(cross), view(cross), 1., view(prior)));
    phase_end(h, PH_PREDICT);
    AB_TRY(download(h, prior, 0, 0, p, p, cov));
  } else {
    phase_end(h, PH_PREDICT);
  }
  cudaEventRecord(h->ev_total_end, h->stream);
  return download(h, m, 0, 0, p, 1, mean);
}

int ab_gp_predict(ab_handle h, ab_factor f, const ab_op *prog, int nops, const double *train_feats,
                  int64_t n, int dim, const double *information, const double *test_feats,
                  int64_t p, int what, double *mean, double *var, double *cov) {
  return gp_predict_impl(h, f, prog, nops, prog, nops, train_feats, n, dim, information,
                         test_feats, p, what, mean, var, cov);
}

int ab_gp_predict2(ab_handle h, ab_factor f, const ab_op *cross_prog, int cross_nops,
                   const ab_op *prior_prog, int prior_nops, const double *train_feats, int64_t n,
                   int dim, const double *information, const double *test_feats, int64_t p,
                   int what, double *mean, double *var, double *cov) {
  return gp_predict_impl(h, f, cross_prog, cross_nops, prior_prog, prior_nops, train_feats, n, dim,
                         information, test_feats, p, what, mean, var, cov);
}

// ---- incremental update -------------------------------------------------------------------

// dinv[leaf] = (lower-triangular LEAF x LEAF diagonal block of L at row/column leaf * LEAF)^-1, identity
// padded for a ragged last leaf; one CTA per leaf, thread c solves column c by forward substitution.
static __global__ void __launch_bounds__(LEAF)
leaf_inverse_kernel(const double *L, int64_t ld, int64_t n, double *dinv) {
  __shared__ double s[LEAF][LEAF + 1];
  const int64_t o = static_cast<int64_t>(blockIdx.x) * LEAF;
  const int nb = static_cast<int>(n - o < LEAF ? n - o : LEAF);
  const int c = threadIdx.x;
  for (int r = 0; r < LEAF; ++r) {
    double v = (r == c) ? 1. : 0.;
    if (r < nb && c < nb && r >= c) {
      v = L[(o + r) + (o + c) * ld];
    }
    s[r][c] = v;
  }
  __syncthreads();
  double *out = dinv + static_cast<int64_t>(blockIdx.x) * LEAF * LEAF;
  double col[LEAF]; // column c of the inverse
  for (int r = 0; r < LEAF; ++r) {
    double acc = (r == c) ? 1. : 0.;
    for (int t = c; t < r; ++t) {
      acc = fma(-s[r][t], col[t], acc);
    }
    col[r] = r >= c ? acc / s[r][r] : 0.;
    out[r + c * LEAF] = col[r];
  }
}

// dst(i, j) = src(j, i): rows x cols of dst
static __global__ void transpose_into_kernel(const double *src, int64_t lds, double *dst, int64_t ldd, int64_t rows,
                                      int64_t cols) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < rows && j < cols) {
    dst[i + j * ldd] = src[j + i * lds];
  }
}

/*
 * Device form of GaussianProcessBase::_update_impl (gp.hpp:386-414) + BlockSymmetric
 * (linalg/block_symmetric.hpp:46-133).  The reference keeps A^-1 B and the Schur complement's LDLT next to the old
 * factor; with a Cholesky factor the same algebra IS the factor of the enlarged matrix,
 *     [K  B; B^T C] = [L 0; X^T L_S] [L 0; X^T L_S]^T,   X = L^-1 B,   L_S L_S^T = C - X^T X,
 * so the update returns an ordinary factor of size n + p: O(n^2 p) work instead of the O((n+p)^3) refit.
 */
int ab_gp_update(ab_handle h, ab_factor old_factor, const ab_op *prog, int nops, const double *train_feats,
                 int64_t n, int dim, const double *information_old, const double *new_feats, int64_t p,
                 const double *y_new, const double *yvar_new, ab_factor *factor, double *information) {
  AB_REQUIRE(h != nullptr && factor != nullptr && train_feats != nullptr && new_feats != nullptr &&
                 y_new != nullptr && n >= 1 && p >= 1 && (information == nullptr || information_old != nullptr),
             "null / sizes");
  Lock lock(h);
  AB_TRY(require_usable(old_factor));
  AB_REQUIRE(old_factor->n == n, "factor size differs from the training set");
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  Scope sc(h);
  timings_reset(h);
  *factor = nullptr;
  const int64_t N = n + p;
  ab_matrix_s *FX = nullptr, *FN = nullptr, *B = nullptr, *Cn = nullptr, *Lnew = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload_features(h, train_feats, n, dim, &FX));
  sc.own(FX);
  AB_TRY(upload_features(h, new_feats, p, dim, &FN));
  sc.own(FN);
  phase_end(h, PH_H2D);
  phase_begin(h, PH_GRAM);
  AB_TRY(gram_cross_device(h, P, FX, FN, &B)); // n x p
  sc.own(B);
  AB_TRY(gram_sym_device(h, P, FN, AB_GRAM_LOWER_ONLY, &Cn));
  sc.own(Cn);
  if (yvar_new != nullptr) {
    void *dv = nullptr;
    AB_TRY(upload_bytes(h, sc, yvar_new, static_cast<size_t>(p) * sizeof(double), &dv));
    AB_TRY(add_diag(h, view(Cn), p, static_cast<double *>(dv)));
  }
  phase_end(h, PH_GRAM);
  phase_begin(h, PH_FACTOR);
  // pivot floor of the new block from its diagonal BEFORE the Schur complement is formed (linalg.cu potrf)
  void *d_floor = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(p) * sizeof(double), &d_floor));
  AB_TRY(pivot_floor(h, view(Cn), p, static_cast<double *>(d_floor)));
  // X = L^-1 B, in place;  S = C - X^T X (lower)
  AB_TRY(trsm_left_lower(h, view(old_factor->m), old_factor->dinv, n, view(B), p));
  AB_TRY(gemm(h, GEMM_TRANS_A | GEMM_LOWER, p, p, n, -1., view(B), view(B), 1., view(Cn)));
  // the enlarged factor: old L, X^T below it, chol(S) in the corner
  AB_TRY(matrix_new(h, N, N, &Lnew));
  AB_CUDA(cudaMemcpy2DAsync(Lnew->d, Lnew->ld * sizeof(double), old_factor->m->d,
                            old_factor->m->ld * sizeof(double), static_cast<size_t>(n) * sizeof(double),
                            static_cast<size_t>(n), cudaMemcpyDeviceToDevice, h->stream));
  {
    const dim3 grid(static_cast<unsigned>((p + 255) / 256), static_cast<unsigned>(n));
    transpose_into_kernel<<<grid, 256, 0, h->stream>>>(B->d, B->ld, Lnew->d + n, Lnew->ld, p, n);
    AB_LAUNCHED(h);
  }
  AB_CUDA(cudaMemcpy2DAsync(Lnew->d + n + n * Lnew->ld, Lnew->ld * sizeof(double), Cn->d,
                            Cn->ld * sizeof(double), static_cast<size_t>(p) * sizeof(double),
                            static_cast<size_t>(p), cudaMemcpyDeviceToDevice, h->stream));
  ab_factor_s *f = nullptr;
  {
    int s = new_factor(h, Lnew, &f);
    if (s != AB_OK) {
      matrix_delete(h, Lnew);
      return s;
    }
  }
  // factor the corner in place (its own leaf grid), then rebuild the leaf inverses on the new matrix's grid
  void *tmp_inv = nullptr;
  const size_t tmp_bytes = static_cast<size_t>((p + LEAF - 1) / LEAF) * LEAF * LEAF * sizeof(double);
  int status = sc.alloc(tmp_bytes, &tmp_inv);
  h->h_flags[0] = INT_MAX;
  if (status == AB_OK) {
    cudaMemcpyAsync(h->d_flags, h->h_flags, sizeof(int), cudaMemcpyHostToDevice, h->stream);
    status = potrf(h, view(Lnew).sub(n, n), p, static_cast<double *>(tmp_inv), h->d_flags,
                   static_cast<double *>(d_floor));
  }
  if (status == AB_OK) {
    leaf_inverse_kernel<<<static_cast<unsigned>((N + LEAF - 1) / LEAF), LEAF, 0, h->stream>>>(Lnew->d, Lnew->ld, N,
                                                                                         f->dinv);
    h->launches++;
    status = download_bytes(h, h->d_flags, sizeof(int), h->h_flags);
  }
  phase_end(h, PH_FACTOR);
  if (status != AB_OK) {
    delete_factor(h, f);
    return status;
  }
  f->bad_pivot = h->h_flags[0] == INT_MAX ? -1 : n + h->h_flags[0];
  *factor = f;
  if (f->bad_pivot >= 0) {
    set_error("updated covariance is not positive definite: pivot %lld", static_cast<long long>(f->bad_pivot));
    cudaEventRecord(h->ev_total_end, h->stream);
    return AB_ERR_NOT_PD;
  }
  if (information != nullptr) {
    // the old targets are not part of a fit (gp.hpp:49-51 keeps features, factor and information): they are
    // recovered as y = K alpha = L (L^T alpha); then information' = K'^-1 [y; y_new]
    ab_matrix_s *x = nullptr, *t = nullptr;
    phase_begin(h, PH_SOLVE);
    AB_TRY(matrix_new(h, N, 1, &x));
    sc.own(x);
    AB_TRY(matrix_new(h, n, 1, &t));
    sc.own(t);
    AB_CUDA(cudaMemcpyAsync(x->d, information_old, static_cast<size_t>(n) * sizeof(double),
                            cudaMemcpyHostToDevice, h->stream));
    AB_CUDA(cudaMemcpyAsync(x->d + n, y_new, static_cast<size_t>(p) * sizeof(double), cudaMemcpyHostToDevice,
                            h->stream));
    AB_TRY(trmm_left_lower(h, view(old_factor->m), n, true, view(x), view(t), 1));
    AB_TRY(trmm_left_lower(h, view(old_factor->m), n, false, view(t), view(x), 1));
    AB_TRY(trsm_left_lower(h, view(f->m), f->dinv, N, view(x), 1));
    AB_TRY(trsm_left_lower_T(h, view(f->m), f->dinv, N, view(x), 1));
    phase_end(h, PH_SOLVE);
    cudaEventRecord(h->ev_total_end, h->stream);
    return download(h, x, 0, 0, N, 1, information);
  }
  cudaEventRecord(h->ev_total_end, h->stream);
  AB_CUDA(cudaStreamSynchronize(h->stream));
  return AB_OK;
}

} // extern "C"

namespace ab {

// Column chunk [j0, j0 + cb) of diag(K^-1): column sums of squares of L^-1 E_J obtained by a
// triangular solve that starts at row j0 (rows above are structurally zero).
static int inverse_diagonal_chunk(ab_handle_s *h, const ab_factor_s *f, int64_t j0, int64_t cb,
                                  ab_matrix_s *G, double *d_a) {
  const int64_t rows = f->n - j0; // j0 is a multiple of LEAF
  AB_TRY(fill(h, view(G), rows, cb, 0.));
  set_diag_one_kernel<<<static_cast<unsigned>((cb + 255) / 256), 256, 0, h->stream>>>(G->d, G->ld,
                                                                                      cb);
  AB_LAUNCHED(h);
  AB_TRY(trsm_left_lower(h, view(f->m).sub(j0, j0), f->dinv + (j0 / LEAF) * LEAF * LEAF, rows,
                         view(G), cb));
  return column_dots(h, view(G), view(G), rows, cb, d_a + j0);
}

// Shared implementation of ab_gp_cv / ab_dist_gp_cv.  Work is restricted to the groups (or, for
// pure leave-one-out, the column chunks) c with c % stride == phase; outputs of the other shards
// are left at zero so that a sum over ranks assembles the full result.  stride == 1: everything,
// through one explicit inverse factor (N^3/3); stride > 1: per-shard triangular solves.
int gp_cv_impl(ab_handle_s *h, ab_factor_s *f, const double *y, const double *information,
               const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
               int phase, int stride, double *mean, double *var, double *joint, double *score,
               double *group_scores) {
  const int64_t n = f->n;
  if (score != nullptr) {
    *score = 0.;
  }
  if (group_scores != nullptr) {
    std::fill(group_scores, group_scores + ngroups, 0.);
  }
  const bool want_score = score != nullptr || group_scores != nullptr;
  if (ngroups == 0 || n == 0) {
    return AB_OK;
  }
  const int64_t total = offsets[ngroups];
  AB_REQUIRE(total <= n, "more held-out indices than observations");
  int64_t maxg = 0;
  for (int64_t g = 0; g < ngroups; ++g) {
    AB_REQUIRE(offsets[g + 1] >= offsets[g], "offsets must be non-decreasing");
    maxg = std::max(maxg, offsets[g + 1] - offsets[g]);
  }
  const bool sharded = stride > 1;
  bool pure_loo = maxg == 1 && total == n && ngroups == n;
  if (pure_loo) {
    // the element-wise fast path indexes host vectors by group: the indices must be a permutation of
    // 0..n-1 (anything else — out of range, duplicates — takes the general path, which range-checks)
    std::vector<char> seen(static_cast<size_t>(n), 0);
    for (int64_t g = 0; g < n && pure_loo; ++g) {
      const int64_t idx = indices[offsets[g]];
      if (idx < 0 || idx >= n || seen[static_cast<size_t>(idx)]) {
        pure_loo = false;
      } else {
        seen[static_cast<size_t>(idx)] = 1;
      }
    }
  }
  Scope sc(h);
  ab_matrix_s *W = nullptr;
  if (!sharded) {
    phase_begin(h, PH_SOLVE);
    AB_TRY(inverse_factor(h, sc, f, &W));
    phase_end(h, PH_SOLVE);
  }

  void *d_y = nullptr, *d_info = nullptr, *d_idx = nullptr, *d_mean = nullptr, *d_var = nullptr;
  const size_t nbytes = static_cast<size_t>(n) * sizeof(double);
  AB_TRY(upload_bytes(h, sc, y, nbytes, &d_y));
  AB_TRY(upload_bytes(h, sc, information, nbytes, &d_info));
  AB_TRY(upload_bytes(h, sc, indices, static_cast<size_t>(total) * sizeof(int64_t), &d_idx));
  AB_TRY(sc.alloc(nbytes, &d_mean));
  AB_TRY(sc.alloc(nbytes, &d_var));
  AB_CUDA(cudaMemsetAsync(d_mean, 0, nbytes, h->stream));
  AB_CUDA(cudaMemsetAsync(d_var, 0, nbytes, h->stream));

  phase_begin(h, PH_PREDICT);
  double total_score = 0.;
  if (pure_loo) {
    // pure leave-one-out: everything is element-wise on diag(K^-1)
    void *d_a = nullptr, *d_terms = nullptr;
    AB_TRY(sc.alloc(nbytes, &d_a));
    AB_TRY(sc.alloc(nbytes, &d_terms));
    AB_CUDA(cudaMemsetAsync(d_terms, 0, nbytes, h->stream));
    if (!sharded) {
      AB_TRY(column_dots(h, view(W), view(W), n, n, static_cast<double *>(d_a)));
      loo_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(
          static_cast<double *>(d_a), static_cast<double *>(d_y), static_cast<double *>(d_info), n,
          static_cast<double *>(d_mean), static_cast<double *>(d_var),
          static_cast<double *>(d_terms));
      AB_LAUNCHED(h);
    } else {
      constexpr int64_t CB = 2048; // column chunk dealt cyclically to the ranks
      ab_matrix_s *G = nullptr;
      AB_TRY(matrix_new(h, n, std::min(CB, n), &G));
      sc.own(G);
      int64_t c = 0;
      for (int64_t j0 = 0; j0 < n; j0 += CB, ++c) {
        if (c % stride != phase) {
          continue;
        }
        const int64_t cb = std::min(CB, n - j0);
        AB_TRY(inverse_diagonal_chunk(h, f, j0, cb, G, static_cast<double *>(d_a)));
        loo_kernel<<<static_cast<unsigned>((cb + 255) / 256), 256, 0, h->stream>>>(
            static_cast<double *>(d_a) + j0, static_cast<double *>(d_y) + j0,
            static_cast<double *>(d_info) + j0, cb, static_cast<double *>(d_mean) + j0,
            static_cast<double *>(d_var) + j0, static_cast<double *>(d_terms) + j0);
        AB_LAUNCHED(h);
      }
    }
    phase_end(h, PH_PREDICT);
    AB_TRY(download_bytes(h, d_mean, nbytes, mean));
    if (what != AB_PREDICT_MEAN) {
      // for 1x1 groups the joint blocks, in key order, are the variances in index order
      std::vector<double> v(n);
      AB_TRY(download_bytes(h, d_var, nbytes, v.data()));
      if (what == AB_PREDICT_MARGINAL) {
        std::copy(v.begin(), v.end(), var);
      } else {
        for (int64_t g = 0; g < n; ++g) {
          joint[g] = v[indices[offsets[g]]];
        }
      }
    }
    if (want_score) {
      std::vector<double> t(n);
      AB_TRY(download_bytes(h, d_terms, nbytes, t.data()));
      for (int64_t g = 0; g < n; ++g) {
        total_score += t[indices[offsets[g]]];
        if (group_scores != nullptr) {
          group_scores[g] = t[indices[offsets[g]]];
        }
      }
      if (score != nullptr) {
        *score = total_score;
      }
    }
    cudaEventRecord(h->ev_total_end, h->stream);
    return AB_OK;
  }

  // general groups: A_g = (K^-1)_gg ; x = A_g^-1 v_g ; mean_g = y_g - x ; cov_g = A_g^-1
  ab_matrix_s *G = nullptr, *A = nullptr, *Wg = nullptr, *Tg = nullptr, *x = nullptr;
  AB_TRY(matrix_new(h, n, maxg, &G));
  sc.own(G);
  AB_TRY(matrix_new(h, maxg, maxg, &A));
  sc.own(A);
  AB_TRY(matrix_new(h, maxg, maxg, &Wg));
  sc.own(Wg);
  const int64_t half = round_up((maxg + 1) / 2, LEAF) + LEAF;
  AB_TRY(matrix_new(h, half, half, &Tg));
  sc.own(Tg);
  AB_TRY(matrix_new(h, maxg, 2, &x));
  sc.own(x);
  void *d_diag = nullptr, *d_gscal = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(maxg) * sizeof(double), &d_diag));
  AB_TRY(sc.alloc(static_cast<size_t>(ngroups) * 2 * sizeof(double), &d_gscal));
  AB_CUDA(cudaMemsetAsync(d_gscal, 0, static_cast<size_t>(ngroups) * 2 * sizeof(double),
                          h->stream));
  ab_matrix_s *Cg = nullptr;
  if (what == AB_PREDICT_JOINT) {
    AB_TRY(matrix_new(h, maxg, maxg, &Cg));
    sc.own(Cg);
  }
  const int64_t nleaf = (maxg + LEAF - 1) / LEAF;
  void *d_ginv = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(nleaf) * LEAF * LEAF * sizeof(double), &d_ginv));
  double *joint_cursor = joint;
  h->h_flags[0] = INT_MAX;
  AB_CUDA(cudaMemcpyAsync(h->d_flags, h->h_flags, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  for (int64_t g = 0; g < ngroups; ++g) {
    const int64_t k = offsets[g + 1] - offsets[g];
    if (k == 0) {
      continue;
    }
    if (g % stride != phase) {
      if (what == AB_PREDICT_JOINT) {
        std::fill(joint_cursor, joint_cursor + k * k, 0.);
        joint_cursor += k * k;
      }
      continue;
    }
    const int64_t *gi = static_cast<int64_t *>(d_idx) + offsets[g];
    AB_TRY(inverse_block_device(h, f, W, gi, indices + offsets[g], k, G, view(A)));
    AB_TRY(potrf(h, view(A), k, static_cast<double *>(d_ginv), h->d_flags));
    // x = A^-1 v_g  (column 0 of x), keep v_g in column 1 for the score
    gather_vec_kernel<<<static_cast<unsigned>((k + 255) / 256), 256, 0, h->stream>>>(
        static_cast<double *>(d_info), gi, k, x->d);
    AB_LAUNCHED(h);
    AB_CUDA(cudaMemcpyAsync(x->d + x->ld, x->d, static_cast<size_t>(k) * sizeof(double),
                            cudaMemcpyDeviceToDevice, h->stream));
    AB_TRY(trsm_left_lower(h, view(A), static_cast<double *>(d_ginv), k, view(x), 1));
    AB_TRY(trsm_left_lower_T(h, view(A), static_cast<double *>(d_ginv), k, view(x), 1));
    double *diag = nullptr;
    if (what != AB_PREDICT_MEAN) {
      // A^-1 = Wg^T Wg with Wg = chol(A)^-1
      AB_TRY(fill(h, view(Wg), k, k, 0.));
      AB_TRY(trtri_rec(h, view(A), static_cast<double *>(d_ginv), k, view(Wg), view(Tg)));
      if (what == AB_PREDICT_MARGINAL) {
        AB_TRY(column_dots(h, view(Wg), view(Wg), k, k, static_cast<double *>(d_diag)));
        diag = static_cast<double *>(d_diag);
      } else {
        AB_TRY(gemm(h, GEMM_TRANS_A, k, k, k, 1., view(Wg), view(Wg), 0., view(Cg)));
      }
    }
    heldout_scatter_kernel<<<static_cast<unsigned>((k + 255) / 256), 256, 0, h->stream>>>(
        static_cast<double *>(d_y), x->d, diag, gi, k, static_cast<double *>(d_mean),
        diag != nullptr ? static_cast<double *>(d_var) : nullptr);
    AB_LAUNCHED(h);
    if (want_score) {
      // NLL(dev = x, cov = A^-1) = 0.5 (-logdet(A) + x^T A x + k log 2pi), x^T A x = x . v_g
      AB_TRY(logdet_chol(h, view(A), k, static_cast<double *>(d_gscal) + 2 * g));
      AB_TRY(dot(h, x->d, x->d + x->ld, k, static_cast<double *>(d_gscal) + 2 * g + 1));
    }
    if (what == AB_PREDICT_JOINT) {
      AB_TRY(download(h, Cg, 0, 0, k, k, joint_cursor));
      joint_cursor += k * k;
    }
  }
  phase_end(h, PH_PREDICT);
  AB_TRY(download_bytes(h, d_mean, nbytes, mean));
  if (what == AB_PREDICT_MARGINAL) {
    AB_TRY(download_bytes(h, d_var, nbytes, var));
  }
  AB_TRY(download_bytes(h, h->d_flags, sizeof(int), h->h_flags));
  if (h->h_flags[0] != INT_MAX) {
    set_error("a held-out block of the inverse covariance is not positive definite");
    return AB_ERR_NOT_PD;
  }
  if (want_score) {
    std::vector<double> gs(static_cast<size_t>(ngroups) * 2);
    AB_TRY(download_bytes(h, d_gscal, gs.size() * sizeof(double), gs.data()));
    for (int64_t g = 0; g < ngroups; ++g) {
      const int64_t k = offsets[g + 1] - offsets[g];
      if (k == 0 || g % stride != phase) {
        continue;
      }
      const double sg =
          0.5 * (-gs[2 * g] + gs[2 * g + 1] + static_cast<double>(k) * std::log(2 * M_PI));
      total_score += sg;
      if (group_scores != nullptr) {
        group_scores[g] = sg;
      }
    }
    if (score != nullptr) {
      *score = total_score;
    }
  }
  cudaEventRecord(h->ev_total_end, h->stream);
  return AB_OK;
}

} // namespace ab

extern "C" {

int ab_gp_cv(ab_handle h, ab_factor f, const double *y, const double *information,
             const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
             double *mean, double *var, double *joint, double *score) {
  AB_REQUIRE(h != nullptr && y != nullptr && information != nullptr && indices != nullptr &&
                 offsets != nullptr && mean != nullptr && ngroups >= 0,
             "null");
  AB_REQUIRE(what == AB_PREDICT_MEAN || (what == AB_PREDICT_MARGINAL && var != nullptr) ||
                 (what == AB_PREDICT_JOINT && joint != nullptr),
             "prediction kind / outputs");
  Lock lock(h);
  AB_TRY(require_usable(f));
  timings_reset(h);
  return gp_cv_impl(h, f, y, information, indices, offsets, ngroups, what, 0, 1, mean, var, joint,
                    score, nullptr);
}

int ab_gp_cv_scores(ab_handle h, ab_factor f, const double *y, const double *information,
                    const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
                    double *mean, double *var, double *joint, double *score,
                    double *group_scores) {
  AB_REQUIRE(h != nullptr && y != nullptr && information != nullptr && indices != nullptr &&
                 offsets != nullptr && mean != nullptr && ngroups >= 0,
             "null");
  AB_REQUIRE(what == AB_PREDICT_MEAN || (what == AB_PREDICT_MARGINAL && var != nullptr) ||
                 (what == AB_PREDICT_JOINT && joint != nullptr),
             "prediction kind / outputs");
  Lock lock(h);
  AB_TRY(require_usable(f));
  timings_reset(h);
  return gp_cv_impl(h, f, y, information, indices, offsets, ngroups, what, 0, 1, mean, var, joint,
                    score, group_scores);
}

int ab_gp_cv_shard(ab_handle h, ab_factor f, const double *y, const double *information,
                   const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
                   int shard, int nshards, double *mean, double *var, double *score) {
  AB_REQUIRE(h != nullptr && y != nullptr && information != nullptr && indices != nullptr &&
                 offsets != nullptr && mean != nullptr && ngroups >= 0,
             "null");
  AB_REQUIRE(what == AB_PREDICT_MEAN || (what == AB_PREDICT_MARGINAL && var != nullptr),
             "sharded CV returns means or marginals");
  AB_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, "shard / nshards");
  Lock lock(h);
  AB_TRY(require_usable(f));
  timings_reset(h);
  // nshards == 1 still takes the per-shard solve path (stride 2 with every unit in phase 0 would
  // change the partition), so run it as "stride = nshards" with a guard for the degenerate case
  if (nshards == 1) {
    return gp_cv_impl(h, f, y, information, indices, offsets, ngroups, what, 0, 1, mean, var,
                      nullptr, score, nullptr);
  }
  return gp_cv_impl(h, f, y, information, indices, offsets, ngroups, what, shard, nshards, mean,
                    var, nullptr, score, nullptr);
}

// ---- dense building block ---------------------------------------------------------------------

int ab_gemm(ab_handle h, uint32_t flags, double alpha, ab_matrix A, ab_matrix B, double beta,
            ab_matrix C) {
  AB_REQUIRE(h != nullptr && A != nullptr && B != nullptr && C != nullptr, "null");
  const bool ta = flags & AB_GEMM_TRANS_A;
  const bool tb = flags & AB_GEMM_TRANS_B;
  const int64_t m = ta ? A->cols : A->rows;
  const int64_t k = ta ? A->rows : A->cols;
  const int64_t kb = tb ? B->cols : B->rows;
  const int64_t n = tb ? B->rows : B->cols;
  AB_REQUIRE(k == kb && C->rows == m && C->cols == n, "GEMM shapes do not conform");
  AB_REQUIRE(C != A && C != B, "ab_gemm does not work in place");
  Lock lock(h);
  timings_reset(h);
  phase_begin(h, PH_FACTOR);
  unsigned f = (ta ? GEMM_TRANS_A : 0u) | (tb ? GEMM_TRANS_B : 0u) |
               ((flags & AB_GEMM_LOWER) ? GEMM_LOWER : 0u);
  int s = gemm(h, f, m, n, k, alpha, view(A), view(B), beta, view(C));
  phase_end(h, PH_FACTOR);
  cudaEventRecord(h->ev_total_end, h->stream);
  return s;
}

// ---- integer contract -------------------------------------------------------------------------

int ab_group_indexers(const int64_t *item_keys, int64_t n, int64_t *keys, int64_t *offsets,
                      int64_t *indices, int64_t *ngroups) {
  AB_REQUIRE(n >= 0 && offsets != nullptr && ngroups != nullptr &&
                 (n == 0 || (item_keys && keys && indices)),
             "null");
  // std::map<Key, std::vector<size_t>> semantics (group_by.hpp:349-376): keys ascending, members in
  // encounter order == ascending index.  A stable sort of indices by key reproduces it exactly.
  std::vector<int64_t> order(static_cast<size_t>(n));
  std::iota(order.begin(), order.end(), int64_t(0));
  std::stable_sort(order.begin(), order.end(),
                   [&](int64_t a, int64_t b) { return item_keys[a] < item_keys[b]; });
  int64_t g = 0;
  offsets[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t key = item_keys[order[static_cast<size_t>(i)]];
    if (i == 0 || key != keys[g - 1]) {
      if (i > 0) {
        offsets[g] = i;
      }
      keys[g++] = key;
    }
    indices[i] = order[static_cast<size_t>(i)];
  }
  offsets[g] = n;
  *ngroups = g;
  return AB_OK;
}

int ab_partition_triangular(int64_t n, int64_t count, int64_t *bounds) {
  AB_REQUIRE(n >= 0 && count >= 1 && bounds != nullptr, "partition_triangular arguments");
  // end_b = rint(n * f_b) with f_b^2 = b / count accumulated the way the reference accumulates it (the
  // rounding of the running area is part of the contract): block.hpp:31-39
  double area = 0.;
  int64_t start = 0;
  for (int64_t b = 0; b < count; ++b) {
    const double end_fraction = std::sqrt(1. / static_cast<double>(count) + area);
    area = end_fraction * end_fraction;
    const int64_t end = static_cast<int64_t>(std::rint(static_cast<double>(n) * end_fraction));
    bounds[2 * b] = start;
    bounds[2 * b + 1] = end;
    start = end;
  }
  bounds[2 * (count - 1) + 1] = std::min(bounds[2 * (count - 1) + 1], n); // block.hpp:41-42
  return AB_OK;
}

} // extern "C"
