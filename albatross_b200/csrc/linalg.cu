// Blocked Cholesky, triangular solves and reductions built on the DMMA GEMM family (gemm.cu).
//
// Factorisation: recursive right-looking blocked Cholesky.  All O(n^3) work is DSYRK/DGEMM on the
// FP64 tensor pipe; the only non-GEMM kernel is the LEAF x LEAF diagonal-block kernel, which also
// produces the block's explicit inverse so that every triangular solve (inside the factorisation
// and afterwards) is a GEMM as well.  Replaces Eigen's unblocked, diagonally pivoted LDLT
// (reference third_party/eigen/Eigen/src/Cholesky/LDLT.h:294-394) with an unpivoted LL^T whose
// D = diag(L)^2; see DESIGN.md for the parity argument.
#include "linalg.cuh"

#include <climits>

namespace ab {

// ------------------------------------------------------------------------------------------------
// LEAF x LEAF: Cholesky of a diagonal block + explicit inverse of its factor
// ------------------------------------------------------------------------------------------------

constexpr int LS = LEAF + 1; // padded shared-memory stride

__global__ void __launch_bounds__(256)
potf2_inv_kernel(double *A, int64_t lda, int nb, double *dinv, int64_t global_offset, int *d_bad) {
  extern __shared__ double sm[];
  double *s = sm;               // s[r * LS + c]  : the block, lower triangle becomes L
  double *inv = sm + LEAF * LS; // inv[r * LS + c]: L^-1
  const int tid = threadIdx.x;

  for (int idx = tid; idx < LEAF * LEAF; idx += blockDim.x) {
    const int r = idx % LEAF;
    const int c = idx / LEAF;
    double v = (r == c) ? 1. : 0.; // identity padding for ragged blocks
    if (r < nb && c < nb && r >= c) {
      v = A[r + c * lda];
    }
    s[r * LS + c] = v;
  }
  __syncthreads();

  for (int j = 0; j < LEAF; ++j) {
    if (tid == 0) {
      const double d = s[j * LS + j];
      if (!(d > 0.) && j < nb) {
        atomicMin(d_bad, static_cast<int>(global_offset + j));
      }
      s[j * LS + j] = sqrt(d);
    }
    __syncthreads();
    const double rdiag = 1. / s[j * LS + j];
    if (tid > j && tid < LEAF) {
      s[tid * LS + j] *= rdiag;
    }
    __syncthreads();
    const int w = LEAF - j - 1;
    for (int idx = tid; idx < w * w; idx += blockDim.x) {
      const int r = j + 1 + idx % w;
      const int c = j + 1 + idx / w;
      if (r >= c) {
        s[r * LS + c] -= s[r * LS + j] * s[c * LS + j];
      }
    }
    __syncthreads();
  }

  // inverse: thread c owns column c of L^-1 (forward substitution on e_c)
  if (tid < LEAF) {
    const int c = tid;
    for (int r = 0; r < c; ++r) {
      inv[r * LS + c] = 0.;
    }
    inv[c * LS + c] = 1. / s[c * LS + c];
    for (int r = c + 1; r < LEAF; ++r) {
      double acc = 0.;
      for (int t = c; t < r; ++t) {
        acc = fma(s[r * LS + t], inv[t * LS + c], acc);
      }
      inv[r * LS + c] = -acc / s[r * LS + r];
    }
  }
  __syncthreads();

  for (int idx = tid; idx < LEAF * LEAF; idx += blockDim.x) {
    const int r = idx % LEAF;
    const int c = idx / LEAF;
    if (r < nb && c < nb && r >= c) {
      A[r + c * lda] = s[r * LS + c];
    }
    dinv[r + c * LEAF] = inv[r * LS + c];
  }
}

static int64_t split(int64_t n) {
  int64_t n1 = round_up((n + 1) / 2, LEAF);
  if (n1 >= n) {
    n1 = round_up(n, LEAF) - LEAF;
  }
  return n1;
}

static MatView leaf_inverse(const double *dinv) {
  return MatView{const_cast<double *>(dinv), LEAF};
}

int trsm_right_lower_T(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView X,
                       int64_t m) {
  if (n <= 0 || m <= 0) {
    return AB_OK;
  }
  if (n <= LEAF) {
    // X <- X * Linv^T, in place (one CTA owns all n columns of its row block)
    return gemm(h, GEMM_TRANS_B, m, n, n, 1., X, leaf_inverse(dinv), 0., X);
  }
  const int64_t n1 = split(n);
  const int64_t n2 = n - n1;
  AB_TRY(trsm_right_lower_T(h, L, dinv, n1, X, m));
  AB_TRY(gemm(h, GEMM_TRANS_B, m, n2, n1, -1., X, L.sub(n1, 0), 1., X.sub(0, n1)));
  return trsm_right_lower_T(h, L.sub(n1, n1), dinv + (n1 / LEAF) * LEAF * LEAF, n2, X.sub(0, n1),
                            m);
}

int trsm_left_lower(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView X,
                    int64_t p) {
  if (n <= 0 || p <= 0) {
    return AB_OK;
  }
  if (n <= LEAF) {
    return gemm(h, 0u, n, p, n, 1., leaf_inverse(dinv), X, 0., X);
  }
  const int64_t n1 = split(n);
  const int64_t n2 = n - n1;
  AB_TRY(trsm_left_lower(h, L, dinv, n1, X, p));
  AB_TRY(gemm(h, 0u, n2, p, n1, -1., L.sub(n1, 0), X, 1., X.sub(n1, 0)));
  return trsm_left_lower(h, L.sub(n1, n1), dinv + (n1 / LEAF) * LEAF * LEAF, n2, X.sub(n1, 0), p);
}

int trsm_left_lower_T(ab_handle_s *h, MatView L, const double *dinv, int64_t n, MatView X,
                      int64_t p) {
  if (n <= 0 || p <= 0) {
    return AB_OK;
  }
  if (n <= LEAF) {
    return gemm(h, GEMM_TRANS_A, n, p, n, 1., leaf_inverse(dinv), X, 0., X);
  }
  const int64_t n1 = split(n);
  const int64_t n2 = n - n1;
  AB_TRY(trsm_left_lower_T(h, L.sub(n1, n1), dinv + (n1 / LEAF) * LEAF * LEAF, n2, X.sub(n1, 0),
                           p));
  AB_TRY(gemm(h, GEMM_TRANS_A, n1, p, n2, -1., L.sub(n1, 0), X.sub(n1, 0), 1., X));
  return trsm_left_lower_T(h, L, dinv, n1, X, p);
}

static int potrf_rec(ab_handle_s *h, MatView A, int64_t n, double *dinv, int64_t offset,
                     int *d_bad) {
  if (n <= LEAF) {
    constexpr size_t smem = 2 * LEAF * LS * sizeof(double);
    static bool configured = false;
    if (!configured) {
      AB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
      configured = true;
    }
    potf2_inv_kernel<<<1, 256, smem, h->stream>>>(A.p, A.ld, static_cast<int>(n), dinv, offset,
                                                  d_bad);
    AB_LAUNCHED(h);
    return AB_OK;
  }
  const int64_t n1 = split(n);
  const int64_t n2 = n - n1;
  AB_TRY(potrf_rec(h, A, n1, dinv, offset, d_bad));
  AB_TRY(trsm_right_lower_T(h, A, dinv, n1, A.sub(n1, 0), n2));
  AB_TRY(gemm(h, GEMM_TRANS_B | GEMM_LOWER, n2, n2, n1, -1., A.sub(n1, 0), A.sub(n1, 0), 1.,
              A.sub(n1, n1)));
  return potrf_rec(h, A.sub(n1, n1), n2, dinv + (n1 / LEAF) * LEAF * LEAF, offset + n1, d_bad);
}

int potrf(ab_handle_s *h, MatView A, int64_t n, double *dinv, int *d_bad) {
  if (n <= 0) {
    return AB_OK;
  }
  AB_REQUIRE(n < INT_MAX, "matrix too large");
  return potrf_rec(h, A, n, dinv, 0, d_bad);
}

// ------------------------------------------------------------------------------------------------
// reductions and small utilities
// ------------------------------------------------------------------------------------------------

template <int THREADS> __device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  double total = 0.;
  if (threadIdx.x < 32) {
    total = threadIdx.x < THREADS / 32 ? red[threadIdx.x] : 0.;
    for (int o = 16; o > 0; o >>= 1) {
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
  }
  __syncthreads();
  return total; // valid in warp 0
}

__global__ void __launch_bounds__(1024) logdet_kernel(const double *L, int64_t ld, int64_t n,
                                                      double *out) {
  double acc = 0.;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    acc += 2. * log(L[i + i * ld]);
  }
  const double total = block_sum<1024>(acc);
  if (threadIdx.x == 0) {
    out[0] = total;
  }
}

// One CTA per column j: out[j] = sum_i A(i,j) * B(i,j).  (column 0 with cols == 1 is a plain dot.)
__global__ void __launch_bounds__(256) column_dots_kernel(const double *A, int64_t lda,
                                                          const double *B, int64_t ldb,
                                                          int64_t rows, double *out) {
  const int64_t j = blockIdx.x;
  const double *a = A + j * lda;
  const double *b = B + j * ldb;
  double acc = 0.;
  for (int64_t i = threadIdx.x; i < rows; i += 256) {
    acc = fma(a[i], b[i], acc);
  }
  const double total = block_sum<256>(acc);
  if (threadIdx.x == 0) {
    out[j] = total;
  }
}

__global__ void add_diag_kernel(double *A, int64_t ld, int64_t n, const double *d) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    A[i + i * ld] += d[i];
  }
}

__global__ void fill_kernel(double *A, int64_t ld, int64_t rows, int64_t cols, double value) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < rows && j < cols) {
    A[i + j * ld] = value;
  }
}

__global__ void set_diag_kernel(double *A, int64_t ld, int64_t n, double value) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    A[i + i * ld] = value;
  }
}

__global__ void __launch_bounds__(256) nan_lower_kernel(const double *A, int64_t ld, int64_t n,
                                                        int *flag) {
  const int64_t j = blockIdx.x;
  bool bad = false;
  for (int64_t i = j + threadIdx.x; i < n; i += 256) {
    const double v = A[i + j * ld];
    bad = bad || (v != v);
  }
  if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) {
    atomicExch(flag, 1);
  }
}

int logdet_chol(ab_handle_s *h, MatView L, int64_t n, double *d_out) {
  logdet_kernel<<<1, 1024, 0, h->stream>>>(L.p, L.ld, n, d_out);
  AB_LAUNCHED(h);
  return AB_OK;
}

int column_dots(ab_handle_s *h, MatView A, MatView B, int64_t rows, int64_t cols, double *d_out) {
  if (cols <= 0) {
    return AB_OK;
  }
  column_dots_kernel<<<static_cast<unsigned>(cols), 256, 0, h->stream>>>(A.p, A.ld, B.p, B.ld,
                                                                         rows, d_out);
  AB_LAUNCHED(h);
  return AB_OK;
}

int dot(ab_handle_s *h, const double *a, const double *b, int64_t n, double *d_out) {
  return column_dots(h, MatView{const_cast<double *>(a), n}, MatView{const_cast<double *>(b), n},
                     n, 1, d_out);
}

int add_diag(ab_handle_s *h, MatView A, int64_t n, const double *d_diag) {
  if (n <= 0) {
    return AB_OK;
  }
  add_diag_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(A.p, A.ld, n,
                                                                                 d_diag);
  AB_LAUNCHED(h);
  return AB_OK;
}

int fill(ab_handle_s *h, MatView A, int64_t rows, int64_t cols, double value) {
  if (rows <= 0 || cols <= 0) {
    return AB_OK;
  }
  // gridDim.y is limited to 65535: tile the columns
  for (int64_t c0 = 0; c0 < cols; c0 += 65535) {
    const int64_t nc = cols - c0 < 65535 ? cols - c0 : 65535;
    const dim3 grid(static_cast<unsigned>((rows + 255) / 256), static_cast<unsigned>(nc));
    const MatView sub = A.sub(0, c0);
    fill_kernel<<<grid, 256, 0, h->stream>>>(sub.p, sub.ld, rows, nc, value);
    AB_LAUNCHED(h);
  }
  return AB_OK;
}

int set_identity(ab_handle_s *h, MatView A, int64_t n) {
  if (n <= 0) {
    return AB_OK;
  }
  AB_TRY(fill(h, A, n, n, 0.));
  set_diag_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(A.p, A.ld, n, 1.);
  AB_LAUNCHED(h);
  return AB_OK;
}

int has_nan_lower(ab_handle_s *h, MatView A, int64_t n, int *d_flag) {
  if (n <= 0) {
    return AB_OK;
  }
  nan_lower_kernel<<<static_cast<unsigned>(n), 256, 0, h->stream>>>(A.p, A.ld, n, d_flag);
  AB_LAUNCHED(h);
  return AB_OK;
}

} // namespace ab
