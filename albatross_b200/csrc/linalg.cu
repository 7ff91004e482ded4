// Blocked Cholesky, triangular solves and reductions built on the DMMA GEMM family (gemm.cu).
//
// Factorisation: recursive right-looking blocked Cholesky.  All O(n^3) work is DSYRK/DGEMM on the
// FP64 tensor pipe; the only non-GEMM kernel is the LEAF x LEAF diagonal-block kernel, which also
// produces the block's explicit inverse so that every triangular solve (inside the factorisation
// and afterwards) is a GEMM as well.  Replaces Eigen's unblocked, diagonally pivoted LDLT
// (reference third_party/eigen/Eigen/src/Cholesky/LDLT.h:294-394) with an unpivoted LL^T whose
// D = diag(L)^2; see DESIGN.md for the parity argument.
#include "linalg.cuh"

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace ab {

// ------------------------------------------------------------------------------------------------
// LEAF x LEAF: Cholesky of a diagonal block + explicit inverse of its factor
// ------------------------------------------------------------------------------------------------

constexpr int LS = LEAF + 1; // padded shared-memory stride

__global__ void __launch_bounds__(256)
potf2_inv_kernel(double *A, int64_t lda, int nb, double *dinv, int64_t global_offset, int *d_bad,
                 const double *floor) {
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
      if (j < nb && !(d > (floor != nullptr ? floor[j] : 0.))) {
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

// Version 2 of the leaf: the block and the growing inverse live in REGISTERS (thread (r, g) owns
// row r, columns g, g + 4, ..., of both), one barrier per column, no divisions or square roots on
// the critical path except one reciprocal per column.
//
// Right-looking elimination on the unscaled Schur complement S (S_0 = block, symmetric) and on
// W (W_0 = I), for j = 0..63 with d_j = S_j[j][j]:
//     S[r][c] -= S[r][j] S[c][j] / d_j      (r > j, c > j)
//     W[r][c] -= S[r][j] W[j][c] / d_j      (r > j, c <= j)          (forward substitution on I)
// after which L[r][c] = S[r][c] / sqrt(d_c) and (L^-1)[r][c] = W[r][c] / sqrt(d_r).  Column j of S and
// row j of W are final when step j starts; they are broadcast through double-buffered shared memory.
// v1 (three barriers, a serial square root and integer divisions per column, then a 64-thread
// serial substitution for the inverse) took 92 us per leaf; this one is bounded by 64 x (barrier +
// LDS + reciprocal + 16 FMA) ~ 6 us.
__global__ void __launch_bounds__(256)
potf2_inv_kernel_v2(double *A, int64_t lda, int nb, double *dinv, int64_t global_offset,
                    int *d_bad, const double *floor) {
  __shared__ double colbuf[2][LEAF];
  __shared__ double rowbuf[2][LEAF];
  __shared__ double dpiv[LEAF];
  const int tid = threadIdx.x;
  const int r = tid & (LEAF - 1);
  const int g = tid >> 6; // 0..3
  constexpr int NI = LEAF / 4;
  double S[NI], W[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = g + 4 * i;
    const int hi = r > c ? r : c;
    const int lo = r > c ? c : r;
    double v = (r == c) ? 1. : 0.; // identity padding for ragged blocks
    if (hi < nb) {
      v = A[hi + lo * lda];
    }
    S[i] = v;
    W[i] = (r == c) ? 1. : 0.;
  }
  if (g == 0) {
    colbuf[0][r] = S[0];
  }
  if (r == 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      rowbuf[0][g + 4 * i] = W[i];
    }
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < LEAF; ++j) {
    const double *col = colbuf[j & 1];
    const double *wr = rowbuf[j & 1];
    const double d = col[j];
    if (tid == 0) {
      dpiv[j] = d;
      // a pivot must clear a floor relative to the ORIGINAL diagonal entry (see potrf): "> 0" alone lets a
      // numerically singular matrix through whenever cancellation leaves +1e-17 instead of -1e-17
      if (j < nb && !(d > (floor != nullptr ? floor[j] : 0.))) {
        atomicMin(d_bad, static_cast<int>(global_offset + j));
      }
    }
    if (r > j) {
      const double lr = col[r] * (1. / d);
      const int ilo = j >> 2; // columns 4 ilo .. 4 ilo + 3 straddle j
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = g + 4 * i;
        if (i < ilo) {
          W[i] = fma(-lr, wr[c], W[i]);
        } else if (i > ilo) {
          S[i] = fma(-lr, col[c], S[i]);
        } else if (c > j) {
          S[i] = fma(-lr, col[c], S[i]);
        } else {
          W[i] = fma(-lr, wr[c], W[i]);
        }
      }
    }
    if (j + 1 < LEAF) {
      if (g == ((j + 1) & 3)) {
        colbuf[(j + 1) & 1][r] = S[(j + 1) >> 2];
      }
      if (r == j + 1) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          rowbuf[(j + 1) & 1][g + 4 * i] = W[i];
        }
      }
    }
    __syncthreads();
  }

  // rs_c = 1 / sqrt(d_c), reusing colbuf[0]
  if (tid < LEAF) {
    colbuf[0][tid] = 1. / sqrt(dpiv[tid]);
  }
  __syncthreads();
  const double rs_r = colbuf[0][r];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = g + 4 * i;
    if (r >= c && r < nb) {
      A[r + c * lda] = S[i] * colbuf[0][c];
    }
    dinv[r + c * LEAF] = (c <= r) ? W[i] * rs_r : 0.;
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
  if (p == 1 && n > LEAF && gemv_fast_ok(L, X.p)) {
    return trsv_lower(h, L, dinv, n, X.p); // one right-hand side: block substitution (trsv.cu)
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
  if (p == 1 && n > LEAF && gemv_fast_ok(L, X.p)) {
    return trsv_lower_T(h, L, dinv, n, X.p);
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

// Y[0:nb, c] = tril(L) X[0:nb, c] or tril(L)^T X[0:nb, c] for one LEAF-sized diagonal block; thread =
// (row, column of a 4-column group).
__global__ void __launch_bounds__(256)
trmm_leaf_kernel(const double *L, int64_t ldl, int nb, int trans, const double *X, int64_t ldx,
                 double *Y, int64_t ldy, int64_t p) {
  __shared__ double s[LEAF * LS];
  for (int idx = threadIdx.x; idx < LEAF * LEAF; idx += blockDim.x) {
    const int r = idx % LEAF, c = idx / LEAF;
    s[r * LS + c] = (r < nb && c <= r) ? L[r + c * ldl] : 0.;
  }
  __syncthreads();
  const int r = threadIdx.x & (LEAF - 1);
  const int64_t c = blockIdx.x * static_cast<int64_t>(4) + (threadIdx.x >> 6);
  if (r >= nb || c >= p) {
    return;
  }
  const double *x = X + c * ldx;
  double acc = 0.;
  if (!trans) {
    for (int t = 0; t <= r; ++t) {
      acc = fma(s[r * LS + t], x[t], acc);
    }
  } else {
    for (int t = r; t < nb; ++t) {
      acc = fma(s[t * LS + r], x[t], acc);
    }
  }
  Y[r + c * ldy] = acc;
}

int trmm_left_lower(ab_handle_s *h, MatView L, int64_t n, bool trans, MatView X, MatView Y,
                    int64_t p) {
  if (n <= 0 || p <= 0) {
    return AB_OK;
  }
  if (n <= LEAF) {
    trmm_leaf_kernel<<<static_cast<unsigned>((p + 3) / 4), 256, 0, h->stream>>>(
        L.p, L.ld, static_cast<int>(n), trans ? 1 : 0, X.p, X.ld, Y.p, Y.ld, p);
    AB_LAUNCHED(h);
    return AB_OK;
  }
  const int64_t n1 = split(n);
  const int64_t n2 = n - n1;
  AB_TRY(trmm_left_lower(h, L, n1, trans, X, Y, p));
  AB_TRY(trmm_left_lower(h, L.sub(n1, n1), n2, trans, X.sub(n1, 0), Y.sub(n1, 0), p));
  if (!trans) { // Y2 += L21 X1
    return gemm(h, 0u, n2, p, n1, 1., L.sub(n1, 0), X, 1., Y.sub(n1, 0));
  }
  // Y1 += L21^T X2
  return gemm(h, GEMM_TRANS_A, n1, p, n2, 1., L.sub(n1, 0), X.sub(n1, 0), 1., Y);
}

// floor: per-pivot thresholds of THIS sub-problem (may be null = 0)
static int potrf_rec(ab_handle_s *h, MatView A, int64_t n, double *dinv, int64_t offset,
                     int *d_bad, const double *floor) {
  if (n <= LEAF) {
    constexpr size_t smem = 2 * LEAF * LS * sizeof(double);
    static PerDeviceOnce once;
    if (once.need(h->device)) {
      AB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    }
#ifdef AB_LEAF_V1
    potf2_inv_kernel<<<1, 256, smem, h->stream>>>(A.p, A.ld, static_cast<int>(n), dinv, offset,
                                                  d_bad, floor);
#else
    potf2_inv_kernel_v2<<<1, 256, 0, h->stream>>>(A.p, A.ld, static_cast<int>(n), dinv, offset,
                                                  d_bad, floor);
#endif
    AB_LAUNCHED(h);
    return AB_OK;
  }
  const int64_t n1 = split(n);
  const int64_t n2 = n - n1;
  AB_TRY(potrf_rec(h, A, n1, dinv, offset, d_bad, floor));
  AB_TRY(trsm_right_lower_T(h, A, dinv, n1, A.sub(n1, 0), n2));
  AB_TRY(gemm(h, GEMM_TRANS_B | GEMM_LOWER, n2, n2, n1, -1., A.sub(n1, 0), A.sub(n1, 0), 1.,
              A.sub(n1, n1)));
  return potrf_rec(h, A.sub(n1, n1), n2, dinv + (n1 / LEAF) * LEAF * LEAF, offset + n1, d_bad,
                   floor != nullptr ? floor + n1 : nullptr);
}

// Right-looking blocked Cholesky with look-ahead 1 on two streams.
//
// The recursion above keeps every O(n^3) flop in large GEMMs, but its bottom levels are a chain of
// ~14 000 dependent single-CTA kernels (64 x 64 leaves, 64-wide triangular solves) during which 147
// of the 148 SMs idle: 10-12 % of the factorisation at N = 65 536 (profiles/r01_ncu_launches_*).
// Here that chain — factor the nb x nb diagonal block, solve the block column below it — runs on a
// high-priority stream WHILE the previous panel's trailing DSYRK fills the machine:
//
//   step k (panel k already factored, block columns > k hold the updates of panels < k):
//     S: update block column k+1 with panel k            (two GEMMs, r x nb x nb)
//     P: wait for it; factor panel k+1                   (leaf / small-GEMM chain, 1-2 % of the SMs)
//     S: update block columns k+2.. with panel k         (one GEMM_LOWER, (r-nb)^2 x nb)  } concurrent
//
// The chain is exposed only for the first panel and once the trailing matrix is too small to cover
// it (r < ~10 000).  Same flops, same kernels, same results as the recursion to rounding.
#ifndef AB_POTRF_NB
#define AB_POTRF_NB 2048
#endif
constexpr int64_t LA_NB = AB_POTRF_NB; // panel width (multiple of LEAF)
constexpr int64_t LA_MIN_N = 4 * LA_NB; // below this the plain recursion is used
static_assert(LA_NB % LEAF == 0, "panel width");

int ensure_panel_stream(ab_handle_s *h) {
  if (h->panel_stream != nullptr) {
    return AB_OK;
  }
  int lo = 0, hi = 0;
  AB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  AB_CUDA(cudaStreamCreateWithPriority(&h->panel_stream, cudaStreamNonBlocking, hi));
  AB_CUDA(cudaEventCreateWithFlags(&h->ev_panel, cudaEventDisableTiming));
  AB_CUDA(cudaEventCreateWithFlags(&h->ev_col, cudaEventDisableTiming));
  return AB_OK;
}

// Every helper launches on h->stream; the panel chain borrows it for the scope.
struct StreamScope {
  StreamScope(ab_handle_s *h, cudaStream_t s) : h_(h), saved_(h->stream) { h->stream = s; }
  ~StreamScope() { h_->stream = saved_; }
  ab_handle_s *h_;
  cudaStream_t saved_;
};

static int factor_panel(ab_handle_s *h, MatView A, int64_t n, int64_t k0, int64_t w, double *dinv,
                        int *d_bad, const double *floor) {
  StreamScope scope(h, h->panel_stream);
  double *dk = dinv + (k0 / LEAF) * LEAF * LEAF;
  AB_TRY(potrf_rec(h, A.sub(k0, k0), w, dk, k0, d_bad, floor != nullptr ? floor + k0 : nullptr));
  const int64_t below = n - k0 - w;
  if (below > 0) {
    AB_TRY(trsm_right_lower_T(h, A.sub(k0, k0), dk, w, A.sub(k0 + w, k0), below));
  }
  AB_CUDA(cudaEventRecord(h->ev_panel, h->panel_stream));
  return AB_OK;
}

static int potrf_lookahead(ab_handle_s *h, MatView A, int64_t n, double *dinv, int *d_bad,
                           const double *floor) {
  AB_TRY(ensure_panel_stream(h));
  cudaStream_t S = h->stream;
  // the panel stream starts behind everything already enqueued on S (the Gram build of A)
  AB_CUDA(cudaEventRecord(h->ev_col, S));
  AB_CUDA(cudaStreamWaitEvent(h->panel_stream, h->ev_col, 0));
  int status = factor_panel(h, A, n, 0, std::min(LA_NB, n), dinv, d_bad, floor);
  for (int64_t k0 = 0; status == AB_OK && k0 + LA_NB < n; k0 += LA_NB) {
    const int64_t w = LA_NB;
    const int64_t next = k0 + w;
    const int64_t wn = std::min(LA_NB, n - next);
    const int64_t rest = n - next - wn;
    AB_CUDA(cudaStreamWaitEvent(S, h->ev_panel, 0)); // panel k is factored
    const MatView Pk = A.sub(next, k0);              // rows next.., the w columns of panel k
    // block column k+1 first ...
    status = gemm(h, GEMM_TRANS_B | GEMM_LOWER, wn, wn, w, -1., Pk, Pk, 1., A.sub(next, next));
    if (status == AB_OK && rest > 0) {
      status = gemm(h, GEMM_TRANS_B, rest, wn, w, -1., A.sub(next + wn, k0), Pk, 1.,
                    A.sub(next + wn, next));
    }
    if (status != AB_OK) {
      break;
    }
    AB_CUDA(cudaEventRecord(h->ev_col, S));
    AB_CUDA(cudaStreamWaitEvent(h->panel_stream, h->ev_col, 0));
    // ... so that its factorisation overlaps the rest of this update
    status = factor_panel(h, A, n, next, wn, dinv, d_bad, floor);
    if (status == AB_OK && rest > 0) {
      const MatView Pr = A.sub(next + wn, k0);
      status = gemm(h, GEMM_TRANS_B | GEMM_LOWER, rest, rest, w, -1., Pr, Pr, 1.,
                    A.sub(next + wn, next + wn));
    }
  }
  // S continues only after the last panel; on an error the streams are still joined
  cudaStreamWaitEvent(S, h->ev_panel, 0);
  return status;
}

// floor[i] = PIVOT_RTOL * A(i, i)
constexpr double PIVOT_RTOL = 64. * 2.220446049250313e-16;
__global__ void pivot_floor_kernel(const double *A, int64_t ld, int64_t n, double *floor) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    floor[i] = PIVOT_RTOL * A[i + i * ld];
  }
}

int pivot_floor(ab_handle_s *h, MatView A, int64_t n, double *d_floor) {
  if (n <= 0) {
    return AB_OK;
  }
  pivot_floor_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(A.p, A.ld, n, d_floor);
  AB_LAUNCHED(h);
  return AB_OK;
}

int potrf(ab_handle_s *h, MatView A, int64_t n, double *dinv, int *d_bad, const double *d_floor) {
  if (n <= 0) {
    return AB_OK;
  }
  AB_REQUIRE(n < INT_MAX, "matrix too large");
  // Pivot floor.  An unpivoted factorisation of a numerically singular matrix (duplicate points without a
  // noise term) does not reliably meet a pivot <= 0: exact cancellation leaves +-1e-17, and a tiny POSITIVE
  // pivot lets the factorisation "succeed" with garbage.  The reference's pivoted LDLT ends with (near) zero
  // pivots that its callers see through is_positive_definite() / log_determinant().  Here a pivot must exceed
  // 64 eps times the diagonal entry the matrix had BEFORE elimination, else AB_ERR_NOT_PD (condition numbers
  // beyond ~1e13 are rejected rather than answered with noise).  Callers that factor a block which was already
  // updated (dist.cu, ab_gp_update) pass the floor of the original diagonal.
  void *own = nullptr;
  const size_t bytes = static_cast<size_t>(n) * sizeof(double);
  if (d_floor == nullptr) {
    AB_TRY(dev_alloc(h, bytes, &own));
    int s = pivot_floor(h, A, n, static_cast<double *>(own));
    if (s != AB_OK) {
      dev_release(h, own, bytes);
      return s;
    }
    d_floor = static_cast<double *>(own);
  }
  // AB_POTRF_RECURSIVE / AB_POTRF_LOOKAHEAD_MIN: test hooks (tests/test_gpu_gp.py compares the two
  // schedules at sizes below the default threshold)
  int64_t min_n = LA_MIN_N;
  if (const char *e = std::getenv("AB_POTRF_LOOKAHEAD_MIN")) {
    min_n = std::atoll(e);
  }
  int status;
  if (n >= min_n && std::getenv("AB_POTRF_RECURSIVE") == nullptr) {
    status = potrf_lookahead(h, A, n, dinv, d_bad, d_floor);
  } else {
    status = potrf_rec(h, A, n, dinv, 0, d_bad, d_floor);
  }
  if (own != nullptr) {
    dev_release(h, own, bytes); // stream-ordered reuse: later users are enqueued behind the factorisation
  }
  return status;
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
