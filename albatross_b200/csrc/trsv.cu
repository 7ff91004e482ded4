// One-right-hand-side triangular solves and the matrix-vector products that feed them (HBM-bound).
//
// The solves of fit / log_likelihood (information = K^-1 y, the quadratic form) replace Eigen's
// LDLT::_solve_impl (reference third_party/eigen/Eigen/src/Cholesky/LDLT.h:558-592) for a single column.
// Round 1 ran them through the GEMM-shaped recursion of linalg.cu: ~4000 dependent single-CTA launches per
// solve at N = 65 536 (41 ms for 34 GB of reads, 14 % of HBM peak in the GEMV that carried them).  Here a
// solve is n / 1024 steps of two kernels:
//   * trsv_block_kernel: one CTA solves a 1024 x 1024 diagonal block, the right-hand side living in shared
//     memory; its 64 x 64 leaves are multiplications by the explicit leaf inverses the factorisation already
//     produced (dinv), the rest of the block is streamed once;
//   * gemv_n_kernel / gemv_t_kernel: the panel below the block times the solved segment (forward) or its
//     transpose times the solved tail (backward), 16-byte loads, eight of them in flight per thread, k-split
//     over the grid with a deterministic second-pass reduction when the matrix is short and fat.
// The distributed solves (dist.cu) are built from the same three kernels.
#include "linalg.cuh"

#include <algorithm>

namespace ab {

constexpr int TRSV_BLOCK = 1024; // diagonal block solved by one CTA (multiple of LEAF)
constexpr int TRSV_THREADS = 512;
static_assert(TRSV_BLOCK % LEAF == 0, "block size");

// Solves L x = b (TRANS == false) or L^T x = b (TRANS == true) for one diagonal block of `nb` <= 1024 rows,
// in place in `x`.  L: the block's lower triangle (column-major, leading dimension ld); dinv: the explicit
// inverses of its LEAF x LEAF diagonal leaves (identity-padded for a ragged last leaf).
template <bool TRANS>
__global__ void __launch_bounds__(TRSV_THREADS)
trsv_block_kernel(const double *__restrict__ L, int64_t ld, const double *__restrict__ dinv, int nb,
                  double *x) {
  __shared__ double xs[TRSV_BLOCK];
  __shared__ double ys[LEAF];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  constexpr int NWARPS = TRSV_THREADS / 32;
  for (int i = tid; i < TRSV_BLOCK; i += TRSV_THREADS) {
    xs[i] = i < nb ? x[i] : 0.;
  }
  __syncthreads();
  const int nleaf = (nb + LEAF - 1) / LEAF;
  if (!TRANS) {
    for (int leaf = 0; leaf < nleaf; ++leaf) {
      const int c0 = leaf * LEAF;
      const double *inv = dinv + static_cast<int64_t>(leaf) * LEAF * LEAF;
      // y = inv * xs[c0 .. c0 + 64): thread r < 64 owns row r (coalesced along r for every column)
      if (tid < LEAF) {
        double acc = 0.;
#pragma unroll 8
        for (int c = 0; c < LEAF; ++c) {
          acc = fma(inv[tid + c * LEAF], xs[c0 + c], acc);
        }
        ys[tid] = acc;
      }
      __syncthreads();
      if (tid < LEAF) {
        xs[c0 + tid] = ys[tid];
      }
      // rows below the leaf inside the block: xs[r] -= L[r, c0 .. c0 + 64) . y
      const int r0 = c0 + LEAF;
      for (int r = r0 + tid; r < nb; r += TRSV_THREADS) {
        const double *row = L + r + static_cast<int64_t>(c0) * ld;
        double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
#pragma unroll 4
        for (int c = 0; c < LEAF; c += 4) {
          a0 = fma(row[static_cast<int64_t>(c) * ld], ys[c], a0);
          a1 = fma(row[static_cast<int64_t>(c + 1) * ld], ys[c + 1], a1);
          a2 = fma(row[static_cast<int64_t>(c + 2) * ld], ys[c + 2], a2);
          a3 = fma(row[static_cast<int64_t>(c + 3) * ld], ys[c + 3], a3);
        }
        xs[r] -= (a0 + a1) + (a2 + a3);
      }
      __syncthreads();
    }
  } else {
    for (int leaf = nleaf - 1; leaf >= 0; --leaf) {
      const int c0 = leaf * LEAF;
      const int r0 = c0 + LEAF;
      const double *inv = dinv + static_cast<int64_t>(leaf) * LEAF * LEAF;
      // t[c] = sum_{r >= r0} L[r, c0 + c] xs[r]: one warp per column (contiguous rows), 4 columns per warp
      for (int c = warp; c < LEAF; c += NWARPS) {
        const double *col = L + static_cast<int64_t>(c0 + c) * ld;
        double acc = 0.;
        for (int r = r0 + lane; r < nb; r += 32) {
          acc = fma(col[r], xs[r], acc);
        }
        for (int o = 16; o > 0; o >>= 1) {
          acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane == 0) {
          ys[c] = xs[c0 + c] - acc;
        }
      }
      __syncthreads();
      // x_leaf = inv^T ys: thread c owns column c of inv (contiguous)
      if (tid < LEAF) {
        const double *col = inv + tid * LEAF;
        double acc = 0.;
#pragma unroll 8
        for (int r = 0; r < LEAF; ++r) {
          acc = fma(col[r], ys[r], acc);
        }
        xs[c0 + tid] = acc;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < nb; i += TRSV_THREADS) {
    x[i] = xs[i];
  }
}

// ---- y[m] = beta y + alpha A[m x k] x[k] ---------------------------------------------------------------
// CTA = 8 warps as WR x WC: a warp covers 64 consecutive rows with one 16-byte load per lane and column,
// the WC warp columns take every WC-th column of this CTA's k range; eight loads in flight per thread.
// grid = (row tiles, k splits).  ksplit == 1: the result goes straight to y; otherwise partial sums go to
// part[split][m] and gemv_reduce_kernel finishes (fixed summation order: deterministic).
constexpr int GV_WR = 2, GV_WC = 4;
constexpr int GV_ROWS = 64 * GV_WR;

__global__ void __launch_bounds__(256)
gemv_n_kernel(int64_t m, int64_t k, double alpha, const double *__restrict__ A, int64_t lda,
              const double *__restrict__ x, double beta, double *y, double *part, int64_t kchunk) {
  __shared__ double2 red[GV_WC][GV_WR * 32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wr = warp % GV_WR;
  const int wc = warp / GV_WR;
  const int64_t row = blockIdx.x * static_cast<int64_t>(GV_ROWS) + wr * 64 + 2 * lane;
  const int64_t k0 = blockIdx.y * kchunk;
  const int64_t k1 = k0 + kchunk < k ? k0 + kchunk : k;
  double2 acc[4] = {{0., 0.}, {0., 0.}, {0., 0.}, {0., 0.}};
  if (row + 1 < m) {
    const double *a = A + row;
    int64_t kk = k0 + wc;
    for (; kk + 7 * GV_WC < k1; kk += 8 * GV_WC) {
      double2 v[8];
      double s[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = *reinterpret_cast<const double2 *>(a + (kk + u * GV_WC) * lda);
        s[u] = x[kk + u * GV_WC];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc[u & 3].x = fma(v[u].x, s[u], acc[u & 3].x);
        acc[u & 3].y = fma(v[u].y, s[u], acc[u & 3].y);
      }
    }
    for (; kk < k1; kk += GV_WC) {
      const double2 v = *reinterpret_cast<const double2 *>(a + kk * lda);
      const double s = x[kk];
      acc[0].x = fma(v.x, s, acc[0].x);
      acc[0].y = fma(v.y, s, acc[0].y);
    }
  } else if (row < m) { // last (odd) row of the matrix
    const double *a = A + row;
    for (int64_t kk = k0 + wc; kk < k1; kk += GV_WC) {
      acc[0].x = fma(a[kk * lda], x[kk], acc[0].x);
    }
  }
  red[wc][wr * 32 + lane] = make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x),
                                         (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
  __syncthreads();
  if (wc == 0 && row < m) {
    double2 total = make_double2(0., 0.);
#pragma unroll
    for (int w = 0; w < GV_WC; ++w) {
      total.x += red[w][wr * 32 + lane].x;
      total.y += red[w][wr * 32 + lane].y;
    }
    if (part != nullptr) {
      double *p = part + blockIdx.y * m + row;
      p[0] = total.x;
      if (row + 1 < m) {
        p[1] = total.y;
      }
      return;
    }
    double v0 = alpha * total.x, v1 = alpha * total.y;
    if (beta != 0.) {
      v0 = fma(beta, y[row], v0);
    }
    y[row] = v0;
    if (row + 1 < m) {
      if (beta != 0.) {
        v1 = fma(beta, y[row + 1], v1);
      }
      y[row + 1] = v1;
    }
  }
}

// y[i] = beta y[i] + alpha sum_s part[s][i]
__global__ void gemv_reduce_kernel(int64_t m, int nsplit, double alpha, const double *__restrict__ part,
                                   double beta, double *y) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= m) {
    return;
  }
  double total = 0.;
  for (int s = 0; s < nsplit; ++s) {
    total += part[s * m + i];
  }
  double v = alpha * total;
  if (beta != 0.) {
    v = fma(beta, y[i], v);
  }
  y[i] = v;
}

// ---- y[k] = beta y + alpha A[m x k]^T x[m] -------------------------------------------------------------
// A column is contiguous.  Tall (m large): one CTA per column, 16-byte loads, four accumulators.
__global__ void __launch_bounds__(256)
gemv_t_tall_kernel(int64_t m, double alpha, const double *__restrict__ A, int64_t lda,
                   const double *__restrict__ x, double beta, double *y) {
  __shared__ double red[8];
  const double *a = A + blockIdx.x * lda;
  double acc0 = 0., acc1 = 0., acc2 = 0., acc3 = 0.;
  const int64_t m2 = m & ~int64_t(1);
  int64_t i = 2 * static_cast<int64_t>(threadIdx.x);
  for (; i + 1536 < m2; i += 2048) {
    const double2 v0 = *reinterpret_cast<const double2 *>(a + i);
    const double2 v1 = *reinterpret_cast<const double2 *>(a + i + 512);
    const double2 v2 = *reinterpret_cast<const double2 *>(a + i + 1024);
    const double2 v3 = *reinterpret_cast<const double2 *>(a + i + 1536);
    const double2 x0 = *reinterpret_cast<const double2 *>(x + i);
    const double2 x1 = *reinterpret_cast<const double2 *>(x + i + 512);
    const double2 x2 = *reinterpret_cast<const double2 *>(x + i + 1024);
    const double2 x3 = *reinterpret_cast<const double2 *>(x + i + 1536);
    acc0 = fma(v0.x, x0.x, fma(v0.y, x0.y, acc0));
    acc1 = fma(v1.x, x1.x, fma(v1.y, x1.y, acc1));
    acc2 = fma(v2.x, x2.x, fma(v2.y, x2.y, acc2));
    acc3 = fma(v3.x, x3.x, fma(v3.y, x3.y, acc3));
  }
  for (; i < m2; i += 512) {
    const double2 v0 = *reinterpret_cast<const double2 *>(a + i);
    const double2 x0 = *reinterpret_cast<const double2 *>(x + i);
    acc0 = fma(v0.x, x0.x, fma(v0.y, x0.y, acc0));
  }
  if (threadIdx.x == 0 && m2 < m) {
    acc1 = fma(a[m2], x[m2], acc1);
  }
  double v = (acc0 + acc1) + (acc2 + acc3);
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      total += red[w];
    }
    double out = alpha * total;
    if (beta != 0.) {
      out = fma(beta, y[blockIdx.x], out);
    }
    y[blockIdx.x] = out;
  }
}

// Short and fat (m <= a few thousand rows, many columns): one warp per column.
__global__ void __launch_bounds__(256)
gemv_t_warp_kernel(int64_t m, int64_t k, double alpha, const double *__restrict__ A, int64_t lda,
                   const double *__restrict__ x, double beta, double *y) {
  const int lane = threadIdx.x & 31;
  const int64_t col = blockIdx.x * static_cast<int64_t>(8) + (threadIdx.x >> 5);
  if (col >= k) {
    return;
  }
  const double *a = A + col * lda;
  double acc0 = 0., acc1 = 0.;
  const int64_t m2 = m & ~int64_t(1);
  int64_t i = 2 * static_cast<int64_t>(lane);
  for (; i + 64 < m2; i += 128) {
    const double2 v0 = *reinterpret_cast<const double2 *>(a + i);
    const double2 v1 = *reinterpret_cast<const double2 *>(a + i + 64);
    const double2 x0 = *reinterpret_cast<const double2 *>(x + i);
    const double2 x1 = *reinterpret_cast<const double2 *>(x + i + 64);
    acc0 = fma(v0.x, x0.x, fma(v0.y, x0.y, acc0));
    acc1 = fma(v1.x, x1.x, fma(v1.y, x1.y, acc1));
  }
  for (; i < m2; i += 64) {
    const double2 v0 = *reinterpret_cast<const double2 *>(a + i);
    const double2 x0 = *reinterpret_cast<const double2 *>(x + i);
    acc0 = fma(v0.x, x0.x, fma(v0.y, x0.y, acc0));
  }
  if (lane == 0 && m2 < m) {
    acc1 = fma(a[m2], x[m2], acc1);
  }
  double v = acc0 + acc1;
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if (lane == 0) {
    double out = alpha * v;
    if (beta != 0.) {
      out = fma(beta, y[col], out);
    }
    y[col] = out;
  }
}

static bool aligned16(const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

bool gemv_fast_ok(MatView A, const double *x) {
  return aligned16(A.p) && A.ld % 2 == 0 && aligned16(x);
}

int gemv_n(ab_handle_s *h, int64_t m, int64_t k, double alpha, MatView A, const double *x, double beta,
           double *y) {
  if (m <= 0) {
    return AB_OK;
  }
  const int64_t row_tiles = (m + GV_ROWS - 1) / GV_ROWS;
  // enough CTAs to fill the machine: split k when the matrix is short
  int64_t nsplit = 1;
  const int64_t want = 4 * static_cast<int64_t>(h->sm_count);
  if (row_tiles < want && k >= 1024) {
    nsplit = std::min<int64_t>({(want + row_tiles - 1) / row_tiles, k / 256, 64});
    nsplit = std::max<int64_t>(nsplit, 1);
  }
  const int64_t kchunk = round_up((k + nsplit - 1) / nsplit, 8 * GV_WC);
  nsplit = (k + kchunk - 1) / kchunk;
  if (nsplit <= 1) {
    gemv_n_kernel<<<dim3(static_cast<unsigned>(row_tiles), 1), 256, 0, h->stream>>>(
        m, k, alpha, A.p, A.ld, x, beta, y, nullptr, k > 0 ? k : 1);
    AB_LAUNCHED(h);
    return AB_OK;
  }
  void *part = nullptr;
  const size_t bytes = static_cast<size_t>(nsplit) * static_cast<size_t>(m) * sizeof(double);
  AB_TRY(dev_alloc(h, bytes, &part));
  gemv_n_kernel<<<dim3(static_cast<unsigned>(row_tiles), static_cast<unsigned>(nsplit)), 256, 0,
                  h->stream>>>(m, k, alpha, A.p, A.ld, x, beta, y, static_cast<double *>(part), kchunk);
  AB_LAUNCHED(h);
  gemv_reduce_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, h->stream>>>(
      m, static_cast<int>(nsplit), alpha, static_cast<double *>(part), beta, y);
  AB_LAUNCHED(h);
  dev_release(h, part, bytes); // stream-ordered reuse: later users are enqueued behind the reduction
  return AB_OK;
}

int gemv_t(ab_handle_s *h, int64_t m, int64_t k, double alpha, MatView A, const double *x, double beta,
           double *y) {
  if (k <= 0) {
    return AB_OK;
  }
  if (m >= 8192) {
    gemv_t_tall_kernel<<<static_cast<unsigned>(k), 256, 0, h->stream>>>(m, alpha, A.p, A.ld, x, beta, y);
  } else {
    gemv_t_warp_kernel<<<static_cast<unsigned>((k + 7) / 8), 256, 0, h->stream>>>(m, k, alpha, A.p, A.ld,
                                                                               x, beta, y);
  }
  AB_LAUNCHED(h);
  return AB_OK;
}

int trsv_block(ab_handle_s *h, bool trans, MatView L, const double *dinv, int64_t nb, double *x) {
  if (nb <= 0) {
    return AB_OK;
  }
  AB_REQUIRE(nb <= TRSV_BLOCK, "trsv block too large");
  if (trans) {
    trsv_block_kernel<true><<<1, TRSV_THREADS, 0, h->stream>>>(L.p, L.ld, dinv, static_cast<int>(nb), x);
  } else {
    trsv_block_kernel<false><<<1, TRSV_THREADS, 0, h->stream>>>(L.p, L.ld, dinv, static_cast<int>(nb), x);
  }
  AB_LAUNCHED(h);
  return AB_OK;
}

// x <- L^-1 x: right-looking over 1024-row blocks.
int trsv_lower(ab_handle_s *h, MatView L, const double *dinv, int64_t n, double *x) {
  for (int64_t k0 = 0; k0 < n; k0 += TRSV_BLOCK) {
    const int64_t w = std::min<int64_t>(TRSV_BLOCK, n - k0);
    AB_TRY(trsv_block(h, false, L.sub(k0, k0), dinv + (k0 / LEAF) * LEAF * LEAF, w, x + k0));
    const int64_t below = n - k0 - w;
    if (below > 0) {
      AB_TRY(gemv_n(h, below, w, -1., L.sub(k0 + w, k0), x + k0, 1., x + k0 + w));
    }
  }
  return AB_OK;
}

// x <- L^-T x: block k needs the dot products of its columns with the solved tail.
int trsv_lower_T(ab_handle_s *h, MatView L, const double *dinv, int64_t n, double *x) {
  if (n <= 0) {
    return AB_OK;
  }
  for (int64_t k0 = (n - 1) / TRSV_BLOCK * TRSV_BLOCK; k0 >= 0; k0 -= TRSV_BLOCK) {
    const int64_t w = std::min<int64_t>(TRSV_BLOCK, n - k0);
    const int64_t below = n - k0 - w;
    if (below > 0) {
      AB_TRY(gemv_t(h, below, w, -1., L.sub(k0 + w, k0), x + k0 + w, 1., x + k0));
    }
    AB_TRY(trsv_block(h, true, L.sub(k0, k0), dinv + (k0 / LEAF) * LEAF * LEAF, w, x + k0));
  }
  return AB_OK;
}

} // namespace ab
