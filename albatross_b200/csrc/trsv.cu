// One-right-hand-side triangular solves and the matrix-vector products that feed them (HBM-bound).
//
// The solves of fit / log_likelihood (information = K^-1 y, the quadratic form) replace Eigen's
// LDLT::_solve_impl (reference third_party/eigen/Eigen/src/Cholesky/LDLT.h:558-592) for a single column.
// Round 1 ran them through the GEMM-shaped recursion of linalg.cu: ~4000 dependent single-CTA launches per
// solve at N = 65 536 (41 ms for 34 GB of reads, 14 % of HBM peak in the GEMV that carried them).  Here a
// solve is n / 512 steps of two kernels:
//   * trsv_block_kernel: one CTA solves a diagonal block (512 rows on one GPU, the 1024-row block columns
//     of the distributed factor), the right-hand side living in shared memory; its 64 x 64 leaves are
//     multiplications by the explicit leaf inverses the factorisation already produced (dinv), the rest of
//     the block is streamed once through a cp.async staging buffer;
//   * gemv_n_kernel / gemv_t_kernel: the panel below the block times the solved segment (forward) or its
//     transpose times the solved tail (backward), 16-byte loads, eight of them in flight per thread, k-split
//     over the grid with a deterministic second-pass reduction when the matrix is short and fat.
// The distributed solves (dist.cu) are built from the same three kernels.
#include "linalg.cuh"

#include <algorithm>

namespace ab {

constexpr int TRSV_BLOCK = 1024; // largest diagonal block one CTA solves (multiple of LEAF)
// Block size of the single-GPU substitution loops: the in-block work of the lone CTA grows with the square
// of the block, the number of (launch-latency bound) steps falls with it; 512 balances the two.
constexpr int TRSV_STEP = 512;
constexpr int TRSV_THREADS = 512;
static_assert(TRSV_BLOCK % LEAF == 0, "block size");

// Solves L x = b (TRANS == false) or L^T x = b (TRANS == true) for one diagonal block of `nb` <= 1024 rows,
// in place in `x`.  L: the block's lower triangle (column-major, leading dimension ld); dinv: the explicit
// inverses of its LEAF x LEAF diagonal leaves (identity-padded for a ragged last leaf).
//
// The CTA runs alone on one SM, so its speed is set by how many loads it keeps in flight.  Plain loads do
// not work: ptxas interleaves them two by two with the FMAs that consume them (to save registers), the
// in-order warps stall on every pair and the block streams at ~10 GB/s (measured: 41 ms per N = 65 536
// solve).  The panel slice below / beside a leaf is therefore staged through shared memory with cp.async —
// fire-and-forget copies, 32 per thread per chunk, all in flight at once — and consumed from there.
constexpr int TRSV_CHUNK = 256;                     // rows per staged chunk
constexpr int TRSV_STAGE = TRSV_CHUNK * LEAF;       // doubles: 128 KB

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// stage[c * TRSV_CHUNK + r] = L[row0 + r, col0 + c] for r < rows (<= TRSV_CHUNK), c < LEAF
__device__ __forceinline__ void stage_chunk(double *stage, const double *__restrict__ L, int64_t ld,
                                            int row0, int col0, int rows, int tid) {
  for (int idx = tid; idx < TRSV_STAGE; idx += TRSV_THREADS) {
    const int r = idx % TRSV_CHUNK;
    const int c = idx / TRSV_CHUNK;
    if (r < rows) {
      cp_async8(stage + idx, L + (row0 + r) + static_cast<int64_t>(col0 + c) * ld);
    }
  }
  cp_async_wait_all();
  __syncthreads();
}

template <bool TRANS>
__global__ void __launch_bounds__(TRSV_THREADS)
trsv_block_kernel(const double *__restrict__ L, int64_t ld, const double *__restrict__ dinv, int nb,
                  double *x) {
  extern __shared__ __align__(16) double trsv_smem[];
  double *stage = trsv_smem;              // TRSV_STAGE
  double *xs = stage + TRSV_STAGE;        // TRSV_BLOCK
  double *ys = xs + TRSV_BLOCK;           // LEAF
  double *part = ys + LEAF;               // TRSV_THREADS
  const int tid = threadIdx.x;
  for (int i = tid; i < TRSV_BLOCK; i += TRSV_THREADS) {
    xs[i] = i < nb ? x[i] : 0.;
  }
  __syncthreads();
  const int nleaf = (nb + LEAF - 1) / LEAF;
  if (!TRANS) {
    for (int leaf = 0; leaf < nleaf; ++leaf) {
      const int c0 = leaf * LEAF;
      const double *inv = dinv + static_cast<int64_t>(leaf) * LEAF * LEAF;
      // y = inv * xs[c0 .. c0 + 64): the 64 x 64 inverse goes through the stage as well
      for (int idx = tid; idx < LEAF * LEAF; idx += TRSV_THREADS) {
        cp_async8(stage + idx, inv + idx);
      }
      cp_async_wait_all();
      __syncthreads();
      { // thread (r, q): row r, columns q * 8 .. q * 8 + 8 (8 partial sums per row)
        const int r = tid & (LEAF - 1), q = tid >> 6;
        double acc = 0.;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc = fma(stage[r + (q * 8 + c) * LEAF], xs[c0 + q * 8 + c], acc);
        }
        part[tid] = acc;
      }
      __syncthreads();
      if (tid < LEAF) {
        double acc = 0.;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          acc += part[tid + q * LEAF];
        }
        ys[tid] = acc;
        xs[c0 + tid] = acc;
      }
      __syncthreads();
      // rows below the leaf inside the block: xs[r] -= L[r, c0 .. c0 + 64) . y, TRSV_CHUNK rows at a time
      for (int r0 = c0 + LEAF; r0 < nb; r0 += TRSV_CHUNK) {
        const int rows = min(TRSV_CHUNK, nb - r0);
        stage_chunk(stage, L, ld, r0, c0, rows, tid);
        const int r = tid % TRSV_CHUNK, hlf = tid / TRSV_CHUNK; // two threads per row, 32 columns each
        double a0 = 0., a1 = 0.;
        if (r < rows) {
#pragma unroll
          for (int c = 0; c < LEAF / 2; c += 2) {
            const int cc = hlf * (LEAF / 2) + c;
            a0 = fma(stage[cc * TRSV_CHUNK + r], ys[cc], a0);
            a1 = fma(stage[(cc + 1) * TRSV_CHUNK + r], ys[cc + 1], a1);
          }
        }
        part[tid] = a0 + a1;
        __syncthreads();
        if (tid < rows) {
          xs[r0 + tid] -= part[tid] + part[tid + TRSV_CHUNK];
        }
        __syncthreads();
      }
    }
  } else {
    for (int leaf = nleaf - 1; leaf >= 0; --leaf) {
      const int c0 = leaf * LEAF;
      const double *inv = dinv + static_cast<int64_t>(leaf) * LEAF * LEAF;
      // t[c] = sum_{r >= c0 + 64} L[r, c0 + c] xs[r], accumulated chunk by chunk in ys (starting at 0)
      if (tid < LEAF) {
        ys[tid] = 0.;
      }
      __syncthreads();
      for (int r0 = c0 + LEAF; r0 < nb; r0 += TRSV_CHUNK) {
        const int rows = min(TRSV_CHUNK, nb - r0);
        stage_chunk(stage, L, ld, r0, c0, rows, tid);
        // a warp owns 4 columns; its lanes walk 32 consecutive rows at a time (conflict-free), then reduce
        const int lane = tid & 31, warp = tid >> 5;
        double acc[4] = {0., 0., 0., 0.};
#pragma unroll
        for (int g = 0; g < TRSV_CHUNK / 32; ++g) {
          const int rr = g * 32 + lane;
          const bool in = rr < rows; // rows beyond the chunk hold stale staging data
          const double xv = in ? xs[r0 + rr] : 0.;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const double lv = in ? stage[(warp * 4 + qq) * TRSV_CHUNK + rr] : 0.;
            acc[qq] = fma(lv, xv, acc[qq]);
          }
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          double t = acc[qq];
          for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_xor_sync(0xffffffffu, t, o);
          }
          if (lane == 0) {
            ys[warp * 4 + qq] += t;
          }
        }
        __syncthreads();
      }
      // x_leaf = inv^T (xs_leaf - t): thread (c, q): column c of inv (contiguous), rows q * 8 .. q * 8 + 8
      for (int idx = tid; idx < LEAF * LEAF; idx += TRSV_THREADS) {
        cp_async8(stage + idx, inv + idx);
      }
      cp_async_wait_all();
      if (tid < LEAF) {
        ys[tid] = xs[c0 + tid] - ys[tid];
      }
      __syncthreads();
      {
        const int q = tid & 7, c = tid >> 3;
        double acc = 0.;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc = fma(stage[c * LEAF + q * 8 + i], ys[q * 8 + i], acc);
        }
        part[tid] = acc;
      }
      __syncthreads();
      if (tid < LEAF) {
        double total = 0.;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          total += part[tid * 8 + g];
        }
        xs[c0 + tid] = total;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < nb; i += TRSV_THREADS) {
    x[i] = xs[i];
  }
}

constexpr size_t TRSV_SMEM = (TRSV_STAGE + TRSV_BLOCK + LEAF + TRSV_THREADS) * sizeof(double);

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
  static PerDeviceOnce once;
  if (once.need(h->device)) {
    AB_CUDA(cudaFuncSetAttribute(trsv_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(TRSV_SMEM)));
    AB_CUDA(cudaFuncSetAttribute(trsv_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(TRSV_SMEM)));
  }
  if (trans) {
    trsv_block_kernel<true><<<1, TRSV_THREADS, TRSV_SMEM, h->stream>>>(L.p, L.ld, dinv,
                                                                        static_cast<int>(nb), x);
  } else {
    trsv_block_kernel<false><<<1, TRSV_THREADS, TRSV_SMEM, h->stream>>>(L.p, L.ld, dinv,
                                                                         static_cast<int>(nb), x);
  }
  AB_LAUNCHED(h);
  return AB_OK;
}

// x <- L^-1 x: right-looking over TRSV_STEP-row blocks.
int trsv_lower(ab_handle_s *h, MatView L, const double *dinv, int64_t n, double *x) {
  for (int64_t k0 = 0; k0 < n; k0 += TRSV_STEP) {
    const int64_t w = std::min<int64_t>(TRSV_STEP, n - k0);
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
  for (int64_t k0 = (n - 1) / TRSV_STEP * TRSV_STEP; k0 >= 0; k0 -= TRSV_STEP) {
    const int64_t w = std::min<int64_t>(TRSV_STEP, n - k0);
    const int64_t below = n - k0 - w;
    if (below > 0) {
      AB_TRY(gemv_t(h, below, w, -1., L.sub(k0 + w, k0), x + k0 + w, 1., x + k0));
    }
    AB_TRY(trsv_block(h, true, L.sub(k0, k0), dinv + (k0 / LEAF) * LEAF * LEAF, w, x + k0));
  }
  return AB_OK;
}

} // namespace ab
