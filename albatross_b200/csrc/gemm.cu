// fp64 GEMM family on the FP64 tensor pipe (DMMA m8n8k4), cp.async multi-stage pipeline.
//
//   C[m x n] = alpha * op(A) * op(B) + beta * C        (column-major)
//
// This is the trailing-update engine of the blocked Cholesky (DSYRK / DGEMM), of the TRSM
// recursions, the predictive-covariance SYRK and the triangular inverse used by LOO-CV.  It replaces
// the GEMV-bound inner loops of Eigen's unblocked LDLT (reference third_party/eigen/Eigen/src/
// Cholesky/LDLT.h:349-355) and Eigen's GEBP triangular solves (LDLT.h:558-592).
//
// sm_100a has no tcgen05 kind for f64; `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4) is the FP64 tensor
// instruction.  The MMA's row dimension is mapped to n and its column dimension to m so that every
// thread owns two CONSECUTIVE rows of C per fragment and the epilogue is 16-byte vectorised.
#include "linalg.cuh"

#include <cstdlib>

namespace ab {

// Tile configuration (overridable at compile time for tools/gemm_sweep.sh; the defaults are the
// measured best).  CTA tile BM x BN x BK, WARPS_M x WARPS_N warps, each owning a
// (BM / WARPS_M) x (BN / WARPS_N) patch as 8 x 8 DMMA fragments.
#ifndef AB_GEMM_BM
#define AB_GEMM_BM 128
#endif
#ifndef AB_GEMM_BN
#define AB_GEMM_BN 64
#endif
#ifndef AB_GEMM_BK
#define AB_GEMM_BK 16
#endif
#ifndef AB_GEMM_STAGES
#define AB_GEMM_STAGES 3
#endif
#ifndef AB_GEMM_WARPS_M
#define AB_GEMM_WARPS_M 4
#endif
#ifndef AB_GEMM_WARPS_N
#define AB_GEMM_WARPS_N 2
#endif
#ifndef AB_GEMM_MIN_CTAS
#define AB_GEMM_MIN_CTAS 2
#endif
constexpr int BM = AB_GEMM_BM;
constexpr int BN = AB_GEMM_BN;
constexpr int BK = AB_GEMM_BK;
constexpr int STAGES = AB_GEMM_STAGES;
constexpr int WARPS_M = AB_GEMM_WARPS_M;
constexpr int WARPS_N = AB_GEMM_WARPS_N;
constexpr int GEMM_THREADS = 32 * WARPS_M * WARPS_N;
constexpr int WTM = BM / WARPS_M; // warp tile
constexpr int WTN = BN / WARPS_N;
constexpr int MF = WTM / 8;       // fragments per warp along m / n
constexpr int NF = WTN / 8;
static_assert(BM % (8 * WARPS_M) == 0 && BN % (8 * WARPS_N) == 0 && BK % 4 == 0, "tile shape");
static_assert((BK * BM / 2) % GEMM_THREADS == 0 && (BK * BN / 2) % GEMM_THREADS == 0, "loader shape");
// the in-place triangular-solve leaves (linalg.cuh) need one CTA to own a whole LEAF-wide operand
static_assert(BN >= LEAF && BM >= LEAF, "CTA tile smaller than the in-place solve leaf");
constexpr int LDK = BK + 4; // [extent][LDK] tile of an operand whose k index is contiguous in memory
// [BK][extent + 4] tile of an operand whose m/n index is contiguous in memory
template <int EXTENT, bool KMAJOR> constexpr int tile_elems() {
  return KMAJOR ? EXTENT * LDK : BK * (EXTENT + 4);
}

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc, int src_bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, int src_bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// Loads one BK-deep tile of an operand into shared memory.
//   KMAJOR == false: element (i, kk) at g[i + kk * ld]   -> smem[kk * (EXT + 4) + i]
//   KMAJOR == true : element (i, kk) at g[kk + i * ld]   -> smem[i * LDK + kk]
// i in [0, EXT) relative to the tile; rows >= extent and k >= kextent are zero-filled (rows beyond
// the extent only ever feed outputs that are not stored, but k beyond kextent must contribute 0).
// vec == false: the operand is only 8-byte aligned (odd row offset of a sub-view or odd leading
// dimension): the same chunks are moved as two predicated 8-byte copies.
template <int EXT, bool KMAJOR>
__device__ __forceinline__ void load_tile(double *smem, const double *g, int64_t ld, int64_t i0,
                                          int64_t extent, int64_t k0, int64_t kextent, int tid,
                                          bool vec) {
  if (!vec) {
#pragma unroll
    for (int it = 0; it < (BK * EXT / 2) / GEMM_THREADS; ++it) {
      const int chunk = tid + it * GEMM_THREADS;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int i, kk;
        if (!KMAJOR) {
          kk = chunk / (EXT / 2);
          i = (chunk % (EXT / 2)) * 2 + e;
        } else {
          i = chunk / (BK / 2);
          kk = (chunk % (BK / 2)) * 2 + e;
        }
        const bool ok = (i0 + i < extent) && (k0 + kk < kextent);
        const double *src =
            ok ? (KMAJOR ? g + (k0 + kk) + (i0 + i) * ld : g + (i0 + i) + (k0 + kk) * ld) : g;
        cp_async8(KMAJOR ? smem + i * LDK + kk : smem + kk * (EXT + 4) + i, src, ok ? 8 : 0);
      }
    }
    return;
  }
  if (!KMAJOR) {
#pragma unroll
    for (int it = 0; it < (BK * EXT / 2) / GEMM_THREADS; ++it) {
      const int chunk = tid + it * GEMM_THREADS;
      const int kk = chunk / (EXT / 2);
      const int ic = (chunk % (EXT / 2)) * 2;
      const bool ok = (i0 + ic < extent) && (k0 + kk < kextent);
      const double *src = ok ? g + (i0 + ic) + (k0 + kk) * ld : g;
      cp_async16(smem + kk * (EXT + 4) + ic, src, ok ? 16 : 0);
    }
  } else {
#pragma unroll
    for (int it = 0; it < (BK * EXT / 2) / GEMM_THREADS; ++it) {
      const int chunk = tid + it * GEMM_THREADS;
      const int i = chunk / (BK / 2);
      const int kc = (chunk % (BK / 2)) * 2;
      const int64_t krem = kextent - (k0 + kc);
      const bool ok = (i0 + i < extent) && (krem > 0);
      const double *src = ok ? g + (k0 + kc) + (i0 + i) * ld : g;
      cp_async16(smem + i * LDK + kc, src, ok ? (krem >= 2 ? 16 : 8) : 0);
    }
  }
}

// AB_GEMM_FASTLOAD (default on; measured 8192^3 NN / TN / NT 30.9 / 32.7 / 31.0 -> 33.3 / 33.4 / 33.5
// TFLOP/s, profiles/r02a_gemm_sweep.txt): the k-loop of load_tile spends 310 non-DMMA instructions per 64
// DMMA, almost all of them the 64-bit address and predicate arithmetic it redoes for every k-tile
// (profiles/r01cdef_ncu_and_probe_summary.md).
// For CTA tiles that lie fully inside the operands (16-byte aligned, k a multiple of BK) everything but
// the k offset is loop invariant: a thread keeps one source pointer and one shared-memory offset per
// operand; chunk `it` of a k-tile is a fixed stride away from chunk 0 in both address spaces.
#ifndef AB_GEMM_FASTLOAD
#define AB_GEMM_FASTLOAD 1
#endif

template <int EXT, bool KMAJOR> struct FastTile {
  static constexpr int ITERS = (BK * EXT / 2) / GEMM_THREADS;
  // chunk = tid + it * GEMM_THREADS
  //   !KMAJOR: kk = chunk / (EXT / 2), ic = 2 (chunk % (EXT / 2)); GEMM_THREADS % (EXT / 2) == 0, so ic is
  //            the same for every it and kk advances by KSTEP = GEMM_THREADS / (EXT / 2)
  //    KMAJOR: i = chunk / (BK / 2), kc = 2 (chunk % (BK / 2)); i advances by ISTEP = GEMM_THREADS / (BK / 2)
  static constexpr int KSTEP = GEMM_THREADS / (EXT / 2);
  static constexpr int ISTEP = GEMM_THREADS / (BK / 2);
  static_assert(KMAJOR || GEMM_THREADS % (EXT / 2) == 0, "loader shape (fast path)");
  static_assert(!KMAJOR || GEMM_THREADS % (BK / 2) == 0, "loader shape (fast path)");
  static constexpr int S_IT = KMAJOR ? ISTEP * LDK : KSTEP * (EXT + 4); // smem elements between chunks

  const double *g0; // chunk 0 of k-tile 0
  int64_t g_it;     // global elements between chunks of one k-tile
  int64_t g_kt;     // global elements between k-tiles
  int s0;           // smem element offset of chunk 0 inside a stage

  __device__ __forceinline__ FastTile(const double *g, int64_t ld, int64_t i0, int tid) {
    if (!KMAJOR) {
      const int kk = tid / (EXT / 2);
      const int ic = (tid % (EXT / 2)) * 2;
      g0 = g + (i0 + ic) + static_cast<int64_t>(kk) * ld;
      g_it = static_cast<int64_t>(KSTEP) * ld;
      g_kt = static_cast<int64_t>(BK) * ld;
      s0 = kk * (EXT + 4) + ic;
    } else {
      const int i = tid / (BK / 2);
      const int kc = (tid % (BK / 2)) * 2;
      g0 = g + kc + (i0 + i) * ld;
      g_it = static_cast<int64_t>(ISTEP) * ld;
      g_kt = BK;
      s0 = i * LDK + kc;
    }
  }
  // all chunks of k-tile kt into the stage at `stage`
  __device__ __forceinline__ void issue(double *stage, int kt) const {
    const double *src = g0 + static_cast<int64_t>(kt) * g_kt;
    double *dst = stage + s0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      cp_async16(dst + it * S_IT, src, 16);
      src += g_it;
    }
  }
};

template <int EXT, bool KMAJOR>
__device__ __forceinline__ double frag(const double *smem, int idx, int kk) {
  return KMAJOR ? smem[idx * LDK + kk] : smem[kk * (EXT + 4) + idx];
}

// TA: op(A) = A^T (A stored k x m, k contiguous).  TB: op(B) = B^T (B stored n x k, n contiguous).
template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS, AB_GEMM_MIN_CTAS)
gemm_kernel(int64_t m, int64_t n, int64_t k, double alpha, const double *A, int64_t lda,
            const double *B, int64_t ldb, double beta, double *C,
            int64_t ldc, int tiles_m, int lower, int vec_flags) {
  const int vec_ok = vec_flags & 1;      // C: 16-byte epilogue
  const bool a_vec = vec_flags & 2;      // A: 16-byte cp.async
  const bool b_vec = vec_flags & 4;      // B: 16-byte cp.async
  constexpr bool A_KMAJOR = TA;
  constexpr bool B_KMAJOR = !TB;
  constexpr int A_ELEMS = tile_elems<BM, A_KMAJOR>();
  constexpr int B_ELEMS = tile_elems<BN, B_KMAJOR>();
  extern __shared__ __align__(16) double smem[];
  double *sA = smem;
  double *sB = smem + STAGES * A_ELEMS;

  int64_t bm, bn;
  if (lower && BM == BN) {
    const int64_t t = blockIdx.x;
    int64_t i = static_cast<int64_t>((sqrt(8. * static_cast<double>(t) + 1.) - 1.) * 0.5);
    while (i * (i + 1) / 2 > t) {
      --i;
    }
    while ((i + 1) * (i + 2) / 2 <= t) {
      ++i;
    }
    bm = i;
    bn = t - i * (i + 1) / 2;
  } else {
    bm = blockIdx.x % tiles_m;
    bn = blockIdx.x / tiles_m;
  }
  const int64_t m0 = bm * BM;
  const int64_t n0 = bn * BN;
  if (BM != BN && lower && m0 + BM <= n0) {
    return; // rectangular tiles: the full grid is launched, tiles strictly above the diagonal exit
  }

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp % WARPS_M;
  const int wn = warp / WARPS_M;
  const int lq = lane >> 2; // 0..7
  const int lr = lane & 3;  // 0..3

  double acc[NF][MF][2];
#pragma unroll
  for (int nf = 0; nf < NF; ++nf) {
#pragma unroll
    for (int mf = 0; mf < MF; ++mf) {
      acc[nf][mf][0] = 0.;
      acc[nf][mf][1] = 0.;
    }
  }

  const int ktiles = static_cast<int>((k + BK - 1) / BK);
  // fast loader: this CTA's tile lies fully inside both operands, 16-byte aligned, no k remainder
  const bool fast = AB_GEMM_FASTLOAD && a_vec && b_vec && (m0 + BM <= m) && (n0 + BN <= n) &&
                    (k % BK == 0);
  const FastTile<BM, A_KMAJOR> fa_tile(A, lda, m0, tid);
  const FastTile<BN, B_KMAJOR> fb_tile(B, ldb, n0, tid);
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) {
      if (fast) {
        fa_tile.issue(sA + s * A_ELEMS, s);
        fb_tile.issue(sB + s * B_ELEMS, s);
      } else {
        load_tile<BM, A_KMAJOR>(sA + s * A_ELEMS, A, lda, m0, m, static_cast<int64_t>(s) * BK, k, tid,
                                a_vec);
        load_tile<BN, B_KMAJOR>(sB + s * B_ELEMS, B, ldb, n0, n, static_cast<int64_t>(s) * BK, k, tid,
                                b_vec);
      }
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nt = kt + STAGES - 1;
      if (nt < ktiles) {
        const int s = nt % STAGES;
        if (fast) {
          fa_tile.issue(sA + s * A_ELEMS, nt);
          fb_tile.issue(sB + s * B_ELEMS, nt);
        } else {
          load_tile<BM, A_KMAJOR>(sA + s * A_ELEMS, A, lda, m0, m, static_cast<int64_t>(nt) * BK, k,
                                  tid, a_vec);
          load_tile<BN, B_KMAJOR>(sB + s * B_ELEMS, B, ldb, n0, n, static_cast<int64_t>(nt) * BK, k,
                                  tid, b_vec);
        }
      }
      cp_async_commit();
    }
    const double *tA = sA + (kt % STAGES) * A_ELEMS;
    const double *tB = sB + (kt % STAGES) * B_ELEMS;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int kk = ks * 4 + lr;
      double fb[NF], fa[MF];
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
        fb[nf] = frag<BN, B_KMAJOR>(tB, wn * WTN + nf * 8 + lq, kk);
      }
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) {
        fa[mf] = frag<BM, A_KMAJOR>(tA, wm * WTM + mf * 8 + lq, kk);
      }
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) {
          dmma_884(acc[nf][mf][0], acc[nf][mf][1], fb[nf], fa[mf]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: thread owns C(m0 + wm*WTM + mf*8 + 2*lr + {0,1}, n0 + wn*WTN + nf*8 + lq)
#pragma unroll
  for (int nf = 0; nf < NF; ++nf) {
    const int64_t col = n0 + wn * WTN + nf * 8 + lq;
    if (col >= n) {
      continue;
    }
#pragma unroll
    for (int mf = 0; mf < MF; ++mf) {
      const int64_t row = m0 + wm * WTM + mf * 8 + 2 * lr;
      if (row >= m) {
        continue;
      }
      double *dst = C + row + col * ldc;
      double v0 = alpha * acc[nf][mf][0];
      double v1 = alpha * acc[nf][mf][1];
      if (row + 1 < m) {
        if (vec_ok) {
          if (beta != 0.) {
            const double2 c = *reinterpret_cast<const double2 *>(dst);
            v0 = fma(beta, c.x, v0);
            v1 = fma(beta, c.y, v1);
          }
          *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
        } else {
          if (beta != 0.) {
            v0 = fma(beta, dst[0], v0);
            v1 = fma(beta, dst[1], v1);
          }
          dst[0] = v0;
          dst[1] = v1;
        }
      } else {
        if (beta != 0.) {
          v0 = fma(beta, dst[0], v0);
        }
        dst[0] = v0;
      }
    }
  }
}

template <bool TA, bool TB>
static int launch(ab_handle_s *h, bool lower, int64_t m, int64_t n, int64_t k, double alpha,
                  MatView A, MatView B, double beta, MatView C) {
  constexpr int A_ELEMS = tile_elems<BM, TA>();
  constexpr int B_ELEMS = tile_elems<BN, !TB>();
  constexpr size_t smem = static_cast<size_t>(STAGES) * (A_ELEMS + B_ELEMS) * sizeof(double);
  static PerDeviceOnce once;
  if (once.need(h->device)) {
    AB_CUDA(cudaFuncSetAttribute(gemm_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  }
  const int64_t tm = (m + BM - 1) / BM;
  const int64_t tn = (n + BN - 1) / BN;
  const int64_t tiles = (lower && BM == BN) ? tm * (tm + 1) / 2 : tm * tn;
  AB_REQUIRE(tiles < (int64_t(1) << 31), "GEMM grid too large");
  const auto aligned16 = [](const MatView &M) {
    return reinterpret_cast<uintptr_t>(M.p) % 16 == 0 && M.ld % 2 == 0;
  };
  const int vec_ok = (aligned16(C) ? 1 : 0) | (aligned16(A) ? 2 : 0) | (aligned16(B) ? 4 : 0);
  gemm_kernel<TA, TB><<<static_cast<unsigned>(tiles), GEMM_THREADS, smem, h->stream>>>(
      m, n, k, alpha, A.p, A.ld, B.p, B.ld, beta, C.p, C.ld, static_cast<int>(tm), lower ? 1 : 0,
      vec_ok);
  AB_LAUNCHED(h);
  return AB_OK;
}

// ------------------------------------------------------------------------------------------------
// n == 1: matrix-vector products (HBM-bound; the 128 x 128 MMA tiling would leave 127/128 of the
// tensor work idle and, for op(A) = A^T, only m/128 CTAs in flight)
// ------------------------------------------------------------------------------------------------

constexpr int GEMV_ROWS = 32;
constexpr int GEMV_WARPS = 8;

// c[m] = alpha * A[m x k] b[k] + beta * c ; A column-major.  CTA = 32 rows; warp w walks the columns
// kk = w, w + 8, ...; lane = row (256 contiguous bytes per warp load); deterministic reduction.
__global__ void __launch_bounds__(GEMV_ROWS *GEMV_WARPS)
gemv_n_unaligned_kernel(int64_t m, int64_t k, double alpha, const double *__restrict__ A, int64_t lda,
              const double *__restrict__ b, double beta, double *c) {
  __shared__ double red[GEMV_WARPS][GEMV_ROWS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t row = blockIdx.x * static_cast<int64_t>(GEMV_ROWS) + lane;
  double acc0 = 0., acc1 = 0., acc2 = 0., acc3 = 0.;
  if (row < m) {
    const double *a = A + row;
    int64_t kk = warp;
    for (; kk + 3 * GEMV_WARPS < k; kk += 4 * GEMV_WARPS) {
      acc0 = fma(a[kk * lda], b[kk], acc0);
      acc1 = fma(a[(kk + GEMV_WARPS) * lda], b[kk + GEMV_WARPS], acc1);
      acc2 = fma(a[(kk + 2 * GEMV_WARPS) * lda], b[kk + 2 * GEMV_WARPS], acc2);
      acc3 = fma(a[(kk + 3 * GEMV_WARPS) * lda], b[kk + 3 * GEMV_WARPS], acc3);
    }
    for (; kk < k; kk += GEMV_WARPS) {
      acc0 = fma(a[kk * lda], b[kk], acc0);
    }
  }
  red[warp][lane] = (acc0 + acc1) + (acc2 + acc3);
  __syncthreads();
  if (warp == 0 && row < m) {
    double total = 0.;
#pragma unroll
    for (int w = 0; w < GEMV_WARPS; ++w) {
      total += red[w][lane];
    }
    double v = alpha * total;
    if (beta != 0.) {
      v = fma(beta, c[row], v);
    }
    c[row] = v;
  }
}

// c[m] = alpha * A^T b + beta * c ; A stored k x m (k contiguous).  One CTA per output element.
__global__ void __launch_bounds__(256)
gemv_t_unaligned_kernel(int64_t k, double alpha, const double *__restrict__ A, int64_t lda,
              const double *__restrict__ b, double beta, double *c) {
  __shared__ double red[8];
  const double *a = A + blockIdx.x * lda;
  double acc0 = 0., acc1 = 0.;
  int64_t i = threadIdx.x;
  for (; i + 256 < k; i += 512) {
    acc0 = fma(a[i], b[i], acc0);
    acc1 = fma(a[i + 256], b[i + 256], acc1);
  }
  for (; i < k; i += 256) {
    acc0 = fma(a[i], b[i], acc0);
  }
  double v = acc0 + acc1;
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
      out = fma(beta, c[blockIdx.x], out);
    }
    c[blockIdx.x] = out;
  }
}

int gemm(ab_handle_s *h, unsigned flags, int64_t m, int64_t n, int64_t k, double alpha, MatView A,
         MatView B, double beta, MatView C) {
  if (m <= 0 || n <= 0) {
    return AB_OK;
  }
  const bool ta = flags & GEMM_TRANS_A;
  const bool tb = flags & GEMM_TRANS_B;
  const bool lower = flags & GEMM_LOWER;
  if (lower) {
    AB_REQUIRE(m == n, "GEMM_LOWER needs a square C");
  }
  if (n == 1 && !tb && m > 1 && C.p != B.p && C.p != A.p && k > 0) {
    // B is a contiguous k-vector (op(B) = B, one column)
    if (gemv_fast_ok(A, B.p)) { // 16-byte loads, k-split grids (trsv.cu)
      return ta ? gemv_t(h, k, m, alpha, A, B.p, beta, C.p) : gemv_n(h, m, k, alpha, A, B.p, beta, C.p);
    }
    if (ta) {
      gemv_t_unaligned_kernel<<<static_cast<unsigned>(m), 256, 0, h->stream>>>(k, alpha, A.p, A.ld, B.p,
                                                                      beta, C.p);
    } else {
      const unsigned blocks = static_cast<unsigned>((m + GEMV_ROWS - 1) / GEMV_ROWS);
      gemv_n_unaligned_kernel<<<blocks, GEMV_ROWS * GEMV_WARPS, 0, h->stream>>>(m, k, alpha, A.p, A.ld, B.p,
                                                                      beta, C.p);
    }
    AB_LAUNCHED(h);
    return AB_OK;
  }
  if (tb && !ta && gemm_tma_enabled()) {
    const int s = gemm_nt_tma(h, lower, m, n, k, alpha, A, B, beta, C);
    if (s != AB_ERR_UNSUPPORTED) {
      return s;
    }
  }
  if (ta && tb) {
    return launch<true, true>(h, lower, m, n, k, alpha, A, B, beta, C);
  } else if (ta) {
    return launch<true, false>(h, lower, m, n, k, alpha, A, B, beta, C);
  } else if (tb) {
    return launch<false, true>(h, lower, m, n, k, alpha, A, B, beta, C);
  }
  return launch<false, false>(h, lower, m, n, k, alpha, A, B, beta, C);
}

} // namespace ab
