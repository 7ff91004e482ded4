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

namespace ab {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int STAGES = 3;
constexpr int GEMM_THREADS = 256;
constexpr int LDMN = BM + 4; // [BK][LDMN] tile of an operand whose m/n index is contiguous in memory
constexpr int LDK = BK + 4;  // [BM][LDK]  tile of an operand whose k index is contiguous in memory
constexpr int TILE_MN_ELEMS = BK * LDMN;
constexpr int TILE_K_ELEMS = BM * LDK;

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc, int src_bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc),
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
//   KMAJOR == false: element (i, kk) at g[i + kk * ld]   -> smem[kk * LDMN + i]
//   KMAJOR == true : element (i, kk) at g[kk + i * ld]   -> smem[i * LDK + kk]
// i in [0, BM) relative to the tile; rows >= extent and k >= kextent are zero-filled (rows beyond
// the extent only ever feed outputs that are not stored, but k beyond kextent must contribute 0).
template <bool KMAJOR>
__device__ __forceinline__ void load_tile(double *smem, const double *g, int64_t ld, int64_t i0,
                                          int64_t extent, int64_t k0, int64_t kextent, int tid) {
  if (!KMAJOR) {
#pragma unroll
    for (int it = 0; it < (BK * BM / 2) / GEMM_THREADS; ++it) {
      const int chunk = tid + it * GEMM_THREADS;
      const int kk = chunk / (BM / 2);
      const int ic = (chunk % (BM / 2)) * 2;
      const bool ok = (i0 + ic < extent) && (k0 + kk < kextent);
      const double *src = ok ? g + (i0 + ic) + (k0 + kk) * ld : g;
      cp_async16(smem + kk * LDMN + ic, src, ok ? 16 : 0);
    }
  } else {
#pragma unroll
    for (int it = 0; it < (BK * BM / 2) / GEMM_THREADS; ++it) {
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

template <bool KMAJOR>
__device__ __forceinline__ double frag(const double *smem, int idx, int kk) {
  return KMAJOR ? smem[idx * LDK + kk] : smem[kk * LDMN + idx];
}

// TA: op(A) = A^T (A stored k x m, k contiguous).  TB: op(B) = B^T (B stored n x k, n contiguous).
template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(int64_t m, int64_t n, int64_t k, double alpha, const double *A, int64_t lda,
            const double *B, int64_t ldb, double beta, double *C,
            int64_t ldc, int tiles_m, int lower, int vec_ok) {
  constexpr bool A_KMAJOR = TA;
  constexpr bool B_KMAJOR = !TB;
  constexpr int A_ELEMS = A_KMAJOR ? TILE_K_ELEMS : TILE_MN_ELEMS;
  constexpr int B_ELEMS = B_KMAJOR ? TILE_K_ELEMS : TILE_MN_ELEMS;
  extern __shared__ __align__(16) double smem[];
  double *sA = smem;
  double *sB = smem + STAGES * A_ELEMS;

  int64_t bm, bn;
  if (lower) {
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

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp & 1;  // 2 warps along m (64 rows each)
  const int wn = warp >> 1; // 4 warps along n (32 cols each)
  const int lq = lane >> 2; // 0..7
  const int lr = lane & 3;  // 0..3

  double acc[4][8][2];
#pragma unroll
  for (int nf = 0; nf < 4; ++nf) {
#pragma unroll
    for (int mf = 0; mf < 8; ++mf) {
      acc[nf][mf][0] = 0.;
      acc[nf][mf][1] = 0.;
    }
  }

  const int ktiles = static_cast<int>((k + BK - 1) / BK);
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) {
      load_tile<A_KMAJOR>(sA + s * A_ELEMS, A, lda, m0, m, static_cast<int64_t>(s) * BK, k, tid);
      load_tile<B_KMAJOR>(sB + s * B_ELEMS, B, ldb, n0, n, static_cast<int64_t>(s) * BK, k, tid);
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
        load_tile<A_KMAJOR>(sA + s * A_ELEMS, A, lda, m0, m, static_cast<int64_t>(nt) * BK, k, tid);
        load_tile<B_KMAJOR>(sB + s * B_ELEMS, B, ldb, n0, n, static_cast<int64_t>(nt) * BK, k, tid);
      }
      cp_async_commit();
    }
    const double *tA = sA + (kt % STAGES) * A_ELEMS;
    const double *tB = sB + (kt % STAGES) * B_ELEMS;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int kk = ks * 4 + lr;
      double fb[4], fa[8];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        fb[nf] = frag<B_KMAJOR>(tB, wn * 32 + nf * 8 + lq, kk);
      }
#pragma unroll
      for (int mf = 0; mf < 8; ++mf) {
        fa[mf] = frag<A_KMAJOR>(tA, wm * 64 + mf * 8 + lq, kk);
      }
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
#pragma unroll
        for (int mf = 0; mf < 8; ++mf) {
          dmma_884(acc[nf][mf][0], acc[nf][mf][1], fb[nf], fa[mf]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: thread owns C(m0 + wm*64 + mf*8 + 2*lr + {0,1}, n0 + wn*32 + nf*8 + lq)
#pragma unroll
  for (int nf = 0; nf < 4; ++nf) {
    const int64_t col = n0 + wn * 32 + nf * 8 + lq;
    if (col >= n) {
      continue;
    }
#pragma unroll
    for (int mf = 0; mf < 8; ++mf) {
      const int64_t row = m0 + wm * 64 + mf * 8 + 2 * lr;
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
  constexpr int A_ELEMS = TA ? TILE_K_ELEMS : TILE_MN_ELEMS;
  constexpr int B_ELEMS = !TB ? TILE_K_ELEMS : TILE_MN_ELEMS;
  constexpr size_t smem = static_cast<size_t>(STAGES) * (A_ELEMS + B_ELEMS) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    AB_CUDA(cudaFuncSetAttribute(gemm_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    configured = true;
  }
  const int64_t tm = (m + BM - 1) / BM;
  const int64_t tn = (n + BN - 1) / BN;
  const int64_t tiles = lower ? tm * (tm + 1) / 2 : tm * tn;
  AB_REQUIRE(tiles < (int64_t(1) << 31), "GEMM grid too large");
  const int vec_ok = (reinterpret_cast<uintptr_t>(C.p) % 16 == 0) && (C.ld % 2 == 0);
  AB_REQUIRE(reinterpret_cast<uintptr_t>(A.p) % 16 == 0 && A.ld % 2 == 0 &&
                 reinterpret_cast<uintptr_t>(B.p) % 16 == 0 && B.ld % 2 == 0,
             "GEMM operands must be 16-byte aligned with even leading dimension");
  gemm_kernel<TA, TB><<<static_cast<unsigned>(tiles), GEMM_THREADS, smem, h->stream>>>(
      m, n, k, alpha, A.p, A.ld, B.p, B.ld, beta, C.p, C.ld, static_cast<int>(tm), lower ? 1 : 0,
      vec_ok);
  AB_LAUNCHED(h);
  return AB_OK;
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
