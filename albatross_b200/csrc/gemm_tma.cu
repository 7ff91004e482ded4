// TMA-fed variant of the DMMA GEMM for the hot shape of the factorisation,
//
//   C[m x n] = alpha * A[m x k] * B[n x k]^T + beta * C         (GEMM_TRANS_B, optionally GEMM_LOWER),
//
// i.e. the trailing DSYRK / DGEMM updates of the blocked Cholesky (reference: the GEMV inner loop of
// third_party/eigen/Eigen/src/Cholesky/LDLT.h:349-355 that this library turns into level-3 work).
//
// Default path of every NT product large enough for one CTA tile (gemm() in gemm.cu falls back to the
// cp.async kernel otherwise); AB_GEMM_TMA=0 (read once) switches it off, which the parity tests use to
// compare the two kernels.  Measured on B200 (profiles/r02a_*): 8192^3 34.2 TFLOP/s vs 31.0 (cuBLAS 35.5),
// 16384^2 x 1024 34.4 vs 30.3, factorisation N = 32 768 31.3 vs 28.2 TFLOP/s.
//
// Why: the cp.async kernel (gemm.cu) keeps the DMMA pipe 84 % busy; its k-loop spends 310 non-DMMA
// instructions per 64 DMMA on per-thread address / predicate arithmetic and meets at a CTA-wide
// barrier every k-tile (profiles/r01cdef_ncu_and_probe_summary.md).  Here
//   * operand tiles are moved by TMA (cp.async.bulk.tensor.2d): one elected thread issues 12 box copies
//     per k-tile, nobody computes a global address, out-of-range rows / k are zero-filled by the hardware;
//   * stages are handed over through mbarriers (full: TMA transaction bytes; empty: one arrival per
//     warp), so there is no __syncthreads in the k-loop and warps drift apart freely;
//   * boxes are 16 rows x 16 k (128-byte rows, SWIZZLE_128B), which makes the m8n8k4 fragment reads
//     2-way bank conflicted at worst (a dense 128-row tile would be 4-way): 8 LDS.64 per 16 DMMA.
// Tiling, warp layout and epilogue are those of gemm.cu: 128 x 64 x 16 CTA tile, 8 warps 4 x 2, warp
// tile 32 x 32, 2 CTAs per SM, 16-byte vectorised read-modify-write of C.
#include "linalg.cuh"

#include <cuda.h>

#include <cstdlib>

namespace ab {

namespace {

constexpr int TBM = 128, TBN = 64, TBK = 16;
constexpr int TSTAGES = 4;
constexpr int TGROUP = 8; // tile columns per rasterisation group
constexpr int TWARPS_M = 4, TWARPS_N = 2;
constexpr int TTHREADS = 32 * TWARPS_M * TWARPS_N;
constexpr int BOX_ROWS = 16;                                 // 16 doubles = 128 bytes: the swizzle span
constexpr int BOX_BYTES = BOX_ROWS * TBK * 8;                // 2048
constexpr int A_BOXES = TBM / BOX_ROWS, B_BOXES = TBN / BOX_ROWS;
constexpr int STAGE_BYTES = (A_BOXES + B_BOXES) * BOX_BYTES; // 24 KB
constexpr int TMF = (TBM / TWARPS_M) / 8, TNF = (TBN / TWARPS_N) / 8;

__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "TMA_GEMM_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra TMA_GEMM_DONE;\n"
               "bra TMA_GEMM_WAIT;\n"
               "TMA_GEMM_DONE:\n"
               "}\n" ::"r"(bar),
               "r"(parity)
               : "memory");
}
// one 16 x 16 box: rows c0 .. c0+15 (inner, contiguous), k = c1 .. c1+15 -> smem, signalling `bar`
__device__ __forceinline__ void tma_box(unsigned dst, const CUtensorMap *map, int c0, int c1,
                                        unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
               "[%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// Byte offset, inside an operand's stage region, of element (row r of the CTA tile, k index kk):
// box r / 16; inside the box row kk is 128 bytes and its 16-byte chunk c sits at c ^ (kk & 7).
__device__ __forceinline__ unsigned swz(int r, int kk) {
  const int ii = r & (BOX_ROWS - 1);
  return static_cast<unsigned>((r >> 4) * BOX_BYTES + kk * 128 + ((((ii >> 1) ^ (kk & 7)) << 4) | ((ii & 1) << 3)));
}

__global__ void __launch_bounds__(TTHREADS, 2)
gemm_nt_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   int64_t m, int64_t n, int64_t k, double alpha, double beta, double *C, int64_t ldc,
                   int tiles_m, int tiles_n, int lower, int vec_ok, int cyc_blk, int64_t cyc_stride,
                   int64_t b_row0) {
  extern __shared__ __align__(1024) unsigned char tsmem[];
  __shared__ __align__(8) unsigned long long bars[2 * TSTAGES]; // full[0..S), empty[S..2S)

  // Tile order: groups of TGROUP tile columns, the tiles of one tile row of a group next to each other, so
  // that the ~296 CTAs resident at any time cover 37 tile rows x 8 tile columns (41 MB of operands) instead
  // of 296 tile rows x 1 tile column (297 MB): the A panel of a trailing update (up to 1 GB) is then streamed
  // from HBM once per GROUP of tile columns instead of once per tile column.
  const int gidx = static_cast<int>(blockIdx.x) / (TGROUP * tiles_m);
  const int gfirst = gidx * TGROUP;
  const int gsize = tiles_n - gfirst < TGROUP ? tiles_n - gfirst : TGROUP;
  const int gin = static_cast<int>(blockIdx.x) - gidx * TGROUP * tiles_m;
  const int64_t m0 = static_cast<int64_t>(gin / gsize) * TBM;
  const int64_t n0 = static_cast<int64_t>(gfirst + gin % gsize) * TBN;
  // Block-cyclic B (CyclicB, linalg.cuh): column block q = n0 / cyc_blk of C multiplies the rows
  // b_row0 + q * cyc_stride + (n0 % cyc_blk) ... of B, and "lower" is measured against that stretched
  // diagonal.  cyc_blk == 0: the plain product (row n0 of B, the ordinary diagonal).
  const int64_t gcol = cyc_blk > 0 ? (n0 / cyc_blk) * cyc_stride + (n0 % cyc_blk) : n0;
  const int64_t brow = b_row0 + gcol;
  if (lower && m0 + TBM <= gcol) {
    return; // strictly above the diagonal (whole CTA exits before any barrier is initialised)
  }
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp % TWARPS_M;
  const int wn = warp / TWARPS_M;
  const int lq = lane >> 2;
  const int lr = lane & 3;
  // 1024-byte aligned stage base (dynamic shared memory is only guaranteed 16-byte aligned)
  const unsigned base = (smem_u32(tsmem) + 1023u) & ~1023u;
  const unsigned char *base_ptr = tsmem + (base - smem_u32(tsmem));
  const unsigned bar0 = smem_u32(bars);

  if (tid == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(bar0 + 8 * s, 1);                                   // full: the producer's expect_tx arrival
      mbar_init(bar0 + 8 * (TSTAGES + s), TWARPS_M * TWARPS_N);     // empty: one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  const int ktiles = static_cast<int>((k + TBK - 1) / TBK);
  auto produce = [&](int kt) { // called by thread 0 only
    const int s = kt % TSTAGES;
    const unsigned full = bar0 + 8 * s;
    const unsigned stage = base + s * STAGE_BYTES;
    mbar_expect_tx(full, STAGE_BYTES);
#pragma unroll
    for (int b = 0; b < A_BOXES; ++b) {
      tma_box(stage + b * BOX_BYTES, &mapA, static_cast<int>(m0) + b * BOX_ROWS, kt * TBK, full);
    }
#pragma unroll
    for (int b = 0; b < B_BOXES; ++b) {
      tma_box(stage + (A_BOXES + b) * BOX_BYTES, &mapB, static_cast<int>(brow) + b * BOX_ROWS, kt * TBK,
              full);
    }
  };
  if (tid == 0) {
    for (int kt = 0; kt < TSTAGES - 1 && kt < ktiles; ++kt) {
      produce(kt);
    }
  }

  // per-thread fragment offsets for k-step 0 of a stage (kk = lr); k-step ks adds ks * 512 bytes and
  // flips chunk bit 2 when ks is odd ((4 ks + lr) & 7 = lr | 4 (ks & 1))
  unsigned offA[TMF], offB[TNF];
#pragma unroll
  for (int mf = 0; mf < TMF; ++mf) {
    offA[mf] = swz(wm * (TBM / TWARPS_M) + mf * 8 + lq, lr);
  }
#pragma unroll
  for (int nf = 0; nf < TNF; ++nf) {
    offB[nf] = A_BOXES * BOX_BYTES + swz(wn * (TBN / TWARPS_N) + nf * 8 + lq, lr);
  }

  double acc[TNF][TMF][2];
#pragma unroll
  for (int nf = 0; nf < TNF; ++nf) {
#pragma unroll
    for (int mf = 0; mf < TMF; ++mf) {
      acc[nf][mf][0] = 0.;
      acc[nf][mf][1] = 0.;
    }
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    const int s = kt % TSTAGES;
    // refill the stage that was consumed in iteration kt - 1 with k-tile kt + STAGES - 1
    if (tid == 0) {
      const int nt = kt + TSTAGES - 1;
      if (nt < ktiles) {
        if (nt >= TSTAGES) { // the stage has been used before: wait until all 8 warps released it
          mbar_wait(bar0 + 8 * (TSTAGES + nt % TSTAGES), ((nt / TSTAGES) - 1) & 1);
        }
        produce(nt);
      }
    }
    __syncwarp();
    mbar_wait(bar0 + 8 * s, (kt / TSTAGES) & 1);
    // plain loads: the "memory" clobber of mbar_wait orders them after the barrier for the compiler,
    // the mbarrier's acquire semantics for the hardware
    const unsigned char *stage = base_ptr + s * STAGE_BYTES;
#pragma unroll
    for (int ks = 0; ks < TBK / 4; ++ks) {
      const unsigned kofs = ks * 512;
      const unsigned flip = (ks & 1) ? 64u : 0u; // chunk ^ 4  ==  byte offset ^ 64
      double fb[TNF], fa[TMF];
#pragma unroll
      for (int nf = 0; nf < TNF; ++nf) {
        fb[nf] = *reinterpret_cast<const double *>(stage + ((offB[nf] ^ flip) + kofs));
      }
#pragma unroll
      for (int mf = 0; mf < TMF; ++mf) {
        fa[mf] = *reinterpret_cast<const double *>(stage + ((offA[mf] ^ flip) + kofs));
      }
#pragma unroll
      for (int nf = 0; nf < TNF; ++nf) {
#pragma unroll
        for (int mf = 0; mf < TMF; ++mf) {
          dmma(acc[nf][mf][0], acc[nf][mf][1], fb[nf], fa[mf]);
        }
      }
    }
    // this warp is done with stage s
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar0 + 8 * (TSTAGES + s));
    }
  }

  // epilogue (as gemm.cu): thread owns C(m0 + wm*32 + mf*8 + 2*lr + {0,1}, n0 + wn*32 + nf*8 + lq)
#pragma unroll
  for (int nf = 0; nf < TNF; ++nf) {
    const int64_t col = n0 + wn * (TBN / TWARPS_N) + nf * 8 + lq;
    if (col >= n) {
      continue;
    }
#pragma unroll
    for (int mf = 0; mf < TMF; ++mf) {
      const int64_t row = m0 + wm * (TBM / TWARPS_M) + mf * 8 + 2 * lr;
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

using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                              CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = []() -> EncodeFn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

// rows x kext operand, element (r, kk) at p[r + kk * ld]: dimension 0 = rows (contiguous)
bool make_map(CUtensorMap *map, const MatView &M, int64_t rows, int64_t kext) {
  EncodeFn fn = encode_fn();
  if (fn == nullptr) {
    return false;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(kext)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(M.ld) * sizeof(double)};
  const cuuint32_t box[2] = {BOX_ROWS, TBK};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, M.p, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace

bool gemm_tma_enabled() {
  static const bool on = []() {
    const char *e = std::getenv("AB_GEMM_TMA");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

// Returns AB_OK when the product was launched, AB_ERR_UNSUPPORTED when the caller must use the
// cp.async kernel (shape / alignment outside what the tensor maps can describe).
int gemm_nt_tma(ab_handle_s *h, bool lower, int64_t m, int64_t n, int64_t k, double alpha, MatView A,
                MatView B, double beta, MatView C, const CyclicB *cyc) {
  const auto aligned16 = [](const MatView &M) {
    return reinterpret_cast<uintptr_t>(M.p) % 16 == 0 && M.ld % 2 == 0;
  };
  if (!aligned16(A) || !aligned16(B) || k < TBK || m < TBM || n < TBN || m >= (int64_t(1) << 31) ||
      n >= (int64_t(1) << 31) || k >= (int64_t(1) << 31) || C.p == A.p || C.p == B.p) {
    return AB_ERR_UNSUPPORTED;
  }
  const bool cyclic = cyc != nullptr && cyc->blk > 0;
  if (cyclic && (cyc->blk % TBN != 0 || cyc->rows >= (int64_t(1) << 31) || cyc->blk >= (int64_t(1) << 31))) {
    return AB_ERR_UNSUPPORTED;
  }
  CUtensorMap mapA, mapB;
  // cyclic: B is the whole packed panel (cyc->rows rows); rows past its end read as zero (TMA fill)
  if (!make_map(&mapA, A, m, k) || !make_map(&mapB, B, cyclic ? cyc->rows : n, k)) {
    return AB_ERR_UNSUPPORTED;
  }
  constexpr size_t smem = static_cast<size_t>(TSTAGES) * STAGE_BYTES + 1024; // + alignment slack
  static PerDeviceOnce once;
  if (once.need(h->device)) {
    AB_CUDA(cudaFuncSetAttribute(gemm_nt_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  }
  const int64_t tm = (m + TBM - 1) / TBM;
  const int64_t tn = (n + TBN - 1) / TBN;
  AB_REQUIRE(tm * tn < (int64_t(1) << 31), "GEMM grid too large");
  const int vec_ok = aligned16(C) ? 1 : 0;
  gemm_nt_tma_kernel<<<static_cast<unsigned>(tm * tn), TTHREADS, smem, h->stream>>>(
      mapA, mapB, m, n, k, alpha, beta, C.p, C.ld, static_cast<int>(tm), static_cast<int>(tn), lower ? 1 : 0, vec_ok,
      cyclic ? static_cast<int>(cyc->blk) : 0, cyclic ? cyc->stride : 0, cyclic ? cyc->row0 : 0);
  AB_LAUNCHED(h);
  return AB_OK;
}

} // namespace ab
