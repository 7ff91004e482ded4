// Store-pattern probe for the Gram kernel (not part of the product): how fast can 296 persistent
// CTAs write an N x N column-major fp64 matrix as T x T tiles, depending on
//   * the tile size T (a column segment of a tile is T * 8 contiguous bytes),
//   * the tile order (column-major over the full grid | lower tile + its mirror, super-block order),
//   * the store width (8 / 16 bytes per lane) and mechanism (st.global | cp.async.bulk from smem).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e = (x);                                                                        \
    if (e != cudaSuccess) {                                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);            \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int THREADS = 256;

__global__ void linear_fill(double2 *out, int64_t n2, double v) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n2; i += stride) {
    out[i] = make_double2(v, v);
  }
}

// CTA-contiguous chunks: CTA b writes chunk b, b + grid, ... of `chunk` bytes each
__global__ void chunk_fill(double2 *out, int64_t n2, int64_t chunk2, double v) {
  const int64_t nchunks = n2 / chunk2;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    double2 *dst = out + c * chunk2;
    for (int64_t i = threadIdx.x; i < chunk2; i += blockDim.x) {
      dst[i] = make_double2(v, v);
    }
  }
}

template <int T, int WIDTH>
__device__ __forceinline__ void write_tile(double *out, int64_t ld, int64_t i0, int64_t j0, double v) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (WIDTH == 16) {
    // a warp writes 512 contiguous bytes; T / 64 such pieces per column
    constexpr int PIECES = T / 64;
    for (int idx = warp; idx < T * PIECES; idx += THREADS / 32) {
      const int c = idx / PIECES, p = idx % PIECES;
      *reinterpret_cast<double2 *>(out + i0 + p * 64 + 2 * lane + (j0 + c) * ld) = make_double2(v, v);
    }
  } else {
    constexpr int PIECES = T / 32;
    for (int idx = warp; idx < T * PIECES; idx += THREADS / 32) {
      const int c = idx / PIECES, p = idx % PIECES;
      out[i0 + p * 32 + lane + (j0 + c) * ld] = v;
    }
  }
}

template <int T>
__device__ __forceinline__ void bulk_tile(double *out, int64_t ld, int64_t i0, int64_t j0,
                                          const double *stage) {
  // one cp.async.bulk of T * 8 bytes per column, issued by T threads (or in rounds)
  for (int c = threadIdx.x; c < T; c += THREADS) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(stage + c * T));
    double *dst = out + i0 + (j0 + c) * ld;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(s),
                 "r"(T * 8)
                 : "memory");
  }
}

__device__ __forceinline__ bool decode_sym(unsigned t, unsigned tiles, unsigned SB, unsigned &I,
                                           unsigned &J) {
  const unsigned sb = t / (SB * SB);
  const unsigned local = t % (SB * SB);
  unsigned i = static_cast<unsigned>((sqrtf(8.f * static_cast<float>(sb) + 1.f) - 1.f) * 0.5f);
  while (i * (i + 1u) / 2u > sb) --i;
  while ((i + 1u) * (i + 2u) / 2u <= sb) ++i;
  I = i * SB + local % SB;
  J = (sb - i * (i + 1u) / 2u) * SB + local / SB;
  return I < tiles && J <= I;
}

// ORDER 0: all tiles, column-major over the tile grid.  ORDER 1: lower tiles in super-block order,
// each followed by its mirror tile.  ORDER 2: as 1 but lower tiles only (half the bytes).
// MECH 0: st.global WIDTH bytes/lane.  MECH 1: cp.async.bulk from shared memory.
// ORDER 3 of the Gram candidates (AB_GRAM_MICRO): 64 x 64 lower tiles + mirrors, but each CTA walks the MB x MB
// tiles of a micro-block one after the other (rows fastest), micro-blocks in super-block order.
__global__ void __launch_bounds__(THREADS) micro_fill(double *out, int64_t n, int64_t ld, double v,
                                                      unsigned SB, unsigned MB) {
  constexpr int T = 64;
  const unsigned tiles = static_cast<unsigned>(n / T);
  const unsigned mt = (tiles + MB - 1) / MB;
  const unsigned nsb = (mt + SB - 1) / SB;
  const unsigned nmicro = nsb * (nsb + 1) / 2 * SB * SB;
  for (unsigned mic = blockIdx.x; mic < nmicro; mic += gridDim.x) {
    unsigned MI, MJ;
    if (!decode_sym(mic, mt, SB, MI, MJ)) continue;
    for (unsigned sub = 0; sub < MB * MB; ++sub) {
      const unsigned I = MB * MI + sub % MB, J = MB * MJ + sub / MB;
      if (I >= tiles || J > I) continue;
      const int64_t i0 = static_cast<int64_t>(I) * T, j0 = static_cast<int64_t>(J) * T;
      write_tile<T, 16>(out, ld, i0, j0, v);
      if (I != J) write_tile<T, 8>(out, ld, j0, i0, v);
    }
  }
}

template <int T, int ORDER, int WIDTH, int MECH>
__global__ void __launch_bounds__(THREADS) tile_fill(double *out, int64_t n, int64_t ld, double v,
                                                     unsigned SB) {
  extern __shared__ __align__(128) double stage[];
  const unsigned tiles = static_cast<unsigned>(n / T);
  if (MECH == 1) {
    for (int i = threadIdx.x; i < T * T; i += THREADS) stage[i] = v;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
  }
  if (ORDER == 0) {
    for (unsigned t = blockIdx.x; t < tiles * tiles; t += gridDim.x) {
      const int64_t i0 = static_cast<int64_t>(t % tiles) * T, j0 = static_cast<int64_t>(t / tiles) * T;
      if (MECH == 0) {
        write_tile<T, WIDTH>(out, ld, i0, j0, v);
      } else {
        bulk_tile<T>(out, ld, i0, j0, stage);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 4;\n" ::: "memory");
      }
    }
  } else {
    const unsigned nsb = (tiles + SB - 1) / SB;
    const unsigned nitems = nsb * (nsb + 1) / 2 * SB * SB;
    for (unsigned t = blockIdx.x; t < nitems; t += gridDim.x) {
      unsigned I, J;
      if (!decode_sym(t, tiles, SB, I, J)) continue;
      const int64_t i0 = static_cast<int64_t>(I) * T, j0 = static_cast<int64_t>(J) * T;
      if (MECH == 0) {
        write_tile<T, 16>(out, ld, i0, j0, v);
        if (ORDER == 1 && I != J) write_tile<T, WIDTH>(out, ld, j0, i0, v);
      } else {
        bulk_tile<T>(out, ld, i0, j0, stage);
        if (ORDER == 1 && I != J) bulk_tile<T>(out, ld, j0, i0, stage);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 4;\n" ::: "memory");
      }
    }
  }
  if (MECH == 1) {
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
}

static double *g_out;
static double *g_flush;
static int64_t g_n, g_ld;

template <class F> static void timeit(const char *name, double bytes, F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaMemsetAsync(g_flush, rep, 512ll << 20)); // > L2
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  printf("%-58s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best * 1e-6);
  fflush(stdout);
}

template <int T, int ORDER, int WIDTH, int MECH> static void run_tile(int grid, unsigned sb_elems) {
  char name[128];
  const unsigned SB = sb_elems / T;
  snprintf(name, sizeof name, "T=%d order=%s width=%d mech=%s grid=%d sb=%u", T,
           ORDER == 0 ? "colmajor" : (ORDER == 1 ? "sym+mirror" : "lower-only"), WIDTH,
           MECH ? "bulk" : "st", grid, sb_elems);
  const size_t smem = MECH ? sizeof(double) * T * T : 0;
  if (smem > 48 * 1024) {
    CK(cudaFuncSetAttribute(tile_fill<T, ORDER, WIDTH, MECH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            static_cast<int>(smem)));
  }
  const double full = 8.0 * g_n * g_n;
  const double bytes = ORDER == 2 ? full * 0.5 * (1.0 + 1.0 * T / g_n) : full;
  timeit(name, bytes, [&] {
    tile_fill<T, ORDER, WIDTH, MECH><<<grid, THREADS, smem>>>(g_out, g_n, g_ld, 1.0, SB);
  });
}

int main(int argc, char **argv) {
  g_n = argc > 1 ? atoll(argv[1]) : 32768;
  g_ld = g_n + 16;
  CK(cudaMalloc(&g_out, sizeof(double) * g_ld * g_n));
  CK(cudaMalloc(&g_flush, 512ll << 20));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, N=%lld ld=%lld\n", prop.name, sms, (long long)g_n, (long long)g_ld);
  const double full = 8.0 * g_n * g_n;
  const int64_t n2 = g_ld * g_n / 2;
  timeit("linear fill 16B grid=8*SMs x 256", 8.0 * g_ld * g_n,
         [&] { linear_fill<<<8 * sms, 256>>>(reinterpret_cast<double2 *>(g_out), n2, 1.0); });
  timeit("linear fill 16B grid=2*SMs x 256", 8.0 * g_ld * g_n,
         [&] { linear_fill<<<2 * sms, 256>>>(reinterpret_cast<double2 *>(g_out), n2, 1.0); });
  timeit("cudaMemset", 8.0 * g_ld * g_n, [&] { CK(cudaMemsetAsync(g_out, 0, sizeof(double) * g_ld * g_n)); });
  for (int64_t chunk : {4096, 32768, 262144}) {
    char name[64];
    snprintf(name, sizeof name, "chunk fill %lld B per CTA, grid=2*SMs", (long long)chunk);
    timeit(name, 8.0 * g_ld * g_n, [&] {
      chunk_fill<<<2 * sms, 256>>>(reinterpret_cast<double2 *>(g_out), n2, chunk / 16, 1.0);
    });
  }
  (void)full;
  const int g2 = 2 * sms, g4 = 4 * sms, g1 = sms;
  // tile size, full grid column-major
  run_tile<64, 0, 16, 0>(g2, 1024);
  run_tile<128, 0, 16, 0>(g2, 1024);
  run_tile<256, 0, 16, 0>(g2, 1024);
  run_tile<64, 0, 8, 0>(g2, 1024);
  run_tile<64, 0, 16, 0>(g4, 1024);
  run_tile<64, 0, 16, 0>(g1, 1024);
  // symmetric order (what the Gram kernel does), super-block size
  run_tile<64, 1, 8, 0>(g2, 1024);
  run_tile<64, 1, 16, 0>(g2, 1024);
  run_tile<64, 1, 16, 0>(g2, 512);
  run_tile<64, 1, 16, 0>(g2, 2048);
  run_tile<64, 1, 16, 0>(g2, 4096);
  run_tile<64, 1, 16, 0>(g4, 1024);
  run_tile<128, 1, 16, 0>(g2, 1024);
  run_tile<128, 1, 16, 0>(g2, 2048);
  run_tile<256, 1, 16, 0>(g2, 2048);
  run_tile<64, 2, 16, 0>(g2, 1024);
  // micro-block order of 64 x 64 tiles (what AB_GRAM_MICRO=2 / 4 would write)
  for (unsigned mb : {1u, 2u, 4u, 8u}) {
    char name[96];
    snprintf(name, sizeof name, "T=64 order=sym+mirror micro-block %ux%u grid=%d sb=16", mb, mb, g2);
    timeit(name, 8.0 * g_n * g_n, [&] { micro_fill<<<g2, THREADS>>>(g_out, g_n, g_ld, 1.0, 16u, mb); });
  }
  // bulk (TMA 1-D) stores from shared memory
  run_tile<64, 0, 16, 1>(g2, 1024);
  run_tile<64, 1, 16, 1>(g2, 1024);
  run_tile<64, 1, 16, 1>(g4, 1024);
  run_tile<128, 1, 16, 1>(g1, 1024);
  run_tile<128, 1, 16, 1>(g1, 2048);
  printf("store_pattern rc=0\n");
  return 0;
}
