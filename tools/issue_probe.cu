// Issue-slot probe for the Gram kernel's FP64 + integer instruction mix (not part of the product):
// does a warp-wide DFMA (2 cycles on the 16-lane FP64 pipe of an SM sub-partition) also hold the issue /
// dispatch port for its second cycle, i.e. do the integer instructions of the evaluator ride for free
// next to the FP64 work or do they add to it?  Also measures the dependent-issue latency of DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe issue_probe.cu
#include <cuda_runtime.h>

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

// NF independent DFMA chains and NI independent IMAD chains per thread, NI_PER integer instructions
// issued after every DFMA (0, 1 or 2).
// integer-side instruction kinds
enum { K_IMAD = 0, K_IADD3 = 1, K_LOP3 = 2, K_SHF = 3, K_FFMA = 4, K_LDS = 5 };

template <int KIND> __device__ __forceinline__ int int_op(int v, int b, int it, const int *sm) {
  if (KIND == K_IMAD) {
    return v * b + it;
  } else if (KIND == K_IADD3) {
    int r;
    asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(b));
    return r;
  } else if (KIND == K_LOP3) {
    int r;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(v), "r"(b), "r"(it));
    return r;
  } else if (KIND == K_SHF) {
    int r;
    asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(b), "r"(it));
    return r;
  } else if (KIND == K_FFMA) {
    return __float_as_int(fmaf(__int_as_float(v), 0.999f, 1.0f));
  } else {
    return sm[(v & 31)] + it;
  }
}

template <int NF, int NI_PER, int KIND = K_IMAD>
__global__ void mix_kernel(double *out, int *iout, int iters, double a, int b, long long *cycles) {
  double f[NF];
  int v[NF * (NI_PER > 0 ? NI_PER : 1)];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    f[i] = threadIdx.x * 1e-3 + i;
  }
#pragma unroll
  for (int i = 0; i < NF * (NI_PER > 0 ? NI_PER : 1); ++i) {
    v[i] = threadIdx.x + i;
  }
  __shared__ int sm[32];
  if (threadIdx.x < 32) {
    sm[threadIdx.x] = threadIdx.x ^ b;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        f[i] = fma(f[i], a, a);
#pragma unroll
        for (int j = 0; j < NI_PER; ++j) {
          v[i * NI_PER + j] = int_op<KIND>(v[i * NI_PER + j], b, it, sm);
        }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.;
  int vs = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    s += f[i];
  }
#pragma unroll
  for (int i = 0; i < NF * (NI_PER > 0 ? NI_PER : 1); ++i) {
    vs += v[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = vs;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *cycles = t1 - t0;
  }
}

// DFMA with three distinct 64-bit register sources per instruction (the Gram evaluator's shape; the
// kernel above re-reads one register pair twice, which the operand-reuse cache serves).
template <int NF>
__global__ void dfma3_kernel(double *out, int iters, double a, long long *cycles) {
  double f[NF], g[NF], h[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    f[i] = threadIdx.x * 1e-3 + i;
    g[i] = 0.999 + 1e-6 * (threadIdx.x + i) * a;
    h[i] = 1e-3 * (i + 1) * a;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        f[i] = fma(f[i], g[(i + rep) % NF], h[(i + 2 * rep + 1) % NF]);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.;
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    s += f[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *cycles = t1 - t0;
  }
}

// Do DMMA (FP64 tensor sub-pipe) and DFMA (FP64 pipe) share hardware?  ND independent m8n8k4 accumulators
// and NF independent DFMA chains per thread; if the two are separate units the mixed loop costs
// max(ND * 16, NF * 2) cycles per iteration and sub-partition, if they are one unit the sum.
template <int ND, int NF>
__global__ void dmma_dfma_kernel(double *out, int iters, double a, long long *cycles) {
  double d[ND > 0 ? ND : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < (ND > 0 ? ND : 1); ++i) {
    d[i][0] = threadIdx.x * 1e-3 + i;
    d[i][1] = threadIdx.x * 2e-3 + i;
  }
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) {
    f[i] = threadIdx.x * 1e-3 + i;
  }
  const double fa = 1.0 + a * 1e-9 * threadIdx.x, fb = 0.999 + a * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
      if (i < ND) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(d[i][0]), "+d"(d[i][1])
                     : "d"(fa), "d"(fb));
      }
      if (i < NF) {
        f[i] = fma(f[i], a, a);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.;
#pragma unroll
  for (int i = 0; i < (ND > 0 ? ND : 1); ++i) {
    s += d[i][0] + d[i][1];
  }
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) {
    s += f[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *cycles = t1 - t0;
  }
}

template <int ND, int NF> static void run_mix(int threads, double *out, long long *cyc) {
  const int iters = 4096;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int rep = 0; rep < 2; ++rep) {
    dmma_dfma_kernel<ND, NF><<<sms, threads>>>(out, iters, 0.999, cyc);
    CK(cudaDeviceSynchronize());
  }
  long long c = 0;
  CK(cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost));
  const double w = threads / 128.0;
  printf("DMMA x%d + DFMA x%d per thread-iteration, warps/SMSP=%.0f : %8.2f cycles per iteration per SMSP "
         "(separate units: %d, one unit: %d)\n",
         ND, NF, w, c / (double)iters, (int)(w * (ND * 16 > NF * 2 ? ND * 16 : NF * 2)),
         (int)(w * (ND * 16 + NF * 2)));
  fflush(stdout);
}

template <int NF> static void run3(int threads, double *out, long long *cyc) {
  const int iters = 2048;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int rep = 0; rep < 2; ++rep) {
    dfma3_kernel<NF><<<sms, threads>>>(out, iters, 0.5, cyc);
    CK(cudaDeviceSynchronize());
  }
  long long c = 0;
  CK(cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost));
  const double warps_per_smsp = threads / 128.0;
  printf("DFMA 3 distinct sources, chains=%2d warps/SMSP=%.0f : %8.3f cycles per warp-DFMA per SMSP\n", NF,
         warps_per_smsp, c / (4.0 * NF * iters * warps_per_smsp));
  fflush(stdout);
}

template <int NF, int NI_PER, int KIND = K_IMAD>
static void run(int threads, double *out, int *iout, long long *cyc) {
  const int iters = 2048;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int rep = 0; rep < 2; ++rep) {
    mix_kernel<NF, NI_PER, KIND><<<sms, threads>>>(out, iout, iters, 0.999, 3, cyc);
    CK(cudaDeviceSynchronize());
  }
  long long c = 0;
  CK(cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost));
  const double warps_per_smsp = threads / 128.0;
  const double dfma_per_smsp = 4.0 * NF * iters * warps_per_smsp;
  static const char *names[] = {"IMAD", "IADD3", "LOP3", "SHF", "FFMA", "LDS"};
  printf("%-5s chains=%2d int/dfma=%d warps/SMSP=%.0f : %8.3f cycles per warp-DFMA per SMSP (%.2f cycles per "
         "issued instruction)\n",
         names[KIND], NF, NI_PER, warps_per_smsp, c / dfma_per_smsp, c / (dfma_per_smsp * (1 + NI_PER)));
  fflush(stdout);
}

int main() {
  double *out;
  int *iout;
  long long *cyc;
  CK(cudaMalloc(&out, 148 * 1024 * sizeof(double) * 2));
  CK(cudaMalloc(&iout, 148 * 1024 * sizeof(int) * 2));
  CK(cudaMalloc(&cyc, sizeof(long long)));
  printf("-- DFMA dependent-issue latency (1 warp per SMSP, 1 chain) and pipe rate (8 chains)\n");
  run<1, 0>(128, out, iout, cyc);
  run<2, 0>(128, out, iout, cyc);
  run<4, 0>(128, out, iout, cyc);
  run<8, 0>(128, out, iout, cyc);
  printf("-- 4 warps per SMSP (the Gram kernel's occupancy), 4 chains per thread\n");
  run<4, 0>(512, out, iout, cyc);
  run<4, 1>(512, out, iout, cyc);
  run<4, 2>(512, out, iout, cyc);
  printf("-- 6 and 8 warps per SMSP\n");
  run<4, 0>(768, out, iout, cyc);
  run<4, 1>(768, out, iout, cyc);
  run<4, 2>(768, out, iout, cyc);
  run<4, 1>(1024, out, iout, cyc);
  run<4, 2>(1024, out, iout, cyc);
  printf("-- 2 chains per thread (less ILP)\n");
  run<2, 0>(512, out, iout, cyc);
  run<2, 1>(512, out, iout, cyc);
  run<2, 2>(512, out, iout, cyc);
  printf("-- which instruction classes ride next to DFMA (4 warps per SMSP, 4 chains, 1 / 2 / 3 per DFMA)\n");
  run<4, 1, K_IADD3>(512, out, iout, cyc);
  run<4, 2, K_IADD3>(512, out, iout, cyc);
  run<4, 3, K_IADD3>(512, out, iout, cyc);
  run<4, 1, K_LOP3>(512, out, iout, cyc);
  run<4, 2, K_LOP3>(512, out, iout, cyc);
  run<4, 3, K_LOP3>(512, out, iout, cyc);
  run<4, 1, K_SHF>(512, out, iout, cyc);
  run<4, 2, K_SHF>(512, out, iout, cyc);
  run<4, 1, K_FFMA>(512, out, iout, cyc);
  run<4, 2, K_FFMA>(512, out, iout, cyc);
  run<4, 1, K_LDS>(512, out, iout, cyc);
  run<4, 2, K_LDS>(512, out, iout, cyc);
  run<4, 3, K_IMAD>(512, out, iout, cyc);
  printf("-- register-file side: three distinct 64-bit sources per DFMA\n");
  run3<4>(128, out, cyc);
  run3<8>(128, out, cyc);
  run3<4>(512, out, cyc);
  run3<8>(512, out, cyc);
  run3<4>(768, out, cyc);
  printf("-- DMMA and DFMA: one unit or two?\n");
  run_mix<8, 0>(512, out, cyc);
  run_mix<0, 64>(512, out, cyc);
  run_mix<8, 64>(512, out, cyc);
  run_mix<8, 32>(512, out, cyc);
  run_mix<8, 16>(512, out, cyc);
  printf("issue_probe rc=0\n");
  return 0;
}
