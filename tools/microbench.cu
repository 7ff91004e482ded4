// Roofline probes for the fp64 path on B200 (not part of the product):
//   * DMMA.8x8x4 issue throughput (the FP64 tensor instruction sm_100a offers), per warps/SM
//   * DFMA throughput
//   * cuBLAS DGEMM / DSYRK at the shapes the factorisation uses  -> the FP64 "measured peak"
//   * fp64 exp() throughput (libdevice)                          -> Gram kernel ALU bound
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e = (x);                                                                        \
    if (e != cudaSuccess) {                                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);            \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

template <int NACC> __global__ void dmma_kernel(double *out, int iters) {
  double acc[NACC][2];
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    acc[i][0] = 0.;
    acc[i][1] = 0.;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1])
                   : "d"(a), "d"(b));
    }
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    s += acc[i][0] + acc[i][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC> __global__ void dfma_kernel(double *out, int iters) {
  double acc[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    acc[i] = i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    s += acc[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_kernel(double *out, int iters) {
  double x[8];
  for (int i = 0; i < 8; ++i) {
    x[i] = -1e-3 * (threadIdx.x + i * 37 + 1);
  }
  double s = 0.;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s += exp(x[i]);
      x[i] -= 1e-4;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void sqrt_kernel(double *out, int iters) {
  double x[8];
  for (int i = 0; i < 8; ++i) {
    x[i] = 1e-3 * (threadIdx.x + i * 37 + 1);
  }
  double s = 0.;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s += sqrt(x[i]);
      x[i] += 1e-4;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  double *out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));

  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    for (int bps : {1, 2}) {
      if (warps * bps > 64) continue;
      float ms = time_ms([&] { dmma_kernel<16><<<sms * bps, warps * 32>>>(out, iters); });
      double flops = 2.0 * 8 * 8 * 4 * 16 * (double)iters * warps * bps * sms;
      printf("DMMA.8x8x4 nacc=16 warps/CTA=%2d CTAs/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms,
             flops / ms * 1e-9);
    }
  }
  for (int warps : {4, 8}) {
    float ms = time_ms([&] { dmma_kernel<4><<<sms, warps * 32>>>(out, iters); });
    double flops = 2.0 * 8 * 8 * 4 * 4 * (double)iters * warps * sms;
    printf("DMMA.8x8x4 nacc=4  warps/CTA=%2d            : %8.3f ms  %7.2f TFLOP/s\n", warps, ms,
           flops / ms * 1e-9);
  }
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dfma_kernel<16><<<sms * 2, warps * 32>>>(out, iters); });
    double flops = 2.0 * 32 * 16 * (double)iters * warps * 2 * sms;
    printf("DFMA nacc=16 warps/CTA=%2d CTAs/SM=2        : %8.3f ms  %7.2f TFLOP/s\n", warps, ms,
           flops / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { exp_kernel<<<sms * 4, 512>>>(out, 1024); });
    double n = 8.0 * 1024 * 512 * 4 * sms;
    printf("exp(double): %8.3f ms  %7.2f Gexp/s\n", ms, n / ms * 1e-6);
    ms = time_ms([&] { sqrt_kernel<<<sms * 4, 512>>>(out, 1024); });
    printf("sqrt(double): %8.3f ms  %7.2f Gsqrt/s\n", ms, n / ms * 1e-6);
  }

  cublasHandle_t cb;
  cublasCreate(&cb);
  struct Shape {
    int m, n, k;
    const char *what;
    int trans_b;
  };
  std::vector<Shape> shapes = {{8192, 8192, 8192, "DGEMM NN 8192^3", 0},
                               {8192, 8192, 8192, "DGEMM NT 8192^3", 1},
                               {16384, 16384, 1024, "DGEMM NT 16384^2 x 1024", 1},
                               {16384, 16384, 4096, "DGEMM NT 16384^2 x 4096", 1},
                               {16384, 16384, 256, "DGEMM NT 16384^2 x 256", 1}};
  const size_t maxel = (size_t)16384 * 16384;
  double *A, *B, *Cm;
  CK(cudaMalloc(&A, maxel * 8));
  CK(cudaMalloc(&B, maxel * 8));
  CK(cudaMalloc(&Cm, maxel * 8));
  CK(cudaMemset(A, 0, maxel * 8));
  CK(cudaMemset(B, 0, maxel * 8));
  CK(cudaMemset(Cm, 0, maxel * 8));
  const double alpha = -1., beta = 1.;
  for (auto &s : shapes) {
    float ms = time_ms([&] {
      cublasDgemm(cb, CUBLAS_OP_N, s.trans_b ? CUBLAS_OP_T : CUBLAS_OP_N, s.m, s.n, s.k, &alpha, A,
                  s.m, B, s.trans_b ? s.n : s.k, &beta, Cm, s.m);
    });
    printf("cuBLAS %-28s: %8.3f ms  %7.2f TFLOP/s\n", s.what, ms,
           2.0 * s.m * s.n * s.k / ms * 1e-9);
  }
  {
    const int n = 16384, k = 4096;
    float ms = time_ms([&] {
      cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, k, &alpha, A, n, &beta, Cm, n);
    });
    printf("cuBLAS DSYRK 16384 x 4096          : %8.3f ms  %7.2f TFLOP/s\n", ms,
           1.0 * n * n * k / ms * 1e-9);
  }
  // sustained: 8192^3 back to back for ~3 s
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int m = 8192;
    int reps = 100;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) {
      cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, m, m, m, &alpha, A, m, B, m, &beta, Cm, m);
    }
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cuBLAS DGEMM NT 8192^3 sustained x%d: %8.3f ms total  %7.2f TFLOP/s\n", reps, ms,
           2.0 * m * m * m * reps / ms * 1e-9);
  }
  return 0;
}
