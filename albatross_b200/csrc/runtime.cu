// Handle lifetime, device-buffer recycling, host<->device transfers and phase timing.
#include "common.cuh"

#include <cstdarg>
#include <cstring>

namespace ab {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// Freed buffers are kept in the handle and handed back for requests of the same size class:
// a 32 GiB cudaMalloc/cudaFree pair per tuner iteration would serialise the device.
int dev_alloc(ab_handle_s *h, size_t bytes, void **out) {
  bytes = static_cast<size_t>(round_up(static_cast<int64_t>(bytes < 256 ? 256 : bytes), 256));
  auto it = h->pool.lower_bound(bytes);
  if (it != h->pool.end() && it->first <= bytes + bytes / 8) {
    *out = it->second;
    h->pool_bytes -= it->first;
    h->pool.erase(it);
    return AB_OK;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {
    // retry once after dropping the cache
    for (auto &kv : h->pool) {
      cudaFree(kv.second);
    }
    h->pool.clear();
    h->pool_bytes = 0;
    cudaGetLastError();
    e = cudaMalloc(out, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return AB_ERR_ALLOC;
  }
  return AB_OK;
}

void dev_release(ab_handle_s *h, void *p, size_t bytes) {
  if (p == nullptr) {
    return;
  }
  bytes = static_cast<size_t>(round_up(static_cast<int64_t>(bytes < 256 ? 256 : bytes), 256));
  // Work touching `p` was enqueued on h->stream; later users are enqueued on the same stream, so
  // stream order makes recycling safe without a synchronize.
  h->pool.emplace(bytes, p);
  h->pool_bytes += bytes;
}

int matrix_new(ab_handle_s *h, int64_t rows, int64_t cols, ab_matrix_s **out) {
  if (rows < 0 || cols < 0) {
    set_error("negative matrix dimension");
    return AB_ERR_INVALID;
  }
  auto *m = new ab_matrix_s();
  m->rows = rows;
  m->cols = cols;
  m->ld = padded_ld(rows);
  m->bytes = static_cast<size_t>(m->ld) * static_cast<size_t>(cols < 1 ? 1 : cols) * sizeof(double);
  void *p = nullptr;
  int s = dev_alloc(h, m->bytes, &p);
  if (s != AB_OK) {
    delete m;
    return s;
  }
  m->d = static_cast<double *>(p);
  *out = m;
  return AB_OK;
}

void matrix_delete(ab_handle_s *h, ab_matrix_s *m) {
  if (m == nullptr) {
    return;
  }
  dev_release(h, m->d, m->bytes);
  delete m;
}

int upload(ab_handle_s *h, const double *host, int64_t rows, int64_t cols, ab_matrix_s **out) {
  ab_matrix_s *m = nullptr;
  AB_TRY(matrix_new(h, rows, cols, &m));
  if (rows > 0 && cols > 0) {
    cudaError_t e = cudaMemcpy2DAsync(m->d, m->ld * sizeof(double), host, rows * sizeof(double),
                                      rows * sizeof(double), cols, cudaMemcpyHostToDevice,
                                      h->stream);
    if (e != cudaSuccess) {
      matrix_delete(h, m);
      set_error("H2D copy failed: %s", cudaGetErrorString(e));
      return AB_ERR_CUDA;
    }
  }
  *out = m;
  return AB_OK;
}

int download(ab_handle_s *h, const ab_matrix_s *m, int64_t row0, int64_t col0, int64_t rows,
             int64_t cols, double *host) {
  if (rows == 0 || cols == 0) {
    return AB_OK;
  }
  AB_REQUIRE(row0 >= 0 && col0 >= 0 && row0 + rows <= m->rows && col0 + cols <= m->cols,
             "block out of range");
  AB_CUDA(cudaMemcpy2DAsync(host, rows * sizeof(double), m->d + row0 + col0 * m->ld,
                            m->ld * sizeof(double), rows * sizeof(double), cols,
                            cudaMemcpyDeviceToHost, h->stream));
  AB_CUDA(cudaStreamSynchronize(h->stream));
  return AB_OK;
}

void phase_begin(ab_handle_s *h, int phase) {
  if (!h->ev_used[phase]) {
    cudaEventRecord(h->ev_begin[phase], h->stream);
    h->ev_used[phase] = true;
  }
}

void phase_end(ab_handle_s *h, int phase) { cudaEventRecord(h->ev_end[phase], h->stream); }

void timings_reset(ab_handle_s *h) {
  for (int p = 0; p < PH_COUNT; ++p) {
    h->ev_used[p] = false;
  }
  h->total_used = true;
  cudaEventRecord(h->ev_total_begin, h->stream);
}

} // namespace ab

using namespace ab;

extern "C" {

const char *ab_last_error(void) { return g_error; }

int ab_version(void) { return AB_VERSION; }

int ab_device_count(int *count) {
  AB_REQUIRE(count != nullptr, "null");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return AB_ERR_CUDA;
  }
  return AB_OK;
}

int ab_create_on_stream(ab_handle *out, int device, void *cuda_stream) {
  AB_REQUIRE(out != nullptr, "null handle pointer");
  int count = 0;
  AB_TRY(ab_device_count(&count));
  if (count <= 0) {
    set_error("no CUDA device visible: albatross_b200 has no CPU fallback");
    return AB_ERR_CUDA;
  }
  AB_REQUIRE(device >= 0 && device < count, "device index out of range");
  AB_CUDA(cudaSetDevice(device));
  auto *h = new ab_handle_s();
  h->device = device;
  cudaDeviceProp prop;
  AB_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
              prop.minor);
    delete h;
    return AB_ERR_CUDA;
  }
  if (cuda_stream != nullptr) {
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
  } else {
    AB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  for (int p = 0; p < PH_COUNT; ++p) {
    AB_CUDA(cudaEventCreate(&h->ev_begin[p]));
    AB_CUDA(cudaEventCreate(&h->ev_end[p]));
  }
  AB_CUDA(cudaEventCreate(&h->ev_total_begin));
  AB_CUDA(cudaEventCreate(&h->ev_total_end));
  AB_CUDA(cudaMalloc(&h->d_scalars, 64 * sizeof(double)));
  AB_CUDA(cudaMalloc(&h->d_flags, 16 * sizeof(int)));
  AB_CUDA(cudaMallocHost(&h->h_scalars, 64 * sizeof(double)));
  AB_CUDA(cudaMallocHost(&h->h_flags, 16 * sizeof(int)));
  *out = h;
  return AB_OK;
}

int ab_create(ab_handle *out, int device) { return ab_create_on_stream(out, device, nullptr); }

int ab_trim(ab_handle h) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  AB_CUDA(cudaStreamSynchronize(h->stream));
  for (auto &kv : h->pool) {
    cudaFree(kv.second);
  }
  h->pool.clear();
  h->pool_bytes = 0;
  return AB_OK;
}

int ab_destroy(ab_handle h) {
  if (h == nullptr) {
    return AB_OK;
  }
  ab_dist_finalize(h);
  ab_trim(h);
  for (int p = 0; p < PH_COUNT; ++p) {
    cudaEventDestroy(h->ev_begin[p]);
    cudaEventDestroy(h->ev_end[p]);
  }
  cudaEventDestroy(h->ev_total_begin);
  cudaEventDestroy(h->ev_total_end);
  cudaFree(h->d_scalars);
  cudaFree(h->d_flags);
  cudaFreeHost(h->h_scalars);
  cudaFreeHost(h->h_flags);
  if (h->h_stage != nullptr) {
    cudaFreeHost(h->h_stage);
  }
  if (h->comm_stream != nullptr) { // left by a world-1 distributed fit (ab_dist_finalize not called)
    cudaStreamDestroy(h->comm_stream);
    cudaEventDestroy(h->ev_bcast[0]);
    cudaEventDestroy(h->ev_bcast[1]);
    cudaEventDestroy(h->ev_ready);
    cudaEventDestroy(h->ev_free);
  }
  if (h->panel_stream != nullptr) {
    cudaStreamDestroy(h->panel_stream);
    cudaEventDestroy(h->ev_panel);
    cudaEventDestroy(h->ev_col);
  }
  if (h->own_stream) {
    cudaStreamDestroy(h->stream);
  }
  delete h;
  return AB_OK;
}

int ab_synchronize(ab_handle h) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  AB_CUDA(cudaStreamSynchronize(h->stream));
  return AB_OK;
}

int ab_reset_counters(ab_handle h) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  h->launches = 0;
  return AB_OK;
}

int ab_timings(ab_handle h, ab_phase_times *out) {
  AB_REQUIRE(h != nullptr && out != nullptr, "null");
  Lock lock(h);
  AB_CUDA(cudaStreamSynchronize(h->stream));
  std::memset(out, 0, sizeof(*out));
  double *slots[PH_COUNT] = {&out->h2d_ms,    &out->gram_ms,    &out->factor_ms, &out->solve_ms,
                             &out->reduce_ms, &out->predict_ms, &out->d2h_ms};
  for (int p = 0; p < PH_COUNT; ++p) {
    if (h->ev_used[p]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->ev_begin[p], h->ev_end[p]) == cudaSuccess) {
        *slots[p] = ms;
      } else {
        cudaGetLastError();
      }
    }
  }
  if (h->total_used) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_total_begin, h->ev_total_end) == cudaSuccess) {
      out->total_ms = ms;
    } else {
      cudaGetLastError();
    }
  }
  out->kernel_launches = h->launches;
  return AB_OK;
}

int ab_matrix_upload(ab_handle h, const double *host, int64_t rows, int64_t cols,
                     ab_matrix *out) {
  AB_REQUIRE(h != nullptr && out != nullptr && (host != nullptr || rows * cols == 0), "null");
  Lock lock(h);
  AB_TRY(upload(h, host, rows, cols, out));
  AB_CUDA(cudaStreamSynchronize(h->stream)); // `host` may be pageable and reused by the caller
  return AB_OK;
}

int ab_matrix_alloc(ab_handle h, int64_t rows, int64_t cols, ab_matrix *out) {
  AB_REQUIRE(h != nullptr && out != nullptr, "null");
  Lock lock(h);
  return matrix_new(h, rows, cols, out);
}

int ab_matrix_download(ab_handle h, ab_matrix m, double *host) {
  AB_REQUIRE(h != nullptr && m != nullptr && host != nullptr, "null");
  Lock lock(h);
  return download(h, m, 0, 0, m->rows, m->cols, host);
}

int ab_matrix_download_block(ab_handle h, ab_matrix m, int64_t row0, int64_t col0, int64_t rows,
                             int64_t cols, double *host) {
  AB_REQUIRE(h != nullptr && m != nullptr && host != nullptr, "null");
  Lock lock(h);
  return download(h, m, row0, col0, rows, cols, host);
}

int ab_matrix_dims(ab_matrix m, int64_t *rows, int64_t *cols) {
  AB_REQUIRE(m != nullptr, "null");
  if (rows != nullptr) {
    *rows = m->rows;
  }
  if (cols != nullptr) {
    *cols = m->cols;
  }
  return AB_OK;
}

int ab_matrix_device_ptr(ab_matrix m, void **ptr, int64_t *ld) {
  AB_REQUIRE(m != nullptr, "null");
  if (ptr != nullptr) {
    *ptr = m->d;
  }
  if (ld != nullptr) {
    *ld = m->ld;
  }
  return AB_OK;
}

int ab_matrix_free(ab_handle h, ab_matrix m) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  matrix_delete(h, m);
  return AB_OK;
}

} // extern "C"
