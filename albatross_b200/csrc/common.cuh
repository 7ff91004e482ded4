// Shared internals of libalbatross_b200.so: handle, device buffers, error plumbing.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/albatross_b200.h"

namespace ab {

void set_error(const char *fmt, ...);

#define AB_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ab::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return AB_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define AB_TRY(expr)                                                                           \
  do {                                                                                         \
    int _s = (expr);                                                                           \
    if (_s != AB_OK) {                                                                         \
      return _s;                                                                               \
    }                                                                                          \
  } while (0)

#define AB_REQUIRE(cond, msg)                                                                  \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      ab::set_error("%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg);    \
      return AB_ERR_INVALID;                                                                   \
    }                                                                                          \
  } while (0)

// Checks the launch of the kernel just enqueued and counts it.
#define AB_LAUNCHED(h)                                                                         \
  do {                                                                                         \
    AB_CUDA(cudaGetLastError());                                                               \
    (h)->launches++;                                                                           \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Kernel function attributes (the dynamic shared-memory opt-in) belong to a device, not to the process: one
// process may drive several GPUs through several handles (the tuner's finite-difference evaluations,
// include/albatross_b200/tune.hpp).  `static PerDeviceOnce once; if (once.need(h->device)) { ... }`.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need(int device) {
    if (device < 0 || device >= 64) {
      return true;
    }
    if (done[device]) {
      return false;
    }
    done[device] = true;
    return true;
  }
};

// Leading dimension of a device matrix with `rows` rows: 16-double (128 B) aligned columns, and
// never a multiple of 1024 doubles so that column walks do not alias in L2/HBM channels.
inline int64_t padded_ld(int64_t rows) {
  int64_t ld = round_up(rows < 1 ? 1 : rows, 16);
  if (rows >= 2048 && ld % 1024 == 0) {
    ld += 16;
  }
  return ld;
}

} // namespace ab

struct ab_matrix_s {
  double *d = nullptr;
  int64_t rows = 0;
  int64_t cols = 0;
  int64_t ld = 0;
  size_t bytes = 0;
};

struct ab_factor_s {
  ab_matrix_s *m = nullptr; // lower triangle holds L (Cholesky, diag = sqrt(D))
  int64_t n = 0;
  int64_t bad_pivot = -1;
  double *dinv = nullptr; // explicit inverses of the LEAF x LEAF diagonal blocks of L
  size_t dinv_bytes = 0;
};

enum ab_phase { PH_H2D = 0, PH_GRAM, PH_FACTOR, PH_SOLVE, PH_REDUCE, PH_PREDICT, PH_D2H, PH_COUNT };

struct ab_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::recursive_mutex mu;
  int64_t launches = 0;
  // recycled device buffers, keyed by size in bytes
  std::multimap<size_t, void *> pool;
  size_t pool_bytes = 0;
  // phase timing: pairs of events recorded on `stream`
  cudaEvent_t ev_begin[PH_COUNT] = {};
  cudaEvent_t ev_end[PH_COUNT] = {};
  bool ev_used[PH_COUNT] = {};
  cudaEvent_t ev_total_begin = nullptr, ev_total_end = nullptr;
  bool total_used = false;
  // small device scratch for scalars / flags
  double *d_scalars = nullptr; // 64 doubles
  int *d_flags = nullptr;      // 16 ints
  double *h_scalars = nullptr; // pinned mirror
  int *h_flags = nullptr;
  // pinned staging for host<->device vector traffic
  void *h_stage = nullptr;
  size_t h_stage_bytes = 0;
  // distributed group (dist.cu): one process per GPU, NCCL communicator over NVLink / NVSwitch
  void *comm = nullptr; // ncclComm_t
  int rank = 0;
  int world = 1;
  cudaStream_t comm_stream = nullptr; // high-priority stream for panel broadcasts
  // single-GPU look-ahead factorisation (linalg.cu): panel chain on a high-priority stream
  cudaStream_t panel_stream = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_col = nullptr;
  cudaEvent_t ev_bcast[2] = {nullptr, nullptr};
  cudaEvent_t ev_ready = nullptr, ev_free = nullptr;
};

namespace ab {

struct Lock {
  explicit Lock(ab_handle_s *h) : h_(h) {
    h_->mu.lock();
    cudaSetDevice(h_->device);
  }
  ~Lock() { h_->mu.unlock(); }
  ab_handle_s *h_;
};

int dev_alloc(ab_handle_s *h, size_t bytes, void **out);
void dev_release(ab_handle_s *h, void *p, size_t bytes);
int matrix_new(ab_handle_s *h, int64_t rows, int64_t cols, ab_matrix_s **out);
void matrix_delete(ab_handle_s *h, ab_matrix_s *m);
int upload(ab_handle_s *h, const double *host, int64_t rows, int64_t cols, ab_matrix_s **out);
int download(ab_handle_s *h, const ab_matrix_s *m, int64_t row0, int64_t col0, int64_t rows,
             int64_t cols, double *host);

void phase_begin(ab_handle_s *h, int phase);
void phase_end(ab_handle_s *h, int phase);
void timings_reset(ab_handle_s *h);

} // namespace ab
