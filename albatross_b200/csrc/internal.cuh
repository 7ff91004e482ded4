// Helpers shared by the GP-level translation units (gp.cu, sparse.cu, dist.cu): RAII scope for
// temporaries, factor objects, host<->device vector moves.
#pragma once

#include "gram.cuh"
#include "linalg.cuh"

#include <algorithm>
#include <climits>
#include <vector>

namespace ab {

int gram_sym_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *feats, uint32_t flags,
                    ab_matrix_s **out);
int gram_cross_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *fx,
                      const ab_matrix_s *fy, ab_matrix_s **out);
int gram_diag_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *feats, double *d_out);
int upload_features(ab_handle_s *h, const double *feats, int64_t n, int dim, ab_matrix_s **out);

// ---- distributed group (dist.cu); all are no-ops / identities on a handle without ab_dist_init ----
int dist_rank(const ab_handle_s *h);
int dist_world(const ab_handle_s *h);
// In-place sum over ranks of `count` doubles in device memory, on the handle's stream.
int dist_allreduce_sum(ab_handle_s *h, double *d_buf, int64_t count);
// Sum over ranks of a host integer (synchronises the stream).
int64_t dist_total(ab_handle_s *h, int64_t local);

// Shared implementation of ab_gp_cv / ab_dist_gp_cv (gp.cu).
int gp_cv_impl(ab_handle_s *h, ab_factor_s *f, const double *y, const double *information,
               const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
               int phase, int stride, double *mean, double *var, double *joint, double *score,
               double *group_scores);

// RAII for temporaries so that every early return recycles device buffers.
struct Scope {
  explicit Scope(ab_handle_s *h) : h(h) {}
  ~Scope() {
    for (auto *m : mats) {
      matrix_delete(h, m);
    }
    for (auto &b : bufs) {
      dev_release(h, b.first, b.second);
    }
  }
  ab_matrix_s *own(ab_matrix_s *m) {
    mats.push_back(m);
    return m;
  }
  void disown(ab_matrix_s *m) { mats.erase(std::remove(mats.begin(), mats.end(), m), mats.end()); }
  int alloc(size_t bytes, void **out) {
    int s = dev_alloc(h, bytes, out);
    if (s == AB_OK) {
      bufs.emplace_back(*out, bytes);
    }
    return s;
  }
  ab_handle_s *h;
  std::vector<ab_matrix_s *> mats;
  std::vector<std::pair<void *, size_t>> bufs;
};

inline MatView view(const ab_matrix_s *m) { return MatView{m->d, m->ld}; }

inline int upload_bytes(ab_handle_s *h, Scope &sc, const void *host, size_t bytes, void **dev) {
  AB_TRY(sc.alloc(bytes, dev));
  if (bytes > 0) {
    AB_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  return AB_OK;
}

inline int download_bytes(ab_handle_s *h, const void *dev, size_t bytes, void *host) {
  if (bytes > 0) {
    AB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  AB_CUDA(cudaStreamSynchronize(h->stream));
  return AB_OK;
}

inline int new_factor(ab_handle_s *h, ab_matrix_s *m, ab_factor_s **out) {
  AB_REQUIRE(m->rows == m->cols, "factorisation needs a square matrix");
  auto *f = new ab_factor_s();
  f->m = m;
  f->n = m->rows;
  const int64_t nblocks = (f->n + LEAF - 1) / LEAF;
  f->dinv_bytes = static_cast<size_t>(nblocks < 1 ? 1 : nblocks) * LEAF * LEAF * sizeof(double);
  void *p = nullptr;
  int s = dev_alloc(h, f->dinv_bytes, &p);
  if (s != AB_OK) {
    delete f;
    return s;
  }
  f->dinv = static_cast<double *>(p);
  *out = f;
  return AB_OK;
}

inline void delete_factor(ab_handle_s *h, ab_factor_s *f) {
  if (f == nullptr) {
    return;
  }
  matrix_delete(h, f->m);
  dev_release(h, f->dinv, f->dinv_bytes);
  delete f;
}

// Factor `m` in place (consumed) and report the first bad pivot.
inline int factorize(ab_handle_s *h, ab_matrix_s *m, ab_factor_s **out) {
  ab_factor_s *f = nullptr;
  int s = new_factor(h, m, &f);
  if (s != AB_OK) {
    matrix_delete(h, m);
    return s;
  }
  h->h_flags[0] = INT_MAX;
  cudaMemcpyAsync(h->d_flags, h->h_flags, sizeof(int), cudaMemcpyHostToDevice, h->stream);
  s = potrf(h, view(m), f->n, f->dinv, h->d_flags);
  if (s == AB_OK) {
    cudaError_t e =
        cudaMemcpyAsync(h->h_flags, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) {
      e = cudaStreamSynchronize(h->stream);
    }
    if (e != cudaSuccess) {
      set_error("factorisation failed: %s", cudaGetErrorString(e));
      s = AB_ERR_CUDA;
    }
  }
  if (s != AB_OK) {
    delete_factor(h, f);
    return s;
  }
  f->bad_pivot = h->h_flags[0] == INT_MAX ? -1 : h->h_flags[0];
  *out = f;
  if (f->bad_pivot >= 0) {
    set_error("matrix is not positive definite: pivot %lld is <= 0 or NaN",
              static_cast<long long>(f->bad_pivot));
    return AB_ERR_NOT_PD;
  }
  return AB_OK;
}

inline int require_usable(const ab_factor_s *f) {
  AB_REQUIRE(f != nullptr && f->m != nullptr, "null factor");
  if (f->bad_pivot >= 0) {
    set_error("factor is not usable: matrix was not positive definite (pivot %lld)",
              static_cast<long long>(f->bad_pivot));
    return AB_ERR_NOT_PD;
  }
  return AB_OK;
}

} // namespace ab
