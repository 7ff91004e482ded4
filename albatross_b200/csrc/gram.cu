// Gram-matrix construction: tiled pairwise distance + fused covariance program.
//
// Replaces compute_covariance_matrix (reference include/albatross/src/covariance_functions/
// callers.hpp:38-166) and the leaf kernels of radial.hpp / noise.hpp / polynomials.hpp.
//
// Layout: output column-major fp64, 64x64 tiles, one CTA (256 threads) per tile.  For the symmetric
// build only tiles on/below the diagonal are evaluated; the transposed tile is written from a
// shared-memory staging buffer so that both orientations are stored as full 256/512-byte rows.
#include "gram.cuh"

#include <cmath>

namespace ab {

// ------------------------------------------------------------------------------------------------
// host: postfix program -> device program
// ------------------------------------------------------------------------------------------------

static bool leaf_to_dev(const ab_op &o, DevOp *d) {
  d->flags = 0;
  switch (o.op) {
  case AB_OP_SQUARED_EXPONENTIAL:
    d->kind = o.p0 > 0. ? DK_SE : DK_ZERO;
    d->c0 = -1. / (o.p0 * o.p0);
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_EXPONENTIAL:
    d->kind = o.p0 > 0. ? DK_EXP : DK_ZERO;
    d->c0 = -1. / o.p0;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_MATERN32:
    d->kind = o.p0 > 0. ? DK_M32 : DK_ZERO;
    d->c0 = std::sqrt(3.) / o.p0;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_MATERN52:
    d->kind = o.p0 > 0. ? DK_M52 : DK_ZERO;
    d->c0 = std::sqrt(5.) / o.p0;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_CONSTANT:
    d->kind = DK_CONST;
    d->c0 = 0.;
    d->amp = o.p0 * o.p0;
    return true;
  case AB_OP_INDEPENDENT_NOISE:
    d->kind = DK_NOISE;
    d->c0 = 0.;
    d->amp = o.p0 * o.p0;
    return true;
  default:
    return false;
  }
}

int compile_program(const ab_op *prog, int nops, DevProg *out) {
  AB_REQUIRE(prog != nullptr && nops >= 1 && nops <= AB_MAX_OPS, "covariance program size");
  // validate postfix shape
  int depth = 0, max_depth = 0;
  for (int k = 0; k < nops; ++k) {
    if (prog[k].op == AB_OP_SUM || prog[k].op == AB_OP_PRODUCT) {
      AB_REQUIRE(depth >= 2, "malformed postfix covariance program");
      depth -= 1;
    } else {
      AB_REQUIRE(prog[k].op >= AB_OP_SQUARED_EXPONENTIAL && prog[k].op <= AB_OP_INDEPENDENT_NOISE,
                 "unknown covariance opcode");
      depth += 1;
    }
    max_depth = depth > max_depth ? depth : max_depth;
  }
  AB_REQUIRE(depth == 1, "malformed postfix covariance program");
  AB_REQUIRE(max_depth <= 8, "covariance program nests deeper than 8");

  out->need_dist = 0;
  out->need_equal = 0;
  for (int k = 0; k < nops; ++k) {
    if (prog[k].op == AB_OP_EXPONENTIAL || prog[k].op == AB_OP_MATERN32 ||
        prog[k].op == AB_OP_MATERN52) {
      out->need_dist = 1;
    }
    if (prog[k].op == AB_OP_INDEPENDENT_NOISE) {
      out->need_equal = 1;
    }
  }

  // Try the stack-free sum-of-products form:  leaf (leaf *)* ( leaf (leaf *)* + )*
  // i.e. a left-associated sum of left-associated products of leaves.
  {
    DevProg sop = *out;
    int n = 0;
    bool ok = true;
    int k = 0;
    int term = 0;
    while (k < nops && ok) {
      // one term: leaf (leaf PRODUCT)*
      if (!leaf_to_dev(prog[k], &sop.ops[n])) {
        ok = false;
        break;
      }
      sop.ops[n].flags = DF_TERM_START;
      ++n;
      ++k;
      while (k + 1 < nops && prog[k + 1].op == AB_OP_PRODUCT && leaf_to_dev(prog[k], &sop.ops[n])) {
        sop.ops[n].flags = 0;
        ++n;
        k += 2;
      }
      sop.ops[n - 1].flags |= DF_TERM_END | (term == 0 ? DF_FIRST_TERM : 0);
      if (term > 0) {
        if (k < nops && prog[k].op == AB_OP_SUM) {
          ++k;
        } else {
          ok = false;
        }
      }
      ++term;
    }
    if (ok && k == nops) {
      sop.nops = n;
      sop.mode = 0;
      *out = sop;
      return AB_OK;
    }
  }
  // generic stack form
  out->mode = 1;
  out->nops = nops;
  for (int k = 0; k < nops; ++k) {
    if (prog[k].op == AB_OP_SUM) {
      out->ops[k] = DevOp{DK_SUM, 0, 0., 0.};
    } else if (prog[k].op == AB_OP_PRODUCT) {
      out->ops[k] = DevOp{DK_PROD, 0, 0., 0.};
    } else {
      leaf_to_dev(prog[k], &out->ops[k]);
    }
  }
  return AB_OK;
}

// ------------------------------------------------------------------------------------------------
// device: evaluation
// ------------------------------------------------------------------------------------------------

// exp(x) for x <= 0 (every radial kernel's argument).  NaN propagates; x < -745.2 underflows to 0.
__device__ __forceinline__ double exp_nonpos(double x) { return exp(x); }

__device__ __forceinline__ double leaf_value(const DevOp &o, double d2, double dist, bool equal) {
  switch (o.kind) {
  case DK_SE:
    return o.amp * exp_nonpos(o.c0 * d2);
  case DK_EXP:
    return o.amp * exp_nonpos(o.c0 * dist);
  case DK_M32: {
    const double s = o.c0 * dist;
    return o.amp * (1. + s) * exp_nonpos(-s);
  }
  case DK_M52: {
    const double s = o.c0 * dist;
    return o.amp * (1. + s + s * s * (1. / 3.)) * exp_nonpos(-s);
  }
  case DK_CONST:
    return o.amp;
  case DK_NOISE:
    return equal ? o.amp : 0.;
  default:
    return 0.;
  }
}

// Sum-of-products evaluation of NP pairs at once: the op loop is uniform across the CTA and its
// decode cost is amortised over the NP pairs a thread owns.
template <int NP>
__device__ __forceinline__ void eval_sop(const DevProg &P, const double (&d2)[NP],
                                         const double (&dist)[NP], unsigned eqmask,
                                         double (&out)[NP]) {
  double prod[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    prod[i] = 0.;
    out[i] = 0.;
  }
  for (int k = 0; k < P.nops; ++k) {
    const DevOp o = P.ops[k];
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      v[i] = leaf_value(o, d2[i], dist[i], (eqmask >> i) & 1u);
    }
    if (o.flags & DF_TERM_START) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        prod[i] = v[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        prod[i] = (prod[i] != 0.) ? prod[i] * v[i] : prod[i]; // covariance_function.hpp:362-366
      }
    }
    if (o.flags & DF_TERM_END) {
      if (o.flags & DF_FIRST_TERM) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          out[i] = prod[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          out[i] += prod[i];
        }
      }
    }
  }
}

__device__ __noinline__ double eval_stack(const DevProg &P, double d2, double dist, bool equal) {
  double stack[8];
  int sp = 0;
  for (int k = 0; k < P.nops; ++k) {
    const DevOp o = P.ops[k];
    if (o.kind == DK_SUM) {
      const double rhs = stack[--sp];
      stack[sp - 1] = stack[sp - 1] + rhs;
    } else if (o.kind == DK_PROD) {
      const double rhs = stack[--sp];
      const double lhs = stack[sp - 1];
      stack[sp - 1] = (lhs != 0.) ? lhs * rhs : lhs;
    } else {
      stack[sp++] = leaf_value(o, d2, dist, equal);
    }
  }
  return stack[0];
}

constexpr int TILE = 64;
constexpr int LDT = TILE + 1;
constexpr int GRAM_THREADS = 256;
constexpr int NPAIR = 16; // 2 rows x 8 columns per thread

// SYM: blockIdx.x enumerates tiles (I >= J) of the lower triangle; otherwise I = b % tiles_i.
template <int DIM, bool SYM>
__global__ void __launch_bounds__(GRAM_THREADS)
gram_kernel(const __grid_constant__ DevProg P, const double *__restrict__ fx, int64_t ldfx,
            int64_t n, const double *__restrict__ fy, int64_t ldfy, int64_t m,
            double *__restrict__ out, int64_t ld, int tiles_i, uint32_t flags) {
  __shared__ double xs[TILE * DIM];
  __shared__ double ys[TILE * DIM];
  __shared__ double stage[SYM ? TILE * LDT : 1];

  int64_t I, J;
  if (SYM) {
    const int64_t t = blockIdx.x;
    int64_t i = static_cast<int64_t>((sqrt(8. * static_cast<double>(t) + 1.) - 1.) * 0.5);
    while (i * (i + 1) / 2 > t) {
      --i;
    }
    while ((i + 1) * (i + 2) / 2 <= t) {
      ++i;
    }
    I = i;
    J = t - i * (i + 1) / 2;
  } else {
    I = blockIdx.x % tiles_i;
    J = blockIdx.x / tiles_i;
  }
  const int64_t i0 = I * TILE;
  const int64_t j0 = J * TILE;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  for (int idx = tid; idx < TILE * DIM; idx += GRAM_THREADS) {
    const int p = idx / DIM;
    const int d = idx - p * DIM;
    xs[idx] = i0 + p < n ? fx[(i0 + p) * ldfx + d] : 0.;
    ys[idx] = j0 + p < m ? fy[(j0 + p) * ldfy + d] : 0.;
  }
  __syncthreads();

  const int r0 = 2 * lane;
  double xi[2][DIM];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      xi[a][d] = xs[(r0 + a) * DIM + d];
    }
  }

  double d2[NPAIR], dist[NPAIR], vals[NPAIR];
  unsigned eqmask = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = warp * 8 + k;
    double yj[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      yj[d] = ys[c * DIM + d];
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      double s = 0.;
      bool eq = true;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double diff = xi[a][d] - yj[d];
        s = fma(diff, diff, s);
        eq = eq && (xi[a][d] == yj[d]);
      }
      d2[2 * k + a] = s;
      if (DIM == 1) {
        dist[2 * k + a] = fabs(xi[a][0] - yj[0]);
      }
      eqmask |= (eq ? 1u : 0u) << (2 * k + a);
    }
  }
  if (DIM != 1) {
    if (P.need_dist) {
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) {
        dist[i] = sqrt(d2[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) {
        dist[i] = 0.;
      }
    }
  }

  if (P.mode == 0) {
    eval_sop<NPAIR>(P, d2, dist, eqmask, vals);
  } else {
#pragma unroll 1
    for (int i = 0; i < NPAIR; ++i) {
      vals[i] = eval_stack(P, d2[i], dist[i], (eqmask >> i) & 1u);
    }
  }

  // direct tile: rows i0 + r0 (+1), columns j0 + c; 16-byte stores, 512 contiguous bytes per warp.
  const int64_t gi = i0 + r0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t gj = j0 + warp * 8 + k;
    if (gj < m) {
      double *dst = out + gi + gj * ld;
      if (gi + 1 < n) {
        *reinterpret_cast<double2 *>(dst) = make_double2(vals[2 * k], vals[2 * k + 1]);
      } else if (gi < n) {
        dst[0] = vals[2 * k];
      }
    }
  }

  if (SYM) {
    if (I != J && !(flags & AB_GRAM_LOWER_ONLY)) {
      // transposed tile through shared memory: stage[c][r], read back with r fixed per warp-store.
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = warp * 8 + k;
        stage[c * LDT + r0] = vals[2 * k];
        stage[c * LDT + r0 + 1] = vals[2 * k + 1];
      }
      __syncthreads();
      // mirror element (row = j0 + c, col = i0 + r) = stage[c][r]
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = warp * 8 + k;
        const int64_t gcol = i0 + r;
        if (gcol < n) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c = lane + 32 * half;
            const int64_t grow = j0 + c;
            if (grow < n) {
              out[grow + gcol * ld] = stage[c * LDT + r];
            }
          }
        }
      }
    }
  }
}

template <int DIM>
__global__ void gram_diag_kernel(const __grid_constant__ DevProg P, const double *__restrict__ f,
                                 int64_t ldf, int64_t n, double *__restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) {
    return;
  }
  // k(x_i, x_i): distance 0, features equal unless a coordinate is NaN (NaN != NaN).
  bool eq = true;
  for (int d = 0; d < DIM; ++d) {
    const double v = f[i * ldf + d];
    eq = eq && (v == v);
  }
  if (P.mode == 0) {
    double d2[1] = {0.}, dist[1] = {0.}, v[1];
    eval_sop<1>(P, d2, dist, eq ? 1u : 0u, v);
    out[i] = v[0];
  } else {
    out[i] = eval_stack(P, 0., 0., eq);
  }
}

template <bool SYM>
static int launch_gram(ab_handle_s *h, const DevProg &P, int dim, const double *fx, int64_t ldfx,
                       int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out,
                       int64_t ld, uint32_t flags) {
  if (n == 0 || m == 0) {
    return AB_OK;
  }
  const int64_t ti = (n + TILE - 1) / TILE;
  const int64_t tj = (m + TILE - 1) / TILE;
  const int64_t tiles = SYM ? ti * (ti + 1) / 2 : ti * tj;
  AB_REQUIRE(tiles < (int64_t(1) << 31), "Gram too large for one launch");
  const dim3 grid(static_cast<unsigned>(tiles));
  const dim3 block(GRAM_THREADS);
#define AB_GRAM_CASE(D)                                                                        \
  case D:                                                                                      \
    gram_kernel<D, SYM><<<grid, block, 0, h->stream>>>(P, fx, ldfx, n, fy, ldfy, m, out, ld,  \
                                                       static_cast<int>(ti), flags);          \
    break;
  switch (dim) {
    AB_GRAM_CASE(1)
    AB_GRAM_CASE(2)
    AB_GRAM_CASE(3)
    AB_GRAM_CASE(4)
    AB_GRAM_CASE(5)
    AB_GRAM_CASE(6)
    AB_GRAM_CASE(7)
    AB_GRAM_CASE(8)
  default:
    set_error("feature dimension %d has no device form (1..%d supported)", dim, AB_MAX_DIM);
    return AB_ERR_UNSUPPORTED;
  }
#undef AB_GRAM_CASE
  AB_LAUNCHED(h);
  return AB_OK;
}

int gram_sym_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *feats, uint32_t flags,
                    ab_matrix_s **out) {
  const int dim = static_cast<int>(feats->rows);
  const int64_t n = feats->cols;
  ab_matrix_s *K = nullptr;
  AB_TRY(matrix_new(h, n, n, &K));
  int s = launch_gram<true>(h, P, dim, feats->d, feats->ld, n, feats->d, feats->ld, n, K->d, K->ld,
                              flags);
  if (s != AB_OK) {
    matrix_delete(h, K);
    return s;
  }
  *out = K;
  return AB_OK;
}

int gram_cross_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *fx,
                      const ab_matrix_s *fy, ab_matrix_s **out) {
  AB_REQUIRE(fx->rows == fy->rows, "feature dimensions differ");
  const int dim = static_cast<int>(fx->rows);
  ab_matrix_s *K = nullptr;
  AB_TRY(matrix_new(h, fx->cols, fy->cols, &K));
  int s = launch_gram<false>(h, P, dim, fx->d, fx->ld, fx->cols, fy->d, fy->ld, fy->cols, K->d,
                               K->ld, 0u);
  if (s != AB_OK) {
    matrix_delete(h, K);
    return s;
  }
  *out = K;
  return AB_OK;
}

int gram_diag_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *feats, double *d_out) {
  const int dim = static_cast<int>(feats->rows);
  const int64_t n = feats->cols;
  if (n == 0) {
    return AB_OK;
  }
  const dim3 block(256);
  const dim3 grid(static_cast<unsigned>((n + 255) / 256));
#define AB_DIAG_CASE(D)                                                                        \
  case D:                                                                                      \
    gram_diag_kernel<D><<<grid, block, 0, h->stream>>>(P, feats->d, feats->ld, n, d_out);     \
    break;
  switch (dim) {
    AB_DIAG_CASE(1)
    AB_DIAG_CASE(2)
    AB_DIAG_CASE(3)
    AB_DIAG_CASE(4)
    AB_DIAG_CASE(5)
    AB_DIAG_CASE(6)
    AB_DIAG_CASE(7)
    AB_DIAG_CASE(8)
  default:
    set_error("feature dimension %d has no device form (1..%d supported)", dim, AB_MAX_DIM);
    return AB_ERR_UNSUPPORTED;
  }
#undef AB_DIAG_CASE
  AB_LAUNCHED(h);
  return AB_OK;
}

// Features live on the device as a dim x n matrix whose columns are PACKED (ld == dim) so that the
// AoS stream of std::vector<X> is reproduced byte for byte.
int upload_features(ab_handle_s *h, const double *feats, int64_t n, int dim, ab_matrix_s **out) {
  AB_REQUIRE(dim >= 1 && dim <= AB_MAX_DIM, "feature dimension");
  auto *m = new ab_matrix_s();
  m->rows = dim;
  m->cols = n;
  m->ld = dim;
  m->bytes = static_cast<size_t>(dim) * static_cast<size_t>(n < 1 ? 1 : n) * sizeof(double);
  void *p = nullptr;
  int s = dev_alloc(h, m->bytes, &p);
  if (s != AB_OK) {
    delete m;
    return s;
  }
  m->d = static_cast<double *>(p);
  if (n > 0) {
    cudaError_t e = cudaMemcpyAsync(m->d, feats, static_cast<size_t>(dim) * n * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) {
      matrix_delete(h, m);
      set_error("feature upload failed: %s", cudaGetErrorString(e));
      return AB_ERR_CUDA;
    }
  }
  *out = m;
  return AB_OK;
}

} // namespace ab
