// Gram-matrix construction: tiled pairwise distance + fused covariance program.
//
// Replaces compute_covariance_matrix (reference include/albatross/src/covariance_functions/
// callers.hpp:38-166) and the leaf kernels of radial.hpp / noise.hpp / polynomials.hpp.
//
// Layout: output column-major fp64, 64x64 tiles, one CTA (256 threads) per tile.  For the symmetric
// build only tiles on/below the diagonal are evaluated; the transposed tile is written from a
// shared-memory staging buffer so that both orientations are stored as full 256/512-byte rows.
//
// Roofline: the kernel must retire one pair per SM-clock to keep up with HBM (8 B per pair), i.e.
// <= 64 FP64-pipe instructions per pair.  libdevice exp()/sqrt() alone cost ~21/17 each, so the
// evaluator uses its own exp (128-entry 2^(j/128) table in shared memory + degree-5 polynomial,
// 10 FP64 ops, <= 1 ulp) and sqrt (MUFU.RSQ64H seed + Goldschmidt, 7 FP64 ops, <= 1 ulp).
#include "gram_kernel.cuh"

#include <algorithm>
#include <vector>

#include <cstring>

#include <cmath>

namespace ab {

// ------------------------------------------------------------------------------------------------
// host: postfix program -> device program
// ------------------------------------------------------------------------------------------------

static bool leaf_to_dev_raw(const ab_op &o, DevOp *d) {
  d->flags = 0;
  d->a2 = d->a1 = d->b2 = d->b1 = 0.;
  const double l = o.p0;
  switch (o.op) {
  case AB_OP_SQUARED_EXPONENTIAL:
    d->kind = l <= 0. ? DK_ZERO : DK_RADIAL; // a NaN length scale propagates (radial.hpp:28-30)
    d->a2 = -1. / (l * l);
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_EXPONENTIAL:
    d->kind = l <= 0. ? DK_ZERO : DK_RADIAL; // a NaN length scale propagates (radial.hpp:28-30)
    d->flags = DF_USES_DIST;
    d->a1 = -1. / l;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_MATERN32:
    d->kind = l <= 0. ? DK_ZERO : DK_RADIAL; // a NaN length scale propagates (radial.hpp:28-30)
    d->flags = DF_USES_DIST | DF_POLY_D1;
    d->a1 = -std::sqrt(3.) / l;
    d->b1 = std::sqrt(3.) / l;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_MATERN52:
    d->kind = l <= 0. ? DK_ZERO : DK_RADIAL; // a NaN length scale propagates (radial.hpp:28-30)
    d->flags = DF_USES_DIST | DF_POLY_D1 | DF_POLY_D2;
    d->a1 = -std::sqrt(5.) / l;
    d->b1 = std::sqrt(5.) / l;
    d->b2 = 5. / (3. * l * l);
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_CONSTANT:
    d->kind = DK_CONST;
    d->amp = o.p0 * o.p0;
    return true;
  case AB_OP_INDEPENDENT_NOISE:
    d->kind = DK_NOISE;
    d->amp = o.p0 * o.p0;
    return true;
  default:
    return false;
  }
}

static bool leaf_to_dev(const ab_op &o, DevOp *d) {
  d->ab1 = d->ab2 = d->a2s = d->a1s = 0.;
  d->lim_hi = d->pad_ = 0;
  if (!leaf_to_dev_raw(o, d)) {
    return false;
  }
  d->ab1 = d->amp * d->b1;
  d->ab2 = d->amp * d->b2;
  constexpr double steps_per_unit = 2954.639443740597; // 2048 / ln 2
  d->a2s = d->a2 * steps_per_unit;
  d->a1s = d->a1 * steps_per_unit;
  const double a = (d->flags & DF_USES_DIST) ? d->a1 : d->a2;
  const double limit = a < 0. ? 708. / -a : 0.;          // a >= 0 cannot occur for a live radial leaf
  uint64_t bits = 0;
  std::memcpy(&bits, &limit, sizeof bits);
  const int64_t hi = static_cast<int64_t>(bits >> 32) - 1; // one high-word step below the limit
  d->lim_hi = (limit > 0. && limit < 1e300 && hi > 0) ? static_cast<int>(hi) : 0;
  d->pad_ = 0;
  return true;
}

static int compile_stationary(const ab_op *prog, int nops, DevProg *out);

// Splits the top-level sum of the postfix program into its summands (index ranges of complete sub-programs).
static bool top_level_summands(const ab_op *prog, int nops, std::vector<std::pair<int, int>> *terms) {
  // span[k] = first index of the sub-expression that ends at k
  std::vector<int> span(static_cast<size_t>(nops)), stack;
  for (int k = 0; k < nops; ++k) {
    if (prog[k].op == AB_OP_SUM || prog[k].op == AB_OP_PRODUCT) {
      if (stack.size() < 2) {
        return false;
      }
      stack.pop_back();
      span[static_cast<size_t>(k)] = span[static_cast<size_t>(stack.back())];
      stack.back() = k;
    } else {
      span[static_cast<size_t>(k)] = k;
      stack.push_back(k);
    }
  }
  if (stack.size() != 1) {
    return false;
  }
  // walk down the SUM spine from the root
  std::vector<int> todo = {nops - 1};
  while (!todo.empty()) {
    const int k = todo.back();
    todo.pop_back();
    if (prog[k].op == AB_OP_SUM) {
      const int rhs_end = k - 1;
      const int lhs_end = span[static_cast<size_t>(rhs_end)] - 1;
      todo.push_back(rhs_end);
      todo.push_back(lhs_end);
    } else {
      terms->emplace_back(span[static_cast<size_t>(k)], k);
    }
  }
  std::sort(terms->begin(), terms->end());
  return true;
}

int compile_program(const ab_op *prog, int nops, DevProg *out) {
  AB_REQUIRE(prog != nullptr && nops >= 1 && nops <= AB_MAX_OPS, "covariance program size");
  out->npoly = 0;
  bool has_poly = false;
  for (int k = 0; k < nops; ++k) {
    has_poly = has_poly || prog[k].op == AB_OP_POLYNOMIAL_TERM;
  }
  if (!has_poly) {
    return compile_stationary(prog, nops, out);
  }
  // polynomial terms must be summands of the top-level sum; everything else is compiled as before
  std::vector<std::pair<int, int>> terms;
  AB_REQUIRE(top_level_summands(prog, nops, &terms), "malformed postfix covariance program");
  std::vector<ab_op> rest;
  int nrest = 0, npoly = 0;
  for (const auto &t : terms) {
    if (t.first == t.second && prog[t.first].op == AB_OP_POLYNOMIAL_TERM) {
      const double deg = prog[t.first].p1;
      if (npoly >= AB_MAX_POLY || !(deg >= 0. && deg <= 16. && deg == std::floor(deg))) {
        set_error("polynomial term: at most %d terms of integer degree 0..16", AB_MAX_POLY);
        return AB_ERR_UNSUPPORTED;
      }
      out->poly_deg[npoly] = static_cast<int>(deg);
      out->poly_s2[npoly] = prog[t.first].p0 * prog[t.first].p0;
      ++npoly;
      continue;
    }
    for (int k = t.first; k <= t.second; ++k) {
      if (prog[k].op == AB_OP_POLYNOMIAL_TERM) {
        set_error("a Polynomial term inside a product has no device form (only as a summand of the covariance)");
        return AB_ERR_UNSUPPORTED;
      }
      rest.push_back(prog[k]);
    }
    if (nrest++ > 0) {
      rest.push_back(ab_op{AB_OP_SUM, 0, 0., 0.});
    }
  }
  if (rest.empty()) {
    rest.push_back(ab_op{AB_OP_CONSTANT, 0, 0., 0.}); // a purely polynomial covariance: 0 + terms
  }
  AB_TRY(compile_stationary(rest.data(), static_cast<int>(rest.size()), out));
  out->npoly = npoly;
  return AB_OK;
}

static int compile_stationary(const ab_op *prog, int nops, DevProg *out) {
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
    if ((prog[k].op == AB_OP_EXPONENTIAL || prog[k].op == AB_OP_MATERN32 ||
         prog[k].op == AB_OP_MATERN52) &&
        prog[k].p0 > 0.) {
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
      sop.ops[n].flags |= DF_TERM_START;
      ++n;
      ++k;
      while (k + 1 < nops && prog[k + 1].op == AB_OP_PRODUCT && leaf_to_dev(prog[k], &sop.ops[n])) {
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
      out->ops[k] = DevOp{DK_SUM, 0, 0., 0., 0., 0., 0.};
    } else if (prog[k].op == AB_OP_PRODUCT) {
      out->ops[k] = DevOp{DK_PROD, 0, 0., 0., 0., 0., 0.};
    } else {
      leaf_to_dev(prog[k], &out->ops[k]);
    }
  }
  return AB_OK;
}

// ------------------------------------------------------------------------------------------------
// device: lean fp64 exp / sqrt
// ------------------------------------------------------------------------------------------------




__device__ __noinline__ double eval_stack(const DevProg &P, double d2, double dist, bool equal) {
  double stack[8];
  int sp = 0;
  for (int k = 0; k < P.nops; ++k) {
    const DevOp &o = P.ops[k];
    if (o.kind == DK_SUM) {
      const double rhs = stack[--sp];
      stack[sp - 1] = stack[sp - 1] + rhs;
    } else if (o.kind == DK_PROD) {
      const double rhs = stack[--sp];
      const double lhs = stack[sp - 1];
      stack[sp - 1] = (lhs != 0.) ? lhs * rhs : lhs;
    } else if (o.kind == DK_RADIAL) {
      double v = exp((o.flags & DF_USES_DIST) ? o.a1 * dist : o.a2 * d2);
      if (o.flags & DF_POLY_D1) {
        double p = fma(o.b1, dist, 1.);
        if (o.flags & DF_POLY_D2) {
          p = fma(o.b2, d2, p);
        }
        v *= p;
      }
      stack[sp++] = o.amp * v;
    } else if (o.kind == DK_NOISE) {
      stack[sp++] = equal ? o.amp : 0.;
    } else {
      stack[sp++] = o.kind == DK_CONST ? o.amp : 0.;
    }
  }
  return stack[0];
}

template <int DIM>
__global__ void gram_diag_kernel(const __grid_constant__ DevProg P, const double *__restrict__ f,
                                 int64_t ldf, int64_t n, double *__restrict__ out) {
  __shared__ double tab[128];
  if (threadIdx.x < 128) {
    tab[threadIdx.x] = EXP_TABLE[threadIdx.x];
  }
  __syncthreads();
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
    eval_sop<1, MODE_SOP>(P, d2, dist, eq ? 1u : 0u, tab, v);
    out[i] = v[0];
  } else {
    out[i] = eval_stack(P, 0., 0., eq);
  }
}

template <bool SYM>
static int launch_gram_base(ab_handle_s *h, const DevProg &P, int dim, const double *fx, int64_t ldfx,
                            int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out,
                            int64_t ld, uint32_t flags) {
  if (n == 0 || m == 0) {
    return AB_OK;
  }
  const int64_t ti = (n + TILE - 1) / TILE;
  const int64_t tj = (m + TILE - 1) / TILE;
  const int64_t tiles = gram_items(SYM, ti, tj); // work items (tiles, plus skipped ones for SYM)
  AB_REQUIRE(tiles < (int64_t(1) << 31), "Gram too large for one launch");
  if (reinterpret_cast<uintptr_t>(out) % 16 != 0 || ld % 2 != 0) {
    flags |= GRAM_UNALIGNED;
  }
  // pick the leanest kernel specialisation the program allows
  int mode = MODE_STACK;
  if (P.mode == 0) {
    mode = P.need_equal ? MODE_SUM_NOISE : MODE_SUM;
    for (int k = 0; k < P.nops; ++k) {
      if ((P.ops[k].flags & (DF_TERM_START | DF_TERM_END)) != (DF_TERM_START | DF_TERM_END)) {
        mode = MODE_SOP;
      }
    }
  }
  const unsigned ntiles = static_cast<unsigned>(tiles);
  if (launch_gram_fixed(h, P, dim, SYM, fx, ldfx, n, fy, ldfy, m, out, ld, static_cast<int>(ti),
                        ntiles, flags)) {
    AB_LAUNCHED(h);
    return AB_OK;
  }
#define AB_GRAM_ARGS h, P, fx, ldfx, n, fy, ldfy, m, out, ld, static_cast<int>(ti), ntiles, flags
#define AB_GRAM_CASE(D)                                                                        \
  case D:                                                                                      \
    switch (mode) {                                                                            \
    case MODE_SUM:                                                                             \
      launch_err = gram_launch<D, SYM, EvalProgram<MODE_SUM>>(AB_GRAM_ARGS);                   \
      break;                                                                                   \
    case MODE_SUM_NOISE:                                                                       \
      launch_err = gram_launch<D, SYM, EvalProgram<MODE_SUM_NOISE>>(AB_GRAM_ARGS);             \
      break;                                                                                   \
    case MODE_SOP:                                                                             \
      launch_err = gram_launch<D, SYM, EvalProgram<MODE_SOP>>(AB_GRAM_ARGS);                   \
      break;                                                                                   \
    default:                                                                                   \
      launch_err = gram_launch<D, SYM, EvalProgram<MODE_STACK>>(AB_GRAM_ARGS);                 \
      break;                                                                                   \
    }                                                                                          \
    break;
  cudaError_t launch_err = cudaSuccess;
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
#undef AB_GRAM_ARGS
#undef AB_GRAM_CASE
  AB_CUDA(launch_err);
  h->launches++;
  return AB_OK;
}

// out(i, j) += sum_p s2_p x_i^p y_j^p (scalar features).  sym: i >= j only, mirrored unless lower_only.
__device__ __forceinline__ double poly_value(const DevProg &P, double x, double y) {
  double v = 0.;
  for (int t = 0; t < P.npoly; ++t) {
    double xp = 1., yp = 1.;
    for (int e = 0; e < P.poly_deg[t]; ++e) {
      xp *= x;
      yp *= y;
    }
    v += P.poly_s2[t] * xp * yp;
  }
  return v;
}

__global__ void poly_add_kernel(const __grid_constant__ DevProg P, const double *__restrict__ fx, int64_t ldfx,
                                int64_t n, const double *__restrict__ fy, int64_t ldfy, int64_t m, double *out,
                                int64_t ld, int sym, int lower_only) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t j = blockIdx.y + static_cast<int64_t>(blockIdx.z) * 65535;
  if (i >= n || j >= m || (sym && i < j)) {
    return;
  }
  const double v = poly_value(P, fx[i * ldfx], fy[j * ldfy]);
  out[i + j * ld] += v;
  if (sym && !lower_only && i != j) {
    out[j + i * ld] += v;
  }
}

__global__ void poly_add_diag_kernel(const __grid_constant__ DevProg P, const double *__restrict__ f, int64_t ldf,
                                     int64_t n, double *out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    out[i] += poly_value(P, f[i * ldf], f[i * ldf]);
  }
}

template <bool SYM>
static int launch_gram(ab_handle_s *h, const DevProg &P, int dim, const double *fx, int64_t ldfx,
                       int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out,
                       int64_t ld, uint32_t flags) {
  AB_TRY((launch_gram_base<SYM>(h, P, dim, fx, ldfx, n, fy, ldfy, m, out, ld, flags)));
  if (P.npoly > 0 && n > 0 && m > 0) {
    if (dim != 1) {
      set_error("Polynomial covariance terms are defined for scalar features (polynomials.hpp:79)");
      return AB_ERR_UNSUPPORTED;
    }
    const dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(m < 65535 ? m : 65535),
                    static_cast<unsigned>((m + 65534) / 65535));
    poly_add_kernel<<<grid, 256, 0, h->stream>>>(P, fx, ldfx, n, fy, ldfy, m, out, ld, SYM ? 1 : 0,
                                                 (flags & AB_GRAM_LOWER_ONLY) ? 1 : 0);
    AB_LAUNCHED(h);
  }
  return AB_OK;
}

int gram_into(ab_handle_s *h, const DevProg &P, int dim, bool sym, const double *fx, int64_t ldfx,
              int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out, int64_t ld,
              uint32_t flags) {
  if (sym) {
    return launch_gram<true>(h, P, dim, fx, ldfx, n, fx, ldfx, n, out, ld, flags);
  }
  return launch_gram<false>(h, P, dim, fx, ldfx, n, fy, ldfy, m, out, ld, 0u);
}

int gram_diag_device(ab_handle_s *h, const DevProg &P, const ab_matrix_s *feats, double *d_out);

int gram_diag_into(ab_handle_s *h, const DevProg &P, int dim, const double *f, int64_t ldf,
                   int64_t n, double *d_out) {
  ab_matrix_s view;
  view.d = const_cast<double *>(f);
  view.rows = dim;
  view.cols = n;
  view.ld = ldf;
  return gram_diag_device(h, P, &view, d_out);
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
  if (P.npoly > 0) {
    if (dim != 1) {
      set_error("Polynomial covariance terms are defined for scalar features (polynomials.hpp:79)");
      return AB_ERR_UNSUPPORTED;
    }
    poly_add_diag_kernel<<<grid, block, 0, h->stream>>>(P, feats->d, feats->ld, n, d_out);
    AB_LAUNCHED(h);
  }
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
