// Compile-time-specialised Gram kernels: sums of up to three leaves whose kinds are template
// parameters (EvalFixed, gram_kernel.cuh), for feature dimensions 1..3.  Everything else — products,
// nesting, more than three terms, radial leaves with a non-positive length scale, dimensions 4..8 —
// runs through the generic program evaluator in gram.cu.  Kept in its own translation unit so that
// the two sets of instantiations compile in parallel.
#include "gram_kernel.cuh"

// overridable for tools/sweep.sh
#ifndef AB_GRAM_COLS
#define AB_GRAM_COLS 2
#endif
#ifndef AB_GRAM_MINB
#define AB_GRAM_MINB 2
#endif

namespace ab {
namespace {

// Leaf kind of a compiled single-leaf term, or -1 if it has no fixed form.
int leaf_sig(const DevOp &o) {
  if ((o.flags & (DF_TERM_START | DF_TERM_END)) != (DF_TERM_START | DF_TERM_END)) {
    return -1; // part of a product
  }
  switch (o.kind) {
  case DK_RADIAL:
    if (!(o.flags & DF_USES_DIST)) {
      return LS_SE;
    }
    if (o.flags & DF_POLY_D2) {
      return LS_M52;
    }
    return (o.flags & DF_POLY_D1) ? LS_M32 : LS_EXP;
  case DK_CONST:
    return LS_CONST;
  case DK_NOISE:
    return LS_NOISE;
  default:
    return -1; // DK_ZERO
  }
}

template <int DIM, bool SYM, int K0, int K1, int K2>
void launch_one(ab_handle_s *h, const DevProg &P, const double *fx, int64_t ldfx, int64_t n,
                const double *fy, int64_t ldfy, int64_t m, double *out, int64_t ld, int tiles_i,
                unsigned tiles, uint32_t flags) {
  // 2 columns per pass (4 pairs in flight per thread) measured faster than 4 for the fixed
  // evaluators: 1.958 ms vs 2.074 ms at N = 32 768, SE + Matern52
  gram_launch<DIM, SYM, EvalFixed<K0, K1, K2>, AB_GRAM_COLS, AB_GRAM_MINB>(
      h, P, fx, ldfx, n, fy, ldfy, m, out, ld, tiles_i, tiles, flags);
}

template <int K0, int K1, int K2>
bool launch_sig(ab_handle_s *h, const DevProg &P, int dim, bool sym, const double *fx, int64_t ldfx,
                int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out, int64_t ld,
                int tiles_i, unsigned tiles, uint32_t flags) {
#define AB_FIXED_ARGS h, P, fx, ldfx, n, fy, ldfy, m, out, ld, tiles_i, tiles, flags
  switch (dim) {
  case 1:
    sym ? launch_one<1, true, K0, K1, K2>(AB_FIXED_ARGS) : launch_one<1, false, K0, K1, K2>(AB_FIXED_ARGS);
    return true;
  case 2:
    sym ? launch_one<2, true, K0, K1, K2>(AB_FIXED_ARGS) : launch_one<2, false, K0, K1, K2>(AB_FIXED_ARGS);
    return true;
  case 3:
    sym ? launch_one<3, true, K0, K1, K2>(AB_FIXED_ARGS) : launch_one<3, false, K0, K1, K2>(AB_FIXED_ARGS);
    return true;
  default:
    return false;
  }
#undef AB_FIXED_ARGS
}

} // namespace

bool launch_gram_fixed(ab_handle_s *h, const DevProg &P, int dim, bool sym, const double *fx,
                       int64_t ldfx, int64_t n, const double *fy, int64_t ldfy, int64_t m,
                       double *out, int64_t ld, int tiles_i, unsigned tiles, uint32_t flags) {
  if (P.mode != 0 || P.nops < 1 || P.nops > 3 || dim > 3) {
    return false;
  }
  int k[3] = {LS_NONE, LS_NONE, LS_NONE};
  for (int i = 0; i < P.nops; ++i) {
    k[i] = leaf_sig(P.ops[i]);
    if (k[i] < 0) {
      return false;
    }
  }
  const int sig = k[0] | (k[1] << 4) | (k[2] << 8);
#define AB_SIG(a, b, c)                                                                         \
  case ((a) | ((b) << 4) | ((c) << 8)):                                                         \
    return launch_sig<a, b, c>(h, P, dim, sym, fx, ldfx, n, fy, ldfy, m, out, ld, tiles_i,     \
                               tiles, flags);
  switch (sig) {
    AB_SIG(LS_SE, LS_NONE, LS_NONE)
    AB_SIG(LS_EXP, LS_NONE, LS_NONE)
    AB_SIG(LS_M32, LS_NONE, LS_NONE)
    AB_SIG(LS_M52, LS_NONE, LS_NONE)
    AB_SIG(LS_SE, LS_NOISE, LS_NONE)
    AB_SIG(LS_EXP, LS_NOISE, LS_NONE)
    AB_SIG(LS_M32, LS_NOISE, LS_NONE)
    AB_SIG(LS_M52, LS_NOISE, LS_NONE)
    AB_SIG(LS_SE, LS_M52, LS_NONE)
    AB_SIG(LS_SE, LS_M52, LS_NOISE)
  default:
    return false;
  }
#undef AB_SIG
}

} // namespace ab
