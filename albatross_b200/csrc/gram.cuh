// Device form of a covariance program and the per-pair evaluator shared by the Gram kernels.
#pragma once

#include "common.cuh"

namespace ab {

// Leaf kinds on the device.  The four radial kernels of radial.hpp share one code path:
//     k = amp * poly * exp(u),   u = a2 * d^2  or  a1 * d   (u <= 0),
//     poly = 1 [+ b1 * d [+ b2 * d^2]]
//   SquaredExponential  a2 = -1/l^2
//   Exponential         a1 = -1/l
//   Matern32            a1 = -sqrt(3)/l, b1 = sqrt(3)/l
//   Matern52            a1 = -sqrt(5)/l, b1 = sqrt(5)/l, b2 = 5/(3 l^2)
enum DevKind : int {
  DK_RADIAL = 1,
  DK_CONST = 2, // amp
  DK_NOISE = 3, // amp if x == y (all coordinates) else 0
  DK_ZERO = 4,  // radial leaf with length_scale <= 0 (radial.hpp:28-30): contributes 0
  DK_SUM = 7,   // stack mode only
  DK_PROD = 8   // stack mode only
};

enum : int {
  DF_TERM_START = 1,
  DF_TERM_END = 2,
  DF_FIRST_TERM = 4,
  DF_USES_DIST = 8, // u = a1 * d (else u = a2 * d^2)
  DF_POLY_D1 = 16,  // poly has the b1 * d term
  DF_POLY_D2 = 32   // poly has the b2 * d^2 term
};

struct DevOp {
  int kind;
  int flags;
  double a2, a1, b2, b1, amp;
  double ab1, ab2; // amp * b1, amp * b2 (amplitude folded into the Matern polynomial)
  // scaled-domain exp: a2 * 2048/ln2, a1 * 2048/ln2, and the high word of the largest argument
  // (d^2 resp. d) for which a * arg >= -708, rounded down (conservative)
  double a2s, a1s;
  int lim_hi;
  int pad_;
};

// mode 0: "sum of products" — expr := term (+ term)*, term := leaf (* leaf)*, evaluated left to
//         right with two accumulators exactly as the reference's left-associated operator+/operator*
//         tree would; no stack.
// mode 1: generic postfix with an evaluation stack (any nesting).
constexpr int AB_MAX_POLY = 8;

struct DevProg {
  int nops;
  int mode;
  int need_dist;  // some leaf needs d = sqrt(d^2)
  int need_equal; // some leaf needs feature equality
  DevOp ops[AB_MAX_OPS];
  // Polynomial terms sigma_p^2 x^p y^p (polynomials.hpp:63-90; scalar features) that stand as top-level
  // summands of the covariance: not distance based, so they are added to the finished block by a second
  // elementwise pass instead of going through the pairwise evaluator.
  int npoly;
  int poly_deg[AB_MAX_POLY];
  double poly_s2[AB_MAX_POLY];
};

int compile_program(const ab_op *prog, int nops, DevProg *out);

constexpr uint32_t GRAM_UNALIGNED = 0x100u; // internal flag: scalar stores (set by the launcher)

// Gram blocks written into an existing device buffer (possibly a sub-view of a larger matrix).
// Features: point i at f[i * ldf .. i * ldf + dim).  sym: lower triangle (+ mirror unless
// AB_GRAM_LOWER_ONLY) of k(fx, fx); otherwise the n x m cross block k(fx, fy).
int gram_into(ab_handle_s *h, const DevProg &P, int dim, bool sym, const double *fx, int64_t ldfx,
              int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out, int64_t ld,
              uint32_t flags);
int gram_diag_into(ab_handle_s *h, const DevProg &P, int dim, const double *f, int64_t ldf,
                   int64_t n, double *d_out);

} // namespace ab
