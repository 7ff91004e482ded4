// Device form of a covariance program and the per-pair evaluator shared by the Gram kernels.
#pragma once

#include "common.cuh"

namespace ab {

// Leaf kinds on the device (host pre-digests hyper-parameters into c0 / amp).
enum DevKind : int {
  DK_SE = 1,    // amp * exp(c0 * d^2),            c0 = -1/l^2
  DK_EXP = 2,   // amp * exp(c0 * d),              c0 = -1/l
  DK_M32 = 3,   // amp * (1 + s) exp(-s),          s = c0 * d, c0 = sqrt(3)/l
  DK_M52 = 4,   // amp * (1 + s + s^2/3) exp(-s),  s = c0 * d, c0 = sqrt(5)/l
  DK_CONST = 5, // amp
  DK_NOISE = 6, // amp if x == y (all coordinates) else 0
  DK_SUM = 7,   // stack mode only
  DK_PROD = 8,  // stack mode only
  DK_ZERO = 9   // radial leaf with length_scale <= 0 (radial.hpp:28-30): contributes 0
};

// flags for the sum-of-products form
enum : int { DF_TERM_START = 1, DF_TERM_END = 2, DF_FIRST_TERM = 4 };

struct DevOp {
  int kind;
  int flags;
  double c0;
  double amp;
};

// mode 0: "sum of products" — expr := term (+ term)*, term := leaf (* leaf)*, evaluated left to
//         right with two accumulators exactly as the reference's left-associated operator+/operator*
//         tree would; no stack.
// mode 1: generic postfix with an evaluation stack (any nesting).
struct DevProg {
  int nops;
  int mode;
  int need_dist;  // some leaf needs d = sqrt(d^2)
  int need_equal; // some leaf needs feature equality
  DevOp ops[AB_MAX_OPS];
};

int compile_program(const ab_op *prog, int nops, DevProg *out);

} // namespace ab
