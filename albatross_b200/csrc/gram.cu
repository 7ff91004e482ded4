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
#include "gram.cuh"

#include <cmath>

namespace ab {

// ------------------------------------------------------------------------------------------------
// host: postfix program -> device program
// ------------------------------------------------------------------------------------------------

static bool leaf_to_dev(const ab_op &o, DevOp *d) {
  d->flags = 0;
  d->a2 = d->a1 = d->b2 = d->b1 = 0.;
  const double l = o.p0;
  switch (o.op) {
  case AB_OP_SQUARED_EXPONENTIAL:
    d->kind = l > 0. ? DK_RADIAL : DK_ZERO;
    d->a2 = -1. / (l * l);
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_EXPONENTIAL:
    d->kind = l > 0. ? DK_RADIAL : DK_ZERO;
    d->flags = DF_USES_DIST;
    d->a1 = -1. / l;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_MATERN32:
    d->kind = l > 0. ? DK_RADIAL : DK_ZERO;
    d->flags = DF_USES_DIST | DF_POLY_D1;
    d->a1 = -std::sqrt(3.) / l;
    d->b1 = std::sqrt(3.) / l;
    d->amp = o.p1 * o.p1;
    return true;
  case AB_OP_MATERN52:
    d->kind = l > 0. ? DK_RADIAL : DK_ZERO;
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

// 2^(j/128), j = 0..127, correctly rounded.
__device__ const double EXP_TABLE[128] = {
    0x1.0000000000000p+0,
    0x1.0163da9fb3335p+0,
    0x1.02c9a3e778061p+0,
    0x1.04315e86e7f85p+0,
    0x1.059b0d3158574p+0,
    0x1.0706b29ddf6dep+0,
    0x1.0874518759bc8p+0,
    0x1.09e3ecac6f383p+0,
    0x1.0b5586cf9890fp+0,
    0x1.0cc922b7247f7p+0,
    0x1.0e3ec32d3d1a2p+0,
    0x1.0fb66affed31bp+0,
    0x1.11301d0125b51p+0,
    0x1.12abdc06c31ccp+0,
    0x1.1429aaea92de0p+0,
    0x1.15a98c8a58e51p+0,
    0x1.172b83c7d517bp+0,
    0x1.18af9388c8deap+0,
    0x1.1a35beb6fcb75p+0,
    0x1.1bbe084045cd4p+0,
    0x1.1d4873168b9aap+0,
    0x1.1ed5022fcd91dp+0,
    0x1.2063b88628cd6p+0,
    0x1.21f49917ddc96p+0,
    0x1.2387a6e756238p+0,
    0x1.251ce4fb2a63fp+0,
    0x1.26b4565e27cddp+0,
    0x1.284dfe1f56381p+0,
    0x1.29e9df51fdee1p+0,
    0x1.2b87fd0dad990p+0,
    0x1.2d285a6e4030bp+0,
    0x1.2ecafa93e2f56p+0,
    0x1.306fe0a31b715p+0,
    0x1.32170fc4cd831p+0,
    0x1.33c08b26416ffp+0,
    0x1.356c55f929ff1p+0,
    0x1.371a7373aa9cbp+0,
    0x1.38cae6d05d866p+0,
    0x1.3a7db34e59ff7p+0,
    0x1.3c32dc313a8e5p+0,
    0x1.3dea64c123422p+0,
    0x1.3fa4504ac801cp+0,
    0x1.4160a21f72e2ap+0,
    0x1.431f5d950a897p+0,
    0x1.44e086061892dp+0,
    0x1.46a41ed1d0057p+0,
    0x1.486a2b5c13cd0p+0,
    0x1.4a32af0d7d3dep+0,
    0x1.4bfdad5362a27p+0,
    0x1.4dcb299fddd0dp+0,
    0x1.4f9b2769d2ca7p+0,
    0x1.516daa2cf6642p+0,
    0x1.5342b569d4f82p+0,
    0x1.551a4ca5d920fp+0,
    0x1.56f4736b527dap+0,
    0x1.58d12d497c7fdp+0,
    0x1.5ab07dd485429p+0,
    0x1.5c9268a5946b7p+0,
    0x1.5e76f15ad2148p+0,
    0x1.605e1b976dc09p+0,
    0x1.6247eb03a5585p+0,
    0x1.6434634ccc320p+0,
    0x1.6623882552225p+0,
    0x1.68155d44ca973p+0,
    0x1.6a09e667f3bcdp+0,
    0x1.6c012750bdabfp+0,
    0x1.6dfb23c651a2fp+0,
    0x1.6ff7df9519484p+0,
    0x1.71f75e8ec5f74p+0,
    0x1.73f9a48a58174p+0,
    0x1.75feb564267c9p+0,
    0x1.780694fde5d3fp+0,
    0x1.7a11473eb0187p+0,
    0x1.7c1ed0130c132p+0,
    0x1.7e2f336cf4e62p+0,
    0x1.80427543e1a12p+0,
    0x1.82589994cce13p+0,
    0x1.8471a4623c7adp+0,
    0x1.868d99b4492edp+0,
    0x1.88ac7d98a6699p+0,
    0x1.8ace5422aa0dbp+0,
    0x1.8cf3216b5448cp+0,
    0x1.8f1ae99157736p+0,
    0x1.9145b0b91ffc6p+0,
    0x1.93737b0cdc5e5p+0,
    0x1.95a44cbc8520fp+0,
    0x1.97d829fde4e50p+0,
    0x1.9a0f170ca07bap+0,
    0x1.9c49182a3f090p+0,
    0x1.9e86319e32323p+0,
    0x1.a0c667b5de565p+0,
    0x1.a309bec4a2d33p+0,
    0x1.a5503b23e255dp+0,
    0x1.a799e1330b358p+0,
    0x1.a9e6b5579fdbfp+0,
    0x1.ac36bbfd3f37ap+0,
    0x1.ae89f995ad3adp+0,
    0x1.b0e07298db666p+0,
    0x1.b33a2b84f15fbp+0,
    0x1.b59728de5593ap+0,
    0x1.b7f76f2fb5e47p+0,
    0x1.ba5b030a1064ap+0,
    0x1.bcc1e904bc1d2p+0,
    0x1.bf2c25bd71e09p+0,
    0x1.c199bdd85529cp+0,
    0x1.c40ab5fffd07ap+0,
    0x1.c67f12e57d14bp+0,
    0x1.c8f6d9406e7b5p+0,
    0x1.cb720dcef9069p+0,
    0x1.cdf0b555dc3fap+0,
    0x1.d072d4a07897cp+0,
    0x1.d2f87080d89f2p+0,
    0x1.d5818dcfba487p+0,
    0x1.d80e316c98398p+0,
    0x1.da9e603db3285p+0,
    0x1.dd321f301b460p+0,
    0x1.dfc97337b9b5fp+0,
    0x1.e264614f5a129p+0,
    0x1.e502ee78b3ff6p+0,
    0x1.e7a51fbc74c83p+0,
    0x1.ea4afa2a490dap+0,
    0x1.ecf482d8e67f1p+0,
    0x1.efa1bee615a27p+0,
    0x1.f252b376bba97p+0,
    0x1.f50765b6e4540p+0,
    0x1.f7bfdad9cbe14p+0,
    0x1.fa7c1819e90d8p+0,
    0x1.fd3c22b8f71f1p+0};


// exp(x) for x <= 0 (the argument of every radial kernel).  x = 128 n ln2/128 + j ln2/128 + r,
// exp(x) = 2^n * T[j] * (1 + r + ... + r^5/120), |r| <= ln2/256.  10 FP64-pipe instructions; the table
// lookup, the index arithmetic and the 2^n scaling run on the LSU / integer pipes.  Valid for
// -708 <= x <= -0; `bad` accumulates (sign bit set) when x is outside that range so that the caller
// can patch the rare cases: x < -708 is flushed to 0 (the reference would return a subnormal
// < 3e-308 there), NaN stays NaN.
__device__ __forceinline__ double exp_nonpos(double x, const double *__restrict__ tab, int &bad) {
  const double t = fma(x, 184.6649652337873, 6755399441055744.0); // x * 128/ln2, round to nearest
  const int m = __double2loint(t);
  const double mf = t - 6755399441055744.0;
  double r = fma(mf, -0x1.62e42fef00000p-8, x);  // ln2/128 high part (32 significant bits)
  r = fma(mf, -0x1.473de6af278edp-41, r);        // ln2/128 low part
  double p = fma(r, 0.008333333333333333, 0.041666666666666664);
  p = fma(p, r, 0.16666666666666666);
  p = fma(p, r, 0.5);
  const double r2 = r * r;
  const double q = fma(p, r2, r); // expm1(r)
  const double tj = tab[m & 127];
  const double res = fma(tj, q, tj);
  // in range  <=>  hi(x) in [0x80000000, 0xC0862000]  (-0 .. -708)
  bad |= 0x40862000 - (__double2hiint(x) ^ 0x80000000);
  return __hiloint2double(__double2hiint(res) + ((m >> 7) << 20), __double2loint(res));
}

// sqrt(a) for positive normal a: MUFU.RSQ64H seed (2^-22) + one Goldschmidt step + one residual
// correction (7 FP64-pipe instructions, <= 1 ulp).  `bad` gets its sign bit set for zero,
// subnormal, infinite, NaN or negative arguments, which the caller patches inline.
__device__ __forceinline__ double sqrt_fast(double a, int &bad) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y;
  double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double e = fma(-g, g, a);
  // positive normal  <=>  hi(a) in [0x00100000, 0x7fefffff]
  const int hi = __double2hiint(a);
  bad |= (hi - 0x00100000) | (0x7fefffff - hi);
  return fma(e, h, g);
}

// ------------------------------------------------------------------------------------------------
// device: covariance program evaluation
// ------------------------------------------------------------------------------------------------

// Kernel specialisations (template parameter MODE): which parts of the generic evaluator exist.
//   0  sum of single radial/constant leaves (e.g. SE + Matern52)            - no equality, no products
//   1  ... plus IndependentNoise leaves (e.g. SE + noise)                   - feature equality needed
//   2  general sum of products                                              - `prod` accumulator
//   3  arbitrary nesting: postfix evaluation with a stack (slow path)
constexpr int MODE_SUM = 0, MODE_SUM_NOISE = 1, MODE_SOP = 2, MODE_STACK = 3;

// Sum-of-products evaluation of NP pairs at once: the op loop is uniform across the CTA and its
// decode cost is amortised over the NP pairs a thread owns.  Single-leaf terms (the common case)
// accumulate straight into `out` with one FMA; `prod` only exists for genuine products.
template <int NP, int MODE>
__device__ __forceinline__ void eval_sop(const DevProg &P, const double (&d2)[NP],
                                         const double (&dist)[NP], unsigned eqmask,
                                         const double *__restrict__ tab, double (&out)[NP]) {
  double prod[MODE >= MODE_SOP ? NP : 1];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    out[i] = 0.;
  }
  for (int k = 0; k < P.nops; ++k) {
    const int kind = P.ops[k].kind;
    const int flags = P.ops[k].flags;
    const double amp = P.ops[k].amp;
    double v[NP];
    if (kind == DK_RADIAL) {
      if (flags & DF_USES_DIST) {
        const double a1 = P.ops[k].a1;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          v[i] = a1 * dist[i];
        }
      } else {
        const double a2 = P.ops[k].a2;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          v[i] = a2 * d2[i];
        }
      }
      int bad = 0;
      double e[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        e[i] = exp_nonpos(v[i], tab, bad);
      }
      if (bad < 0) { // rare: some argument outside [-708, -0]: underflow -> 0, NaN -> NaN
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          e[i] = (v[i] < -708.0) ? 0. : ((v[i] != v[i]) ? v[i] : e[i]);
        }
      }
      if (flags & DF_POLY_D1) {
        const double b1 = P.ops[k].b1;
        if (flags & DF_POLY_D2) {
          const double b2 = P.ops[k].b2;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            v[i] = e[i] * fma(b2, d2[i], fma(b1, dist[i], 1.));
          }
        } else {
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            v[i] = e[i] * fma(b1, dist[i], 1.);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          v[i] = e[i];
        }
      }
    } else if (MODE >= MODE_SUM_NOISE && kind == DK_NOISE) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[i] = ((eqmask >> i) & 1u) ? 1. : 0.;
      }
    } else {
      const double c = kind == DK_CONST ? 1. : 0.;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[i] = c;
      }
    }
    if (MODE < MODE_SOP ||
        (flags & (DF_TERM_START | DF_TERM_END)) == (DF_TERM_START | DF_TERM_END)) {
      // single-leaf term: out += amp * v  (out starts at +0, so the first term is exact)
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        out[i] = fma(amp, v[i], out[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[i] *= amp;
      }
      if (!(flags & DF_TERM_START)) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const double pr = prod[MODE >= MODE_SOP ? i : 0];
          v[i] = (pr != 0.) ? pr * v[i] : pr; // covariance_function.hpp:362-366
        }
      }
      if (flags & DF_TERM_END) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          out[i] += v[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          prod[MODE >= MODE_SOP ? i : 0] = v[i];
        }
      }
    }
  }
}

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

constexpr int TILE = 64;
constexpr int LDT = TILE + 1;
constexpr int GRAM_THREADS = 256;
constexpr int NPAIR = 8; // 2 rows x 4 columns per thread and pass; 2 passes cover the 64 columns

// SYM: blockIdx.x enumerates tiles (I >= J) of the lower triangle; otherwise I = b % tiles_i.
template <int DIM, bool SYM, int MODE>
__global__ void __launch_bounds__(GRAM_THREADS, 2)
gram_kernel(const __grid_constant__ DevProg P, const double *__restrict__ fx, int64_t ldfx,
            int64_t n, const double *__restrict__ fy, int64_t ldfy, int64_t m,
            double *__restrict__ out, int64_t ld, int tiles_i, uint32_t flags) {
  __shared__ double tab[128];
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

  if (tid < 128) {
    tab[tid] = EXP_TABLE[tid];
  }
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
  const bool mirror = SYM && I != J && !(flags & AB_GRAM_LOWER_ONLY);
  const bool unaligned = flags & GRAM_UNALIGNED; // output sub-view that is only 8-byte aligned
  const int64_t gi = i0 + r0;

#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int cbase = pass * 32 + warp * 4;
    double d2[NPAIR], dist[NPAIR], vals[NPAIR];
    unsigned eqmask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = cbase + k;
      double yj[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        yj[d] = ys[c * DIM + d];
      }
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double s = 0.;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          const double diff = xi[a][d] - yj[d];
          s = fma(diff, diff, s);
        }
        d2[2 * k + a] = s;
        dist[2 * k + a] = DIM == 1 ? fabs(xi[a][0] - yj[0]) : 0.;
        if (MODE != MODE_SUM) {
          bool eq = true;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            eq = eq && (xi[a][d] == yj[d]);
          }
          eqmask |= (eq ? 1u : 0u) << (2 * k + a);
        }
      }
    }
    if (DIM != 1 && P.need_dist) {
      int bad = 0;
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) {
        dist[i] = sqrt_fast(d2[i], bad);
      }
      if (bad < 0) { // rare: a zero / subnormal / non-finite squared distance (e.g. the diagonal)
#pragma unroll
        for (int i = 0; i < NPAIR; ++i) {
          const double a = d2[i];
          if (!(a >= 2.2250738585072014e-308 && a < INFINITY)) {
            int ignored = 0;
            // 0, inf, NaN map to themselves; subnormals are rescaled by 2^108 first
            dist[i] = (a > 0. && a < INFINITY)
                          ? sqrt_fast(a * 3.2451855365842673e32, ignored) * 5.551115123125783e-17
                          : a;
          }
        }
      }
    }

    if (MODE != MODE_STACK) {
      eval_sop<NPAIR, MODE>(P, d2, dist, eqmask, tab, vals);
    } else {
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) {
        vals[i] = eval_stack(P, d2[i], dist[i], (eqmask >> i) & 1u);
      }
    }

    // direct tile: rows i0 + r0 (+1), columns j0 + c; 16-byte stores, 512 contiguous bytes per warp.
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t gj = j0 + cbase + k;
      if (gj < m) {
        double *dst = out + gi + gj * ld;
        if (gi + 1 < n && !unaligned) {
          *reinterpret_cast<double2 *>(dst) = make_double2(vals[2 * k], vals[2 * k + 1]);
        } else if (gi < n) {
          dst[0] = vals[2 * k];
          if (gi + 1 < n) {
            dst[1] = vals[2 * k + 1];
          }
        }
      }
    }
    if (mirror) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = cbase + k;
        stage[c * LDT + r0] = vals[2 * k];
        stage[c * LDT + r0 + 1] = vals[2 * k + 1];
      }
    }
  }

  if (mirror) {
    // transposed tile through shared memory: element (row = j0 + c, col = i0 + r) = stage[c][r];
    // each warp-store covers 32 consecutive rows (256 contiguous bytes).
    __syncthreads();
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
#define AB_GRAM_CASE(D)                                                                        \
  case D:                                                                                      \
    switch (mode) {                                                                            \
    case MODE_SUM:                                                                             \
      gram_kernel<D, SYM, MODE_SUM><<<grid, block, 0, h->stream>>>(                            \
          P, fx, ldfx, n, fy, ldfy, m, out, ld, static_cast<int>(ti), flags);                  \
      break;                                                                                   \
    case MODE_SUM_NOISE:                                                                       \
      gram_kernel<D, SYM, MODE_SUM_NOISE><<<grid, block, 0, h->stream>>>(                      \
          P, fx, ldfx, n, fy, ldfy, m, out, ld, static_cast<int>(ti), flags);                  \
      break;                                                                                   \
    case MODE_SOP:                                                                             \
      gram_kernel<D, SYM, MODE_SOP><<<grid, block, 0, h->stream>>>(                            \
          P, fx, ldfx, n, fy, ldfy, m, out, ld, static_cast<int>(ti), flags);                  \
      break;                                                                                   \
    default:                                                                                   \
      gram_kernel<D, SYM, MODE_STACK><<<grid, block, 0, h->stream>>>(                          \
          P, fx, ldfx, n, fy, ldfy, m, out, ld, static_cast<int>(ti), flags);                  \
      break;                                                                                   \
    }                                                                                          \
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
