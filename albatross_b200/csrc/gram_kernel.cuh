// The Gram tile kernel (version 2) and its two evaluator families.
//
// One CTA (256 threads, 2 CTAs / SM) computes one 64 x 64 tile of K_ij = k(x_i, y_j); a thread owns
// 2 rows x 4 columns per pass (8 pairs in flight, 2 passes).  For the symmetric build only tiles on
// or below the diagonal are evaluated and the transposed tile is written from a shared-memory
// staging buffer, so that both orientations leave the SM as full 256/512-byte row segments.
//
// The round-1 profile (profiles/r01_ncu_gram_gemm_summary.md) showed the first version issue-bound:
// 122 warp instructions per pair of which only 40 are FP64 arithmetic.  This version removes the
// rest as far as the algorithm allows:
//   * interior tiles take a path without any bounds checks and with pointer-increment addressing;
//   * the exp / sqrt argument range checks are one signed min/max per 8 values (VIMNMX3) instead
//     of three integer instructions per value; out-of-range batches take a patch-up branch;
//   * the 2^n scaling of exp is LOP3 + IMAD, the table index LOP3 + LEA;
//   * the triangular tile decode is 32-bit (fp32 sqrt + integer fix-up);
//   * EvalFixed<K0, K1, K2> evaluates a sum of up to three leaves whose kinds are template
//     parameters (the reference composes covariance functions at compile time too,
//     covariance_functions/covariance_function.hpp:222-420): no op loop, no decode branches.
//     EvalProgram<MODE> keeps the generic run-time program (sum of products / postfix stack).
#pragma once

#include "exp_table.cuh"
#include "gram.cuh"

namespace ab {

constexpr int TILE = 64;
// Mirror staging buffer, element (c, r) of the tile (column c, row r): stage[c * 65 + r], 8-byte
// accesses (producer 2-way conflicted).  An XOR-swizzled layout with 16-byte accesses measured slower
// (2.00 vs 1.95 ms, profiles/r01d_sweeps.txt) and was removed in round 2.
constexpr int LDT = TILE + 1;
constexpr int GRAM_THREADS = 256;
// A thread owns 2 rows x COLS columns per pass (2 * COLS pairs in flight); 8 / COLS passes cover the
// 64 columns of the tile.

// ------------------------------------------------------------------------------------------------
// lean fp64 exp / sqrt
// ------------------------------------------------------------------------------------------------

// What the fixed-evaluator kernels ship with (the kernel is issue-bound, not HBM-bound: a warp-wide DFMA
// holds the FP64 pipe 2 cycles, 3 with three distinct sources; IMAD / LOP3 / SHF cost 1-2 cycles each
// next to FP64 work, tools/issue_probe.cu).  Every choice was timed at N = 32 768, SE + Matern52, full
// symmetric, after passing tests/test_gpu_gram.py on the same box (profiles/r02a_gram_sweep.txt):
//   * one range check per pass (AB_GRAM_ONECHECK, default on; =0 keeps the round-1 check-and-patch
//     branches for comparison): a zero / subnormal / non-finite squared distance or an out-of-range exp
//     argument triggers a cold re-evaluation of the pass with the checked evaluator;
//   * exp by range reduction in table steps straight from the distance: with as = a * 2048/ln2 (host
//     side), t = fma(d, as, magic), r' = fma(d, as, -m) (one rounding, |r'| <= 1/2),
//     exp = 2^(m/2048) (1 + r' (c1 + r' (c2 + r' c3))): 7 FP64 instructions per exp instead of 9.  Same
//     error bound against the exact exp(a d) as a two-step reduction, 1.57 (1 + |x|) 2^-53 (CPU emulation,
//     tools/exp_emulation.c), up to 2 |x| 2^-53 away from libm's exp(fl(a d));
//   * 2^n scaling as shift + multiply-add; amplitude folded into the Matern polynomial coefficients;
//   * interior-tile store addresses as 64-bit tile base + 32-bit element offset.
// Together 1.956 -> 1.821 ms (4392 -> 4717 GB/s, 0.670 -> 0.720 of the measured HBM peak).  Measured and
// removed in round 2: exp constants from constant memory (2.000 ms), expm1 as r * fma(r, p, 1) (1.954),
// 2x2 / 4x4 micro-block tile order (1.997 / 2.122), 4 columns per pass at 1 CTA/SM (2.296), loop-carried
// store pointers (2.17, round 1), XOR-swizzled staging (2.00, round 1).
#ifndef AB_GRAM_ONECHECK
#define AB_GRAM_ONECHECK 1
#endif

// res * 2^(m >> SHIFT) for a normal result (no overflow: the argument range is checked)
template <int SHIFT> __device__ __forceinline__ double exp_scale(double res, int m) {
  int hi;
  asm("mad.lo.s32 %0, %1, 1048576, %2;" : "=r"(hi) : "r"(m >> SHIFT), "r"(__double2hiint(res)));
  return __hiloint2double(hi, __double2loint(res));
}

// exp(x) for -708 <= x <= -0 (the argument of every radial kernel): x = (128 n + j) ln2/128 + r,
// exp(x) = 2^n T[j] (1 + r + ... + r^5/120), |r| <= ln2/256; 10 FP64-pipe instructions, <= 1 ulp.
// hi_max accumulates the signed maximum of the arguments' high words: the batch is in range iff
// hi_max <= (int)0xC0862000 (-708.0); +0, positive values, x < -708, +-inf and NaN all compare
// greater and are patched by the caller.
__device__ __forceinline__ double exp_core(double x, const double *__restrict__ tab, int &hi_max) {
  const double t = fma(x, 184.6649652337873, 6755399441055744.0); // x * 128/ln2, round to nearest
  const int m = __double2loint(t);
  const double mf = t - 6755399441055744.0;
  double r = fma(mf, -0x1.62e42fef00000p-8, x); // ln2/128 high part (32 significant bits)
  r = fma(mf, -0x1.473de6af278edp-41, r);       // ln2/128 low part
  double p = fma(r, 0.008333333333333333, 0.041666666666666664);
  p = fma(p, r, 0.16666666666666666);
  p = fma(p, r, 0.5);
  const double r2 = r * r;
  const double q = fma(p, r2, r); // expm1(r)
  const double tj = tab[m & 127];
  const double res = fma(tj, q, tj);
  hi_max = max(hi_max, __double2hiint(x));
  // + n * 2^20 on the high word, n = m >> 7 (arithmetic): (m & ~127) * 2^13
  return exp_scale<7>(res, m);
}

// Same with the 2048-entry table: |r| <= ln2/4096, degree-3 polynomial, 8 FP64-pipe instructions,
// <= 1.3 ulp (tools/make_exp_table.py generates both tables).
__device__ __forceinline__ double exp_core_big(double x, const double *__restrict__ tab,
                                               int &hi_max) {
  const double t = fma(x, 2954.639443740597, 6755399441055744.0); // x * 2048/ln2
  const int m = __double2loint(t);
  const double mf = t - 6755399441055744.0;
  double r = fma(mf, -0x1.62e42fef00000p-12, x);
  r = fma(mf, -0x1.473de6af278edp-45, r);
  const double p = fma(r, 0.16666666666666666, 0.5);
  const double r2 = r * r;
  const double q = fma(p, r2, r); // expm1(r)
  const double tj = tab[m & 2047];
  const double res = fma(tj, q, tj);
  hi_max = max(hi_max, __double2hiint(x));
  return exp_scale<11>(res, m);
}

// Same with a 256-entry table held in shared memory in 16 interleaved copies, copy (lane & 15) at
// tab[16 j + (lane & 15)]: the 16 lanes of a half-warp always hit 16 different 8-byte banks, so the
// data-dependent lookup costs the minimum two wavefronts.  (The plain 2048-entry table costs ~6:
// ncu counted 1.6e8 bank-conflict cycles at N = 32 768, 28 % of all SM cycles, with L1/shared the
// busiest unit of the kernel at 75 %.)  |r| <= ln2/512, degree-4 polynomial, 9 FP64-pipe
// instructions, relative error <= 2.2 * 2^-53.  `tab_lane` = table base + (lane & 15).
constexpr int EXP_REPL = 16;
__device__ __forceinline__ double exp_core_r256(double x, const double *__restrict__ tab_lane,
                                                int &hi_max) {
  const double t = fma(x, 369.3299304675746, 6755399441055744.0); // x * 256/ln2
  const int m = __double2loint(t);
  const double mf = t - 6755399441055744.0;
  double r = fma(mf, -0x1.62e42fef00000p-9, x);
  r = fma(mf, -0x1.473de6af278edp-42, r);
  double p = fma(r, 0.041666666666666664, 0.16666666666666666);
  p = fma(p, r, 0.5);
  const double r2 = r * r;
  const double q = fma(p, r2, r); // expm1(r)
  const double tj = tab_lane[(m & 255) * EXP_REPL];
  const double res = fma(tj, q, tj);
  hi_max = max(hi_max, __double2hiint(x));
  return exp_scale<8>(res, m);
}

constexpr int EXP_HI_LIMIT = static_cast<int>(0xC0862000u);


__device__ __forceinline__ double exp_patch(double x, double fast) {
  // x == +0 or in range: the fast value is right; x < -708: the reference's exp() underflows to a
  // subnormal < 3e-308 or 0, flushed to 0 here; NaN propagates.  x > 0 cannot occur: every radial
  // argument is (negative coefficient) * (distance >= 0).
  return (x < -708.0) ? 0. : ((x != x) ? x : fast);
}

// exp(a d) from d and as = a * 2048/ln2, for 0 <= d <= 708 / |a| (checked by the caller).
static __constant__ double EXP_SCALED_C[3] = {0x1.62e42fefa39efp-12,                       // ln2/2048
                                              0x1.62e42fefa39efp-12 * 0x1.62e42fefa39efp-12 * 0.5,
                                              0x1.62e42fefa39efp-12 * 0x1.62e42fefa39efp-12 *
                                                  0x1.62e42fefa39efp-12 / 6.0};
__device__ __forceinline__ double exp_scaled(double d, double as, const double *__restrict__ tab) {
  const double t = fma(d, as, 6755399441055744.0);
  const int m = __double2loint(t);
  const double mf = t - 6755399441055744.0;
  const double r = fma(d, as, -mf); // in table steps, exact up to one rounding
  double u = fma(r, EXP_SCALED_C[2], EXP_SCALED_C[1]);
  u = fma(r, u, EXP_SCALED_C[0]);
  const double q = r * u; // expm1(r ln2/2048)
  const double tj = tab[m & 2047];
  const double res = fma(tj, q, tj);
  return exp_scale<11>(res, m);
}
// the same for a single lane whose distance is outside the fast range (cold path)
static __device__ __noinline__ double exp_scaled_special(double d, double a) {
  const double x = a * d;
  return (x < -708.0) ? 0. : ((x != x) ? x : exp(x));
}

// sqrt(a) for positive normal a: MUFU.RSQ64H seed (~2^-22) + one Goldschmidt step on g ~ sqrt(a)
// + residual correction with the un-refined h ~ 1/(2 sqrt(a)) (its 2^-22 error only scales a
// 2^-44 correction): 6 FP64-pipe instructions, 0.51 ulp measured.  Valid iff 0x00100000 <= hi(a) <= 0x7fefffff (signed), which
// the caller checks once per batch through hi_min / hi_max.
__device__ __forceinline__ double sqrt_core(double a, int &hi_min, int &hi_max) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y;
  double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  const double e = fma(-g, g, a);
  const int hi = __double2hiint(a);
  hi_min = min(hi_min, hi);
  hi_max = max(hi_max, hi);
  return fma(e, h, g);
}

// the same without range bookkeeping (AB_GRAM_ONECHECK): a zero, subnormal or non-finite argument
// yields NaN (0 * inf, inf - inf), which the exp range check of the caller catches
__device__ __forceinline__ double sqrt_core_nocheck(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y;
  const double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  const double e = fma(-g, g, a);
  return fma(e, h, g);
}

__device__ __forceinline__ double sqrt_patch(double a, double fast) {
  if (a >= 2.2250738585072014e-308 && a < INFINITY) {
    return fast;
  }
  if (a > 0. && a < INFINITY) { // subnormal: rescale by 2^108
    int lo = 0x7fffffff, hi = 0;
    return sqrt_core(a * 3.2451855365842673e32, lo, hi) * 5.551115123125783e-17;
  }
  return a; // 0, inf, NaN map to themselves (negative cannot occur: a is a sum of squares)
}

// ------------------------------------------------------------------------------------------------
// evaluators
// ------------------------------------------------------------------------------------------------

constexpr int MODE_SUM = 0, MODE_SUM_NOISE = 1, MODE_SOP = 2, MODE_STACK = 3;

// exp of NP arguments with one range check for the batch
// table kinds: plain 128 entries (degree 5) | plain 2048 entries (degree 3) | 256 entries x 16 copies
constexpr int TAB_128 = 0, TAB_2048 = 1, TAB_R256 = 2;

// exp of NP arguments, no check: hi_acc accumulates the range information for the caller
template <int NP, int TAB>
__device__ __forceinline__ void exp_batch_nocheck(const double (&v)[NP],
                                                  const double *__restrict__ tab, double (&e)[NP],
                                                  int &hi_acc) {
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    e[i] = TAB == TAB_2048   ? exp_core_big(v[i], tab, hi_acc)
           : TAB == TAB_R256 ? exp_core_r256(v[i], tab, hi_acc)
                             : exp_core(v[i], tab, hi_acc);
  }
}

template <int NP, int TAB = TAB_128>
__device__ __forceinline__ void exp_batch(const double (&v)[NP], const double *__restrict__ tab,
                                          double (&e)[NP]) {
  int hi_max = EXP_HI_LIMIT;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    e[i] = TAB == TAB_2048   ? exp_core_big(v[i], tab, hi_max)
           : TAB == TAB_R256 ? exp_core_r256(v[i], tab, hi_max)
                             : exp_core(v[i], tab, hi_max);
  }
  if (hi_max > EXP_HI_LIMIT) { // rare
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      e[i] = exp_patch(v[i], e[i]);
    }
  }
}

// Run-time program, sum-of-products form (gram.cuh): the op loop is uniform across the CTA and its
// decode cost is amortised over the NP pairs a thread owns.
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
      double e[NP];
      exp_batch<NP>(v, tab, e);
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

__device__ __noinline__ double eval_stack(const DevProg &P, double d2, double dist, bool equal);

template <int MODE> struct EvalProgram {
  static constexpr bool NEED_EQ = MODE != MODE_SUM;
  static constexpr bool ONE_CHECK = false;
  template <int NP>
  __device__ static __forceinline__ bool run_fast(const DevProg &, const double (&)[NP],
                                                  const double (&)[NP], unsigned, const double *,
                                                  double (&)[NP]) {
    return true;
  }
  static constexpr int TABLE = 128; // doubles of shared memory
  __device__ static __forceinline__ const double *table() { return EXP_TABLE; }
  __device__ static __forceinline__ void fill_table(double *smem, int tid) {
    for (int idx = tid; idx < TABLE; idx += GRAM_THREADS) {
      smem[idx] = EXP_TABLE[idx];
    }
  }
  __device__ static __forceinline__ const double *lane_table(const double *smem, int) { return smem; }
  __device__ static __forceinline__ bool need_dist(const DevProg &P) { return P.need_dist != 0; }
  template <int NP>
  __device__ static __forceinline__ void run(const DevProg &P, const double (&d2)[NP],
                                             const double (&dist)[NP], unsigned eqmask,
                                             const double *__restrict__ tab, double (&out)[NP]) {
    if (MODE != MODE_STACK) {
      eval_sop<NP, MODE>(P, d2, dist, eqmask, tab, out);
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        out[i] = eval_stack(P, d2[i], dist[i], (eqmask >> i) & 1u);
      }
    }
  }
};

// Compile-time leaf kinds of EvalFixed.
enum LeafSig : int { LS_NONE = 0, LS_SE = 1, LS_EXP = 2, LS_M32 = 3, LS_M52 = 4, LS_CONST = 5, LS_NOISE = 6 };

template <int KIND, int NP, int TAB, bool CHECK = true>
__device__ __forceinline__ void fixed_term(const DevOp &o, const double (&d2)[NP],
                                           const double (&dist)[NP], unsigned eqmask,
                                           const double *__restrict__ tab, double (&out)[NP],
                                           int &hi_acc) {
  if constexpr (KIND == LS_NONE) {
    return;
  } else if constexpr (KIND == LS_CONST) {
    const double amp = o.amp;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      out[i] = fma(amp, 1., out[i]);
    }
  } else if constexpr (KIND == LS_NOISE) {
    const double amp = o.amp;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      out[i] = fma(amp, ((eqmask >> i) & 1u) ? 1. : 0., out[i]);
    }
  } else {
    double v[NP], e[NP];
    static_assert(TAB == TAB_2048, "the scaled-domain exp uses the 2048-entry table");
    {
      const double as = KIND == LS_SE ? o.a2s : o.a1s;
      const unsigned lim = static_cast<unsigned>(o.lim_hi);
      unsigned worst = 0u;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[i] = KIND == LS_SE ? d2[i] : dist[i];
        const unsigned hi = static_cast<unsigned>(__double2hiint(v[i]));
        if constexpr (CHECK) { // per lane: the same bits as the branch-free path when in range
          e[i] = hi <= lim ? exp_scaled(v[i], as, tab)
                           : exp_scaled_special(v[i], KIND == LS_SE ? o.a2 : o.a1);
        } else {
          e[i] = exp_scaled(v[i], as, tab);
          worst = max(worst, hi);
        }
      }
      if constexpr (!CHECK) {
        // fold into the caller's signed accumulator: any value above EXP_HI_LIMIT means "re-evaluate"
        hi_acc = max(hi_acc, worst > lim ? 0x7fffffff : EXP_HI_LIMIT);
      }
    }
    const double amp = o.amp;
    if constexpr (KIND == LS_M32 || KIND == LS_M52) {
      // amp * (1 + b1 d [+ b2 d^2]) * e as fma(poly', e, out), poly' = amp + (amp b1) d [+ (amp b2) d^2]
      const double ab1 = o.ab1, ab2 = o.ab2;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        double poly = fma(ab1, dist[i], amp);
        if constexpr (KIND == LS_M52) {
          poly = fma(ab2, d2[i], poly);
        }
        out[i] = fma(poly, e[i], out[i]);
      }
      return;
    }
    if constexpr (KIND == LS_M32) {
      const double b1 = o.b1;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        e[i] = e[i] * fma(b1, dist[i], 1.);
      }
    } else if constexpr (KIND == LS_M52) {
      const double b1 = o.b1, b2 = o.b2;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        e[i] = e[i] * fma(b2, d2[i], fma(b1, dist[i], 1.));
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      out[i] = fma(amp, e[i], out[i]);
    }
  }
}

// Sum of up to three single leaves with compile-time kinds; arithmetic identical to
// eval_sop<MODE_SUM / MODE_SUM_NOISE> on the same program (same operations in the same order).
// AB_GRAM_TABLE (tools/sweep.sh) selects the table of the fixed evaluators; AB_GRAM_TABLE_GLOBAL
// reads a plain table through L1 instead of shared memory (measured slower: L1/shared is the
// kernel's busiest unit).
#ifndef AB_GRAM_TABLE
#define AB_GRAM_TABLE 1
#endif
constexpr int FIXED_TABLE = AB_GRAM_TABLE;
constexpr bool TABLE_IN_SMEM = true;

template <int K0, int K1, int K2, int TAB = FIXED_TABLE> struct EvalFixed {
  static constexpr bool NEED_EQ = K0 == LS_NOISE || K1 == LS_NOISE || K2 == LS_NOISE;
  static constexpr int TABLE = TAB == TAB_2048 ? 2048 : (TAB == TAB_R256 ? 256 * EXP_REPL : 128);
  __device__ static __forceinline__ const double *table() {
    return TAB == TAB_128 ? EXP_TABLE : EXP_TABLE_BIG;
  }
  __device__ static __forceinline__ void fill_table(double *smem, int tid) {
    for (int idx = tid; idx < TABLE; idx += GRAM_THREADS) {
      // 2^(j/256) = EXP_TABLE_BIG[8 j]
      smem[idx] = TAB == TAB_R256 ? EXP_TABLE_BIG[(idx / EXP_REPL) * 8] : table()[idx];
    }
  }
  __device__ static __forceinline__ const double *lane_table(const double *smem, int lane) {
    return TAB == TAB_R256 ? smem + (lane & (EXP_REPL - 1)) : smem;
  }
  static constexpr bool NEED_DIST = (K0 >= LS_EXP && K0 <= LS_M52) || (K1 >= LS_EXP && K1 <= LS_M52) ||
                                    (K2 >= LS_EXP && K2 <= LS_M52);
  __device__ static __forceinline__ bool need_dist(const DevProg &) { return NEED_DIST; }
  template <int NP>
  __device__ static __forceinline__ void run(const DevProg &P, const double (&d2)[NP],
                                             const double (&dist)[NP], unsigned eqmask,
                                             const double *__restrict__ tab, double (&out)[NP]) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      out[i] = 0.;
    }
    int unused = 0;
    fixed_term<K0, NP, TAB>(P.ops[0], d2, dist, eqmask, tab, out, unused);
    fixed_term<K1, NP, TAB>(P.ops[1], d2, dist, eqmask, tab, out, unused);
    fixed_term<K2, NP, TAB>(P.ops[2], d2, dist, eqmask, tab, out, unused);
  }
  // The same arithmetic without any branch; returns true when some exp argument was out of range
  // (then `out` is garbage and the caller re-evaluates the pass through run()).  Because every use
  // of `dist` is an exp argument a * dist with a finite a, a NaN distance (the branch-free sqrt of a
  // zero, subnormal or non-finite squared distance) is caught by the same test.
  static constexpr bool ONE_CHECK = AB_GRAM_ONECHECK != 0;
  template <int NP>
  __device__ static __forceinline__ bool run_fast(const DevProg &P, const double (&d2)[NP],
                                                  const double (&dist)[NP], unsigned eqmask,
                                                  const double *__restrict__ tab,
                                                  double (&out)[NP]) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      out[i] = 0.;
    }
    int hi_acc = EXP_HI_LIMIT;
    fixed_term<K0, NP, TAB, false>(P.ops[0], d2, dist, eqmask, tab, out, hi_acc);
    fixed_term<K1, NP, TAB, false>(P.ops[1], d2, dist, eqmask, tab, out, hi_acc);
    fixed_term<K2, NP, TAB, false>(P.ops[2], d2, dist, eqmask, tab, out, hi_acc);
    return hi_acc > EXP_HI_LIMIT;
  }
};

// ------------------------------------------------------------------------------------------------
// the tile kernel
// ------------------------------------------------------------------------------------------------

// Work item t of the launch -> tile (I, J); returns false for items that carry no tile.
//   SYM: the lower triangle is walked in SB x SB super-blocks of tiles (row by row over the
//        super-blocks, t = (B_I (B_I + 1) / 2 + B_J) * SB^2 + local), rows fastest inside a
//        super-block.  The ~300 tiles in flight at any time then cover a compact 2-D patch of K, so
//        that both the direct stores (contiguous along rows) and the mirrored ones stay within a
//        few hundred 2 MB pages instead of sweeping every column of K for each row of tiles
//        (measured store-bound rates before: 5.3 TB/s at N = 16 384 falling to 4.5 TB/s at 49 152).
//        Items above the diagonal or beyond the last tile row are skipped.  32-bit arithmetic is
//        exact: the launcher caps the item count at 2^31.
//   otherwise column-major over the tiles_i x tiles_j grid (rows fastest: contiguous stores).
#ifndef AB_GRAM_SB
#define AB_GRAM_SB 16
#endif
constexpr unsigned SB = AB_GRAM_SB;

template <bool SYM>
__device__ __forceinline__ bool decode_tile(unsigned t, unsigned tiles_i, unsigned &I, unsigned &J) {
  if (SYM) {
    const unsigned sb = t / (SB * SB);
    const unsigned local = t % (SB * SB);
    unsigned i = static_cast<unsigned>((sqrtf(8.f * static_cast<float>(sb) + 1.f) - 1.f) * 0.5f);
    while (i * (i + 1u) / 2u > sb) {
      --i;
    }
    while ((i + 1u) * (i + 2u) / 2u <= sb) {
      ++i;
    }
    I = i * SB + local % SB;
    J = (sb - i * (i + 1u) / 2u) * SB + local / SB;
    return I < tiles_i && J <= I;
  }
  I = t % tiles_i;
  J = t / tiles_i;
  return true;
}

// First work item >= t (stepping by `step`) that carries a tile, decoded into (I, J).
template <bool SYM>
__device__ __forceinline__ unsigned next_tile(unsigned t, unsigned step, unsigned nitems,
                                              unsigned tiles_i, unsigned &I, unsigned &J) {
  while (t < nitems && !decode_tile<SYM>(t, tiles_i, I, J)) {
    t += step;
  }
  return t;
}

// Number of work items of a launch (see decode_tile).
inline int64_t gram_items(bool sym, int64_t tiles_i, int64_t tiles_j) {
  if (!sym) {
    return tiles_i * tiles_j;
  }
  const int64_t nsb = (tiles_i + SB - 1) / SB;
  return nsb * (nsb + 1) / 2 * SB * SB;
}

// Interior-tile stores (evict-first marking measured no gain, profiles/r01d_sweeps.txt).
template <class T> __device__ __forceinline__ void gram_store(T *dst, const T &v) {
  *dst = v;
}

constexpr int FEAT_SLOTS = 2; // TILE * AB_MAX_DIM / GRAM_THREADS feature elements per thread

// This thread's share of the x / y features of tile (I, J), straight from global memory.
template <int DIM>
__device__ __forceinline__ void fetch_features(const double *__restrict__ fx, int64_t ldfx,
                                               int64_t n, const double *__restrict__ fy,
                                               int64_t ldfy, int64_t m, unsigned I, unsigned J,
                                               int tid, double (&px)[FEAT_SLOTS],
                                               double (&py)[FEAT_SLOTS]) {
  const int64_t i0 = static_cast<int64_t>(I) * TILE;
  const int64_t j0 = static_cast<int64_t>(J) * TILE;
#pragma unroll
  for (int e = 0; e < FEAT_SLOTS; ++e) {
    const int idx = tid + e * GRAM_THREADS;
    px[e] = 0.;
    py[e] = 0.;
    if (idx < TILE * DIM) {
      const int p = idx / DIM;
      const int d = idx - p * DIM;
      if (i0 + p < n) {
        px[e] = fx[(i0 + p) * ldfx + d];
      }
      if (j0 + p < m) {
        py[e] = fy[(j0 + p) * ldfy + d];
      }
    }
  }
}

// One slice (1 / PARTS) of the transposed copy of a finished tile: element (row = j0 + c,
// col = i0 + r) = stage[c][r]; each warp-store covers 32 consecutive rows (256 contiguous bytes).
template <int PARTS>
__device__ __forceinline__ void mirror_slice(const double *__restrict__ stage,
                                             double *__restrict__ out, int64_t ld, int64_t n,
                                             unsigned I, unsigned J, bool interior, int part,
                                             int lane, int warp) {
  constexpr int KS = 8 / PARTS;
  const int64_t i0 = static_cast<int64_t>(I) * TILE;
  const int64_t j0 = static_cast<int64_t>(J) * TILE;
  const int rbase = warp * 8 + part * KS;
  if (interior) {
    double *dst = out + (j0 + lane) + (i0 + rbase) * ld;
    const double *src = stage + lane * LDT + rbase;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      gram_store(dst, src[k]);
      gram_store(dst + 32, src[32 * LDT + k]);
      dst += ld;
    }
  } else {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = rbase + k;
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

// Hides the provenance of a pointer from the optimiser.
__device__ __forceinline__ double *opaque_ptr(double *p) {
  unsigned long long v = reinterpret_cast<unsigned long long>(p);
  asm volatile("" : "+l"(v));
  double *q = reinterpret_cast<double *>(v);
  __builtin_assume(__isGlobal(q)); // keep st.global (a pointer of unknown provenance stores generically)
  return q;
}

// The same slice of an interior tile addressed as tile base + 32-bit element offset: mbase = &out[j0, i0];
// one IMAD.WIDE per address instead of 64-bit index arithmetic.
template <int PARTS>
__device__ __forceinline__ void mirror_slice_off(const double *__restrict__ stage,
                                                 double *__restrict__ mbase, unsigned ld32, int part,
                                                 int lane, int warp) {
  constexpr int KS = 8 / PARTS;
  __builtin_assume(__isGlobal(mbase));
  const int rbase = warp * 8 + part * KS;
  const double *src = stage + lane * LDT + rbase;
  unsigned off = static_cast<unsigned>(lane) + static_cast<unsigned>(rbase) * ld32;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    double *dst = mbase + off;
    gram_store(dst, src[k]);
    gram_store(dst + 32, src[32 * LDT + k]);
    off += ld32;
  }
}

// Persistent CTAs: each CTA walks the tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
//   * the features of the next tile are fetched into registers while the current one is evaluated
//     (global-load latency, table load and the store drain at exit are paid once per CTA);
//   * the transposed copy of tile k is written while tile k+1 is being evaluated: the staging
//     buffer is double-buffered and one slice of the pending mirror is issued after every pass, so
//     the LDS/STG traffic of the mirror hides behind FP64 work instead of forming a phase of its own
//     (measured: the separate mirror phase cost 0.53 ms of 2.28 ms at N = 32 768).
template <int DIM, bool SYM, class EV, int COLS = 4, int MINB = 2>
__global__ void __launch_bounds__(GRAM_THREADS, MINB)
gram_kernel(const __grid_constant__ DevProg P, const double *__restrict__ fx, int64_t ldfx,
            int64_t n, const double *__restrict__ fy, int64_t ldfy, int64_t m,
            double *__restrict__ out, int64_t ld, int tiles_i, unsigned ntiles, uint32_t flags) {
  constexpr int NPAIR = 2 * COLS;
  constexpr int PASSES = 8 / COLS;
  constexpr int STAGE = TILE * LDT;
  // dynamic shared memory (gram_smem_bytes): exp table | x features | y features | 2 x mirror staging
  extern __shared__ __align__(16) double gram_smem[];
  constexpr int TAB_SMEM = TABLE_IN_SMEM ? EV::TABLE : 0;
  double *xs = gram_smem + TAB_SMEM;
  double *ys = xs + TILE * DIM;
  double *stage0 = ys + TILE * DIM;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int r0 = 2 * lane;
  if (TABLE_IN_SMEM) {
    EV::fill_table(gram_smem, tid);
  }
  const double *tab = TABLE_IN_SMEM ? EV::lane_table(gram_smem, lane) : EV::table();
  const bool need_dist = DIM != 1 && EV::need_dist(P);

  unsigned t = blockIdx.x;
  unsigned I = 0, J = 0;
  double px[FEAT_SLOTS], py[FEAT_SLOTS];
  t = next_tile<SYM>(t, gridDim.x, ntiles, static_cast<unsigned>(tiles_i), I, J);
  if (t < ntiles) {
    fetch_features<DIM>(fx, ldfx, n, fy, ldfy, m, I, J, tid, px, py);
  }
  // pending mirror of the previous tile: bit 0 = pending, bit 1 = interior, bit 2 = staging buffer
  unsigned pend = 0, pI = 0, pJ = 0;
  unsigned buf = 0;
  // interior-tile stores are addressed as 64-bit tile base + 32-bit element offset (one IMAD.WIDE per
  // address instead of 64-bit index arithmetic in every pass: -18 instructions per pass, measured
  // 1.956 -> 1.913 ms, profiles/r02a_gram_sweep.txt)
  const unsigned ld32 = static_cast<unsigned>(ld);
  double *mbase = out; // &out[j0, i0] of the pending mirror

  while (t < ntiles) {
    const unsigned cI = I, cJ = J;
    const int64_t i0 = static_cast<int64_t>(cI) * TILE;
    const int64_t j0 = static_cast<int64_t>(cJ) * TILE;
    const bool mirror = SYM && cI != cJ && !(flags & AB_GRAM_LOWER_ONLY);
    // interior tile of a 16-byte aligned output: no bounds checks anywhere below
    const bool interior = (i0 + TILE <= n) && (j0 + TILE <= m) && !(flags & GRAM_UNALIGNED) &&
                          (ld >> 24) == 0; // 64 * ld elements fit a 32-bit offset
    double *stage = stage0 + buf * STAGE;

    // every reader of xs / ys of the previous tile is done, every slice of the mirror before the
    // pending one has been read, and the pending tile's staging buffer is completely written
    __syncthreads();
#pragma unroll
    for (int e = 0; e < FEAT_SLOTS; ++e) {
      const int idx = tid + e * GRAM_THREADS;
      if (idx < TILE * DIM) {
        xs[idx] = px[e];
        ys[idx] = py[e];
      }
    }
    __syncthreads();

    // prefetch the next tile's features; they are consumed at the top of the next iteration
    t = next_tile<SYM>(t + gridDim.x, gridDim.x, ntiles, static_cast<unsigned>(tiles_i), I, J);
    if (t < ntiles) {
      fetch_features<DIM>(fx, ldfx, n, fy, ldfy, m, I, J, tid, px, py);
    }

    double xi[2][DIM];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        xi[a][d] = xs[(r0 + a) * DIM + d];
      }
    }
    const int64_t gi = i0 + r0;
    // &out[i0, j0]; opaque: keeps base + 32-bit offset (else ptxas re-derives 64-bit indices)
    double *tbase = opaque_ptr(out + i0 + j0 * ld);

#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
      const int cbase = pass * (8 * COLS) + warp * COLS;
      double d2[NPAIR], dist[NPAIR], vals[NPAIR];
      unsigned eqmask = 0;
#pragma unroll
      for (int k = 0; k < COLS; ++k) {
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
          if (EV::NEED_EQ) {
            bool eq = true;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
              eq = eq && (xi[a][d] == yj[d]);
            }
            eqmask |= (eq ? 1u : 0u) << (2 * k + a);
          }
        }
      }
      if constexpr (EV::ONE_CHECK) {
        if (need_dist) {
#pragma unroll
          for (int i = 0; i < NPAIR; ++i) {
            dist[i] = sqrt_core_nocheck(d2[i]);
          }
        }
        if (EV::template run_fast<NPAIR>(P, d2, dist, eqmask, tab, vals)) {
          // rare (the diagonal, coincident points, underflowing or non-finite arguments): repair the
          // distances lane by lane and take the checked evaluator; in-range lanes get the same bits
          if (need_dist) {
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
              dist[i] = sqrt_patch(d2[i], dist[i]);
            }
          }
          EV::template run<NPAIR>(P, d2, dist, eqmask, tab, vals);
        }
      } else {
        if (need_dist) {
          int hi_min = 0x00100000, hi_max = 0x7fefffff;
#pragma unroll
          for (int i = 0; i < NPAIR; ++i) {
            dist[i] = sqrt_core(d2[i], hi_min, hi_max);
          }
          if (hi_min < 0x00100000 || hi_max > 0x7fefffff) {
            // rare: a zero / subnormal / non-finite squared distance (e.g. the diagonal)
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
              dist[i] = sqrt_patch(d2[i], dist[i]);
            }
          }
        }
        EV::template run<NPAIR>(P, d2, dist, eqmask, tab, vals);
      }

      // direct tile: rows i0 + r0 (+1), columns j0 + c; 16-byte stores, 512 contiguous bytes/warp
      if (interior) {
        unsigned off = static_cast<unsigned>(r0) + static_cast<unsigned>(cbase) * ld32;
#pragma unroll
        for (int k = 0; k < COLS; ++k) {
          gram_store(reinterpret_cast<double2 *>(tbase + off), make_double2(vals[2 * k], vals[2 * k + 1]));
          off += ld32;
        }
      } else {
        const bool unaligned = flags & GRAM_UNALIGNED;
#pragma unroll
        for (int k = 0; k < COLS; ++k) {
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
      }
      if (mirror) {
#pragma unroll
        for (int k = 0; k < COLS; ++k) {
          const int c = cbase + k;
          stage[c * LDT + r0] = vals[2 * k];
          stage[c * LDT + r0 + 1] = vals[2 * k + 1];
        }
      }
      if (SYM && (pend & 1u)) {
        if (pend & 2u) {
          mirror_slice_off<PASSES>(stage0 + ((pend >> 2) & 1u) * STAGE, mbase, ld32, pass, lane, warp);
        } else {
          mirror_slice<PASSES>(stage0 + ((pend >> 2) & 1u) * STAGE, out, ld, n, pI, pJ, pend & 2u,
                               pass, lane, warp);
        }
      }
    }

    if (SYM) {
      pend = mirror ? (1u | (interior ? 2u : 0u) | (buf << 2)) : 0u;
      pI = cI;
      pJ = cJ;
      mbase = opaque_ptr(out + j0 + i0 * ld);
      if (mirror) {
        buf ^= 1u;
      }
    }
  }

  if (SYM && (pend & 1u)) { // the last tile's mirror has no next tile to hide behind
    __syncthreads();
#pragma unroll 1
    for (int part = 0; part < PASSES; ++part) {
      if (pend & 2u) {
        mirror_slice_off<PASSES>(stage0 + ((pend >> 2) & 1u) * STAGE, mbase, ld32, part, lane, warp);
      } else {
        mirror_slice<PASSES>(stage0 + ((pend >> 2) & 1u) * STAGE, out, ld, n, pI, pJ, pend & 2u,
                             part, lane, warp);
      }
    }
  }
}

template <int DIM, bool SYM, class EV> constexpr size_t gram_smem_bytes() {
  return sizeof(double) * ((TABLE_IN_SMEM ? EV::TABLE : 0) + 2 * TILE * DIM + (SYM ? 2 * TILE * LDT : 0));
}

// Launches the persistent kernel: MINB CTAs per SM (or one per tile when there are fewer tiles).
template <int DIM, bool SYM, class EV, int COLS = 4, int MINB = 2>
inline cudaError_t gram_launch(ab_handle_s *h, const DevProg &P, const double *fx, int64_t ldfx,
                               int64_t n, const double *fy, int64_t ldfy, int64_t m, double *out,
                               int64_t ld, int tiles_i, unsigned ntiles, uint32_t flags) {
  constexpr size_t smem = gram_smem_bytes<DIM, SYM, EV>();
  static PerDeviceOnce once;
  if (smem > 48 * 1024 && once.need(h->device)) {
    cudaError_t e = cudaFuncSetAttribute(gram_kernel<DIM, SYM, EV, COLS, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      return e;
    }
    if (MINB > 2) { // three CTAs per SM only fit with the maximum shared-memory carveout
      e = cudaFuncSetAttribute(gram_kernel<DIM, SYM, EV, COLS, MINB>,
                               cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) {
        return e;
      }
    }
  }
  const int64_t resident = static_cast<int64_t>(MINB) * h->sm_count;
  const unsigned grid = static_cast<unsigned>(ntiles < resident ? ntiles : resident);
  gram_kernel<DIM, SYM, EV, COLS, MINB><<<grid, GRAM_THREADS, smem, h->stream>>>(
      P, fx, ldfx, n, fy, ldfy, m, out, ld, tiles_i, ntiles, flags);
  return cudaGetLastError();
}

// Launch of the compile-time-specialised kernels (gram_fixed.cu).  Returns true when the program
// has a specialisation (and the launch was issued), false when the generic kernel must be used.
bool launch_gram_fixed(ab_handle_s *h, const DevProg &P, int dim, bool sym, const double *fx,
                       int64_t ldfx, int64_t n, const double *fy, int64_t ldfy, int64_t m,
                       double *out, int64_t ld, int tiles_i, unsigned tiles, uint32_t flags);

} // namespace ab
