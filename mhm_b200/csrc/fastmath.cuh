// fastmath.cuh -- fp64 log / exp / pow for the fast cell kernel.
//
// The cascade spends half of its instructions in exp(b * log(x)) (infiltration shape factor,
// mo_soil_moisture.f90:233; slow interflow S**(1+alpha), mo_runoff.f90:134).  libdevice's
// general-purpose exp/log carry special-case handling (negative, zero, infinite, subnormal
// arguments) and materialise every polynomial coefficient with two 32-bit moves.  The
// arguments here are known to be positive, finite and normal, and the exponents bounded, so the
// textbook algorithms are enough:
//   log : x = 2^e * m, m in [sqrt(1/2), sqrt(2)); s = f / (2 + f), f = m - 1;
//         log(m) = 2s + s*R(s^2) with the degree-7 even polynomial of fdlibm's e_log.c
//         (error < 1 ulp), assembled with the hi/lo split of ln 2;
//   exp : k = rint(x / ln 2), r = x - k ln2 (two-constant Cody-Waite, |r| <= 0.3466),
//         degree-13 Taylor polynomial (truncation error 4e-18), scaled by 2^k through the
//         exponent field.
// Coefficients live in __constant__ memory so that they enter DFMA as constant-bank operands.
// Accuracy (tests/test_fastmath.py, 2e6 samples against glibc): log <= 1 ulp, exp <= 1 ulp,
// pow_pos relative error <= 2.3e-16 * (1 + |y log x|).
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define MHM_HD __host__ __device__ __forceinline__
#else
#define MHM_HD inline
#endif

#include "fastmath_tables.h"

#ifndef MHM_FM_ESTRIN
#define MHM_FM_ESTRIN 1
#endif

namespace mhm {
namespace fm {

struct Coef {
  double lg[7];       // Lg1..Lg7 of fdlibm e_log.c
  double ln2_hi, ln2_lo, inv_ln2;
  double ex[12];      // 1/2! .. 1/13!
  // table-driven versions: Taylor coefficients that are not exact in the high word, range
  // reduction constants, and epsilon(1.0_dp).  Kept in the constant bank because a 64-bit
  // literal costs two moves into a uniform register at every use.
  double third, fifth, msixth, c3, c4, c5;
  double ln2hi42, ln2lo42, invln2n, mln2hin, mln2lon;
  double eps;
};

#if defined(__CUDA_ARCH__)
#define MHM_FM_COEF c_coef
#else
#define MHM_FM_COEF h_coef
#endif

static const Coef h_coef = {
    {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
     2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
     1.479819860511658591e-01},
    6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00,
    {1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
     1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0},
    0.33333333333333333, 0.2, -0.16666666666666666, 0.16666666666666666, 0.041666666666666664,
    0.0083333333333333332, kLn2Hi42, kLn2Lo42, kInvLn2N, -kLn2HiN, -kLn2LoN,
    2.220446049250313e-16};

#if defined(__CUDACC__)
__constant__ Coef c_coef = {
    {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
     2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
     1.479819860511658591e-01},
    6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00,
    {1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
     1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0},
    0.33333333333333333, 0.2, -0.16666666666666666, 0.16666666666666666, 0.041666666666666664,
    0.0083333333333333332, kLn2Hi42, kLn2Lo42, kInvLn2N, -kLn2HiN, -kLn2LoN,
    2.220446049250313e-16};
#endif

MHM_HD int hi_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  int64_t b;
  std::memcpy(&b, &x, 8);
  return (int)(b >> 32);
#endif
}
MHM_HD int lo_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  int64_t b;
  std::memcpy(&b, &x, 8);
  return (int)(b & 0xffffffff);
#endif
}
MHM_HD double make_double(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  int64_t b = ((int64_t)hi << 32) | (uint32_t)lo;
  double x;
  std::memcpy(&x, &b, 8);
  return x;
#endif
}
MHM_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
// reciprocal to full precision for y in [1.7, 2.5): hardware seed + two Newton steps
MHM_HD double recip(double y) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
#else
  double r = (double)(1.0f / (float)y);
#endif
  r = fma_(fma_(-y, r, 1.0), r, r);
  r = fma_(fma_(-y, r, 1.0), r, r);
  return r;
}

// natural logarithm of a positive, finite, normal double
MHM_HD double log_pos(double x) {
  int hi = hi_word(x);
  const int lo = lo_word(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;  // m in [1, 2)
  if (hi >= 0x3ff6a09f) {               // m >= sqrt(2): halve it
    hi -= 0x00100000;
    e += 1;
  }
  const double f = make_double(hi, lo) - 1.0;
  const double s = f * recip(2.0 + f);
  const double z = s * s, w = z * z;
  const Coef& c = MHM_FM_COEF;
  const double t1 = w * fma_(w, fma_(w, c.lg[5], c.lg[3]), c.lg[1]);
  const double t2 = z * fma_(w, fma_(w, fma_(w, c.lg[6], c.lg[4]), c.lg[2]), c.lg[0]);
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)e;
  // e*ln2_hi - ((hfsq - (s*(hfsq + R) + e*ln2_lo)) - f)
  return fma_(dk, c.ln2_hi, -((hfsq - fma_(s, hfsq + R, dk * c.ln2_lo)) - f));
}

// exp(x) for |x| < 700
MHM_HD double exp_bounded(double x) {
  const Coef& c = MHM_FM_COEF;
#if defined(__CUDA_ARCH__)
  const double kd = rint(x * c.inv_ln2);
#else
  const double kd = __builtin_rint(x * c.inv_ln2);
#endif
  double r = fma_(-kd, c.ln2_hi, x);
  r = fma_(-kd, c.ln2_lo, r);
#if MHM_FM_ESTRIN
  // Estrin's scheme: 14 operations in a dependency chain of depth 5 instead of 11 of depth 11
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double a0 = fma_(c.ex[1], r, c.ex[0]), a1 = fma_(c.ex[3], r, c.ex[2]);
  const double a2 = fma_(c.ex[5], r, c.ex[4]), a3 = fma_(c.ex[7], r, c.ex[6]);
  const double a4 = fma_(c.ex[9], r, c.ex[8]), a5 = fma_(c.ex[11], r, c.ex[10]);
  const double b0 = fma_(a1, r2, a0), b1 = fma_(a3, r2, a2), b2 = fma_(a5, r2, a4);
  double p = fma_(b2, r8, fma_(b1, r4, b0));
  p = fma_(p, r2, r);  // r + r^2 * P(r)
#else
  double p = c.ex[11];
#pragma unroll
  for (int i = 10; i >= 0; --i) p = fma_(p, r, c.ex[i]);
  p = fma_(p * r, r, r);  // r + r^2 * P(r)
#endif
  const int k = (int)kd;
  // 2^k through the exponent field; split in two factors so that k down to -1070 stays exact
  const int k1 = k / 2, k2 = k - k1;
  const double s1 = make_double((k1 + 1023) << 20, 0), s2 = make_double((k2 + 1023) << 20, 0);
  return fma_(p, s1, s1) * s2;  // (1 + p) * 2^k1 * 2^k2
}

// x ** y for x > 0 (finite, normal), |y log x| < 700
MHM_HD double pow_pos(double x, double y) { return exp_bounded(y * log_pos(x)); }

// x ** (2/3) for x > 0 (canopy evaporation, mo_canopy_interc.f90:117): r = x ** (-1/3) from a
// single-precision seed (relative error < 2^-20) refined by two Newton steps of
// f(r) = r^-3 - x -- division-free, r <- r + r (1 - x r^3) / 3, error -> 2 e^2 -- then x * r.
MHM_HD double pow23_core(double x);
MHM_HD double pow23_pos(double x) {
  if (x > 1.0e30) return pow_pos(x, 0.6666666666666666666666666666666666667);
  if (x < 1.0e-30) {  // outside the single-precision seed's range: x = m 2^(3k), result m^(2/3) 2^(2k)
    const int k = ((hi_word(x) >> 20) - 1023) / 3;  // k < 0
    const double m = x * make_double((1023 - 3 * k) << 20, 0);
    return pow23_core(m) * make_double((1023 + 2 * k) << 20, 0);
  }
  return pow23_core(x);
}
// pow23_pos for 0 < x <= 1e30 without a branch (the same values): the rescaling of tiny
// arguments is selected instead of skipped, so the call can sit inside straight-line code
MHM_HD double pow23_sel(double x) {
  int k = ((hi_word(x) >> 20) - 1023) / 3;
  k = x < 1.0e-30 ? k : 0;  // k = 0: both scale factors are exactly 1
  const double m = x * make_double((1023 - 3 * k) << 20, 0);
  return pow23_core(m) * make_double((1023 + 2 * k) << 20, 0);
}
// x ** (2/3) for x >= 1.2e-38 (single-precision seed in range), NaN below it (also for x = 0):
// the canopy's relative storage is at most 1, and an evaporation of less than 1e-25 * pet is
// discarded by the caller's `> 0` select together with the NaN
MHM_HD double pow23_nz(double x) { return pow23_core(x); }
MHM_HD double pow23_core(double x) {
#if defined(__CUDA_ARCH__)
  float l, r0;
  const float xf = (float)x;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(xf));
  l *= -0.333333333f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(l));
#else
  const float r0 = __builtin_exp2f(-0.333333333f * __builtin_log2f((float)x));
#endif
  double r = (double)r0;
  const double third = 0.33333333333333333333;
  double e = fma_(-x * (r * r), r, 1.0);
  r = fma_(r * e, third, r);
  e = fma_(-x * (r * r), r, 1.0);
  r = fma_(r * e, third, r);
  return x * r;
}

// ---- table-driven log / exp (128 entries each, 4 KB; the cell kernel stages them in shared
// memory).  Half the operations and less than half the dependency depth of the polynomial-only
// versions above: no division in log, degree-5 polynomials.  Absolute error of log_tab
// <= 0.6 ulp(|log x|) + 6e-17, relative error of exp_tab <= 0.6 ulp; see tests/test_fastmath.py.
struct LogEntry {
  double invc, logc;
};
struct ExpEntry {
  double tail;
  unsigned long long sbits;
};
struct Tables {
  LogEntry lg[128];
  ExpEntry ex[128];
};
static const Tables h_tables = {{MHM_FM_LOG_TABLE}, {MHM_FM_EXP_TABLE}};
#if defined(__CUDACC__)
__device__ const Tables d_tables = {{MHM_FM_LOG_TABLE}, {MHM_FM_EXP_TABLE}};
#endif

// natural logarithm of a positive, finite, normal double
MHM_HD double log_tab(const Tables& T, double x) {
  const int hi = hi_word(x), lo = lo_word(x);
  const int tmp = hi - 0x3fe60000;         // high word of bits(x) - OFF (the low word of OFF is 0)
  const int i = (tmp >> 13) & 127;         // bits 45..51 of the offset
  const int k = tmp >> 20;                 // arithmetic shift: exponent relative to [0.6875, 1.375)
  const double z = make_double(hi - (tmp & (int)0xfff00000), lo);
  const LogEntry e = T.lg[i];
  const double r = fma_(z, e.invc, -1.0);  // exact argument of log1p, |r| < 2^-7.9
  const double kd = (double)k;
  const Coef& c = MHM_FM_COEF;
  const double h = fma_(kd, c.ln2hi42, e.logc);
  const double r2 = r * r;
  // log1p(r) = r + r^2 (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6), truncation < 2e-18
  const double q = fma_(r2, fma_(r2, c.msixth, fma_(c.fifth, r, -0.25)), fma_(c.third, r, -0.5));
  const double l = fma_(kd, c.ln2lo42, r2 * q);
  return h + (r + l);
}

// exp(x) for |x| < 700
MHM_HD double exp_tab(const Tables& T, double x) {
  const Coef& c = MHM_FM_COEF;
#if defined(__CUDA_ARCH__)
  const double kd = rint(x * c.invln2n);
#else
  const double kd = __builtin_rint(x * c.invln2n);
#endif
  const int ki = (int)kd;
  double r = fma_(kd, c.mln2hin, x);
  r = fma_(kd, c.mln2lon, r);              // |r| <= ln2 / 256
  const ExpEntry e = T.ex[ki & 127];
  // 2^(ki / 128): the table's bit pattern plus ki << 45 (exponent and index in one addition)
  const double scale = make_double((int)(e.sbits >> 32) + (ki << 13), (int)(e.sbits & 0xffffffffu));
  const double r2 = r * r;
  // e^r - 1 = r + r^2/2 + r^3/6 + r^4/24 + r^5/120, truncation < 1e-18
  double t = fma_(r2, fma_(c.c3, r, 0.5), e.tail + r);
  t = fma_(r2 * r2, fma_(c.c5, r, c.c4), t);
  return fma_(scale, t, scale);
}

// x ** y for x > 0 (finite, normal), |y log x| < 700
MHM_HD double pow_tab(const Tables& T, double x, double y) { return exp_tab(T, y * log_tab(T, x)); }

}  // namespace fm
}  // namespace mhm
