// Device exp / log / pow that return, bit for bit, what the host libm of the reference build returns.
//
// Why: the reference (g++ -O2, x86-64) takes exp / log / pow from glibc.  CUDA's libdevice versions are as accurate
// but round differently in the last place in a few per cent of calls, and the Blended template branches on the sign of a
// 1e-17 residual that such a last bit decides (src/SnowBalance.cpp:507, DESIGN.md "Numerical fidelity").  These three
// functions are therefore restated here operation by operation.
//
// What is restated: glibc >= 2.28's table-driven double-precision routines (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c,
// e_pow.c - S. Nagy's algorithms), in the form the x86-64 multiarch build dispatches to on every CPU with FMA + AVX2
// (the `__exp_fma / __log_fma / __pow_fma` ifunc variants, compiled with -mfma so that `__FP_FAST_FMA` selects the fma
// branches and the compiler contracts a*b+c where it chose to).  The contraction pattern is not defined by the C source,
// so every statement below was checked against the instruction sequence of the libm.so.6 of this image (glibc 2.39);
// the comments give the C-source expression each fused operation belongs to.  Tables: rvn_glibc_tables.h, read out of
// that same libm by tools/extract_glibc_tables.py.
//
// The file compiles for the device (nvcc) and for the host (gcc -ffp-contract=off, used only by tests/test_glibc_math.py
// to compare this restatement with the running libm on >= 1e8 arguments before any GPU is involved).
// errno / floating-point exception side effects of the originals are not reproduced (nothing on the path reads them).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RG_FN __device__ __forceinline__
#define RG_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RG_MUL(a, b) __dmul_rn((a), (b))
#define RG_ADD(a, b) __dadd_rn((a), (b))
#define RG_SUB(a, b) __dsub_rn((a), (b))
#define RG_DIV(a, b) __ddiv_rn((a), (b))
#define RG_BITS(x) ((uint64_t)__double_as_longlong(x))
#define RG_DBL(u) __longlong_as_double((long long)(u))
#define RG_I2D(k) __int2double_rn(k)
#define RG_LD(p) __ldg(p)
#else
#include <string.h>
#define RG_FN static inline
#define RG_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define RG_MUL(a, b) ((a) * (b))
#define RG_ADD(a, b) ((a) + (b))
#define RG_SUB(a, b) ((a) - (b))
#define RG_DIV(a, b) ((a) / (b))
static inline uint64_t rg_bits_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double rg_dbl_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define RG_BITS(x) rg_bits_(x)
#define RG_DBL(u) rg_dbl_(u)
#define RG_I2D(k) ((double)(k))
#define RG_LD(p) (*(p))
#endif

#if defined(__CUDACC__)
#define RVN_GLIBC_TABLE_QUAL static __device__ const
#endif
#include "rvn_glibc_tables.h"

#define RG_EXPC(i) RG_DBL(RG_LD(&rvn_glibc_exp_c[i]))
#define RG_LOGC(i) RG_DBL(RG_LD(&rvn_glibc_log_c[i]))
#define RG_POWC(i) RG_DBL(RG_LD(&rvn_glibc_pow_c[i]))

#define RG_INF 0x7ff0000000000000ULL
#define RG_ONE 0x3ff0000000000000ULL

RG_FN double rg_nan_() { return RG_DBL(0xfff8000000000000ULL); }   // (x-x)/0 of __math_invalid on x86 (default NaN)

// specialcase() of e_exp.c / e_pow.c: the scale factor would over- or underflow.  `neg_one` is pow's signed variant.
RG_FN double rg_exp_special_(double tmp, uint64_t sbits, uint64_t ki, int signed_scale) {
  if ((ki & 0x80000000ULL) == 0) {   // k > 0
    sbits -= 1009ULL << 52;
    const double scale = RG_DBL(sbits);
    return RG_MUL(0x1p1009, RG_FMA(scale, tmp, scale));
  }
  sbits += 1022ULL << 52;
  const double scale = RG_DBL(sbits);
  const double st = RG_MUL(tmp, scale);       // scale * tmp is formed once and shared (not contracted here)
  double y = RG_ADD(scale, st);
  const double ay = signed_scale ? RG_DBL(RG_BITS(y) & 0x7fffffffffffffffULL) : y;
  if (ay < 1.0) {
    const double one = (signed_scale && y < 0.0) ? -1.0 : 1.0;
    double lo = RG_ADD(RG_SUB(scale, y), st);
    const double hi = RG_ADD(y, one);
    lo = RG_ADD(RG_ADD(RG_SUB(one, hi), y), lo);
    y = RG_SUB(RG_ADD(lo, hi), one);
    if (y == 0.0) y = signed_scale ? RG_DBL(sbits & 0x8000000000000000ULL) : 0.0;
  }
  return RG_MUL(y, 0x1p-1022);
}

// exp_inline() of e_pow.c (xtail != 0 path) and the body of __exp (has_tail = 0); sign_bias is pow's.
RG_FN double rg_exp_core_(double x, double xtail, uint64_t sign_bias, int has_tail) {
  uint32_t abstop = (uint32_t)(RG_BITS(x) >> 52) & 0x7ff;
  if (abstop - 0x3c9u > 0x3eu) {
    if ((int32_t)(abstop - 0x3c9u) < 0) {   // |x| < 2^-54: 1 + x (sign from the bias)
      const double one = RG_ADD(x, 1.0);
      return (has_tail && sign_bias) ? -one : one;
    }
    if (abstop > 0x408u) {                   // |x| >= 1024
      if (!has_tail) {
        if (RG_BITS(x) == 0xfff0000000000000ULL) return 0.0;
        if (abstop == 0x7ffu) return RG_ADD(x, 1.0);
      }
      if (RG_BITS(x) >> 63) return sign_bias ? -0.0 : 0.0;                         // __math_uflow
      return sign_bias ? RG_DBL(0xfff0000000000000ULL) : RG_DBL(RG_INF);           // __math_oflow
    }
    abstop = 0;
  }
  const double shift = RG_EXPC(1);
  double kd = RG_FMA(x, RG_EXPC(0), shift);            // z = InvLn2N * x; kd = z + Shift
  const uint64_t ki = RG_BITS(kd);
  kd = RG_SUB(kd, shift);
  double r = RG_FMA(kd, RG_EXPC(2), x);                // r = x + kd * NegLn2hiN
  r = RG_FMA(kd, RG_EXPC(3), r);                       //       + kd * NegLn2loN
  if (has_tail) r = RG_ADD(xtail, r);
  const uint32_t idx = 2u * (uint32_t)(ki & 127u);
  const uint64_t top = (ki + sign_bias) << 45;
  const double tail = RG_DBL(RG_LD(&rvn_glibc_exp_tab[idx]));
  const uint64_t sbits = RG_LD(&rvn_glibc_exp_tab[idx + 1]) + top;
  const double p23 = RG_FMA(r, RG_EXPC(5), RG_EXPC(4));   // C2 + r * C3
  const double tr = RG_ADD(r, tail);                      // tail + r
  const double r2 = RG_MUL(r, r);
  const double p45 = RG_FMA(r, RG_EXPC(7), RG_EXPC(6));   // C4 + r * C5
  double tmp = RG_FMA(p23, r2, tr);
  tmp = RG_FMA(RG_MUL(r2, r2), p45, tmp);                 // ... + r2 * r2 * (C4 + r * C5)
  if (abstop == 0) return rg_exp_special_(tmp, sbits, ki, has_tail);
  const double scale = RG_DBL(sbits);
  return RG_FMA(scale, tmp, scale);                       // scale + scale * tmp
}

RG_FN double rvn_exp(double x) { return rg_exp_core_(x, 0.0, 0, 0); }

RG_FN double rvn_log(double x) {
  uint64_t ix = RG_BITS(x);
  const uint32_t top = (uint32_t)(ix >> 48);
  if (ix - 0x3fee000000000000ULL <= 0x308ffffffffffULL) {   // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (ix == RG_ONE) return 0.0;
    const double r = RG_SUB(x, 1.0);
    const double b12 = RG_FMA(r, RG_LOGC(9), RG_LOGC(8));    // B[1] + r * B[2]
    const double b45 = RG_FMA(r, RG_LOGC(12), RG_LOGC(11));  // B[4] + r * B[5]
    const double r2 = RG_MUL(r, r);
    const double b78 = RG_FMA(r, RG_LOGC(15), RG_LOGC(14));  // B[7] + r * B[8]
    const double q1 = RG_FMA(r2, RG_LOGC(10), b12);          // + r2 * B[3]
    const double q2 = RG_FMA(r2, RG_LOGC(13), b45);          // + r2 * B[6]
    const double r3 = RG_MUL(r, r2);
    double p = RG_FMA(r2, RG_LOGC(16), b78);                 // + r2 * B[9]
    p = RG_FMA(r3, RG_LOGC(17), p);                          // + r3 * B[10]
    p = RG_FMA(p, r3, q2);
    p = RG_FMA(p, r3, q1);                                   // y = r3 * p, folded into the last fma below
    const double rw = RG_FMA(r, 0x1p27, r);                  // r + w,  w = r * 2^27
    const double rhi = RG_FMA(-0x1p27, r, rw);               // r + w - w
    const double b0 = RG_LOGC(7);                            // B[0] = -0.5
    const double rhi2 = RG_MUL(rhi, rhi);
    const double rlo = RG_SUB(r, rhi);
    const double hi = RG_FMA(rhi2, b0, r);                   // r + rhi * rhi * B[0]
    double lo = RG_FMA(rhi2, b0, RG_SUB(r, hi));             // r - hi + w
    lo = RG_FMA(RG_MUL(b0, rlo), RG_ADD(r, rhi), lo);        // + B[0] * rlo * (rhi + r)
    const double y = RG_FMA(p, r3, lo);
    return RG_ADD(hi, y);
  }
  if (top - 0x0010u > 0x7fdfu) {
    if (ix * 2 == 0) return RG_DBL(0xfff0000000000000ULL);   // -inf
    if (ix == RG_INF) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return (ix << 1) > (RG_INF << 1) ? RG_ADD(x, x) : rg_nan_();
    ix = RG_BITS(RG_MUL(x, 0x1p52));
    ix -= 52ULL << 52;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ULL;
  const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
  const double invc = RG_DBL(RG_LD(&rvn_glibc_log_tab[2 * i]));
  const double logc = RG_DBL(RG_LD(&rvn_glibc_log_tab[2 * i + 1]));
  const double z = RG_DBL(iz);
  const double kd = RG_I2D(k);
  const double w = RG_FMA(kd, RG_LOGC(0), logc);             // kd * Ln2hi + logc
  const double r = RG_FMA(z, invc, -1.0);
  const double a12 = RG_FMA(r, RG_LOGC(4), RG_LOGC(3));      // A[1] + r * A[2]
  const double hi = RG_ADD(r, w);
  const double r2 = RG_MUL(r, r);
  double lo = RG_ADD(RG_SUB(w, hi), r);
  lo = RG_FMA(kd, RG_LOGC(1), lo);                           // + kd * Ln2lo
  const double r3 = RG_MUL(r, r2);
  const double a34 = RG_FMA(r, RG_LOGC(6), RG_LOGC(5));      // A[3] + r * A[4]
  lo = RG_FMA(r2, RG_LOGC(2), lo);                           // lo + r2 * A[0]
  const double p = RG_FMA(a34, r2, a12);
  return RG_ADD(RG_FMA(r3, p, lo), hi);
}

RG_FN int rg_issignaling_(uint64_t i) { return 2 * (i ^ 0x0008000000000000ULL) > 2 * 0x7ff8000000000000ULL; }

// checkint() of e_pow.c: 0 = not an integer, 1 = odd, 2 = even
RG_FN int rg_checkint_(uint64_t iy) {
  const int e = (int)(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ULL << (0x3ff + 52 - e))) return 1;
  return 2;
}

RG_FN double rvn_pow(double x, double y) {
  uint64_t sign_bias = 0;
  uint64_t ix = RG_BITS(x);
  const uint64_t iy = RG_BITS(y);
  uint32_t topx = (uint32_t)(ix >> 52);
  const uint32_t topy = (uint32_t)(iy >> 52);
  if (topx - 1u > 0x7fdu || (topy & 0x7ffu) - 0x3beu > 0x7fu) {
    if (2 * iy - 1 >= 2 * RG_INF - 1) {   // y is 0, inf or nan
      if (2 * iy == 0) return rg_issignaling_(ix) ? RG_ADD(x, y) : 1.0;
      if (ix == RG_ONE) return rg_issignaling_(iy) ? RG_ADD(x, y) : 1.0;
      if (2 * ix > 2 * RG_INF || 2 * iy > 2 * RG_INF) return RG_ADD(x, y);
      if (2 * ix == 2 * RG_ONE) return 1.0;
      if ((2 * ix < 2 * RG_ONE) == !(iy >> 63)) return 0.0;
      return RG_MUL(y, y);
    }
    if (2 * ix - 1 >= 2 * RG_INF - 1) {   // x is 0, inf or nan
      double x2 = RG_MUL(x, x);
      if ((ix >> 63) && rg_checkint_(iy) == 1) { x2 = -x2; sign_bias = 1; }
      if (2 * ix == 0 && (iy >> 63)) return sign_bias ? RG_DBL(0xfff0000000000000ULL) : RG_DBL(RG_INF);
      return (iy >> 63) ? RG_DIV(1.0, x2) : x2;
    }
    if (ix >> 63) {                       // finite x < 0
      const int yint = rg_checkint_(iy);
      if (yint == 0) return rg_nan_();
      if (yint == 1) sign_bias = 0x40000;
      ix &= 0x7fffffffffffffffULL;
      topx &= 0x7ff;
    }
    if ((topy & 0x7ffu) - 0x3beu > 0x7fu) {
      if (ix == RG_ONE) return 1.0;
      if ((topy & 0x7ffu) < 0x3beu) return ix > RG_ONE ? RG_ADD(y, 1.0) : RG_SUB(1.0, y);
      if ((ix > RG_ONE) == (topy < 0x800u)) return RG_DBL(RG_INF);
      return 0.0;
    }
    if (topx == 0) {
      ix = RG_BITS(RG_MUL(x, 0x1p52));
      ix &= 0x7fffffffffffffffULL;
      ix -= 52ULL << 52;
    }
  }
  // log_inline(): hi + lo ~ log(x) to 68 bits
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
  const double z = RG_DBL(iz);
  const double kd = RG_I2D(k);
  const double invc = RG_DBL(RG_LD(&rvn_glibc_pow_tab[3 * i]));
  const double logc = RG_DBL(RG_LD(&rvn_glibc_pow_tab[3 * i + 1]));
  const double logctail = RG_DBL(RG_LD(&rvn_glibc_pow_tab[3 * i + 2]));
  const double t1 = RG_FMA(kd, RG_POWC(0), logc);           // kd * Ln2hi + logc
  const double lo1 = RG_FMA(kd, RG_POWC(1), logctail);      // kd * Ln2lo + logctail
  const double r = RG_FMA(z, invc, -1.0);
  const double ar = RG_MUL(r, RG_POWC(2));                  // A[0] * r
  const double q12 = RG_FMA(r, RG_POWC(4), RG_POWC(3));     // A[1] + r * A[2]
  const double q34 = RG_FMA(r, RG_POWC(6), RG_POWC(5));     // A[3] + r * A[4]
  const double t2 = RG_ADD(r, t1);
  const double lo2 = RG_ADD(RG_SUB(t1, t2), r);
  const double ar2 = RG_MUL(r, ar);
  const double ar3 = RG_MUL(r, ar2);
  const double lo3 = RG_FMA(ar, r, -ar2);
  const double hi0 = RG_ADD(t2, ar2);
  const double q56 = RG_FMA(r, RG_POWC(8), RG_POWC(7));     // A[5] + r * A[6]
  const double lo4 = RG_ADD(RG_SUB(t2, hi0), ar2);
  const double pp = RG_FMA(ar2, RG_FMA(q56, ar2, q34), q12);
  double lo = RG_ADD(RG_ADD(RG_ADD(lo1, lo2), lo3), lo4);
  lo = RG_FMA(ar3, pp, lo);                                 // lo1 + lo2 + lo3 + lo4 + p
  const double hi = RG_ADD(hi0, lo);
  const double tail = RG_ADD(RG_SUB(hi0, hi), lo);
  const double ehi = RG_MUL(y, hi);
  const double elo = RG_FMA(y, tail, RG_FMA(hi, y, -ehi));  // y * lo + fma(y, hi, -ehi)
  return rg_exp_core_(ehi, elo, sign_bias, 1);
}

// ---- expm1 / tanh: fdlibm's s_expm1.c and s_tanh.c as glibc ships them (sysdeps/ieee754/dbl-64), expm1 in the `__expm1_fma`
// ifunc variant (a*b+c contracted where the compiler chose to; checked against the instruction sequence of this image's libm).
RG_FN double rg_set_high_(double y, uint32_t hi) { return RG_DBL(((uint64_t)hi << 32) | (RG_BITS(y) & 0xffffffffULL)); }
RG_FN double rvn_expm1(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
  const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05,
               Q4 = 4.00821782732936239552e-06, Q5 = -2.01099218183624371326e-07;
  const uint64_t ix = RG_BITS(x);
  uint32_t hx = (uint32_t)(ix >> 32);
  const uint32_t xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  double hi, lo, c = 0.0, t;
  int k;
  if (hx >= 0x4043687Au) {                 // |x| >= 56 ln2
    if (hx >= 0x40862E42u) {               // |x| >= 709.78
      if (hx >= 0x7ff00000u) {
        if (((hx & 0xfffffu) | (uint32_t)ix) != 0) return RG_ADD(x, x);   // NaN
        return (xsb == 0) ? x : -1.0;
      }
      if (x > 7.09782712893383973096e+02) return RG_DBL(RG_INF);          // huge * huge
    }
    if (xsb != 0) return RG_SUB(1.0e-300, 1.0);                           // tiny - one
  }
  if (hx > 0x3fd62e42u) {                  // |x| > 0.5 ln2
    if (hx < 0x3FF0A2B2u) {                // and |x| < 1.5 ln2
      if (xsb == 0) { hi = RG_SUB(x, ln2_hi); lo = ln2_lo; k = 1; }
      else { hi = RG_ADD(x, ln2_hi); lo = -ln2_lo; k = -1; }
    } else {
      k = (int)RG_ADD((xsb == 0) ? 0.5 : -0.5, RG_MUL(x, invln2));
      t = RG_I2D(k);
      hi = RG_FMA(-t, ln2_hi, x);          // x - t * ln2_hi (the product is exact)
      lo = RG_MUL(t, ln2_lo);
    }
    x = RG_SUB(hi, lo);
    c = RG_SUB(RG_SUB(hi, x), lo);
  } else if (hx < 0x3c900000u) {           // |x| < 2^-54
    return x;
  } else k = 0;
  const double hfx = RG_MUL(x, 0.5);
  const double hxs = RG_MUL(x, hfx);
  const double R2 = RG_FMA(hxs, Q3, Q2);
  const double R3 = RG_FMA(hxs, Q5, Q4);
  const double h2 = RG_MUL(hxs, hxs);
  const double R1 = RG_FMA(hxs, Q1, 1.0);
  const double h4 = RG_MUL(h2, h2);
  const double r1 = RG_FMA(h4, R3, RG_FMA(h2, R2, R1));
  t = RG_FMA(-r1, hfx, 3.0);                                   // 3.0 - r1 * hfx
  double e = RG_MUL(RG_DIV(RG_SUB(r1, t), RG_FMA(-x, t, 6.0)), hxs);   // hxs * ((r1 - t) / (6.0 - x * t))
  if (k == 0) return RG_SUB(x, RG_FMA(e, x, -hxs));            // x - (x * e - hxs)
  e = RG_FMA(RG_SUB(e, c), x, -c);                             // x * (e - c) - c
  e = RG_SUB(e, hxs);
  if (k == -1) return RG_FMA(0.5, RG_SUB(x, e), -0.5);         // 0.5 * (x - e) - 0.5
  if (k == 1) {
    if (x < -0.25) return RG_MUL(RG_SUB(e, RG_ADD(x, 0.5)), -2.0);
    return RG_FMA(RG_SUB(x, e), 2.0, 1.0);                     // one + 2.0 * (x - e)
  }
  double y;
  if (k <= -2 || k > 56) {
    y = RG_SUB(1.0, RG_SUB(e, x));
    y = rg_set_high_(y, (uint32_t)(RG_BITS(y) >> 32) + ((uint32_t)k << 20));
    return RG_SUB(y, 1.0);
  }
  if (k < 20) {
    t = RG_DBL((uint64_t)(0x3ff00000u - (0x200000u >> k)) << 32);   // 1 - 2^-k
    y = RG_SUB(t, RG_SUB(e, x));
  } else {
    t = RG_DBL((uint64_t)((uint32_t)(0x3ff - k) << 20) << 32);      // 2^-k
    y = RG_SUB(x, RG_ADD(e, t));
    y = RG_ADD(y, 1.0);
  }
  return rg_set_high_(y, (uint32_t)(RG_BITS(y) >> 32) + ((uint32_t)k << 20));
}

RG_FN double rvn_tanh(double x) {
  const uint64_t bits = RG_BITS(x);
  const int32_t jx = (int32_t)(bits >> 32);
  const uint32_t ix = (uint32_t)jx & 0x7fffffffu;
  double z;
  if (ix >= 0x7ff00000u) return (jx >= 0) ? RG_ADD(RG_DIV(1.0, x), 1.0) : RG_SUB(RG_DIV(1.0, x), 1.0);   // inf / NaN
  if (ix < 0x40360000u) {                          // |x| < 22
    if ((ix | (uint32_t)bits) == 0) return x;      // +-0
    if (ix < 0x3c800000u) return RG_MUL(x, RG_ADD(1.0, x));   // |x| < 2^-55
    const double ax = RG_DBL(bits & 0x7fffffffffffffffULL);
    if (ix >= 0x3ff00000u) {                       // |x| >= 1
      const double t = rvn_expm1(RG_ADD(ax, ax));
      z = RG_SUB(1.0, RG_DIV(2.0, RG_ADD(t, 2.0)));
    } else {
      const double t = rvn_expm1(RG_MUL(ax, -2.0));
      z = RG_DIV(-t, RG_ADD(t, 2.0));
    }
  } else z = RG_SUB(1.0, 1.0e-300);                // |x| >= 22
  return (jx >= 0) ? z : -z;
}

