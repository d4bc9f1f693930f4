/*
 * fa_jsmath.h -- deterministic double-precision log / log10 / pow with JavaScript (V8) semantics.
 *
 * Why this exists: the reference's noise gate turns the adaptive maximum `y` into the integer
 * threshold `v` through  parseInt(Math.pow(10, Math.log10(y) - 2) / 2)  and friends
 * (/root/reference/dist/main.js:2@B28506, function C).  For y a multiple of 200 the result sits
 * exactly on a parseInt boundary, so the last bit of log10/pow decides `v`, and `v` decides segment
 * boundaries.  V8 implements Math.log10 / Math.log / Math.pow with its fdlibm port
 * (v8/src/base/ieee754.cc; third-party, not under /root/reference).  This header restates that
 * published algorithm (Sun fdlibm e_log.c, e_log10.c, e_pow.c) using only IEEE-754 +,-,*,/ and bit
 * manipulation, so that the host oracle, the host library and the sm_100a kernels produce the same
 * bits.  It must be compiled without floating-point contraction (gcc: -ffp-contract=off,
 * nvcc: --fmad=false); tests/test_jsmath.py checks it against glibc (<= 1 ulp) and on exact cases.
 *
 * Usable from C, C++ and CUDA device code.
 *
 * The algorithms (argument reduction, polynomial coefficients, the hi / lo splitting of the constants) are those of FDLIBM's
 * e_log.c, e_log10.c and e_pow.c, whose notice is preserved here as it asks:
 *
 * ====================================================
 * Copyright (C) 1993, 2004 by Sun Microsystems, Inc. All rights reserved.
 *
 * Developed at SunSoft / SunPro, a Sun Microsystems, Inc. business.
 * Permission to use, copy, modify, and distribute this
 * software is freely granted, provided that this notice
 * is preserved.
 * ====================================================
 *
 * How far this pins the reference's Math.log10 / Math.pow: tests/test_gate_enumeration.py (exhaustive over every reachable y of
 * the noise gate).
 */
#ifndef FA_JSMATH_H_
#define FA_JSMATH_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#define FA_HD_NOINLINE static __host__ __device__
#else
#define FA_HD static inline
#define FA_HD_NOINLINE static
#endif

FA_HD int32_t fa_hi_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32);
#endif
}
FA_HD uint32_t fa_lo_word(double x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}
FA_HD double fa_from_words(int32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, (int)lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x;
#endif
}
FA_HD double fa_set_hi(double x, int32_t hi) { return fa_from_words(hi, fa_lo_word(x)); }
FA_HD double fa_clear_lo(double x) { return fa_from_words(fa_hi_word(x), 0u); }

/* IEEE sqrt: correctly rounded on both sides. */
FA_HD double fa_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}

/* Math.log(x): fdlibm __ieee754_log. */
FA_HD_NOINLINE double fa_js_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               two54 = 1.80143985094819840000e+16,
               Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  const double zero = 0.0;
  double hfsq, f, s, z, R, w, t1, t2, dk;
  int32_t k, hx, i, j;
  uint32_t lx;
  hx = fa_hi_word(x);
  lx = fa_lo_word(x);
  k = 0;
  if (hx < 0x00100000) { /* x < 2**-1022 */
    if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero; /* log(+-0) = -inf */
    if (hx < 0) return (x - x) / zero;                       /* log(-#) = NaN */
    k -= 54;
    x *= two54;
    hx = fa_hi_word(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  i = (hx + 0x95f64) & 0x100000;
  x = fa_set_hi(x, hx | (i ^ 0x3ff00000)); /* normalize x or x/2 */
  k += (i >> 20);
  f = x - 1.0;
  if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2**-20 */
    if (f == zero) {
      if (k == 0) return zero;
      dk = (double)k;
      return dk * ln2_hi + dk * ln2_lo;
    }
    R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  s = f / (2.0 + f);
  dk = (double)k;
  z = s * s;
  i = hx - 0x6147a;
  w = z * z;
  j = 0x6b851 - hx;
  t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  i |= j;
  R = t2 + t1;
  if (i > 0) {
    hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

/* Math.log10(x): fdlibm __ieee754_log10 (the variant V8 ships). */
FA_HD_NOINLINE double fa_js_log10(double x) {
  const double two54 = 1.80143985094819840000e+16, ivln10 = 4.34294481903251816668e-01,
               log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
  const double zero = 0.0;
  double y, z;
  int32_t i, k, hx;
  uint32_t lx;
  hx = fa_hi_word(x);
  lx = fa_lo_word(x);
  k = 0;
  if (hx < 0x00100000) {
    if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero;
    if (hx < 0) return (x - x) / zero;
    k -= 54;
    x *= two54;
    hx = fa_hi_word(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
  hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
  y = (double)(k + i);
  x = fa_set_hi(x, hx);
  z = y * log10_2lo + ivln10 * fa_js_log(x);
  return z + y * log10_2hi;
}

FA_HD double fa_scalbn_small(double z, int n) {
  /* only reached for subnormal results of pow; exact power-of-two scaling in two safe steps */
  double r = z;
  while (n < -1000) { r *= fa_from_words(0x3ff00000 - (1000 << 20), 0); n += 1000; }
  while (n > 1000) { r *= fa_from_words(0x3ff00000 + (1000 << 20), 0); n -= 1000; }
  return r * fa_from_words(0x3ff00000 + (n << 20), 0);
}

/* Math.pow(x, y): fdlibm __ieee754_pow (V8 base::ieee754::pow). */
FA_HD_NOINLINE double fa_js_pow(double x, double y) {
  const double dp_h1 = 5.84962487220764160156e-01, dp_l1 = 1.35003920212974897128e-08;
  const double zero = 0.0, one = 1.0, two = 2.0, two53 = 9007199254740992.0, huge = 1.0e300,
               tiny = 1.0e-300,
               L1 = 5.99999999999994648725e-01, L2 = 4.28571428578550184252e-01,
               L3 = 3.33333329818377432918e-01, L4 = 2.72728123808534006489e-01,
               L5 = 2.30660745775561754067e-01, L6 = 2.06975017800338417784e-01,
               P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
               P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
               P5 = 4.13813679705723846039e-08, lg2 = 6.93147180559945286227e-01,
               lg2_h = 6.93147182464599609375e-01, lg2_l = -1.90465429995776804525e-09,
               ovt = 8.0085662595372944372e-0017, cp = 9.61796693925975554329e-01,
               cp_h = 9.61796700954437255859e-01, cp_l = -7.02846165095275826516e-09,
               ivln2 = 1.44269504088896338700e+00, ivln2_h = 1.44269502162933349609e+00,
               ivln2_l = 1.92596299112661746887e-08;
  double z, ax, z_h, z_l, p_h, p_l;
  double y1, t1, t2, r, s, t, u, v, w;
  int32_t i, j, k, yisint, n;
  int32_t hx, hy, ix, iy;
  uint32_t lx, ly;

  hx = fa_hi_word(x); lx = fa_lo_word(x);
  hy = fa_hi_word(y); ly = fa_lo_word(y);
  ix = hx & 0x7fffffff;
  iy = hy & 0x7fffffff;

  if ((iy | ly) == 0) return one; /* y == 0 */
  /* +-NaN return x+y */
  if (ix > 0x7ff00000 || ((ix == 0x7ff00000) && (lx != 0)) || iy > 0x7ff00000 ||
      ((iy == 0x7ff00000) && (ly != 0)))
    return x + y;

  /* determine if y is an odd int when x < 0 */
  yisint = 0;
  if (hx < 0) {
    if (iy >= 0x43400000) yisint = 2;
    else if (iy >= 0x3ff00000) {
      k = (iy >> 20) - 0x3ff;
      if (k > 20) {
        j = (int32_t)(ly >> (52 - k));
        if (((uint32_t)j << (52 - k)) == ly) yisint = 2 - (j & 1);
      } else if (ly == 0) {
        j = iy >> (20 - k);
        if ((j << (20 - k)) == iy) yisint = 2 - (j & 1);
      }
    }
  }

  if (ly == 0) {
    if (iy == 0x7ff00000) { /* y is +-inf */
      if (((ix - 0x3ff00000) | lx) == 0) return y - y; /* JS: (+-1)**+-inf is NaN */
      else if (ix >= 0x3ff00000) return (hy >= 0) ? y : zero;
      else return (hy < 0) ? -y : zero;
    }
    if (iy == 0x3ff00000) { /* y is +-1 */
      if (hy < 0) return one / x;
      return x;
    }
    if (hy == 0x40000000) return x * x; /* y is 2 */
    if (hy == 0x3fe00000) {             /* y is 0.5 */
      if (hx >= 0) return fa_sqrt(x);
    }
  }

  ax = x < 0 ? -x : x;
  if (hx < 0 && ix == 0 && lx == 0) ax = 0.0; /* fabs(-0) */
  if (lx == 0) {
    if (ix == 0x7ff00000 || ix == 0 || ix == 0x3ff00000) {
      z = ax;
      if (hy < 0) z = one / z;
      if (hx < 0) {
        if (((ix - 0x3ff00000) | yisint) == 0) z = (z - z) / (z - z);
        else if (yisint == 1) z = -z;
      }
      return z;
    }
  }

  n = (int32_t)((uint32_t)hx >> 31) - 1; /* 0 for x<0, -1 for x>=0 (FreeBSD form) */
  if ((n | yisint) == 0) return (x - x) / (x - x); /* (x<0)**(non-int) is NaN */
  s = one;
  if ((n | (yisint - 1)) == 0) s = -one;

  if (iy > 0x41e00000) { /* |y| > 2**31 */
    if (iy > 0x43f00000) {
      if (ix <= 0x3fefffff) return (hy < 0) ? huge * huge : tiny * tiny;
      if (ix >= 0x3ff00000) return (hy > 0) ? huge * huge : tiny * tiny;
    }
    if (ix < 0x3fefffff) return (hy < 0) ? s * huge * huge : s * tiny * tiny;
    if (ix > 0x3ff00000) return (hy > 0) ? s * huge * huge : s * tiny * tiny;
    t = ax - one;
    w = (t * t) * (0.5 - t * (0.3333333333333333333333 - t * 0.25));
    u = ivln2_h * t;
    v = t * ivln2_l - w * ivln2;
    t1 = fa_clear_lo(u + v);
    t2 = v - (t1 - u);
  } else {
    double ss, s2, s_h, s_l, t_h, t_l, bpk, dphk, dplk;
    n = 0;
    if (ix < 0x00100000) { ax *= two53; n -= 53; ix = fa_hi_word(ax); }
    n += ((ix) >> 20) - 0x3ff;
    j = ix & 0x000fffff;
    ix = j | 0x3ff00000;
    if (j <= 0x3988E) k = 0;
    else if (j < 0xBB67A) k = 1;
    else { k = 0; n += 1; ix -= 0x00100000; }
    ax = fa_set_hi(ax, ix);
    bpk = k ? 1.5 : 1.0;
    dphk = k ? dp_h1 : 0.0;
    dplk = k ? dp_l1 : 0.0;

    u = ax - bpk;
    v = one / (ax + bpk);
    ss = u * v;
    s_h = fa_clear_lo(ss);
    t_h = fa_from_words(((ix >> 1) | 0x20000000) + 0x00080000 + (k << 18), 0);
    t_l = ax - (t_h - bpk);
    s_l = v * ((u - s_h * t_h) - s_h * t_l);
    s2 = ss * ss;
    r = s2 * s2 * (L1 + s2 * (L2 + s2 * (L3 + s2 * (L4 + s2 * (L5 + s2 * L6)))));
    r += s_l * (s_h + ss);
    s2 = s_h * s_h;
    t_h = fa_clear_lo(3.0 + s2 + r);
    t_l = r - ((t_h - 3.0) - s2);
    u = s_h * t_h;
    v = s_l * t_h + t_l * ss;
    p_h = fa_clear_lo(u + v);
    p_l = v - (p_h - u);
    z_h = cp_h * p_h;
    z_l = cp_l * p_h + p_l * cp + dplk;
    t = (double)n;
    t1 = fa_clear_lo(((z_h + z_l) + dphk) + t);
    t2 = z_l - (((t1 - t) - dphk) - z_h);
  }

  y1 = fa_clear_lo(y);
  p_l = (y - y1) * t1 + y * t2;
  p_h = y1 * t1;
  z = p_l + p_h;
  j = fa_hi_word(z);
  i = (int32_t)fa_lo_word(z);
  if (j >= 0x40900000) {
    if (((j - 0x40900000) | i) != 0) return s * huge * huge;
    if (p_l + ovt > z - p_h) return s * huge * huge;
  } else if ((j & 0x7fffffff) >= 0x4090cc00) {
    if (((j - (int32_t)0xc090cc00) | i) != 0) return s * tiny * tiny;
    if (p_l <= z - p_h) return s * tiny * tiny;
  }
  i = j & 0x7fffffff;
  k = (i >> 20) - 0x3ff;
  n = 0;
  if (i > 0x3fe00000) {
    n = j + (0x00100000 >> (k + 1));
    k = ((n & 0x7fffffff) >> 20) - 0x3ff;
    t = fa_from_words(n & ~(0x000fffff >> k), 0);
    n = ((n & 0x000fffff) | 0x00100000) >> (20 - k);
    if (j < 0) n = -n;
    p_h -= t;
  }
  t = fa_clear_lo(p_l + p_h);
  u = t * lg2_h;
  v = (p_l - (t - p_h)) * lg2 + t * lg2_l;
  z = u + v;
  w = v - (z - u);
  t = z * z;
  t1 = z - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  r = (z * t1) / (t1 - two) - (w + z * w);
  z = one - (r - z);
  j = fa_hi_word(z);
  j += (int32_t)((uint32_t)n << 20);
  if ((j >> 20) <= 0) z = fa_scalbn_small(z, n);
  else z = fa_set_hi(z, j);
  return s * z;
}

/* JS parseInt(x) for a finite double whose decimal form has no exponent (|x| < 1e21): truncation. */
FA_HD double fa_js_parse_int(double x) {
  if (x != x) return x;
#if defined(__CUDA_ARCH__)
  return trunc(x);
#else
  return __builtin_trunc(x);
#endif
}

#endif /* FA_JSMATH_H_ */
