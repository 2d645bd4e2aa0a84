// det_math.hpp -- TEST INFRASTRUCTURE (CPU oracle), NOT PRODUCT CODE.
//
// Deterministic elementary functions for the oracle: log, exp, pow and the modified
// Bessel functions I0, I1, K0, K1, written with IEEE-754 +,-,*,/ and sqrt only, in a
// fixed operation order and without fused multiply-add.
//
// Why: the reference calls libm (log, exp, **) and GSL (gsl_sf_bessel_*; call sites
// source/bessel_functions.f90:33-69, source/rotationCurves.f90:75-82,203-215,
// source/surface_density.f90:37,82, source/profiles.f90:321, source/random.f90:104) with
// UNPINNED library versions, and its results depend on their last-bit behaviour: GSL's
// Brent solver returns a bisection midpoint (error up to 5e-8 relative) or a converged
// root depending on the SIGN of the residual at rounding-noise level
// (source/root_finder.f90:171-184 with gsl roots/brent.c), and xder amplifies that into
// 1e-6 differences of the shear, the scale height and finally B.  Two faithful builds of
// the reference (e.g. with and without FMA contraction) therefore differ by up to 1e-5.
// A 1e-8 parity statement is only meaningful between implementations whose elementary
// functions agree bit for bit, so the oracle pins them here (<= 1 ulp for log/exp, a few
// ulp for pow, <= 2e-14 relative for Bessel; checked against the host libm / scipy in
// tests/test_oracle_numerics.py) and the CUDA path carries a line-for-line twin
// (magnetizer_b200/csrc/mag_det_math.cuh; tests/test_det_math_twin.py checks the two
// texts agree, a GPU test checks the values agree bit for bit).
//
// log/exp follow the classic argument reductions and minimax polynomials of Sun's fdlibm
// (e_log.c / e_exp.c), restated.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "bessel_coeffs.hpp"

#define DET_FN static inline
#define DET_BIG static inline   /* the large functions; the CUDA twin can build them out of line */
#define DET_CONST static const

namespace detmath {

DET_FN int32_t det_hi(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
DET_FN uint32_t det_lo(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (uint32_t)u; }
DET_FN double det_words(int32_t hi, uint32_t lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | lo;
  double x;
  std::memcpy(&x, &u, 8);
  return x;
}
DET_FN double det_sqrt(double x) { return std::sqrt(x); }
DET_FN double det_inf() { return det_words(0x7ff00000, 0u); }

// ==== BEGIN TWIN (identical text in magnetizer_b200/csrc/mag_det_math.cuh) ====
// natural logarithm, x > 0 finite (0 -> -inf, negative -> NaN)
DET_BIG double det_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01;
  const double Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01;
  const double Lg7 = 1.479819860511658591e-01;
  int32_t hx = det_hi(x);
  uint32_t lx = det_lo(x);
  int32_t k = 0;
  if (hx < 0x00100000) {
    if (((hx & 0x7fffffff) | (int32_t)lx) == 0) return -det_inf();
    if (hx < 0) return (x - x) / (x - x);
    k -= 54;
    x *= 1.80143985094819840000e+16;
    hx = det_hi(x);
    lx = det_lo(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int32_t i = (hx + 0x95f64) & 0x100000;
  x = det_words(hx | (i ^ 0x3ff00000), lx);
  k += (i >> 20);
  const double f = x - 1.0;
  const double dk = (double)k;
  if ((0x000fffff & (2 + hx)) < 3) {
    if (f == 0.0) {
      if (k == 0) return 0.0;
      return dk * ln2_hi + dk * ln2_lo;
    }
    const double R0 = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R0;
    return dk * ln2_hi - ((R0 - dk * ln2_lo) - f);
  }
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  const double R = t2 + t1;
  const int32_t ii = (hx - 0x6147a) | (0x6b851 - hx);
  if (ii > 0) {
    const double hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

// exponential
DET_BIG double det_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10;
  const double invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05;
  const double P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double twom1000 = 9.33263618503218878990e-302;
  int32_t hx = det_hi(x);
  const int32_t xsb = (int32_t)(((uint32_t)hx >> 31) & 1u);
  hx &= 0x7fffffff;
  if (hx >= 0x40862E42) {
    if (hx >= 0x7ff00000) {
      if (((hx & 0xfffff) | (int32_t)det_lo(x)) != 0) return x + x;
      return (xsb == 0) ? x : 0.0;
    }
    if (x > 7.09782712893383973096e+02) return det_inf();
    if (x < -7.45133219101941108420e+02) return 0.0;
  }
  double hi = 0.0, lo = 0.0;
  int32_t k = 0;
  if (hx > 0x3fd62e42) {
    if (hx < 0x3FF0A2B2) {
      if (xsb == 0) { hi = x - ln2HI; lo = ln2LO; k = 1; } else { hi = x + ln2HI; lo = -ln2LO; k = -1; }
    } else {
      k = (int32_t)(invln2 * x + (xsb == 0 ? 0.5 : -0.5));
      const double t = (double)k;
      hi = x - t * ln2HI;
      lo = t * ln2LO;
    }
    x = hi - lo;
  } else if (hx < 0x3e300000) {
    return 1.0 + x;
  }
  const double t = x * x;
  const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
  const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  if (k >= -1021) return det_words(det_hi(y) + (k << 20), det_lo(y));
  return det_words(det_hi(y) + ((k + 1000) << 20), det_lo(y)) * twom1000;
}

// x**y for x >= 0 (the only use on the path: positive bases, real exponents)
DET_FN double det_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (x == 0.0) return (y > 0.0) ? 0.0 : det_inf();
  return det_exp(y * det_log(x));
}

// Clenshaw evaluation of sum_k c_k T_k(t)
DET_FN double det_cheb(const double* c, int n, double t) {
  double b1 = 0.0, b2 = 0.0;
  const double t2 = 2.0 * t;
  for (int k = n - 1; k >= 1; k--) {
    const double b0 = t2 * b1 + (c[k] - b2);
    b2 = b1;
    b1 = b0;
  }
  return t * b1 + (c[0] - b2);
}

// modified Bessel functions I0, I1, K0, K1 of x > 0, all four at once
DET_BIG void det_bessel_ik01(double x, double* i0, double* i1, double* k0, double* k1) {
  const double euler = 0.57721566490153286061;
  const double q = 0.25 * x * x;
  if (x <= 8.0) {
    double t0 = 1.0, s0 = 1.0, t1 = 1.0, s1 = 1.0;
    for (int k = 1; k < 40; k++) {
      t0 *= q * DET_INV_KK[k];
      t1 *= q * DET_INV_KK1[k];
      s0 += t0;
      s1 += t1;
      if (t0 < 1e-18 * s0) break;
    }
    *i0 = s0;
    *i1 = s1 * 0.5 * x;
  } else {
    const double t = 16.0 / x - 1.0;
    const double e = det_exp(x) / det_sqrt(x);
    *i0 = e * det_cheb(DET_CHEB_I0, DET_CHEB_I0_N, t);
    *i1 = e * det_cheb(DET_CHEB_I1, DET_CHEB_I1_N, t);
  }
  if (x <= 2.0) {
    const double lg = det_log(0.5 * x);
    double t0 = 1.0, H = 0.0, s0 = 0.0;
    double t1 = 1.0, psi1 = -euler, psi2 = 1.0 - euler, s1 = (psi1 + psi2);
    for (int k = 1; k < 25; k++) {
      t0 *= q * DET_INV_KK[k];
      H += DET_INV_K[k];
      s0 += t0 * H;
      t1 *= q * DET_INV_KK1[k];
      psi1 += DET_INV_K[k];
      psi2 += DET_INV_K[k + 1];
      s1 += (psi1 + psi2) * t1;
    }
    *k0 = -(lg + euler) * (*i0) + s0;
    *k1 = 1.0 / x + lg * (*i1) - 0.25 * x * s1;
  } else {
    const double t = 4.0 / x - 1.0;
    const double e = det_exp(-x) / det_sqrt(x);
    *k0 = e * det_cheb(DET_CHEB_K0, DET_CHEB_K0_N, t);
    *k1 = e * det_cheb(DET_CHEB_K1, DET_CHEB_K1_N, t);
  }
}
// ==== END TWIN ====

}  // namespace detmath
