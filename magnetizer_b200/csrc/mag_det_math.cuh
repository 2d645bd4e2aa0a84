// mag_det_math.cuh -- deterministic elementary functions for the B200 dynamo kernels:
// log, exp, pow and the modified Bessel functions I0, I1, K0, K1 with IEEE-754 +,-,*,/ and
// sqrt only, fixed operation order, no fused multiply-add (this translation unit is
// compiled with --fmad=false; the hot loop uses explicit fma()).
//
// The reference gets these from libm and GSL (source/bessel_functions.f90:33-69,
// source/rotationCurves.f90:75-82,203-215, source/surface_density.f90:37,82,
// source/profiles.f90:321, source/random.f90:104), whose versions it does not pin, while
// its results depend on their last bit: GSL's Brent solver (source/root_finder.f90:171-184)
// returns either a converged root or a bisection midpoint (5e-8 off) depending on the sign
// of the residual at rounding-noise level, and the 6th-order derivative of Omega amplifies
// that to 1e-6 in the shear and the scale height.  A 1e-8 parity statement therefore needs
// elementary functions that agree bit for bit with the checker's; the block between the
// TWIN markers is the same text as in the CPU oracle (tests/test_det_math_twin.py).
// log/exp use the classic fdlibm argument reductions and minimax polynomials, restated.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mag_bessel_coeffs.cuh"

#define DET_FN __device__ __forceinline__
// log, exp and the Bessel series are called from ~45 / ~36 / ~10 places of k_profiles.  Inlined
// they make 770 KB of code and the profile phase instruction-cache bound (round 1: 2.7
// stall_no_instruction per issue); out of line the results are bit-identical (same arithmetic,
// --fmad=false) and the profile phase is 25 % faster (round 2: 84 vs 111 ms per 12 500 galaxies,
// profiles/r2_bench.md).  -DMAG_DET_INLINE restores the inlined build for A/B runs.
#ifdef MAG_DET_INLINE
#define DET_BIG __device__ __forceinline__
#else
#define DET_BIG __device__ __noinline__
#endif

namespace detmath {

DET_FN int32_t det_hi(double x) { return __double2hiint(x); }
DET_FN uint32_t det_lo(double x) { return (uint32_t)__double2loint(x); }
DET_FN double det_words(int32_t hi, uint32_t lo) { return __hiloint2double(hi, (int)lo); }
DET_FN double det_sqrt(double x) { return sqrt(x); }
DET_FN double det_inf() { return __hiloint2double(0x7ff00000, 0); }

// ==== BEGIN TWIN (identical text in oracle/det_math.hpp) ====
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
