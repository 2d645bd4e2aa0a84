// mag_device_math.cuh -- device-side numerics for the B200 dynamo kernels:
// physical constants and code units (reference source/constants.f90), the GSL root finders
// the reference drives from source/root_finder.f90 (Brent bracket solver, secant +
// central-difference fallback), libgfortran's RANDOM_NUMBER streams (source/random.f90)
// and warp-level reductions.  Elementary functions: mag_det_math.cuh.
//
// This translation unit is compiled with --fmad=false: every expression outside the
// explicit fma() calls of the hot loop is evaluated in source order with IEEE rounding,
// which is what makes the root-finder decisions reproducible (see mag_det_math.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "mag_det_math.cuh"

namespace magdev {

// ---------------------------------------------------------------- constants
// constants.f90:20-37.  MAG_PI is the reference's value, not pi.
#define MAG_PI 3.14156295358
#define MAG_KM_KPC 3.08567758e16
#define MAG_CM_KM 1.e5
#define MAG_MKG_G 1.e6
#define MAG_MP_G 1.67262158e-24
#define MAG_G_SI 6.673e-11
#define MAG_MSUN_SI 1.98892e30
#define MAG_KPC_SI (3.08567758135e16 * 1e3)
#define MAG_HMASS 1.67372e-24
#define MAG_DISK_SCALE_TO_HALFMASS (1.0 / 1.678346990)
#define MAG_C_U 0.25   // constants.f90:88, :90
#define MAG_C_A 1.0
#define MAG_R_IN 1.0e-7
#define MAG_NGHOST 3

struct Units {  // constants.f90:40-74 evaluated once on the host, passed by value
  double s_Gyr, t0_Gyr, t0_s, h0_km, h0_cm, n0_cm3;
};

__host__ __device__ inline Units make_units() {
  Units u;
  u.s_Gyr = 1.e9 * 365.25 * 24 * 3600;
  const double etat0_kmskpc = 1.0 / 3 * 0.1 * 10.0;
  const double t0_kpcskm = 1.0 / etat0_kmskpc;
  u.t0_Gyr = t0_kpcskm / u.s_Gyr * MAG_KM_KPC;
  u.t0_s = t0_kpcskm * MAG_KM_KPC;
  u.h0_km = MAG_KM_KPC;
  u.h0_cm = u.h0_km * MAG_CM_KM;
  u.n0_cm3 = (1.0 / MAG_MKG_G) * (1.0 / MAG_MKG_G) / (u.h0_cm * u.h0_cm) * (u.t0_s * u.t0_s) / MAG_MP_G;
  return u;
}

// ---------------------------------------------------------------- warp helpers
#define FULL 0xffffffffu
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// minloc with Fortran semantics (first minimum): (value, index) lexicographic
__device__ __forceinline__ void warp_argmin_first(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(FULL, v, o);
    int oi = __shfl_xor_sync(FULL, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
__device__ __forceinline__ void warp_argmax_first(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(FULL, v, o);
    int oi = __shfl_xor_sync(FULL, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

// ---------------------------------------------------------------- root finders
// gsl_root_fsolver_brent as driven by FindRoot (root_finder.f90:110-189; GSL
// roots/brent.c, roots/convergence.c): REL_TOL 1e-7, ABS_TOL 0, at most 200 iterations.
template <class F>
__device__ inline bool brent_root(F f, double x_lower, double x_upper, double* root_out) {
  const double DBL_EPS = 2.2204460492503131e-16;
  const double REL_TOL = 1.0e-7;
  if (x_lower > x_upper) return false;
  double fa = f(x_lower);
  if (!isfinite(fa)) return false;
  double fb = f(x_upper);
  if (!isfinite(fb)) return false;
  double a = x_lower, b = x_upper, c = x_upper, fc = fb;
  double d = x_upper - x_lower, e = d;
  if ((fa < 0.0 && fb < 0.0) || (fa > 0.0 && fb > 0.0)) return false;
  for (int it = 1;; it++) {
    double xl, xu, root;
    bool ac_equal = false;
    if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) {
      ac_equal = true;
      c = a; fc = fa; d = b - a; e = b - a;
    }
    if (fabs(fc) < fabs(fb)) {
      ac_equal = true;
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    const double tol = 0.5 * DBL_EPS * fabs(b);
    double m = 0.5 * (c - b);
    if (fb == 0) {
      root = b; xl = b; xu = b;
    } else if (fabs(m) <= tol) {
      root = b;
      if (b < c) { xl = b; xu = c; } else { xl = c; xu = b; }
    } else {
      if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
        d = m; e = m;
      } else {
        double p, q, r;
        const double s = fb / fa;
        if (ac_equal) {
          p = 2 * m * s;
          q = 1 - s;
        } else {
          q = fa / fc;
          r = fb / fc;
          p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
          q = (q - 1) * (r - 1) * (s - 1);
        }
        if (p > 0) q = -q; else p = -p;
        if (2 * p < fmin(3 * m * q - fabs(tol * q), fabs(e * q))) {
          e = d; d = p / q;
        } else {
          d = m; e = m;
        }
      }
      a = b; fa = fb;
      if (fabs(d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
      fb = f(b);
      if (!isfinite(fb)) return false;
      root = b;
      double cc = c;
      if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) cc = a;
      if (b < cc) { xl = b; xu = cc; } else { xl = cc; xu = b; }
    }
    if (it > 200) return false;
    *root_out = root;
    double min_abs;
    if ((xl > 0.0 && xu > 0.0) || (xl < 0.0 && xu < 0.0)) min_abs = fmin(fabs(xl), fabs(xu)); else min_abs = 0;
    if (fabs(xu - xl) < REL_TOL * min_abs) return true;
  }
}

// gsl_deriv_central (deriv/deriv.c)
template <class F>
__device__ inline void central_deriv_(F f, double x, double h, double* result, double* err_round, double* err_trunc) {
  const double DBL_EPS = 2.2204460492503131e-16;
  const double fm1 = f(x - h), fp1 = f(x + h), fmh = f(x - h / 2), fph = f(x + h / 2);
  const double r3 = 0.5 * (fp1 - fm1);
  const double r5 = (4.0 / 3.0) * (fph - fmh) - (1.0 / 3.0) * r3;
  const double e3 = (fabs(fp1) + fabs(fm1)) * DBL_EPS;
  const double e5 = 2.0 * (fabs(fph) + fabs(fmh)) * DBL_EPS + e3;
  const double dy = fmax(fabs(r3 / h), fabs(r5 / h)) * (fabs(x) / h) * DBL_EPS;
  *result = r5 / h;
  *err_trunc = fabs((r5 - r3) / h);
  *err_round = fabs(e5 / h) + dy;
}
template <class F>
__device__ inline double deriv_central(F f, double x, double h) {
  double r0, round, trunc;
  central_deriv_(f, x, h, &r0, &round, &trunc);
  double error = round + trunc;
  if (round < trunc && (round > 0 && trunc > 0)) {
    double ro, roundo, trunco;
    const double h_opt = h * detmath::det_pow(round / (2.0 * trunc), 1.0 / 3.0);
    central_deriv_(f, x, h_opt, &ro, &roundo, &trunco);
    const double error_opt = roundo + trunco;
    if (error_opt < error && fabs(ro - r0) < 4.0 * error) r0 = ro;
  }
  return r0;
}
// gsl_root_fdfsolver_secant as driven by FindRoot_deriv (root_finder.f90:191-276; GSL
// roots/secant.c): REL_TOL 1e-5, first derivative from gsl_deriv_central(step)
template <class F>
__device__ inline bool secant_root(F f, double guess, double step, double* root_out) {
  const double REL_TOL = 1.0e-5;
  double root = -1.7976931348623157e308;
  double x = guess;
  double fx = f(x);
  double df = deriv_central(f, x, step);
  bool success = true;
  for (int i = 1;; i++) {
    bool bad = false;
    if (fx == 0.0) {
    } else if (df == 0.0) {
      bad = true;
    } else {
      const double x_new = x - (fx / df);
      const double f_new = f(x_new);
      const double df_new = df * ((fx - f_new) / fx);
      x = x_new; fx = f_new; df = df_new;
      if (!isfinite(f_new) || !isfinite(df_new)) bad = true;
    }
    if (bad || i > 200) { success = false; break; }
    const double ri = root;
    root = x;
    if (fabs(root - ri) < REL_TOL * fabs(root) || root == ri) break;
  }
  *root_out = root;
  return success;
}

// ---------------------------------------------------------------- RNG
// libgfortran intrinsics/random.c as seeded by set_random_seed (random.f90:39-64): all
// PUT words equal.  The state lives in shared memory, one per warp; lane 0 advances it.
struct RngState {
  uint64_t s[16];
  int p;
  int kind;
};
__device__ __constant__ const uint64_t RNG_XOR_KEYS[16] = {
    0xbd0c5b6e50c2df49ULL, 0xd46061cd46e1df38ULL, 0xbb4f4d4ed6103544ULL, 0x114a583d0756ad39ULL,
    0x4b5ad8623d0aaab6ULL, 0x3f2ed7afbe0c0f21ULL, 0xdec83fd65f113445ULL, 0x3824f8fbc4f10d24ULL,
    0x5d9025af05878911ULL, 0x500bc46b540340e9ULL, 0x8bd53298e0d00530ULL, 0x57886e40a952e06aULL,
    0x926e76c88e31cdb6ULL, 0xbd0724dac0a3a5f9ULL, 0xc5c8981b858ab796ULL, 0xbb12ab2694c2b32cULL};

__device__ inline void rng_seed(RngState* r, int32_t igal, int32_t seed, int kind) {
  // default-integer arithmetic of gfortran wraps: seed**2 - igal**3
  uint32_t w = (uint32_t)seed * (uint32_t)seed - (uint32_t)igal * (uint32_t)igal * (uint32_t)igal;
  if (w == 0u) {  // where (seed==0) seed = p_random_seed**igal
    uint32_t acc = 1u, b = (uint32_t)seed, e = (uint32_t)igal;
    while (e) {
      if (e & 1u) acc *= b;
      b *= b;
      e >>= 1;
    }
    w = acc;
  }
  const uint64_t word = (uint64_t)w | ((uint64_t)w << 32);
  const int n = kind == 0 ? 4 : 16;
  for (int i = 0; i < n; i++) r->s[i] = word ^ RNG_XOR_KEYS[i];
  r->p = kind == 0 ? 0 : (int)(w & 15u);
  r->kind = kind;
}
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ inline double rng_uniform_lane0(RngState* r) {
  uint64_t v;
  if (r->kind == 0) {  // xoshiro256** (GCC >= 9)
    uint64_t* s = r->s;
    v = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
  } else {  // xorshift1024* (GCC 7/8)
    int p = r->p;
    const uint64_t s0 = r->s[p];
    uint64_t s1 = r->s[p = (p + 1) & 15];
    s1 ^= s1 << 31;
    r->s[p] = s1 ^ s0 ^ (s1 >> 11) ^ (s0 >> 30);
    r->p = p;
    v = r->s[p] * 1181783497276652981ULL;
  }
  v &= ~(uint64_t)0 << 11;
  return (double)v * 5.421010862427522170e-20;  // 2^-64
}
// random_normal (random.f90:68-110), lane 0 only
__device__ inline double rng_normal_lane0(RngState* r) {
  const double s = (double)0.449871f, t = (double)-0.386595f, a = (double)0.19600f, b = (double)0.25472f;
  const double r1 = (double)0.27597f, r2 = (double)0.27846f;
  double u, v;
  for (;;) {
    u = rng_uniform_lane0(r);
    v = rng_uniform_lane0(r);
    v = (double)1.7156f * (v - 0.5);
    const double x = u - s;
    const double y = fabs(v) - t;
    const double q = x * x + y * (a * y - b * x);
    if (q < r1) break;
    if (q > r2) continue;
    if (v * v < -4.0 * detmath::det_log(u) * (u * u)) break;
  }
  return v / u;
}

}  // namespace magdev
