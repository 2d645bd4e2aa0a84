// oracle_numerics.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the third-party arithmetic the reference's hot path calls but
// which is NOT under /root/reference:
//   * GSL (C library, reached through FGSL; version unpinned by the reference,
//     README.md:22-25, Makefile:9): gsl_root_fsolver_brent, gsl_root_test_interval,
//     gsl_root_fdfsolver_secant, gsl_root_test_delta, gsl_deriv_central.  Restated
//     from the published GSL 2.x algorithms (roots/brent.c, roots/convergence.c,
//     roots/secant.c, deriv/deriv.c).  Call sites: source/root_finder.f90:158-183
//     (FindRoot) and :246-269 (FindRoot_deriv).
//   * libgfortran's intrinsic RANDOM_NUMBER / RANDOM_SEED (version unpinned):
//     xoshiro256** for GCC >= 9, xorshift1024* for GCC 7/8
//     (libgfortran/intrinsics/random.c).  Call sites: source/random.f90:39,64,90-91,126.
//   * modified Bessel functions I0,I1,K0,K1 (gsl_sf_bessel_*; call sites
//     source/bessel_functions.f90:33-69) and libm's log/exp/pow: pinned deterministic
//     implementations in det_math.hpp (which explains why library calls cannot be used).
//
// PARITY UNPINNED against the real reference: the reference cannot be built in this
// image (no gfortran/MPI/HDF5/GSL/FGSL) and its own tests pin no number.  What IS
// pinned here (tests/test_oracle_numerics.py): the xorshift1024* stream against the
// vendored GCC-8 libgfortran.so.5 driven through ctypes (tests/golden/rng_*.json),
// Bessel values against scipy/mpmath, Brent on analytic functions.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "det_math.hpp"

namespace oracle {

// ---------------------------------------------------------------- RNG -------
// libgfortran random.c.  Seeding as done by set_random_seed (random.f90:39-64): every
// element of the PUT array is the same default-integer value `w`.
struct Rng {
  int kind = 0;         // 0 xoshiro256**, 1 xorshift1024*
  uint64_t s[16];
  int p = 0;

  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

  void seed_all_words(int32_t w32, int kind_) {
    static const uint64_t keys[16] = {
        0xbd0c5b6e50c2df49ULL, 0xd46061cd46e1df38ULL, 0xbb4f4d4ed6103544ULL,
        0x114a583d0756ad39ULL, 0x4b5ad8623d0aaab6ULL, 0x3f2ed7afbe0c0f21ULL,
        0xdec83fd65f113445ULL, 0x3824f8fbc4f10d24ULL, 0x5d9025af05878911ULL,
        0x500bc46b540340e9ULL, 0x8bd53298e0d00530ULL, 0x57886e40a952e06aULL,
        0x926e76c88e31cdb6ULL, 0xbd0724dac0a3a5f9ULL, 0xc5c8981b858ab796ULL,
        0xbb12ab2694c2b32cULL};
    kind = kind_;
    const uint64_t w = (uint64_t)(uint32_t)w32;
    const uint64_t word = w | (w << 32);
    const int n = kind == 0 ? 4 : 16;
    for (int i = 0; i < n; i++) s[i] = word ^ keys[i];
    p = kind == 0 ? 0 : (int)(w32 & 15);
  }

  inline uint64_t next() {
    if (kind == 0) {  // xoshiro256**
      const uint64_t result = rotl(s[1] * 5, 7) * 9;
      const uint64_t t = s[1] << 17;
      s[2] ^= s[0];
      s[3] ^= s[1];
      s[1] ^= s[2];
      s[0] ^= s[3];
      s[2] ^= t;
      s[3] = rotl(s[3], 45);
      return result;
    } else {  // xorshift1024*
      const uint64_t s0 = s[p];
      uint64_t s1 = s[p = (p + 1) & 15];
      s1 ^= s1 << 31;
      s[p] = s1 ^ s0 ^ (s1 >> 11) ^ (s0 >> 30);
      return s[p] * 1181783497276652981ULL;
    }
  }

  // random_r8: keep the 53 high bits, scale by 2^-64
  inline double uniform() {
    uint64_t v = next();
    v &= ~(uint64_t)0 << 11;
    return (double)v * 0x1.0p-64;
  }
};

// ---------------------------------------------------------------- GSL -------
constexpr double GSL_DBL_EPSILON_ = 2.2204460492503131e-16;

// gsl_root_fsolver_brent, driven exactly like FindRoot (root_finder.f90:110-189):
// REL_TOL 1e-7, ABS_TOL 0, itmax 200.  Returns success; root in *root.
template <class F>
bool find_root_brent(F&& f, double x_lower, double x_upper, double* root_out, int* iters_out = nullptr) {
  const double REL_TOL = 1.0e-7;
  const int itmax = 200;
  double root = -std::numeric_limits<double>::max();
  *root_out = root;
  // gsl_root_fsolver_set
  if (x_lower > x_upper) return false;
  double f_lower = f(x_lower);
  if (!std::isfinite(f_lower)) return false;
  double f_upper = f(x_upper);
  if (!std::isfinite(f_upper)) return false;
  double a = x_lower, fa = f_lower, b = x_upper, fb = f_upper, c = x_upper, fc = f_upper;
  double d = x_upper - x_lower, e = x_upper - x_lower;
  if ((f_lower < 0.0 && f_upper < 0.0) || (f_lower > 0.0 && f_upper > 0.0)) return false;

  int i = 0;
  for (;;) {
    i++;
    // ---- brent_iterate
    double xl, xu;
    bool ok = true;
    {
      double tol, m;
      int ac_equal = 0;
      if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) {
        ac_equal = 1;
        c = a; fc = fa; d = b - a; e = b - a;
      }
      if (std::fabs(fc) < std::fabs(fb)) {
        ac_equal = 1;
        a = b; b = c; c = a;
        fa = fb; fb = fc; fc = fa;
      }
      tol = 0.5 * GSL_DBL_EPSILON_ * std::fabs(b);
      m = 0.5 * (c - b);
      if (fb == 0) {
        root = b; xl = b; xu = b;
      } else if (std::fabs(m) <= tol) {
        root = b;
        if (b < c) { xl = b; xu = c; } else { xl = c; xu = b; }
      } else {
        if (std::fabs(e) < tol || std::fabs(fa) <= std::fabs(fb)) {
          d = m; e = m;
        } else {
          double pp, q, r;
          double s = fb / fa;
          if (ac_equal) {
            pp = 2 * m * s;
            q = 1 - s;
          } else {
            q = fa / fc;
            r = fb / fc;
            pp = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
            q = (q - 1) * (r - 1) * (s - 1);
          }
          if (pp > 0) q = -q; else pp = -pp;
          double t1 = 3 * m * q - std::fabs(tol * q), t2 = std::fabs(e * q);
          if (2 * pp < (t1 < t2 ? t1 : t2)) {
            e = d; d = pp / q;
          } else {
            d = m; e = m;
          }
        }
        a = b; fa = fb;
        if (std::fabs(d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
        fb = f(b);
        if (!std::isfinite(fb)) ok = false;  // SAFE_FUNC_CALL -> GSL_EBADFUNC
        root = b;
        double cc = c;
        if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) cc = a;  // bounds only
        if (b < cc) { xl = b; xu = cc; } else { xl = cc; xu = b; }
      }
    }
    if (!ok || i > itmax) {
      if (iters_out) *iters_out = i;
      return false;  // FindRoot keeps the root it read in the previous iteration (*root_out)
    }
    *root_out = root;
    // gsl_root_test_interval(xlo, xhi, 0, REL_TOL)
    double min_abs;
    if ((xl > 0.0 && xu > 0.0) || (xl < 0.0 && xu < 0.0))
      min_abs = std::fabs(xl) < std::fabs(xu) ? std::fabs(xl) : std::fabs(xu);
    else
      min_abs = 0;
    if (std::fabs(xu - xl) < REL_TOL * min_abs) break;
  }
  if (iters_out) *iters_out = i;
  return true;
}

// gsl_deriv_central (deriv/deriv.c)
template <class F>
static void central_deriv_(F&& f, double x, double h, double* result, double* abserr_round, double* abserr_trunc) {
  double fm1 = f(x - h);
  double fp1 = f(x + h);
  double fmh = f(x - h / 2);
  double fph = f(x + h / 2);
  double r3 = 0.5 * (fp1 - fm1);
  double r5 = (4.0 / 3.0) * (fph - fmh) - (1.0 / 3.0) * r3;
  double e3 = (std::fabs(fp1) + std::fabs(fm1)) * GSL_DBL_EPSILON_;
  double e5 = 2.0 * (std::fabs(fph) + std::fabs(fmh)) * GSL_DBL_EPSILON_ + e3;
  double a1 = std::fabs(r3 / h), a2 = std::fabs(r5 / h);
  double dy = (a1 > a2 ? a1 : a2) * (std::fabs(x) / h) * GSL_DBL_EPSILON_;
  *result = r5 / h;
  *abserr_trunc = std::fabs((r5 - r3) / h);
  *abserr_round = std::fabs(e5 / h) + dy;
}

template <class F>
double deriv_central(F&& f, double x, double h) {
  double r_0, round, trunc, error;
  central_deriv_(f, x, h, &r_0, &round, &trunc);
  error = round + trunc;
  if (round < trunc && (round > 0 && trunc > 0)) {
    double r_opt, round_opt, trunc_opt, error_opt;
    double h_opt = h * detmath::det_pow(round / (2.0 * trunc), 1.0 / 3.0);
    central_deriv_(f, x, h_opt, &r_opt, &round_opt, &trunc_opt);
    error_opt = round_opt + trunc_opt;
    if (error_opt < error && std::fabs(r_opt - r_0) < 4.0 * error) {
      r_0 = r_opt;
      error = error_opt;
    }
  }
  return r_0;
}

// gsl_root_fdfsolver_secant driven like FindRoot_deriv (root_finder.f90:191-276):
// REL_TOL 1e-5, derivative at the guess from gsl_deriv_central(step).
template <class F>
bool find_root_secant(F&& f, double guess, double step, double* root_out) {
  const double REL_TOL = 1.0e-5;
  const int itmax = 200;
  double root = -std::numeric_limits<double>::max();
  // secant_init: f and df at the guess
  double x = guess;
  double fx = f(x);
  double df = deriv_central(f, x, step);
  int i = 0;
  bool success = true;
  for (;;) {
    i++;
    // secant_iterate
    int status = 0;
    if (fx == 0.0) {
      status = 0;
    } else if (df == 0.0) {
      status = 1;  // GSL_EZERODIV
    } else {
      double x_new = x - (fx / df);
      double f_new = f(x_new);
      double df_new = df * ((fx - f_new) / fx);
      x = x_new;
      fx = f_new;
      df = df_new;
      if (!std::isfinite(f_new)) status = 1;
      else if (!std::isfinite(df_new)) status = 1;
    }
    if (status != 0 || i > itmax) { success = false; break; }
    double ri = root;
    root = x;
    // gsl_root_test_delta(root, ri, 0, REL_TOL)
    double tolerance = REL_TOL * std::fabs(root);
    if (std::fabs(root - ri) < tolerance || root == ri) break;
  }
  *root_out = root;
  return success;
}

// ---------------------------------------------------------------- Bessel ----
// det_math.hpp explains why these are pinned implementations rather than library calls.
// -DORACLE_LIBM swaps in the host's libm / libstdc++ special functions: the build the
// sensitivity experiment (tools/sensitivity_experiment.py, profiles/r2_sensitivity.md) uses to
// measure how far the reference's own results move with its unpinned libraries.
#ifdef ORACLE_LIBM
inline double bessel_I0(double x) { return std::cyl_bessel_i(0.0, x); }
inline double bessel_I1(double x) { return std::cyl_bessel_i(1.0, x); }
inline double bessel_K0(double x) { return std::cyl_bessel_k(0.0, x); }
inline double bessel_K1(double x) { return std::cyl_bessel_k(1.0, x); }
inline double m_log(double x) { return std::log(x); }
inline double m_exp(double x) { return std::exp(x); }
inline double m_pow(double x, double y) { return std::pow(x, y); }
#else
inline double bessel_I0(double x) { double a, b, c, d; detmath::det_bessel_ik01(x, &a, &b, &c, &d); return a; }
inline double bessel_I1(double x) { double a, b, c, d; detmath::det_bessel_ik01(x, &a, &b, &c, &d); return b; }
inline double bessel_K0(double x) { double a, b, c, d; detmath::det_bessel_ik01(x, &a, &b, &c, &d); return c; }
inline double bessel_K1(double x) { double a, b, c, d; detmath::det_bessel_ik01(x, &a, &b, &c, &d); return d; }
inline double m_log(double x) { return detmath::det_log(x); }
inline double m_exp(double x) { return detmath::det_exp(x); }
inline double m_pow(double x, double y) { return detmath::det_pow(x, y); }
#endif

}  // namespace oracle
