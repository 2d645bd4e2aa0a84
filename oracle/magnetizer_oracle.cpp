// magnetizer_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU oracle: a literal, one-galaxy-at-a-time C++ restatement of the reference's
// per-galaxy galactic-dynamo path (dynamo_run, source/dynamo.f90:39-300, and
// everything it calls).  Every function cites the reference file:line it follows
// (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, load or call this file;
// the product (magnetizer_b200/) never does.
//
// PARITY UNPINNED against the real reference: the Fortran+MPI+HDF5+GSL/FGSL build
// cannot be produced in this image and the reference's own tests hold no golden
// numbers.  The third-party pieces are pinned separately, see
// oracle_numerics.hpp.  Order-dependent/undefined reference behaviour that is
// consciously NOT reproduced is listed in DESIGN.md ("excluded quirks").
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC (oracle/Makefile).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../include/magnetizer_b200.h"
#include "oracle_numerics.hpp"

namespace oracle {

typedef std::vector<double> vec;

// ------------------------------------------------------------------ constants
// source/constants.f90:20-104.  NB pi is the reference's (wrong) value.
static const double pi = 3.14156295358;
static const double mp_g = 1.67262158e-24;
static const double cm_kpc = 3.08567758e21;
static const double km_kpc = 3.08567758e16;
static const double cm_km = 1.e5;
static const double s_Gyr = 1.e9 * 365.25 * 24 * 3600;
static const double mkG_G = 1.e6;
static const double G_SI = 6.673e-11;              // GSL_CONST_MKSA_GRAVITATIONAL_CONSTANT
static const double Msun_SI = 1.98892e30;          // GSL_CONST_MKSA_SOLAR_MASS
static const double kpc_SI = 3.08567758135e16 * 1e3;  // GSL_CONST_MKSA_PARSEC*1d3
static const double density_SI_to_cgs = 1e-3;
static const double convertPressureSItoGaussian = 10;
// units (constants.f90:40-74)
static const double etat0 = 1.0, h0 = 1.0, B0 = 1.0;
static const double t0 = h0 * h0 / etat0;
static const double n0 = B0 * B0 / (h0 * h0) * (t0 * t0);
static const double etat0_kmskpc = 1.0 / 3 * 0.1 * 10.0;
static const double h0_kpc = 1.0;
static const double h0_km = h0_kpc * km_kpc;
static const double h0_cm = h0_km * cm_km;
static const double B0_mkG = 1.0;
static const double t0_kpcskm = h0_kpc * h0_kpc / etat0_kmskpc;
static const double t0_Gyr = t0_kpcskm / s_Gyr * km_kpc;
static const double t0_s = t0_kpcskm * km_kpc;
static const double n0_cm3 = (B0_mkG / mkG_G) * (B0_mkG / mkG_G) / (h0_cm * h0_cm) * (t0_s * t0_s) / mp_g;
// input_constants (constants.f90:77-104)
static const double ctau = 1.0, C_U = 0.25, C_a = 1.0;
static const double C_d = -(pi * pi);
static const double Rm_inv = 0.0, alpceil = 1.0;
static const double constDiskScaleToHalfMassRatio = 1.0 / 1.678346990;
static const double Hmass = 1.67372e-24;

static const int nxghost = 3;          // grid.f90:27
static const double r_in = 1.0e-7;     // grid.f90:30
static const double INVALID = -99999.0;  // tsDataObj.f90:5

// ------------------------------------------------------------------ defaults
// source/global_input_parameters.f90, float-rounded where the literal has no d0
// (SURVEY.md App. A).
static void params_default(mag_params_t* p) {
  std::memset(p, 0, sizeof(*p));
  p->abi_version = MAG_ABI_VERSION;
  p->nsteps_0 = 1000;                                   // :71
  p->p_nsteps_max = 20000;                              // :75
  p->p_MAX_FAILS = 6;                                   // :76
  p->p_random_seed = 17;                                // :83
  p->p_no_magnetic_fields_test_run = 0;                 // :80
  p->p_courant_v = (double)0.09f;                       // :72
  p->p_courant_eta = (double)0.09f;                     // :73
  p->p_nx_ref = 101;                                    // :106
  p->p_nx_MAX = 10000;                                  // :109
  p->p_rmax_over_rdisk = 2.75;                          // :115
  p->frac_seed = 0.01;                                  // :142
  p->C_alp = 1.0;                                       // :173
  p->C_floor = 1.0;                                     // :179
  p->p_floor_kappa = 10.0;                              // :182
  p->fmag = 0.5;                                        // :191
  p->R_kappa = 1.0;                                     // :194
  p->p_ISM_sound_speed_km_s = 10.0;                     // :209
  p->p_ISM_kappa = 1.0;                                 // :211
  p->p_ISM_xi = (double)0.25f;                          // :214
  p->p_ISM_gamma = 5.0 / 3.0;                           // :217
  p->p_ISM_turbulent_length = (double)0.1f;             // :219
  p->p_stellarHeightToRadiusScale = 1.0 / (double)7.3f; // :224
  p->p_molecularHeightToRadiusScale = (double)0.032f;   // :227
  p->p_gasScaleRadiusToStellarScaleRadius_ratio = 1.0;  // :230
  p->p_Rmol_alpha = 0.92;                               // :234
  p->p_Rmol_P0 = 4.787e-12;                             // :236
  p->p_rreg_to_rdisk = (double)0.1f;                    // :258
  p->p_minimum_density = 1e-30;                         // :265
  p->p_rdisk_min = (double)0.5f;                        // :268
  p->Mgas_disk_min = 1e4;                               // :269
  p->rmin_over_rmax = (double)0.001f;                   // :270
  p->p_ignore_satellite_DM_haloes = 0;                  // :262
  p->p_use_Pdm = 1;                                     // :272
  p->p_use_Pbulge = 1;                                  // :273
  // the generator that reproduces the reference's published Bmax values (tests/test_oracle_golden.py)
  p->rng_kind = MAG_RNG_XORSHIFT1024S;
  p->p_variable_timesteps = 1;                          // :74
  p->p_simplified_pressure = 1;                         // :252
  p->p_halo_contraction = 1;                            // :274
  p->p_enable_P2 = 0;                                   // :254
  p->Dyn_quench = 1;                                    // :132
  p->Alg_quench = 0;                                    // :135
  p->Damp = 0;                                          // :138
  p->lFloor = 1;                                        // :176
  p->p_space_varying_floor = 1;                         // :184
  p->p_time_varying_floor = 1;                          // :188
  p->p_limit_turbulent_scale = 1;                       // :221
  p->p_allow_positive_shears = 0;                       // :243
  p->seed_choice = MAG_SEED_RANDOM;                     // :150 'random'
  p->p_neumann_boundary_condition_rmax = 0;             // :170
  p->p_use_fixed_physical_grid = 0;                     // :112
  p->p_nn_seed = 2;                                     // :152
  p->p_r_seed_decay = 15.0;                             // :151
  p->outflow_type = MAG_OUTFLOW_NONE;                   // :306 'no_outflow'
  // outflow_parameters, :303-325 (literals without d0 are single precision)
  p->p_outflow_Lsn = 1e38;                              // :308
  p->p_outflow_fOB = (double)0.7f;                      // :310
  p->p_outflow_etaSN = 9.4e-3;                          // :312
  p->p_tOB = 3.0;                                       // :314
  p->p_N_SN1OB = 40.0;                                  // :316
  p->p_outflow_hot_gas_density = 1.7e-27;               // :318
  p->p_outflow_nu0 = (double)0.5f;                      // :320
  p->p_outflow_alphahot = (double)-3.2f;                // :322
  p->p_outflow_Vhot = 425.0;                            // :324
}

static bool params_supported(const mag_params_t& p) {
  return p.abi_version == MAG_ABI_VERSION && p.p_variable_timesteps == 1 && p.p_simplified_pressure == 1 &&
         p.p_halo_contraction == 1 && p.p_enable_P2 == 0 && (p.Dyn_quench == 0 || p.Dyn_quench == 1) && (p.Alg_quench == 0 || p.Alg_quench == 1) &&
         p.Damp == 0 && p.lFloor == 1 && p.p_space_varying_floor == 1 && p.p_time_varying_floor == 1 &&
         p.p_limit_turbulent_scale == 1 && p.p_allow_positive_shears == 0 && p.seed_choice >= MAG_SEED_RANDOM && p.seed_choice <= MAG_SEED_DECAYING &&
         p.outflow_type >= MAG_OUTFLOW_NONE && p.outflow_type <= MAG_OUTFLOW_SUPERBUBBLE && p.p_nx_ref >= 8 &&
         (p.rng_kind == 0 || p.rng_kind == 1);
}

// ------------------------------------------------------------------ interpolation
// source/interpolation.f90:25-76 (x increasing branch is the only one reached from
// rescale_array).
static void interpolate(const double* x, const double* y, int n, const double* xi, double* yi, int ni) {
  int j = 0;  // 0-based version of j=1
  for (int i = 0; i < ni; i++) {
    while (xi[i] > x[j + 1]) {
      j = j + 1;
      if (j + 1 >= n) {  // j >= size(x) in 1-based
        j = n - 2;
        break;
      }
    }
    yi[i] = y[j] + (y[j + 1] - y[j]) * (xi[i] - x[j]) / (x[j + 1] - x[j]);
  }
}

// source/interpolation.f90:128-167
static void rescale_array(const double* y, int n, double* y_new, int n_new) {
  if (n == n_new) {
    for (int i = 0; i < n; i++) y_new[i] = y[i];
    return;
  }
  vec x(n), x_new(n_new);
  for (int i = 1; i <= n; i++) x[i - 1] = (double)(i - 1) / (double)(n - 1);
  for (int i = 1; i <= n_new; i++) x_new[i - 1] = (double)(i - 1) / (double)(n_new - 1);
  interpolate(x.data(), y, n, x_new.data(), y_new, n_new);
}

// ------------------------------------------------------------------ per-run stats
struct Stats {
  int64_t point_stages = 0;
  uint32_t flags = 0;
};

// ------------------------------------------------------------------ the operator
struct Galaxy {
  const mag_params_t& P;
  Rng rng;
  Stats stats;

  // ---- input_params module state (input_parameters.f90:25-63)
  int nz = 0;
  const double* output_times = nullptr;  // t (Gyr) [nz]
  std::vector<double> galaxy_data;       // [nz][13]
  int iread = 0;
  bool last_output = false;
  int init_it = -1, max_it = -1;
  double tsnap = 0, dt = 0, t = 0, time_between_inputs = 0;
  int nsteps = 20;
  double t_Gyr = 0, r_max_kpc_history = 0;
  double r_disk = 0, v_disk = 0, r_bulge = 0, v_bulge = 0, r_halo = 0, v_halo = 0, nfw_cs1 = 0;
  double Mstars_disk = 0, Mgas_disk = 0, SFR = 0, Mhalo = 0, Mstars_bulge = 0;
  bool is_central_galaxy = false;

  // ---- grid module state (grid.f90:26-45)
  int nx = -1;
  double dx = 0, lambda = 0, dr_kpc = 0, r_max_kpc = 0, previous_r_max_kpc = -1;
  vec x, r_kpc;

  // ---- profiles module state (profiles.f90:22-30)
  vec h, n, l, l_kpc, v, etat, tau, Beq, alp_k, Uz, Ur, Om, G, Om_d, Om_b, G_b, Om_h, G_h;
  double delta_r_floor = 0;

  // ---- floor_field module state (floor_field.f90:29-32)
  double Bfloor_sign = 1, t_last_sign_choice = 0;
  vec sign_array;
  bool refresh_sign_array = true;

  // ---- dynamo / equ / bzcalc module state
  vec f;  // [3][nx] (var-major, like Fortran f(nx,nvar))
  vec f_snapshot_beginning;
  vec alp, Bzmod;
  char status_code = 'M';
  bool test_run = false;
  bool nx_overflow = false;

  // ---- ts_arrays
  std::vector<vec> ts_prof;    // [MAG_NPROFILE][nz*nx_ref]
  std::vector<vec> ts_scal;    // [MAG_NSCALAR][nz]
  std::vector<char> ts_status; // [nz]
  std::vector<int32_t> hist_nsteps, hist_nx;  // [nz] diagnostics: nsteps / nx used per snapshot

  explicit Galaxy(const mag_params_t& p) : P(p) {}

  double* F(int var) { return f.data() + (size_t)var * nx; }

  // -------------------------------------------------------------- random.f90
  // set_random_seed, random.f90:27-65 (default-integer arithmetic wraps)
  void set_random_seed(int32_t igal, int32_t seed) {
    uint32_t s2 = (uint32_t)seed * (uint32_t)seed;
    uint32_t g3 = (uint32_t)igal * (uint32_t)igal * (uint32_t)igal;
    int32_t w = (int32_t)(s2 - g3);
    if (w == 0) {  // where (seed==0) seed = p_random_seed**igal
      uint32_t r = 1;
      uint32_t b = (uint32_t)seed;
      uint32_t e = (uint32_t)igal;
      // gfortran integer power by repeated squaring; wrapping arithmetic gives the same
      // residue whatever the multiplication order
      while (e) {
        if (e & 1) r *= b;
        b *= b;
        e >>= 1;
      }
      w = (int32_t)r;
    }
    rng.seed_all_words(w, P.rng_kind);
  }
  // random_sign, random.f90:121-129
  double random_sign() {
    double u = rng.uniform();
    return std::copysign(1.0, u - 0.5);
  }
  // random_normal, random.f90:68-110 (single-precision literals widened)
  double random_normal() {
    const double s = (double)0.449871f, tt = (double)-0.386595f;
    const double a = (double)0.19600f, b = (double)0.25472f;
    const double half = 0.5, r1 = (double)0.27597f, r2 = (double)0.27846f;
    double u, vv, xq, y, q;
    for (;;) {
      u = rng.uniform();
      vv = rng.uniform();
      vv = (double)1.7156f * (vv - half);
      xq = u - s;
      y = std::fabs(vv) - tt;
      q = xq * xq + y * (a * y - b * xq);
      if (q < r1) break;
      if (q > r2) continue;
      if (vv * vv < -4.0 * m_log(u) * (u * u)) break;
    }
    return vv / u;
  }

  // -------------------------------------------------------------- input_parameters.f90
  // load_input_parameters, :136-192.  `inputs_g` points at this galaxy's 13 columns:
  // column c, snapshot i at inputs_g[c*col_stride + i].
  void load_input_parameters(const double* inputs_g, size_t col_stride) {
    galaxy_data.assign((size_t)nz * 13, -1.0);
    for (int c = 0; c < 13; c++)
      for (int i = 0; i < nz; i++) galaxy_data[(size_t)i * 13 + c] = inputs_g[(size_t)c * col_stride + i];
    double mx = galaxy_data[0];
    for (int i = 1; i < nz; i++) mx = std::max(mx, galaxy_data[(size_t)i * 13]);
    r_max_kpc_history = mx * P.p_rmax_over_rdisk;
    init_it = -1;
    max_it = nz;
    for (int i = 1; i <= nz; i++) {
      double rd = galaxy_data[(size_t)(i - 1) * 13];
      if (rd >= P.p_rdisk_min && init_it < 0) init_it = i;
      if (rd < 0 && init_it > 0) {
        max_it = i - 1;
        break;
      }
    }
    iread = init_it - 1;
  }

  // set_input_params, :202-255
  void set_input_params(bool* error) {
    *error = false;
    iread = iread + 1;
    if (iread < 0 || max_it - init_it == 0) {
      *error = true;
      status_code = 'p';
      return;
    }
    double current_time_input = output_times[iread - 1];
    if (iread < max_it) {
      double next_time_input = output_times[iread];
      time_between_inputs = next_time_input - current_time_input;
      t = 0;
    } else {
      last_output = true;
    }
    const double* g = &galaxy_data[(size_t)(iread - 1) * 13];
    t_Gyr = current_time_input;
    r_disk = g[0]; v_disk = g[1]; r_bulge = g[2]; v_bulge = g[3]; r_halo = g[4]; v_halo = g[5];
    nfw_cs1 = g[6]; Mgas_disk = g[7]; Mstars_disk = g[8]; SFR = g[9]; Mhalo = g[10]; Mstars_bulge = g[11];
    is_central_galaxy = g[12] > 0.1;
  }

  // set_timestep, :65-134 (p_variable_timesteps = .true. branch only)
  bool set_timestep(bool have_profiles) {
    bool success = true;
    tsnap = time_between_inputs / t0_Gyr;
    if (have_profiles) {
      double minh = 1e10;
      for (int i = 0; i < nx; i++)
        if (h[i] > 0 && h[i] < minh) minh = h[i];
      double maxv = v[0], maxe = etat[0];
      for (int i = 1; i < nx; i++) {
        maxv = std::max(maxv, v[i]);
        maxe = std::max(maxe, etat[i]);
      }
      if (std::fabs(maxv) > 1e-20 && std::fabs(maxe) > 1e-20) {
        double d1 = P.p_courant_v * dx / lambda / maxv;
        double d2 = P.p_courant_eta * (minh * minh) / maxe;
        dt = std::min(d1, d2);
        double nsteps_f = tsnap / dt;
        nsteps = (int)std::min((double)1e9f, nsteps_f);
      } else {
        nsteps = P.p_nsteps_max;
        success = false;
      }
    } else {
      // "missing arguments. Falling back to fixed number of timesteps scheme": dt is
      // NOT recomputed
      nsteps = P.nsteps_0;
    }
    if (nsteps > P.p_nsteps_max) {
      nsteps = P.p_nsteps_max;
      success = false;
    }
    return success;
  }

  // -------------------------------------------------------------- grid.f90
  // construct_grid, :48-81
  void construct_grid() {
    int nxphys = P.p_nx_ref;
    nx = nxphys + 2 * nxghost;
    dx = (1.0 - r_in) / (double)(nxphys - 1);
    x.assign(nx, 0.0);
    r_kpc.assign(nx, 0.0);
    for (int i = 1; i <= nx; i++) x[i - 1] = r_in - nxghost * dx + ((double)i - 1.0) * dx;
    r_max_kpc = P.p_use_fixed_physical_grid ? r_max_kpc_history : P.p_rmax_over_rdisk * r_disk;   // :69-73
    for (int i = 0; i < nx; i++) r_kpc[i] = x[i] * r_max_kpc;
    dr_kpc = dx * r_max_kpc;
    lambda = h0_kpc / r_max_kpc;
  }

  // adjust_grid, :143-297 (default switches)
  void adjust_grid() {
    if (P.p_use_fixed_physical_grid) return;   // :159-162
    const double GRID_ADJ_THRES = (double)0.05f;
    const int nvar = 3;
    const int nx_def = P.p_nx_ref + 2 * nxghost;
    if (nx != nx_def) {  // p_scale_back_f_array
      int nx_old = nx;
      nx = nx_def;
      vec tmp = r_kpc;
      r_kpc.assign(nx, 0.0);
      rescale_array(tmp.data() + nxghost, nx_old - nxghost, r_kpc.data() + nxghost, nx - nxghost);
      dr_kpc = r_kpc[nxghost + 1] - r_kpc[nxghost];
      for (int i = nxghost; i >= 1; i--) r_kpc[i - 1] = r_kpc[i] - dr_kpc;
      // rescale_f_array (interpolation.f90:170-198): inner ghost zone left unset
      vec f_old = f;
      f.assign((size_t)nx * nvar, 0.0);
      for (int var = 0; var < nvar; var++)
        rescale_array(f_old.data() + (size_t)var * nx_old + nxghost, nx_old - nxghost,
                      f.data() + (size_t)var * nx + nxghost, nx - nxghost);
      x.assign(nx, 0.0);
      for (int i = 0; i < nx; i++) x[i] = r_kpc[i] / r_max_kpc;
      // dx is NOT updated here
    }
    previous_r_max_kpc = r_max_kpc;
    r_max_kpc = P.p_rmax_over_rdisk * r_disk;
    double relative_variation = std::fabs(r_max_kpc - previous_r_max_kpc) / previous_r_max_kpc;
    if (relative_variation < GRID_ADJ_THRES) {
      r_max_kpc = previous_r_max_kpc;
      return;
    }
    if (r_max_kpc >= previous_r_max_kpc) {
      // extend the grid keeping the resolution (p_rescale_field_for_expanding_disks=F)
      int nx_new = nx;
      double r_max_candidate = previous_r_max_kpc;
      while (r_max_candidate < r_max_kpc) {
        nx_new = nx_new + 1;
        r_max_candidate = (r_in / dx - nxghost + ((double)(nx_new - nxghost) - 1.0)) * dr_kpc;
        if (nx_new > P.p_nx_MAX + 1) break;  // keeps a runaway loop bounded; flagged below
      }
      if (nx_new > P.p_nx_MAX) {  // reference: stop
        nx_overflow = true;
        return;
      }
      r_max_kpc = r_max_candidate;
      r_kpc.resize(nx_new, 0.0);
      for (int i = nx; i <= nx_new; i++) r_kpc[i - 1] = r_kpc[i - 2] + dr_kpc;
      x.assign(nx_new, 0.0);
      for (int i = 0; i < nx_new; i++) x[i] = r_kpc[i] / r_max_kpc;
      dx = x[1] - x[0];
      vec ftmp = f;
      f.assign((size_t)nx_new * nvar, 0.0);
      const int keep = nx - nxghost - 1;  // f(:nx-nxghost-1,:)
      for (int var = 0; var < nvar; var++) {
        for (int i = 1; i <= keep; i++) f[(size_t)var * nx_new + i - 1] = ftmp[(size_t)var * nx + i - 1];
        for (int i = nx - nxghost; i <= nx_new; i++)
          f[(size_t)var * nx_new + i - 1] = ftmp[(size_t)var * nx + keep - 1] * r_kpc[keep - 1] / r_kpc[i - 1];
      }
      nx = nx_new;
    } else {
      // shrinking: rescale units, keep f (p_rescale_field_for_shrinking_disks=T)
      for (int i = 0; i < nx; i++) r_kpc[i] = r_kpc[i] * r_max_kpc / previous_r_max_kpc;
      dr_kpc = dr_kpc * r_max_kpc / previous_r_max_kpc;
    }
    lambda = h0_kpc / r_max_kpc;
  }

  // -------------------------------------------------------------- deriv.f90
  // xder, :22-54
  void xder(const double* ff, double* out) const {
    const double fac = 1.0 / (60. * (x[1] - x[0]));
    for (int ix = 4; ix <= nx - 3; ix++) {
      const double* p = ff + ix - 1;
      out[ix - 1] = fac * (+45. * (p[1] - p[-1]) - 9. * (p[2] - p[-2]) + (p[3] - p[-3]));
    }
    const double* a = ff;  // a[k] = f(k+1)
    out[0] = fac * (-147. * a[0] + 360. * a[1] - 450. * a[2] + 400. * a[3] - 225. * a[4] + 72. * a[5] - 10. * a[6]);
    out[1] = fac * (-10. * a[0] - 77. * a[1] + 150. * a[2] - 100. * a[3] + 50. * a[4] - 15. * a[5] + 2. * a[6]);
    out[2] = fac * (2. * a[0] - 24. * a[1] - 35. * a[2] + 80. * a[3] - 30. * a[4] + 8. * a[5] - a[6]);
    const double* b = ff + nx - 1;  // b[-k] = f(nx-k)
    out[nx - 1] = fac * (147. * b[0] - 360. * b[-1] + 450. * b[-2] - 400. * b[-3] + 225. * b[-4] - 72. * b[-5] + 10. * b[-6]);
    out[nx - 2] = fac * (10. * b[0] + 77. * b[-1] - 150. * b[-2] + 100. * b[-3] - 50. * b[-4] + 15. * b[-5] - 2. * b[-6]);
    out[nx - 3] = fac * (-2. * b[0] + 24. * b[-1] + 35. * b[-2] - 80. * b[-3] + 30. * b[-4] - 8. * b[-5] + b[-6]);
  }
  // xder2, :56-89
  void xder2(const double* ff, double* out) const {
    const double d = x[1] - x[0];
    const double fac = 1.0 / (180. * (d * d));
    for (int ix = 4; ix <= nx - 3; ix++) {
      const double* p = ff + ix - 1;
      out[ix - 1] = fac * (-490. * p[0] + 270. * (p[1] + p[-1]) - 27. * (p[2] + p[-2]) + 2. * (p[3] + p[-3]));
    }
    const double* a = ff;
    out[0] = fac * (812. * a[0] - 3132. * a[1] + 5265. * a[2] - 5080. * a[3] + 2970. * a[4] - 972. * a[5] + 137. * a[6]);
    out[1] = fac * (137. * a[0] - 147. * a[1] - 255. * a[2] + 470. * a[3] - 285. * a[4] + 93. * a[5] - 13. * a[6]);
    out[2] = fac * (-13. * a[0] + 228. * a[1] - 420. * a[2] + 200. * a[3] + 15. * a[4] - 12. * a[5] + 2. * a[6]);
    const double* b = ff + nx - 1;
    out[nx - 1] = fac * (812. * b[0] - 3132. * b[-1] + 5265. * b[-2] - 5080. * b[-3] + 2970. * b[-4] - 972. * b[-5] + 137. * b[-6]);
    out[nx - 2] = fac * (137. * b[0] - 147. * b[-1] - 255. * b[-2] + 470. * b[-3] - 285. * b[-4] + 93. * b[-5] - 13. * b[-6]);
    out[nx - 3] = fac * (-13. * b[0] + 228. * b[-1] - 420. * b[-2] + 200. * b[-3] + 15. * b[-4] - 12. * b[-5] + 2. * b[-6]);
  }

  // -------------------------------------------------------------- rotationCurves.f90
  // disk_rotation_curve for one radius, :27-97
  double disk_omega(double r, double rmin) const {
    const double TOO_SMALL = (double)2e-7f;
    double rs = r_disk * constDiskScaleToHalfMassRatio;
    double y50 = 1.0 / constDiskScaleToHalfMassRatio;
    if (r_disk < rmin) return 0.0;
    double y;
    if (std::fabs(r) > TOO_SMALL) y = std::fabs(r) / rs; else y = TOO_SMALL / rs;
    double A = (bessel_I0(y) * bessel_K0(y) - bessel_I1(y) * bessel_K1(y)) /
               (bessel_I0(y50) * bessel_K0(y50) - bessel_I1(y50) * bessel_K1(y50));
    A = std::sqrt(A);
    return A * v_disk / r_disk;
  }
  // bulge_rotation_curve for one radius, :99-143
  double bulge_omega(double r) const {
    const double a_to_r50 = std::sqrt(2.0) - 1.0;
    if (v_bulge < (double)1e-3f) return 0.0;
    double a = a_to_r50 * r_bulge;
    double y = std::fabs(r) / a;
    double y50 = 1.0 / a_to_r50;
    double Constant = v_bulge * (1.0 + y50) / std::sqrt(y50);
    double vv = Constant * std::sqrt(y) / (y + 1.0);
    return vv / std::fabs(r);
  }
  // halo_velocity, :192-221
  double halo_velocity(double r) const {
    double y = std::fabs(r) / r_halo;
    double c = 1.0 / nfw_cs1;
    double A = v_halo * v_halo / (m_log(1.0 + c) - c / (1.0 + c));
    double B = (m_log(1.0 + c * y) - c * y / (1.0 + c * y)) / y;
    double v2 = A * B;
    return std::sqrt(std::fabs(v2));
  }
  // halo_velocity_contracted, :223-288, with halo_velocity_contracted_fun :290-322
  double halo_velocity_contracted(double r, double f_b, double rmin, bool at_centre) {
    // Vb2 and Vd2 depend on r only: the reference recomputes them at every function
    // evaluation (8 Bessel calls each); the values are identical, so compute once.
    double ob = bulge_omega(r);
    double Vb2 = (ob * r) * (ob * r);
    double od = disk_omega(r, rmin);
    double Vd2 = (od * r) * (od * r);
    auto fun = [&](double r0) {
      double hv = halo_velocity(r0);
      double V02 = hv * hv;
      return r0 * r0 * V02 - r * r * (Vd2 + Vb2 + (1.0 - f_b) * V02 * (r0 / r));
    };
    double fun_min = fun(r);
    if (fun_min > 0.0) {
      // Reference uses an UNINITIALISED r0 here (rotationCurves.f90:266-269): undefined.
      // This implementation uses r0 = r.
      // The galaxy is flagged when this happens anywhere but at the centre point, where
      // regularize() overwrites Omega with Omega(r_xi) whatever V was (the only place it
      // happens in the example input).
      if (!at_centre) stats.flags |= MAG_FLAG_UNINIT_R0;
      return halo_velocity(r);
    }
    int i;
    double p15 = 1.0;  // 1.5d0**i: exact in binary floating point for i <= 11
    for (i = 1; i <= 10; i++) {
      p15 *= 1.5;
      if (fun(r * p15) / fun_min < 0) break;
    }
    if (i > 10) p15 *= 1.5;  // loop index after a completed DO loop is 11
    double r0;
    bool ok = find_root_brent(fun, r, r * p15, &r0);
    if (!ok) {
      // reference: error_message(abort=.true.) -> MPI_ABORT.  Propagate as NaN.
      return std::numeric_limits<double>::quiet_NaN();
    }
    return std::sqrt(r0 / r) * halo_velocity(r0);
  }

  // -------------------------------------------------------------- profiles.f90
  // regularize, :275-331 (Shear absent)
  void regularize(const vec& rx, int i_xi /*0-based*/, vec& Omega) const {
    const double small_factor = (double)1e-15f;
    double r_xi = rx[i_xi];
    double Omega_xi = Omega[i_xi];
    for (int i = 0; i < nx; i++) {
      double rxi_over_r_k, e;
      if (rx[i] > small_factor * r_xi) {
        double q = r_xi / rx[i];
        rxi_over_r_k = q * q;
      } else {
        Omega[i] = 0.0;
        rxi_over_r_k = 0.0;
      }
      if (rxi_over_r_k < (double)1e12f) e = m_exp(-rxi_over_r_k); else e = 0;
      Omega[i] = e * (Omega[i] - Omega_xi) + Omega_xi;
    }
  }

  // -------------------------------------------------------------- surface_density.f90
  // exp_surface_density, :26-38
  static double exp_surface_density(double rs, double r, double M) {
    return M / 2. / pi / (rs * rs) * m_exp(-r / rs);
  }
  // midplane_pressure_Elmegreen, :86-128 and molecular_to_diffuse_ratio, :69-83
  double molecular_to_diffuse_ratio(double rdisk, double Sigma_g, double Sigma_star) const {
    double h_star_SI = P.p_stellarHeightToRadiusScale * constDiskScaleToHalfMassRatio * rdisk * kpc_SI;
    double v_gas_SI = P.p_ISM_sound_speed_km_s * 1e3;
    double Sigma_g_SI = Sigma_g * Msun_SI / (kpc_SI * kpc_SI);
    double Sigma_star_SI = Sigma_star * Msun_SI / (kpc_SI * kpc_SI);
    double v_star_SI = std::sqrt(pi * G_SI * h_star_SI * Sigma_star_SI);
    if (v_star_SI < v_gas_SI) v_star_SI = v_gas_SI;
    double Pnot = pi / 2.0 * G_SI * Sigma_g_SI * (Sigma_g_SI + (v_gas_SI / v_star_SI) * Sigma_star_SI) *
                  convertPressureSItoGaussian;
    return m_pow(Pnot / P.p_Rmol_P0, P.p_Rmol_alpha);
  }

  // -------------------------------------------------------------- pressureEquilibrium.f90
  // density_to_scaleheight, :253-259
  static double density_to_scaleheight(double h_d, double Sigma_d) {
    return Sigma_d * Msun_SI / kpc_SI / kpc_SI / h_d / 2.0 / kpc_SI * density_SI_to_cgs;
  }
  // computes_midplane_ISM_pressure_P1, :425-559 (default analytic branches)
  double pressure_P1(double Sigma_d, double Sigma_star, double R_m, double h_d, double Om_h_, double G_h_,
                     double Om_b_, double G_b_) const {
    const double I_d = 0.5, km_SI = 1e3, Mstars_bulge_min = 1e3;
    double h_star = r_disk * constDiskScaleToHalfMassRatio * P.p_stellarHeightToRadiusScale;
    double h_m = r_disk * constDiskScaleToHalfMassRatio * P.p_molecularHeightToRadiusScale;
    double Sigma_star_SI = (Sigma_star * Msun_SI / kpc_SI / kpc_SI);
    double Sigma_d_SI = (Sigma_d * Msun_SI / kpc_SI / kpc_SI);
    double preFactor = pi * G_SI * Sigma_d_SI * convertPressureSItoGaussian;
    double Pd = preFactor * Sigma_d_SI * 1.0 * I_d;
    double I_stars = h_d / (h_star + h_d);
    double I_m = h_d / (h_m + h_d);
    double Pm = preFactor * Sigma_d_SI * R_m * I_m;
    double Pstars = preFactor * Sigma_star_SI * I_stars;
    double Pbulge, Pdm;
    if (P.p_use_Pbulge && Mstars_bulge > Mstars_bulge_min) {
      Pbulge = Om_b_ * (1.5 * Om_b_ + G_b_) * ((km_SI / kpc_SI) * (km_SI / kpc_SI));
      Pbulge = Pbulge * Sigma_d_SI * h_d * kpc_SI * convertPressureSItoGaussian;
    } else {
      Pbulge = 0.0;
    }
    if (P.p_use_Pdm) {
      Pdm = Om_h_ * (1.5 * Om_h_ + G_h_) * ((km_SI / kpc_SI) * (km_SI / kpc_SI));
      Pdm = Pdm * Sigma_d_SI * h_d * kpc_SI * convertPressureSItoGaussian;
    } else {
      Pdm = 0.0;
    }
    return Pd + Pm + Pstars + Pbulge + Pdm;
  }
  // computes_midplane_ISM_pressure_P2, :575-594
  static double pressure_P2(double Sigma_d, double Om_, double S, double h_d) {
    const double km_SI = 1e3;
    double Sigma_d_SI = (Sigma_d * Msun_SI / kpc_SI / kpc_SI);
    double Pp = -Om_ * (Om_ + S) * ((km_SI / kpc_SI) * (km_SI / kpc_SI));
    return Pp * Sigma_d_SI * h_d * kpc_SI * convertPressureSItoGaussian;
  }
  // computes_midplane_ISM_pressure_from_B_and_rho, :561-573
  double pressure_gas(double B, double rho) const {
    double cs = P.p_ISM_sound_speed_km_s * 1e5;
    return (B * 1e-6) * (B * 1e-6) / 4.0 / pi +
           (P.p_ISM_kappa / 3.0 * (1. + 2. * P.p_ISM_xi) + 1.0 / P.p_ISM_gamma) * (cs * cs) * rho;
  }

  // solve_hydrostatic_equilibrium_numerical, :51-251 (nghost = 3)
  void solve_hydrostatic(const vec& r, const vec& B, const vec& Om_, const vec& G_, vec& rho_d, vec& h_d, vec& Rm,
                         vec& Sigma_star, vec& Sigma_d) {
    const int n_ = nx, nghost_actual = nxghost;
    const double GUESS_INTERVAL = 25;
    const int ERROR_COUNT_MAX = 4;
    int error_count = 0;
    double rs = constDiskScaleToHalfMassRatio * r_disk;
    double rs_g = P.p_gasScaleRadiusToStellarScaleRadius_ratio * rs;
    vec Sigma_g(n_);
    for (int i = 0; i < n_; i++) {
      Sigma_g[i] = exp_surface_density(rs_g, r[i], Mgas_disk);
      Sigma_star[i] = exp_surface_density(rs, r[i], Mstars_disk);
      Rm[i] = molecular_to_diffuse_ratio(r_disk, Sigma_g[i], Sigma_star[i]);
      Sigma_d[i] = Sigma_g[i] / (Rm[i] + 1.0);
    }
    double minimum_h = 1e-3 * r_disk;
    double maximum_h = 2.0 * r_disk;
    double guess_h = (double)0.050f;
    for (int i = 0; i < n_; i++) h_d[i] = -17.0;

    for (int i = nghost_actual + 2; i <= n_; i++) {
      const int k = i - 1;
      const double p2 = Sigma_d[k], p3 = Sigma_star[k], p4 = Rm[k];
      const double p5 = 0.0, p6 = 0.0;  // p_enable_P2 = F
      const double p7 = Om_h[k], p8 = G_h[k], p9 = Om_b[k], p10 = G_b[k], p11 = B[k];
      (void)Om_; (void)G_;
      // pressure_equation, :262-311
      auto pressure_equation = [&](double hh) {
        double P1 = pressure_P1(p2, p3, p4, hh, p7, p8, p9, p10);
        double P2 = pressure_P2(p2, p5, p6, hh);
        double rho_cgs = density_to_scaleheight(hh, p2);
        double Pgas = pressure_gas(p11, rho_cgs);
        return P1 + P2 - Pgas;
      };
      double pe_min = pressure_equation(minimum_h);
      double pe_max = pressure_equation(maximum_h);
      if (std::copysign(1.0, pe_min) != std::copysign(1.0, pe_max)) {
        bool success = find_root_brent(pressure_equation, minimum_h, maximum_h, &h_d[k]);
        if (!success) {
          for (int j = 0; j < n_; j++) { h_d[j] = -1; rho_d[j] = 0.0; }
          return;
        }
      } else {
        stats.flags |= MAG_FLAG_SECANT;
        bool success = find_root_secant(pressure_equation, guess_h, 1e-6, &h_d[k]);
        if (!success || h_d[k] < 0) {
          if (error_count < ERROR_COUNT_MAX && i > nghost_actual + 2 + 2) {
            status_code = 'P';
            h_d[k] = (m_log(h_d[k - 1]) - m_log(h_d[k - 2])) / (r[k - 1] - r[k - 2]) * (r[k] - r[k - 1]);
            h_d[k] = h_d[k] + m_log(h_d[k - 1]);
            h_d[k] = m_exp(h_d[k]);
            error_count = error_count + 1;
          } else {
            for (int j = 0; j < n_; j++) { h_d[j] = -1; rho_d[j] = 0.0; }
            return;
          }
        }
      }
      rho_d[k] = density_to_scaleheight(h_d[k], Sigma_d[k]);
      if (i != n_) {
        guess_h = h_d[k] * m_exp((r[k + 1] - r[k]) / rs);
        maximum_h = guess_h * (1.0 + GUESS_INTERVAL / 100.);
        minimum_h = guess_h * (1.0 - GUESS_INTERVAL / 100.);
      }
      if (rho_d[k] < P.p_minimum_density) {
        rho_d[k] = P.p_minimum_density;
        h_d[k] = density_to_scaleheight(rho_d[k], Sigma_d[k]);
      }
    }
    for (int i = 1; i <= nghost_actual; i++) h_d[i - 1] = h_d[2 * nghost_actual + 2 - i - 1];
    {
      const int c = nghost_actual;  // 0-based index of point nghost+1
      h_d[c] = h_d[c + 1] + (r[c] - r[c + 1]) * (h_d[c + 1] - h_d[c + 2]) / (r[c + 1] - r[c + 2]);
    }
    for (int i = 1; i <= nghost_actual + 1; i++) rho_d[i - 1] = density_to_scaleheight(h_d[i - 1], Sigma_d[i - 1]);
  }

  // outflow_speed, outflow.f90:24-121, at one radius.  r, h in kpc, rho in g/cm^3, vt in km/s;
  // result in km/s.  Literals without d0 are single precision in the reference.
  double outflow_speed(double r, double rho, double hk, double vt, double Rm_i) const {
    switch (P.outflow_type) {
      case MAG_OUTFLOW_NONE: return r * 0.0;                                      // :65-67
      case MAG_OUTFLOW_VTURB: return vt;                                          // :68-70
      case MAG_OUTFLOW_WIND: {                                                    // :71-88
        const double VDISK_MIN = (double)0.01f;                                   // :56
        const double fm = Rm_i / (1.0 + Rm_i);
        const double beta = m_pow(std::max(VDISK_MIN, v_disk) / P.p_outflow_Vhot, P.p_outflow_alphahot);
        double u = 0.5 * P.p_outflow_nu0 * beta * hk * fm;
        u = u * km_kpc / s_Gyr;
        return u;
      }
      default: break;
    }
    // superbubble[_simple], :89-119
    const double rs = r_disk * constDiskScaleToHalfMassRatio;
    const double nn = rho / Hmass;
    if (!(nn > 0.0)) return 0.0;
    const double constant = (double)51.3f * m_pow(P.p_outflow_Lsn / 1e38, (double)(1.f / 3.f)) * P.p_outflow_fOB / 0.7 *
                            (P.p_outflow_etaSN / (double)9.4e-3f) * (P.p_tOB / 3.0) * (40.0 / P.p_N_SN1OB);
    const double q = rs / 3.0;
    double u = constant * (1.0 / (q * q)) * SFR;                                  // (rs/3.0)**(-2): integer power
    u = u * m_pow(nn, (double)(-1.f / 3.f));
    u = u * m_pow(hk / (double)0.2f, (double)(4.f / 3.f));
    u = u * m_exp(-r / rs);
    if (P.outflow_type == MAG_OUTFLOW_SUPERBUBBLE) u = (P.p_outflow_hot_gas_density / rho) * u;
    return u;
  }

  // construct_profiles, profiles.f90:33-272 (B absent => 0)
  bool construct_profiles() {
    const vec* all[] = {&h, &n, &l, &l_kpc, &v, &etat, &tau, &Beq, &alp_k, &Uz, &Ur, &Om, &G, &Om_d, &Om_b, &G_b, &Om_h, &G_h};
    for (auto* a : all) const_cast<vec*>(a)->resize(nx, 0.0);  // check_allocate
    bool result = true;
    double r_disk_min = r_max_kpc * P.rmin_over_rmax;
    delta_r_floor = P.p_floor_kappa * P.p_ISM_turbulent_length / r_max_kpc;

    vec Om_kmskpc(nx), G_kmskpc(nx), absr(nx), tmp(nx);
    for (int i = 0; i < nx; i++) absr[i] = std::fabs(r_kpc[i]);
    for (int i = 0; i < nx; i++) Om_d[i] = disk_omega(r_kpc[i], r_disk_min);
    for (int i = 0; i < nx; i++) Om_b[i] = bulge_omega(r_kpc[i]);
    if (is_central_galaxy || !P.p_ignore_satellite_DM_haloes) {
      double baryon_fraction = (Mstars_disk + Mstars_bulge + Mgas_disk) / Mhalo;
      // halo_rotation_curve, rotationCurves.f90:145-190: v from abs(r), Omega = v/r (signed r)
      for (int i = 0; i < nx; i++) {
        double vv = halo_velocity_contracted(absr[i], baryon_fraction, r_disk_min, i == nxghost);
        Om_h[i] = vv / r_kpc[i];
      }
    } else {
      for (int i = 0; i < nx; i++) Om_h[i] = 0.0;
    }
    for (int i = 0; i < nx; i++) Om_kmskpc[i] = std::sqrt(Om_d[i] * Om_d[i] + Om_b[i] * Om_b[i] + Om_h[i] * Om_h[i]);

    // ireg = minloc(abs(r_kpc - p_rreg_to_rdisk*r_disk),1)
    int ireg = 0;
    {
      double best = std::fabs(r_kpc[0] - P.p_rreg_to_rdisk * r_disk);
      for (int i = 1; i < nx; i++) {
        double d = std::fabs(r_kpc[i] - P.p_rreg_to_rdisk * r_disk);
        if (d < best) { best = d; ireg = i; }
      }
    }
    regularize(absr, ireg, Om_kmskpc);
    regularize(absr, ireg, Om_h);
    regularize(absr, ireg, Om_b);

    xder(Om_kmskpc.data(), tmp.data());
    for (int i = 0; i < nx; i++) G_kmskpc[i] = tmp[i] * x[i];
    xder(Om_h.data(), tmp.data());
    for (int i = 0; i < nx; i++) G_h[i] = tmp[i] * x[i];
    xder(Om_b.data(), tmp.data());
    for (int i = 0; i < nx; i++) G_b[i] = tmp[i] * x[i];
    for (int i = 0; i < nx; i++)
      if (G_kmskpc[i] > 0.0) G_kmskpc[i] = 0.0;  // p_allow_positive_shears = F

    for (int i = 0; i < nx; i++) {
      Om[i] = Om_kmskpc[i] / h0_km * h0_kpc * t0_s / t0;
      G[i] = G_kmskpc[i] / h0_km * h0_kpc * t0_s / t0;
      Ur[i] = 0.0;
    }
    double v_kms = P.p_ISM_sound_speed_km_s * P.p_ISM_kappa;
    for (int i = 0; i < nx; i++) v[i] = v_kms / h0_km * h0 * t0_s / t0;

    // negligible-disk trap, :131-140
    if (r_disk < r_disk_min || Mgas_disk < P.Mgas_disk_min) {
      for (int i = 0; i < nx; i++) { n[i] = 0; l[i] = 0; h[i] = 0; Uz[i] = 0; etat[i] = 0; alp_k[i] = 0; }
      return true;
    }

    vec B_actual(nx, 0.0), rho_cgs(nx, 0.0), h_kpc(nx, 0.0), Rm(nx), Sigma_star(nx), Sigma_d(nx);
    solve_hydrostatic(absr, B_actual, Om_kmskpc, G_kmskpc, rho_cgs, h_kpc, Rm, Sigma_star, Sigma_d);

    bool anyneg = false;
    for (int i = 0; i < nx; i++) if (h_kpc[i] < 0) anyneg = true;
    if (anyneg) {
      status_code = 'H';
      for (int i = 0; i < nx; i++) rho_cgs[i] = 0.0;
      result = false;
    }
    int i_halfmass = 0;
    {
      double best = std::fabs(r_kpc[0] - r_disk);
      for (int i = 1; i < nx; i++) {
        double d = std::fabs(r_kpc[i] - r_disk);
        if (d < best) { best = d; i_halfmass = i; }
      }
    }
    if (h_kpc[i_halfmass] > r_disk) status_code = 'h';

    for (int i = 0; i < nx; i++) {
      double n_cm3 = rho_cgs[i] / Hmass;
      n[i] = n_cm3 / n0_cm3 * n0;
      l_kpc[i] = P.p_ISM_turbulent_length;
      if (l_kpc[i] > h_kpc[i]) l_kpc[i] = h_kpc[i];  // p_limit_turbulent_scale
      l[i] = l_kpc[i] * h0 / h0_kpc;
      Beq[i] = std::sqrt(4.0 * pi * n[i]) * v[i];
      h[i] = h_kpc[i] * h0 / h0_kpc;
      double Uz_kms = result ? outflow_speed(r_kpc[i], rho_cgs[i], h_kpc[i], v_kms, Rm[i]) : 0.0;  // profiles.f90:225-232
      Uz[i] = Uz_kms / h0_km * h0 * t0_s / t0;
      etat[i] = 1.0 / 3.0 * l[i] * v[i];
      tau[i] = ctau * l[i] / v[i];
      if (h[i] > 1e-70) alp_k[i] = P.C_alp * (l[i] * l[i]) / h[i] * Om[i]; else alp_k[i] = alpceil * v[i];  // Krause
      if (alp_k[i] > alpceil * v[i]) alp_k[i] = alpceil * v[i];  // Alp_ceiling
    }
    return result;
  }

  // -------------------------------------------------------------- floor_field.f90
  // initialize_floor_sign, :116-123
  void initialize_floor_sign() {
    t_last_sign_choice = 0.0;
    Bfloor_sign = random_sign();
    refresh_sign_array = true;
  }
  // update_floor_sign, :99-114 (argument tau*t0_Gyr)
  void update_floor_sign(double time) {
    double mn = tau[0] * t0_Gyr;
    for (int i = 1; i < nx; i++) mn = std::min(mn, tau[i] * t0_Gyr);
    if (time - t_last_sign_choice > P.p_floor_kappa * mn) {
      Bfloor_sign = random_sign();
      t_last_sign_choice = time;
      refresh_sign_array = true;
    }
  }
  // compute_floor_target_field, :34-81
  void compute_floor_target_field(double Delta_r, double* B_floor) {
    for (int i = 0; i < nx; i++) {
      double r = x[i];
      double Ncells = std::fabs(3.0 * r * Delta_r * h[i] / (l[i] * l[i] * l[i]) / (lambda * lambda));
      double brms = P.fmag * Beq[i];
      double b = m_exp(-Delta_r / 2. / std::fabs(r)) * brms / std::sqrt(Ncells) * l[i] / Delta_r * lambda / 3;
      B_floor[i] = b * Bfloor_sign * P.C_floor;
    }
    if ((int)sign_array.size() != nx) refresh_sign_array = true;
    if (refresh_sign_array) {
      sign_array.assign(nx, Bfloor_sign);
      double next_layer = Delta_r;
      for (int i = 1; i <= nx; i++) {
        if (x[i - 1] > next_layer) {
          next_layer = next_layer + Delta_r;
          double s = random_sign();
          for (int j = 0; j < i - 1; j++) sign_array[j] = sign_array[j] * s;
        }
      }
      refresh_sign_array = false;
    }
    for (int i = 0; i < nx; i++) B_floor[i] = B_floor[i] * sign_array[i];
  }

  // -------------------------------------------------------------- seed_field.f90
  // x**n with integer n as gfortran compiles it (libgcc __powidf2: square and multiply from the low bit)
  static double pow_int(double xx, int mm) {
    unsigned int n = mm < 0 ? (unsigned int)(-mm) : (unsigned int)mm;
    double y = (n % 2) ? xx : 1.0;
    while (n >>= 1) {
      xx = xx * xx;
      if (n % 2) y *= xx;
    }
    return mm < 0 ? 1.0 / y : y;
  }
  // init_seed_field, :27-63 ('random', 'fraction', 'decaying')
  void init_seed_field() {
    std::fill(f.begin(), f.end(), 0.0);
    if (P.seed_choice == MAG_SEED_FRACTION) {          // :38-41
      for (int i = 0; i < nx; i++) {
        const double Bseed = P.frac_seed * Beq[i];
        f[i] = -Bseed;
        f[(size_t)nx + i] = Bseed;
      }
      return;
    }
    if (P.seed_choice == MAG_SEED_DECAYING) {          // :42-46
      const double r1 = P.p_r_seed_decay / r_max_kpc;
      for (int i = 0; i < nx; i++) {
        const double Bseed = P.frac_seed * Beq[i];
        const double r = x[i];
        // (1-r)**p_nn_seed: integer power, gfortran multiplies (pow_int below)
        const double shape = -Bseed * r * pow_int(1.0 - r, P.p_nn_seed) * m_exp(-r / r1);
        f[i] = shape;
        f[(size_t)nx + i] = Bseed * r * pow_int(1.0 - r, P.p_nn_seed) * m_exp(-r / r1);
      }
      return;
    }
    for (int var = 0; var < 2; var++)
      for (int i = 0; i < nx; i++) {
        double Bseed = P.frac_seed * Beq[i];
        f[(size_t)var * nx + i] = Bseed * random_normal();
      }
  }

  // -------------------------------------------------------------- gutsdynamo.f90
  // impose_bc, :55-86 (Dirichlet at both ends for Br,Bp; Neumann for alp_m)
  void impose_bc(double* ff) const {
    double* Br = ff; double* Bp = ff + nx; double* am = ff + 2 * (size_t)nx;
    const bool neumann = P.p_neumann_boundary_condition_rmax != 0;
    Br[nxghost] = 0.0; Bp[nxghost] = 0.0;
    if (!neumann) { Br[nx - nxghost - 1] = 0.0; Bp[nx - nxghost - 1] = 0.0; }
    for (int ix = 1; ix <= nxghost; ix++) {
      int a = ix - 1, ma = 2 * (nxghost + 1) - ix - 1;
      int b = nx + 1 - ix - 1, mb = nx + 1 - 2 * (nxghost + 1) + ix - 1;
      Br[a] = -Br[ma]; Bp[a] = -Bp[ma];
      if (neumann) { Br[b] = Br[mb]; Bp[b] = Bp[mb]; }   // :70-72, symmetric about r = R
      else { Br[b] = -Br[mb]; Bp[b] = -Bp[mb]; }
      if (P.Dyn_quench) {                                // :77-82
        am[a] = am[ma];
        am[b] = am[mb];
      }
    }
  }
  // estimate_Bzmod, :101-114
  void estimate_Bzmod() {
    Bzmod.resize(nx, 0.0);
    vec dBrdr(nx);
    xder(F(0), dBrdr.data());
    const double* Br = F(0);
    for (int i = 0; i < nx; i++) Bzmod[i] = lambda * h[i] * std::fabs(Br[i] / x[i] + dBrdr[i]);
    Bzmod[nxghost] = lambda * h[nxghost] * std::fabs(2.0 * dBrdr[nxghost]);
  }

  vec w_dBr, w_d2Br, w_dBp, w_d2Bp, w_dam, w_d2am, w_Bfloor;
  // pde, :130-392 (default branch: FOSA, Dyn_quench, lFloor, no alpha^2)
  void pde(double* ff, double* dfdt) {
    impose_bc(ff);
    alp.resize(nx, 0.0);
    w_dBr.resize(nx); w_d2Br.resize(nx); w_dBp.resize(nx); w_d2Bp.resize(nx); w_dam.resize(nx); w_d2am.resize(nx);
    w_Bfloor.resize(nx);
    const double* Br = ff; const double* Bp = ff + nx; const double* alp_m = ff + 2 * (size_t)nx;
    xder(Br, w_dBr.data()); xder2(Br, w_d2Br.data());
    xder(Bp, w_dBp.data()); xder2(Bp, w_d2Bp.data());
    xder(alp_m, w_dam.data()); xder2(alp_m, w_d2am.data());
    compute_floor_target_field(delta_r_floor, w_Bfloor.data());
    w_dam[nxghost] = 0.0;
    const double lam2 = lambda * lambda;
    const double pi32 = std::pow(pi, 3.0 / 2.0);
    const double R_kappa = P.R_kappa;
    double* d1 = dfdt; double* d2 = dfdt + nx; double* d3 = dfdt + 2 * (size_t)nx;
    for (int i = 0; i < nx; i++) {
      const double r = x[i], hh = h[i], et = etat[i], uz = Uz[i];
      // gutsdynamo.f90:199-209: Bsqtot = Br**2 + Bp**2 + Bzmod**2; Alg_quench takes precedence over Dyn_quench
      const double Bsqtot = Br[i] * Br[i] + Bp[i] * Bp[i] + Bzmod[i] * Bzmod[i];
      const double a = P.Alg_quench ? alp_k[i] / (1.0 + Bsqtot / (Beq[i] * Beq[i])) : (P.Dyn_quench ? alp_k[i] + alp_m[i] : alp_k[i]);
      alp[i] = a;
      const double detatdr = 1.0 / 3 * l[i] * 0.0 + 1.0 / 3 * v[i] * 0.0;  // dvdr = dldr = 0
      const double Dyn_gen = G[i] * a * (hh * hh * hh) / (et * et);
      // compute_floor_source_coefficient, floor_field.f90:83-97
      const double R_U = uz * hh / et;
      const double pi2 = pi / 2;
      const double tU = (1.0 + 4 * C_U * R_U / (pi * pi));
      const double Dyn_crit = -(pi2 * pi2 * pi2 * pi2 * pi2) * (tU * tU);
      const double Afloor = (pi2 * pi2) * et / (hh * hh) * (1.0 + 4 * C_U * R_U / (pi * pi)) * (1.0 - Dyn_gen / Dyn_crit);
      const double Bfloor = w_Bfloor[i];
      d1[i] = -C_U * uz * Br[i] / hh - 2.0 / pi / hh * ctau * a * Bp[i] - pi * pi / 4 / (hh * hh) * (ctau + Rm_inv) * et * Br[i] +
              (ctau + Rm_inv) * et * lam2 * (-Br[i] / (r * r) + w_dBr[i] / r + w_d2Br[i]);
      double v2 = G[i] * Br[i] - C_U * uz * Bp[i] / hh - 2.0 / pi / hh * ctau * a * Br[i] -
                  pi * pi / 4 / (hh * hh) * (ctau + Rm_inv) * et * Bp[i] +
                  (ctau + Rm_inv) * et * lam2 * (-Bp[i] / (r * r) + w_dBp[i] / r + w_d2Bp[i]) +
                  ctau * detatdr * lam2 * (Bp[i] / r + w_dBp[i]) + Afloor * Bfloor;
      v2 = v2 + 2.0 / pi / hh * ctau * a * Br[i];  // .not.Alp_squared
      d2[i] = v2;
      const double lk = h0_kpc / l_kpc[i];
      const double beq2 = Beq[i] * Beq[i];
      d3[i] = -2 * (lk * lk) * et *
                  (ctau * a * (Br[i] * Br[i] + Bp[i] * Bp[i] + 0.0 * (Bzmod[i] * Bzmod[i])) / beq2 +
                   ctau * 3 * et / pi32 / hh * std::sqrt(std::fabs(Dyn_gen)) * Br[i] * Bp[i] / beq2 + Rm_inv * alp_m[i]) -
              C_a * alp_m[i] * uz / hh + R_kappa * et * C_d / (hh * hh) * alp_m[i] +
              R_kappa * et * lam2 * (w_d2am[i] + w_dam[i] / r) + R_kappa * detatdr * lam2 * w_dam[i];
      if (!P.Dyn_quench) d3[i] = 0.0;   // nvar = 2 (var module, :28-43): no alp_m equation; the third array stays at its initial zero
    }
  }

  // check_f_error, :450-463
  void check_f_error(const double* ff, bool* ok) const {
    double m = ff[0];
    for (int i = 1; i < 2 * nx; i++) m = std::max(m, ff[i]);
    if (m > 1e8) *ok = false;
    double m3 = ff[2 * (size_t)nx];
    for (int i = 1; i < nx; i++) m3 = std::max(m3, ff[2 * (size_t)nx + i]);
    if (P.Dyn_quench && m3 > 1000) *ok = false;
  }

  vec w_pdef, w_ftmp;
  // rk, :404-448
  void rk(bool* ok) {
    const double gam1 = 8.0 / 15, gam2 = 5.0 / 12, gam3 = 3.0 / 4, zet1 = -17.0 / 60, zet2 = -5.0 / 12;
    const size_t N = (size_t)3 * nx;
    w_pdef.resize(N); w_ftmp.resize(N);
    double* ff = f.data(); double* pdef = w_pdef.data(); double* ftmp = w_ftmp.data();
    double ttmp;
    stats.point_stages += nx;
    pde(ff, pdef);
    for (size_t i = 0; i < N; i++) ff[i] = ff[i] + dt * gam1 * pdef[i];
    t = t + dt * gam1;
    for (size_t i = 0; i < N; i++) ftmp[i] = ff[i] + dt * zet1 * pdef[i];
    ttmp = t + dt * zet1;
    check_f_error(ff, ok);
    if (!*ok) return;
    stats.point_stages += nx;
    pde(ff, pdef);
    for (size_t i = 0; i < N; i++) ff[i] = ftmp[i] + dt * gam2 * pdef[i];
    t = ttmp + dt * gam2;
    for (size_t i = 0; i < N; i++) ftmp[i] = ff[i] + dt * zet2 * pdef[i];
    ttmp = t + dt * zet2;
    check_f_error(ff, ok);
    if (!*ok) return;
    stats.point_stages += nx;
    pde(ff, pdef);
    for (size_t i = 0; i < N; i++) ff[i] = ftmp[i] + dt * gam3 * pdef[i];
    t = ttmp + dt * gam3;
    check_f_error(ff, ok);
  }

  // -------------------------------------------------------------- ts_arrays.f90
  void reset_ts_arrays() {
    const int nr = P.p_nx_ref;
    ts_prof.assign(MAG_NPROFILE, vec((size_t)nz * nr, INVALID));
    ts_scal.assign(MAG_NSCALAR, vec(nz, INVALID));
    // The reference marks init_it:max_it with '0' only for the FIRST galaxy a rank
    // processes (ts_arrays.f90:99-121); every later galaxy starts from '-'.  The
    // order-independent behaviour ('-') is used.
    ts_status.assign(nz, '-');
    hist_nsteps.assign(nz, 0);
    hist_nx.assign(nz, 0);
  }
  void ts_set(int q, int it, const double* src_interior, int n_int, vec* keep = nullptr) {
    const int nr = P.p_nx_ref;
    rescale_array(src_interior, n_int, &ts_prof[q][(size_t)(it - 1) * nr], nr);
    if (keep) std::copy(&ts_prof[q][(size_t)(it - 1) * nr], &ts_prof[q][(size_t)(it - 1) * nr] + nr, keep->begin());
  }
  // make_ts_arrays, :76-203
  void make_ts_arrays(int it) {
    const int nr = P.p_nx_ref;
    const int ni = nx - 2 * nxghost;
    alp.resize(nx, 0.0);  // only matters when no pde call happened on this grid (nsteps = 0)
    ts_status[it - 1] = status_code;
    ts_scal[MAG_S_DT][it - 1] = t * t0_Gyr;
    vec tmp(nr);
    ts_set(MAG_Q_BR, it, F(0) + nxghost, ni);
    ts_set(MAG_Q_BP, it, F(1) + nxghost, ni);
    ts_set(MAG_Q_BZMOD, it, Bzmod.data() + nxghost, ni);
    if (P.Dyn_quench) ts_set(MAG_Q_ALP_M, it, F(2) + nxghost, ni);   // ts_arrays.f90:132-138
    ts_set(MAG_Q_H, it, h.data() + nxghost, ni);
    ts_set(MAG_Q_OMEGA, it, Om.data() + nxghost, ni);
    ts_set(MAG_Q_SHEAR, it, G.data() + nxghost, ni);
    ts_set(MAG_Q_L, it, l.data() + nxghost, ni);
    ts_set(MAG_Q_V, it, v.data() + nxghost, ni);
    ts_set(MAG_Q_ETAT, it, etat.data() + nxghost, ni);
    ts_set(MAG_Q_TAU, it, tau.data() + nxghost, ni);
    ts_set(MAG_Q_ALP_K, it, alp_k.data() + nxghost, ni);
    ts_set(MAG_Q_UZ, it, Uz.data() + nxghost, ni);
    ts_set(MAG_Q_UR, it, Ur.data() + nxghost, ni);
    ts_set(MAG_Q_N, it, n.data() + nxghost, ni);
    ts_set(MAG_Q_BEQ, it, Beq.data() + nxghost, ni);
    ts_set(MAG_Q_RKPC, it, r_kpc.data() + nxghost, ni);
    ts_set(MAG_Q_OMEGA_H, it, Om_h.data() + nxghost, ni);
    ts_set(MAG_Q_OMEGA_B, it, Om_b.data() + nxghost, ni);
    ts_set(MAG_Q_OMEGA_D, it, Om_d.data() + nxghost, ni);

    const double* oBr = &ts_prof[MAG_Q_BR][(size_t)(it - 1) * nr];
    const double* oBp = &ts_prof[MAG_Q_BP][(size_t)(it - 1) * nr];
    const double* oBz = &ts_prof[MAG_Q_BZMOD][(size_t)(it - 1) * nr];
    const double* rkpc = &ts_prof[MAG_Q_RKPC][(size_t)(it - 1) * nr];
    const double* oh = &ts_prof[MAG_Q_H][(size_t)(it - 1) * nr];
    vec Btot(nr);
    for (int i = 0; i < nr; i++) Btot[i] = std::sqrt(oBr[i] * oBr[i] + oBp[i] * oBp[i] + oBz[i] * oBz[i]);
    int Bmax_idx = 0;
    for (int i = 1; i < nr; i++) if (Btot[i] > Btot[Bmax_idx]) Bmax_idx = i;  // maxloc: first maximum
    ts_scal[MAG_S_BMAX][it - 1] = Btot[Bmax_idx];
    ts_scal[MAG_S_RMAX][it - 1] = rkpc[Bmax_idx];
    ts_scal[MAG_S_BMAX_IDX][it - 1] = (double)(Bmax_idx + 1);
    double s0 = 0, s1 = 0;
    for (int i = 0; i < nr; i++) s0 += rkpc[i] * oh[i];
    double tmp_scalar;
    if (s0 > 0) {
      for (int i = 0; i < nr; i++) s1 += Btot[i] * Btot[i] * rkpc[i] * oh[i];
      tmp_scalar = std::sqrt(s1 / s0);
    } else {
      tmp_scalar = INVALID;
    }
    ts_scal[MAG_S_BEAVG][it - 1] = tmp_scalar;
    double sa = 0, sb = 0;
    for (int i = 0; i < nr; i++) { sa += Btot[i] * rkpc[i]; sb += rkpc[i]; }
    ts_scal[MAG_S_BAVG][it - 1] = sa / sb;
    // ':200-201': the set('alp') is OUTSIDE the if; tmp holds the h profile otherwise
    if (status_code == 'M' || status_code == 'm') {
      ts_set(MAG_Q_ALP, it, alp.data() + nxghost, ni);
    } else {
      std::copy(oh, oh + nr, &ts_prof[MAG_Q_ALP][(size_t)(it - 1) * nr]);
    }
  }

  // -------------------------------------------------------------- dynamo.f90
  // dynamo_run, :39-280.  Returns error (true => nothing is written for the galaxy).
  bool run(int32_t gal_id, int nz_, const double* t_out, const double* inputs_g, size_t col_stride) {
    nz = nz_;
    output_times = t_out;
    test_run = P.p_no_magnetic_fields_test_run != 0;
    bool elliptical = false, ok = true, error = false;
    set_random_seed(gal_id, P.p_random_seed);
    initialize_floor_sign();
    status_code = test_run ? 't' : 'M';
    reset_ts_arrays();
    last_output = false;
    load_input_parameters(inputs_g, col_stride);
    set_input_params(&error);
    if (error) return true;
    construct_grid();
    f.assign((size_t)3 * nx, 0.0);
    alp.assign(nx, 0.0);
    Bzmod.assign(nx, 0.0);
    sign_array.clear();

    bool able = construct_profiles();
    int it = init_it;
    if (!able) {
      finish(it);
      return false;
    }
    if (Mgas_disk < P.Mgas_disk_min) stats.flags |= MAG_FLAG_STALE_PROFILES;
    if (r_disk > P.p_rdisk_min && Mgas_disk > P.Mgas_disk_min) {
      init_seed_field();
      estimate_Bzmod();
      impose_bc(f.data());
    }
    for (it = init_it; it <= max_it; it++) {
      double this_t = t_Gyr;
      (void)this_t;
      if (r_disk < P.p_rdisk_min || Mgas_disk < P.Mgas_disk_min) {
        elliptical = true;
        std::fill(f.begin(), f.end(), 0.0);
      } else {
        if (elliptical) {
          able = construct_profiles();
          init_seed_field();
          estimate_Bzmod();
          impose_bc(f.data());
        }
        if (able) elliptical = false;
      }
      bool next_output = false;
      if (it != init_it && !elliptical) {
        adjust_grid();
        if (nx_overflow) return true;
        impose_bc(f.data());
        able = construct_profiles();
        if (!able) {
          last_output = true;
          break;
        }
      }
      set_timestep(true);
      hist_nsteps[it - 1] = nsteps;
      hist_nx[it - 1] = nx;
      f_snapshot_beginning = f;
      for (int fail_count = 0; fail_count <= P.p_MAX_FAILS; fail_count++) {
        ok = true;
        if (test_run || elliptical || next_output) {
          alp.assign(nx, 0.0);
          Bzmod.assign(nx, 0.0);
          break;
        }
        for (int jt = 1; jt <= nsteps; jt++) {
          this_t = t_Gyr + dt * t0_Gyr * jt;
          estimate_Bzmod();
          update_floor_sign(this_t);
          rk(&ok);
          impose_bc(f.data());
          if (!ok) {
            make_ts_arrays(it);
            break;
          }
        }
        if (ok) break;
        status_code = 'm';
        stats.flags |= MAG_FLAG_RETRY;
        bool timestep_ok = set_timestep(false);
        if (!timestep_ok) {
          status_code = 't';
          ok = false;
          break;
        }
        f = f_snapshot_beginning;
      }
      if (!ok) {
        status_code = 'i';
        make_ts_arrays(it);
        last_output = true;
        std::fill(f.begin(), f.end(), 0.0);
        init_seed_field();
      }
      impose_bc(f.data());
      estimate_Bzmod();
      make_ts_arrays(it);
      if (last_output) break;
      status_code = test_run ? 't' : 'M';
      set_input_params(&error);
      if (error) break;
    }
    if (it > max_it) it = max_it;  // unreachable in practice: last_output exits at max_it
    finish(it);
    return false;
  }
  // finish, :282-300
  void finish(int it) {
    estimate_Bzmod();
    make_ts_arrays(it);
  }
};

}  // namespace oracle

// ====================================================================== C API
extern "C" {

void oracle_params_default(mag_params_t* p) { oracle::params_default(p); }

// Same buffer layouts as mag_solve_batch (include/magnetizer_b200.h).  nthreads<=0:
// hardware concurrency.  Galaxies are farmed dynamically from one atomic counter, the
// same policy as the reference's MPI master/worker (source/distributor.f90:253-384).
int oracle_solve_batch(const mag_params_t* params, int32_t ngal, const int32_t* gal_ids, int32_t nz,
                       const double* t_gyr, const double* inputs, int32_t n_profile_q, const int32_t* profile_q,
                       int32_t n_scalar_q, const int32_t* scalar_q, const mag_outputs_t* out, int32_t nthreads,
                       int32_t* nsteps_hist /* [ngal][nz] or NULL */, int32_t* nx_hist /* [ngal][nz] or NULL */) {
  if (!oracle::params_supported(*params)) return MAG_ERR_UNSUPPORTED;
  for (int q = 0; q < n_profile_q; q++)
    if (profile_q[q] < 0 || profile_q[q] >= MAG_NPROFILE) return MAG_ERR_BAD_ARGUMENT;
  for (int q = 0; q < n_scalar_q; q++)
    if (scalar_q[q] < 0 || scalar_q[q] >= MAG_NSCALAR) return MAG_ERR_BAD_ARGUMENT;
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads <= 0) nthreads = 1;
  if (nthreads > ngal) nthreads = ngal > 0 ? ngal : 1;
  const int nr = params->p_nx_ref;
  std::atomic<int> next(0);
  std::atomic<int> rc(0);
  auto worker = [&]() {
    for (;;) {
      int g = next.fetch_add(1);
      if (g >= ngal) break;
      oracle::Galaxy gal(*params);
      bool err = gal.run(gal_ids[g], nz, t_gyr, inputs + (size_t)g * nz, (size_t)ngal * nz);
      if (gal.nx_overflow) rc = MAG_ERR_NX_OVERFLOW;
      if (out->error) out->error[g] = err ? 1 : 0;
      if (out->nsteps) out->nsteps[g] = gal.nsteps;
      if (out->point_stages) out->point_stages[g] = gal.stats.point_stages;
      if (out->flags) out->flags[g] = gal.stats.flags;
      if (out->status) std::memcpy(out->status + (size_t)g * nz, gal.ts_status.data(), nz);
      if (out->profiles)
        for (int q = 0; q < n_profile_q; q++)
          std::memcpy(out->profiles + ((size_t)q * ngal + g) * nz * nr, gal.ts_prof[profile_q[q]].data(),
                      sizeof(double) * nz * nr);
      if (out->scalars)
        for (int q = 0; q < n_scalar_q; q++)
          std::memcpy(out->scalars + ((size_t)q * ngal + g) * nz, gal.ts_scal[scalar_q[q]].data(), sizeof(double) * nz);
      if (nsteps_hist) std::memcpy(nsteps_hist + (size_t)g * nz, gal.hist_nsteps.data(), sizeof(int32_t) * nz);
      if (nx_hist) std::memcpy(nx_hist + (size_t)g * nz, gal.hist_nx.data(), sizeof(int32_t) * nz);
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++) th.emplace_back(worker);
  for (auto& t : th) t.join();
  return rc.load();
}

// ---- hooks for pinning the third-party restatements (tests/test_oracle_numerics.py)
void oracle_rng_uniforms(int32_t seed_word, int32_t kind, int32_t n, double* out) {
  oracle::Rng r;
  r.seed_all_words(seed_word, kind);
  for (int i = 0; i < n; i++) out[i] = r.uniform();
}
void oracle_rng_normals(int32_t gal_id, int32_t p_random_seed, int32_t kind, int32_t n, double* out) {
  mag_params_t p;
  oracle::params_default(&p);
  p.rng_kind = kind;
  oracle::Galaxy g(p);
  g.set_random_seed(gal_id, p_random_seed);
  for (int i = 0; i < n; i++) out[i] = g.random_normal();
}
// elementary functions of det_math.hpp: 0 log, 1 exp, 2 pow(x,y)
void oracle_det_math(int32_t which, int32_t n, const double* x, const double* y, double* out) {
  for (int i = 0; i < n; i++)
    out[i] = which == 0 ? detmath::det_log(x[i]) : which == 1 ? detmath::det_exp(x[i]) : detmath::det_pow(x[i], y[i]);
}
void oracle_bessel_all(int32_t n, const double* x, double* out /* [4][n] */) {
  for (int i = 0; i < n; i++) detmath::det_bessel_ik01(x[i], out + i, out + n + i, out + 2 * n + i, out + 3 * n + i);
}
double oracle_bessel(int32_t which, double x) {
  switch (which) {
    case 0: return oracle::bessel_I0(x);
    case 1: return oracle::bessel_I1(x);
    case 2: return oracle::bessel_K0(x);
    default: return oracle::bessel_K1(x);
  }
}
// Brent on f(x) = x^p - c over [lo, hi]
int oracle_brent_power(double p, double c, double lo, double hi, double* root, int32_t* iters) {
  int it = 0;
  bool ok = oracle::find_root_brent([&](double x) { return std::pow(x, p) - c; }, lo, hi, root, &it);
  *iters = it;
  return ok ? 1 : 0;
}
int oracle_secant_power(double p, double c, double guess, double step, double* root) {
  return oracle::find_root_secant([&](double x) { return std::pow(x, p) - c; }, guess, step, root) ? 1 : 0;
}
// xder / xder2 on a uniform grid of spacing dx
void oracle_stencils(int32_t nx, double dx, const double* f, double* d1, double* d2) {
  mag_params_t p;
  oracle::params_default(&p);
  oracle::Galaxy g(p);
  g.nx = nx;
  g.x.resize(nx);
  for (int i = 0; i < nx; i++) g.x[i] = i * dx;
  g.xder(f, d1);
  g.xder2(f, d2);
}
void oracle_rescale_array(const double* y, int32_t n, double* y_new, int32_t n_new) {
  oracle::rescale_array(y, n, y_new, n_new);
}
double oracle_constant(int32_t which) {
  switch (which) {
    case 0: return oracle::pi;
    case 1: return oracle::t0_Gyr;
    case 2: return oracle::t0_s;
    case 3: return oracle::n0_cm3;
    case 4: return oracle::h0_km;
    default: return 0;
  }
}
}
