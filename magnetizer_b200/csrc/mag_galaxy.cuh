// mag_galaxy.cuh -- one galaxy history on one warp (sm_100a, FP64).
//
// B200-native re-design of the reference operator dynamo_run (source/dynamo.f90:39-300)
// and everything it calls.  One warp owns one galaxy from its first to its last snapshot:
// radial arrays live in a per-warp workspace (HBM, L1/L2 resident; ~1 KB per array at the
// default grid), radial points are distributed over the 32 lanes, reductions are warp
// shuffles, scalars are warp-uniform registers.  The inherently serial pieces of the
// reference (the hydrostatic chain in r, the RNG stream) run on lane 0.  The hot loop --
// nsteps x 3 Runge-Kutta stages -- has a register/shared-memory fast path in
// mag_rk_fast.cuh; this file holds the generic path that is valid for any nx <= nxcap.
//
// The RHS is evaluated in a hoisted form: everything that is constant during a snapshot
// is folded into per-point coefficients when the ISM profiles are refreshed (all
// divisions, sqrt and exp leave the time loop).  With the reference defaults Uz = 0
// (p_outflow_type='no_outflow'), detatdr = 0, Rm_inv = 0, ctau = 1.
#pragma once
#include "../../include/magnetizer_b200.h"
#include "mag_device_math.cuh"

namespace magdev {

// optional phase clocks (debug builds, -DMAG_PHASE_CLOCKS): cycles lane 0 of every evolve team
// spends in each phase of a snapshot, summed over the launch
#ifdef MAG_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define PHASE_BEGIN long long _pc0 = clock64()
#define PHASE_END(id, lane) do { const long long _pc1 = clock64(); if ((lane) == 0) atomicAdd(&g_phase_clk[id], (unsigned long long)(_pc1 - _pc0)); _pc0 = _pc1; } while (0)
#else
#define PHASE_BEGIN do {} while (0)
#define PHASE_END(id, lane) do {} while (0)
#endif

struct SnapRec;
struct GalHead;
struct ChainJob;

// ------------------------------------------------------------------ workspace
enum WsArray {
  A_X = 0, A_RKPC, A_H, A_N, A_L, A_LKPC, A_ETAT, A_TAU, A_BEQ, A_ALPK, A_OM, A_G, A_OMD, A_OMB, A_OMH, A_GB, A_GH,
  A_F0, A_F1, A_F2, A_FT0, A_FT1, A_FT2, A_P0, A_P1, A_P2, A_FB0, A_FB1, A_FB2, A_BZ, A_ALP, A_SIGN,
  A_OMK, A_GK, A_SIGD, A_SIGS, A_RM, A_RHO, A_HKPC, A_T0, A_T1, A_T2, A_UZ,
  // hoisted RHS coefficients (per snapshot)
  C_IR, C_C1, C_CA, C_CR, C_KDC, C_FB, C_P3, C_QC, C_W,
  // folded coefficients of the register fast path (mag_rk_fast.cuh), zero at ghost points
  C_W0, C_W1, C_W2, C_W3, C_W4, C_W5, C_W6, C_W0A, F_CA, F_GS, C_FBS, F_KDC, F_P3, F_QC, F_AK,
  A_COUNT
};

struct KernelArgs {
  mag_params_t P;
  Units U;
  int32_t nz;
  const double* t_gyr;
  // Every per-galaxy array below is addressed with the CATALOGUE index of the galaxy through a
  // VIRTUAL base pointer: the address the array would start at if its first row belonged to
  // catalogue galaxy 0.  The host subtracts (first galaxy held) * (row size) from the real
  // pointer, so that one kernel serves whole arrays, chunk-local staging buffers and the
  // caller's pinned arrays alike (mag_solver.cu: bind_*).  A null pointer skips that output.
  const int32_t* gal_ids;            // [g]
  const double* in[MAG_NINPUT];      // per input column: [g][nz]
  double* prof[MAG_NPROFILE];        // per REQUESTED profile quantity qq: [g][nz][p_nx_ref]
  double* scal[MAG_NSCALAR];         // per requested scalar quantity: [g][nz]
  char* o_status;                    // [g][nz]
  int32_t* o_nsteps;                 // [g]
  int32_t* o_error;
  uint32_t* o_flags;
  int64_t* o_point_stages;
  const int32_t* order;  // [chunk] processing order of the evolve phase: heavy class first, each class cost sorted
  int32_t n_profile_q, n_scalar_q;
  int32_t profile_q[MAG_NPROFILE];
  int32_t scalar_q[MAG_NSCALAR];
  int32_t g0, g1;        // galaxy range of this chunk (catalogue indices)
  double* ws;            // Phase A: [nwarps][A_COUNT][nxcap]
  double* ws2;           // Phase B, one-warp teams: [nteams][A_COUNT][nxcap]
  double* ws3;           // Phase B, four-warp teams
  int32_t nxcap;
  int32_t* counter;      // Phase A work queue head
  int32_t* counter2;     // Phase B work queue head, one-warp teams (light class)
  int32_t* counter3;     // Phase B work queue head, four-warp teams (heavy class)
  const int32_t* n_heavy;  // galaxies of the heavy class = the first *n_heavy entries of `order`
  unsigned long long* cost_total;  // sum of GalHead::cost over the chunk (profile phase)
  int32_t* n_invalidated;  // galaxies whose evolve phase rewrote rows of the profile phase (status 'i')
  int32_t* fail;         // device-side error flag (nx overflow, arena exhausted)
  int32_t use_fast;      // 1: use the register/smem fast path where it applies
  int32_t* replays;      // diagnostic: fast-path blocks replayed with the generic path
  // snapshot records exchanged between the profile phase and the evolve phase (mag_phases.cuh)
  SnapRec* recs;         // [ngal][nz]
  GalHead* heads;        // [ngal]
  double* arena;         // record arrays
  unsigned long long* arena_top;
  long long arena_cap;   // in doubles
  // hydrostatic-chain jobs exchanged between the two profile passes and k_chains
  ChainJob* jobs;        // [job_cap]
  int32_t* job_count;
  int32_t* jobidx;       // [chunk][nz][2] -> index into jobs or -1
  int32_t job_cap;
  int32_t chain_mode;    // CHAIN_MODE_* of this k_profiles launch
  int32_t pdl_join;      // k_evolve was launched as the programmatic dependent of k_evolve_heavy
};
#define MAG_ERR_ARENA (-6)  // internal: record arena exhausted (the host retries with smaller chunks)

// Per-warp scalar state.  Every lane holds an identical copy ("module variables" of the
// reference: input_params, grid, floor_field, messages::status_code).
struct Gal {
  const KernelArgs* a;
  double* W;
  RngState* rng;
  int lane, g;
  // input_params
  int iread, init_it, max_it, nsteps;
  bool last_output;
  double tsnap, dt, t, time_between_inputs, t_Gyr;
  double r_disk, v_disk, r_bulge, v_bulge, r_halo, v_halo, nfw_cs1, Mgas_disk, Mstars_disk, SFR, Mhalo, Mstars_bulge;
  bool is_central;
  // grid
  int nx;
  double dx, lambda, dr_kpc, r_max_kpc;
  double seed_r1;             // p_r_seed_decay / r_max_kpc of the grid the next seed field is drawn on (from the snapshot record)
  double r_max_kpc_history;   // p_rmax_over_rdisk x the largest r_disk of the input rows (input_parameters.f90:173)
  // profiles
  double delta_r_floor, v_code, tau_min, minh, maxetat;
  // floor
  double t_last_sign_choice;
  bool refresh_sign, sign_valid;
  int sign_nx;
  // hydrostatic chain handling (mag_phases.cuh): inline / export jobs / import results
  int chain_mode, chain_it, chain_kind, chain_job;
  bool chain_diverged;
  // first / last snapshot whose output rows this phase has written (fill_unwritten)
  int w0, w1;
  // status / stats
  char status;
  uint32_t flags;
  long long point_stages;

  __device__ __forceinline__ double* A(int id) const { return W + (size_t)id * a->nxcap; }
};

// ------------------------------------------------------------------ small pieces
// input column c at snapshot (1-based) it
__device__ __forceinline__ double input_at(const Gal& G, int c, int it) {
  const KernelArgs* a = G.a;
  return a->in[c][(size_t)G.g * a->nz + (it - 1)];
}

// load_input_parameters (input_parameters.f90:136-192): init_it / max_it
__device__ inline void load_input_parameters(Gal& G) {
  const int nz = G.a->nz;
  G.init_it = -1;
  G.max_it = nz;
  double mx = input_at(G, MAG_IN_R_DISK, 1);
  for (int i = 2; i <= nz; i++) mx = fmax(mx, input_at(G, MAG_IN_R_DISK, i));
  G.r_max_kpc_history = mx * G.a->P.p_rmax_over_rdisk;
  for (int i = 1; i <= nz; i++) {
    const double rd = input_at(G, MAG_IN_R_DISK, i);
    if (rd >= G.a->P.p_rdisk_min && G.init_it < 0) G.init_it = i;
    if (rd < 0 && G.init_it > 0) { G.max_it = i - 1; break; }
  }
  G.iread = G.init_it - 1;
}

// set_input_params (input_parameters.f90:202-255)
__device__ inline bool set_input_params(Gal& G) {
  G.iread += 1;
  if (G.iread < 0 || G.max_it - G.init_it == 0) { G.status = 'p'; return true; }
  const double cur = G.a->t_gyr[G.iread - 1];
  if (G.iread < G.max_it) {
    G.time_between_inputs = __dadd_rn(G.a->t_gyr[G.iread], -cur);
    G.t = 0;
  } else {
    G.last_output = true;
  }
  G.t_Gyr = cur;
  const int it = G.iread;
  G.r_disk = input_at(G, MAG_IN_R_DISK, it);   G.v_disk = input_at(G, MAG_IN_V_DISK, it);
  G.r_bulge = input_at(G, MAG_IN_R_BULGE, it); G.v_bulge = input_at(G, MAG_IN_V_BULGE, it);
  G.r_halo = input_at(G, MAG_IN_R_HALO, it);   G.v_halo = input_at(G, MAG_IN_V_HALO, it);
  G.nfw_cs1 = input_at(G, MAG_IN_NFW_CS1, it); G.Mgas_disk = input_at(G, MAG_IN_MGAS_DISK, it);
  G.Mstars_disk = input_at(G, MAG_IN_MSTARS_DISK, it); G.SFR = input_at(G, MAG_IN_SFR, it);
  G.Mhalo = input_at(G, MAG_IN_MHALO, it);     G.Mstars_bulge = input_at(G, MAG_IN_MSTARS_BULGE, it);
  G.is_central = input_at(G, MAG_IN_CENTRAL, it) > 0.1;
  return false;
}

// construct_grid (grid.f90:48-81)
__device__ inline void construct_grid(Gal& G) {
  const int nxphys = G.a->P.p_nx_ref;
  G.nx = nxphys + 2 * MAG_NGHOST;
  G.dx = (1.0 - MAG_R_IN) / (double)(nxphys - 1);
  // p_use_fixed_physical_grid (grid.f90:69-73): the largest disc of the history sets the grid once and for all
  G.r_max_kpc = G.a->P.p_use_fixed_physical_grid ? G.r_max_kpc_history : __dmul_rn(G.a->P.p_rmax_over_rdisk, G.r_disk);
  double* x = G.A(A_X);
  double* rk = G.A(A_RKPC);
  for (int i = G.lane; i < G.nx; i += 32) {
    const double xi = __dadd_rn(__dadd_rn(MAG_R_IN, -__dmul_rn((double)MAG_NGHOST, G.dx)), __dmul_rn((double)i, G.dx));
    x[i] = xi;
    rk[i] = __dmul_rn(xi, G.r_max_kpc);
  }
  G.dr_kpc = __dmul_rn(G.dx, G.r_max_kpc);
  G.lambda = 1.0 / G.r_max_kpc;
  __syncwarp();
}

// rescale_array (interpolation.f90:128-167 + interpolate :25-76): value i of the n_new
// resampling of y[0..n-1].  Index search restated in closed form: the reference's
// running `j` is the number of old abscissae (excluding the first) strictly below xi.
__device__ __forceinline__ double rescale_at(const double* y, int n, int i, int n_new) {
  if (n == n_new) return y[i];
  const double dn = (double)(n - 1);
  const double xi = (double)i / (double)(n_new - 1);
  int j = (int)(xi * dn);
  if (j > n - 2) j = n - 2;
  while (j + 1 <= n - 1 && (double)(j + 1) / dn < xi) j++;
  while (j >= 1 && (double)j / dn >= xi) j--;
  if (j > n - 2) j = n - 2;
  const double x0 = (double)j / dn, x1 = (double)(j + 1) / dn;
  return __dadd_rn(y[j], __dmul_rn(y[j + 1] - y[j], xi - x0) / (x1 - x0));
}

// impose_bc (gutsdynamo.f90:55-86): Dirichlet for Br,Bp at both ends, Neumann for alp_m
__device__ inline void impose_bc(const Gal& G, double* f0, double* f1, double* f2) {
  const int nx = G.nx, l = G.lane;
  // p_neumann_boundary_condition_rmax (:63-74): Br, Bp symmetric about r = R and the outer point left alone;
  // without Dyn_quench there is no alp_m column (nvar = 2, :77-82): f2 stays at its initial zero
  const bool neumann = G.a->P.p_neumann_boundary_condition_rmax != 0, dyn = G.a->P.Dyn_quench != 0;
  if (l < MAG_NGHOST) {
    const int ix = l + 1;                                   // 1-based ghost index
    const int a = ix - 1, ma = 2 * (MAG_NGHOST + 1) - ix - 1;
    const int b = nx - ix, mb = nx + 1 - 2 * (MAG_NGHOST + 1) + ix - 1;
    f0[a] = -f0[ma]; f1[a] = -f1[ma];
    if (neumann) { f0[b] = f0[mb]; f1[b] = f1[mb]; }
    else { f0[b] = -f0[mb]; f1[b] = -f1[mb]; }
    if (dyn) { f2[a] = f2[ma]; f2[b] = f2[mb]; }
  } else if (l == MAG_NGHOST) {
    f0[MAG_NGHOST] = 0.0; f1[MAG_NGHOST] = 0.0;
    if (!neumann) { f0[nx - MAG_NGHOST - 1] = 0.0; f1[nx - MAG_NGHOST - 1] = 0.0; }
  }
  __syncwarp();
}
// NB the mirror sources (interior points) are never BC targets, and the zeroed points are
// never mirror sources, so the two branches above commute like the reference's loop.

// xder / xder2 at point i (deriv.f90:22-89); fac1 = 1/(60 dx), fac2 = 1/(180 dx^2)
__device__ __forceinline__ double xder_at(const double* f, int i, int nx, double fac) {
  if (i >= 3 && i < nx - 3)
    return fac * (45. * (f[i + 1] - f[i - 1]) - 9. * (f[i + 2] - f[i - 2]) + (f[i + 3] - f[i - 3]));
  if (i == 0) return fac * (-147. * f[0] + 360. * f[1] - 450. * f[2] + 400. * f[3] - 225. * f[4] + 72. * f[5] - 10. * f[6]);
  if (i == 1) return fac * (-10. * f[0] - 77. * f[1] + 150. * f[2] - 100. * f[3] + 50. * f[4] - 15. * f[5] + 2. * f[6]);
  if (i == 2) return fac * (2. * f[0] - 24. * f[1] - 35. * f[2] + 80. * f[3] - 30. * f[4] + 8. * f[5] - f[6]);
  const double* b = f + nx - 1;
  if (i == nx - 1) return fac * (147. * b[0] - 360. * b[-1] + 450. * b[-2] - 400. * b[-3] + 225. * b[-4] - 72. * b[-5] + 10. * b[-6]);
  if (i == nx - 2) return fac * (10. * b[0] + 77. * b[-1] - 150. * b[-2] + 100. * b[-3] - 50. * b[-4] + 15. * b[-5] - 2. * b[-6]);
  return fac * (-2. * b[0] + 24. * b[-1] + 35. * b[-2] - 80. * b[-3] + 30. * b[-4] - 8. * b[-5] + b[-6]);
}
__device__ __forceinline__ double xder2_at(const double* f, int i, int nx, double fac) {
  if (i >= 3 && i < nx - 3)
    return fac * (-490. * f[i] + 270. * (f[i + 1] + f[i - 1]) - 27. * (f[i + 2] + f[i - 2]) + 2. * (f[i + 3] + f[i - 3]));
  if (i == 0) return fac * (812. * f[0] - 3132. * f[1] + 5265. * f[2] - 5080. * f[3] + 2970. * f[4] - 972. * f[5] + 137. * f[6]);
  if (i == 1) return fac * (137. * f[0] - 147. * f[1] - 255. * f[2] + 470. * f[3] - 285. * f[4] + 93. * f[5] - 13. * f[6]);
  if (i == 2) return fac * (-13. * f[0] + 228. * f[1] - 420. * f[2] + 200. * f[3] + 15. * f[4] - 12. * f[5] + 2. * f[6]);
  const double* b = f + nx - 1;
  if (i == nx - 1) return fac * (812. * b[0] - 3132. * b[-1] + 5265. * b[-2] - 5080. * b[-3] + 2970. * b[-4] - 972. * b[-5] + 137. * b[-6]);
  if (i == nx - 2) return fac * (137. * b[0] - 147. * b[-1] - 255. * b[-2] + 470. * b[-3] - 285. * b[-4] + 93. * b[-5] - 13. * b[-6]);
  return fac * (-13. * b[0] + 228. * b[-1] - 420. * b[-2] + 200. * b[-3] + 15. * b[-4] - 12. * b[-5] + 2. * b[-6]);
}

// ------------------------------------------------------------------ rotation curves
// rotationCurves.f90.  Disk: :27-97, bulge :99-143, NFW halo :192-221, adiabatic
// contraction :223-322 (one Brent solve per radius).
struct RotCtx {
  double r_disk, v_disk, r_bulge, v_bulge, r_halo, v_halo, nfw_cs1, rmin, f_b;
  double disk_norm;   // I0K0-I1K1 at y50
  double halo_A;      // v_halo^2/(ln(1+c) - c/(1+c))
  double halo_c;
};
__device__ inline double disk_omega(const RotCtx& c, double r) {
  const double TOO_SMALL = (double)2e-7f;
  if (c.r_disk < c.rmin) return 0.0;
  const double rs = c.r_disk * MAG_DISK_SCALE_TO_HALFMASS;
  const double y = (fabs(r) > TOO_SMALL) ? fabs(r) / rs : TOO_SMALL / rs;
  double i0, i1, k0, k1;
  detmath::det_bessel_ik01(y, &i0, &i1, &k0, &k1);
  const double A = sqrt((i0 * k0 - i1 * k1) / c.disk_norm);
  return A * c.v_disk / c.r_disk;
}
__device__ inline double bulge_omega(const RotCtx& c, double r) {
  if (c.v_bulge < (double)1e-3f) return 0.0;
  const double a_to_r50 = sqrt(2.0) - 1.0;
  const double a = a_to_r50 * c.r_bulge;
  const double y = fabs(r) / a;
  const double y50 = 1.0 / a_to_r50;
  const double Constant = c.v_bulge * (1.0 + y50) / sqrt(y50);
  return Constant * sqrt(y) / (y + 1.0) / fabs(r);
}
__device__ __forceinline__ double halo_velocity(const RotCtx& c, double r) {
  const double y = fabs(r) / c.r_halo;
  const double cy = c.halo_c * y;
  const double B = (detmath::det_log(1.0 + cy) - cy / (1.0 + cy)) / y;
  return sqrt(fabs(c.halo_A * B));
}
__device__ inline double halo_velocity_contracted(const RotCtx& c, double r, double od, double ob, bool at_centre, uint32_t* flags) {
  // Vd2, Vb2 depend on |r| only: the reference recomputes disk_/bulge_rotation_curve at every
  // residual evaluation (rotationCurves.f90:314-320); the values are bit-identical to the
  // Omega_d, Omega_b of this radius (both use abs(r)), which the caller passes in
  const double Vb2 = (ob * r) * (ob * r), Vd2 = (od * r) * (od * r);
  auto fun = [&](double r0) {
    const double hv = halo_velocity(c, r0);
    const double V02 = hv * hv;
    return r0 * r0 * V02 - r * r * (Vd2 + Vb2 + (1.0 - c.f_b) * V02 * (r0 / r));
  };
  const double fun_min = fun(r);
  if (fun_min > 0.0) {
    // reference reads an uninitialised r0 here (rotationCurves.f90:266-269); r0 = r is used
    if (!at_centre) *flags |= MAG_FLAG_UNINIT_R0;
    return halo_velocity(c, r);
  }
  int i;
  double p15 = 1.0;
  for (i = 1; i <= 10; i++) {
    p15 *= 1.5;
    if (fun(r * p15) / fun_min < 0) break;
  }
  if (i > 10) p15 *= 1.5;
  double r0 = r;
  if (!brent_root(fun, r, r * p15, &r0)) return CUDART_NAN;  // reference: MPI_ABORT
  return sqrt(r0 / r) * halo_velocity(c, r0);
}

// regularize (profiles.f90:275-331) at one point
__device__ __forceinline__ double regularize_at(double rx, double r_xi, double Omega_xi, double Omega) {
  const double small_factor = (double)1e-15f;
  double k;
  if (rx > small_factor * r_xi) {
    const double q = r_xi / rx;
    k = q * q;
  } else {
    Omega = 0.0;
    k = 0.0;
  }
  const double e = (k < (double)1e12f) ? detmath::det_exp(-k) : 0.0;
  return e * (Omega - Omega_xi) + Omega_xi;
}

// ------------------------------------------------------------------ hydrostatic equilibrium
// pressureEquilibrium.f90:51-311,425-594 with surface_density.f90:26-128.  The pressure
// equation is algebraic with the defaults: f(h) = A + B h/(h_m+h) + C h/(h_*+h) + D h - E/h.
struct HydroConst {
  double h_star, h_m, Kgas;  // Kgas: Pgas = Kgas * rho
};
__device__ __forceinline__ double density_to_scaleheight(double h_d, double Sigma_d) {
  return Sigma_d * MAG_MSUN_SI / MAG_KPC_SI / MAG_KPC_SI / h_d / 2.0 / MAG_KPC_SI * 1e-3;
}
__device__ inline double exp_surface_density(double rs, double r, double M) {
  return M / 2. / MAG_PI / (rs * rs) * detmath::det_exp(-r / rs);
}
__device__ inline double molecular_to_diffuse_ratio(const mag_params_t& P, double rdisk, double Sigma_g, double Sigma_star) {
  const double h_star_SI = P.p_stellarHeightToRadiusScale * MAG_DISK_SCALE_TO_HALFMASS * rdisk * MAG_KPC_SI;
  const double v_gas_SI = P.p_ISM_sound_speed_km_s * 1e3;
  const double Sg = Sigma_g * MAG_MSUN_SI / (MAG_KPC_SI * MAG_KPC_SI);
  const double Ss = Sigma_star * MAG_MSUN_SI / (MAG_KPC_SI * MAG_KPC_SI);
  double v_star_SI = sqrt(MAG_PI * MAG_G_SI * h_star_SI * Ss);
  if (v_star_SI < v_gas_SI) v_star_SI = v_gas_SI;
  const double Pnot = MAG_PI / 2.0 * MAG_G_SI * Sg * (Sg + (v_gas_SI / v_star_SI) * Ss) * 10.0;
  return detmath::det_pow(Pnot / P.p_Rmol_P0, P.p_Rmol_alpha);
}

// One hydrostatic-equilibrium chain (solve_hydrostatic_equilibrium_numerical,
// pressureEquilibrium.f90:51-251): serial over radius (the bracket of radius i+1 comes from
// the root at radius i), therefore executed by ONE thread: lane 0 of the galaxy's warp
// (inline mode) or one thread per (galaxy, snapshot) job of k_chains, 32 chains per warp.
struct ChainIn {
  int n;                       // grid size
  double r_disk, Mstars_bulge;
  const double *r;             // abs(r_kpc)
  const double *Sigma_d, *Sigma_star, *Rm, *Om_h, *G_h, *Om_b, *G_b;
  double *h_d, *rho_d;         // out
};
#define CHAIN_OK 1        // the solve did not give up and no scale height is negative
#define CHAIN_STATUS_P 2  // status 'P' was raised (log-extrapolated scale height)
#define CHAIN_SECANT 4    // the secant fallback was used
// Returns CHAIN_* bits.
__device__ inline int hydro_chain(const mag_params_t& P, const ChainIn& c) {
  const int n = c.n, ng = MAG_NGHOST;
  const double *Sigma_d = c.Sigma_d, *Sigma_star = c.Sigma_star, *Rm = c.Rm;
  const double *Om_h = c.Om_h, *G_h = c.G_h, *Om_b = c.Om_b, *G_b = c.G_b, *r = c.r;
  double *h_d = c.h_d, *rho_d = c.rho_d;
  int bits = 0;
  const double rs = MAG_DISK_SCALE_TO_HALFMASS * c.r_disk;
  const double h_star = c.r_disk * MAG_DISK_SCALE_TO_HALFMASS * P.p_stellarHeightToRadiusScale;
  const double h_m = c.r_disk * MAG_DISK_SCALE_TO_HALFMASS * P.p_molecularHeightToRadiusScale;
  const double cs = P.p_ISM_sound_speed_km_s * 1e5;
  // computes_midplane_ISM_pressure_from_B_and_rho with B = 0 (pressureEquilibrium.f90:561-573)
  const double Kgas = (P.p_ISM_kappa / 3.0 * (1. + 2. * P.p_ISM_xi) + 1.0 / P.p_ISM_gamma) * (cs * cs);
  const double kmkpc2 = (1e3 / MAG_KPC_SI) * (1e3 / MAG_KPC_SI);
  const bool use_bulge = P.p_use_Pbulge && c.Mstars_bulge > 1e3;
  const bool use_dm = P.p_use_Pdm != 0;
  int error_count = 0;
  double minimum_h = 1e-3 * c.r_disk, maximum_h = 2.0 * c.r_disk, guess_h = (double)0.050f;
  bool gave_up = false;
  for (int i = ng + 2; i <= n && !gave_up; i++) {
    const int k = i - 1;
    // pressure_equation (pressureEquilibrium.f90:262-311) = P1 + P2 - Pgas in the reference's
    // operation order; only leading factors that the reference itself evaluates first are
    // hoisted out of the root-finder iteration, so every residual is bit-identical to a
    // literal evaluation (P2 is exactly -0 with p_enable_P2 = F)
    const double Sd_SI = Sigma_d[k] * MAG_MSUN_SI / MAG_KPC_SI / MAG_KPC_SI;
    const double Ss_SI = Sigma_star[k] * MAG_MSUN_SI / MAG_KPC_SI / MAG_KPC_SI;
    const double pre = MAG_PI * MAG_G_SI * Sd_SI * 10.0;
    const double Pd = pre * Sd_SI * 1.0 * 0.5;
    const double cPm = pre * Sd_SI * Rm[k];
    const double cPs = pre * Ss_SI;
    const double cPb = use_bulge ? Om_b[k] * (1.5 * Om_b[k] + G_b[k]) * kmkpc2 * Sd_SI : 0.0;
    const double cPh = use_dm ? Om_h[k] * (1.5 * Om_h[k] + G_h[k]) * kmkpc2 * Sd_SI : 0.0;
    auto pe = [&](double h) {
      const double I_stars = h / (h_star + h);
      const double I_m = h / (h_m + h);
      const double Pm = cPm * I_m;
      const double Pstars = cPs * I_stars;
      const double Pbulge = use_bulge ? cPb * h * MAG_KPC_SI * 10.0 : 0.0;
      const double Pdm = use_dm ? cPh * h * MAG_KPC_SI * 10.0 : 0.0;
      const double P1 = Pd + Pm + Pstars + Pbulge + Pdm;
      const double rho_cgs = Sd_SI / h / 2.0 / MAG_KPC_SI * 1e-3;
      return P1 - Kgas * rho_cgs;
    };
    const double pe_min = pe(minimum_h), pe_max = pe(maximum_h);
    double root = -17.0;
    if (signbit(pe_min) != signbit(pe_max)) {
      if (!brent_root(pe, minimum_h, maximum_h, &root)) { gave_up = true; break; }
    } else {
      bits |= CHAIN_SECANT;
      const bool ok = secant_root(pe, guess_h, 1e-6, &root);
      if (!ok || root < 0) {
        if (error_count < 4 && i > ng + 2 + 2) {
          bits |= CHAIN_STATUS_P;
          double e = (detmath::det_log(h_d[k - 1]) - detmath::det_log(h_d[k - 2])) / (r[k - 1] - r[k - 2]) * (r[k] - r[k - 1]);
          e = e + detmath::det_log(h_d[k - 1]);
          root = detmath::det_exp(e);
          error_count++;
        } else {
          gave_up = true;
          break;
        }
      }
    }
    h_d[k] = root;
    rho_d[k] = density_to_scaleheight(root, Sigma_d[k]);
    if (i != n) {
      guess_h = root * detmath::det_exp((r[k + 1] - r[k]) / rs);
      maximum_h = guess_h * (1.0 + 25.0 / 100.);
      minimum_h = guess_h * (1.0 - 25.0 / 100.);
    }
    if (rho_d[k] < P.p_minimum_density) {
      rho_d[k] = P.p_minimum_density;
      h_d[k] = density_to_scaleheight(rho_d[k], Sigma_d[k]);
    }
  }
  if (gave_up) {
    for (int j = 0; j < n; j++) { h_d[j] = -1; rho_d[j] = 0.0; }
    return bits;
  }
  for (int i = 1; i <= ng; i++) h_d[i - 1] = h_d[2 * ng + 2 - i - 1];
  h_d[ng] = h_d[ng + 1] + (r[ng] - r[ng + 1]) * (h_d[ng + 1] - h_d[ng + 2]) / (r[ng + 1] - r[ng + 2]);
  for (int i = 1; i <= ng + 1; i++) rho_d[i - 1] = density_to_scaleheight(h_d[i - 1], Sigma_d[i - 1]);
  // construct_profiles tests any(h_kpc < 0) (profiles.f90:155-160): the linear extrapolation to the centre above can be
  // negative although every root was found (a scale height that rises steeply from a tiny central value).  The arrays
  // stay as they are -- the reference writes them out -- and the caller raises status 'H' as for a solve that gave up.
  if (h_d[ng] < 0) return bits;
  return bits | CHAIN_OK;
}

#define CHAIN_MODE_INLINE 0
#define CHAIN_MODE_EXPORT 1
#define CHAIN_MODE_IMPORT 2
struct Gal;
__device__ void chain_export(Gal& G, const double* absr, uint32_t rot_flags);   // mag_phases.cuh
__device__ int chain_lookup(Gal& G);                          // job of the current construct_profiles call or -1
__device__ const double* chain_job_array(const Gal& G, int k);
__device__ uint32_t chain_job_flags(const Gal& G);
__device__ int chain_import(Gal& G);                          // < 0: not available

// outflow_speed (outflow.f90:24-121) at one radius: r, h in kpc, rho in g/cm^3, vt in km/s -> km/s.
// Literals without d0 are single precision in the reference.
struct OutflowCtx { double r_disk, v_disk, SFR; };
__device__ inline double outflow_speed(const mag_params_t& P, const OutflowCtx& c, double r, double rho, double hk, double vt, double Rm) {
  switch (P.outflow_type) {
    case MAG_OUTFLOW_NONE: return r * 0.0;
    case MAG_OUTFLOW_VTURB: return vt;
    case MAG_OUTFLOW_WIND: {
      const double VDISK_MIN = (double)0.01f;
      const double fm = Rm / (1.0 + Rm);
      const double beta = detmath::det_pow(fmax(VDISK_MIN, c.v_disk) / P.p_outflow_Vhot, P.p_outflow_alphahot);
      double u = 0.5 * P.p_outflow_nu0 * beta * hk * fm;
      u = u * MAG_KM_KPC / (1.e9 * 365.25 * 24 * 3600);
      return u;
    }
    default: break;
  }
  const double rs = c.r_disk * MAG_DISK_SCALE_TO_HALFMASS;
  const double nn = rho / MAG_HMASS;
  if (!(nn > 0.0)) return 0.0;
  const double constant = (double)51.3f * detmath::det_pow(P.p_outflow_Lsn / 1e38, (double)(1.f / 3.f)) * P.p_outflow_fOB / 0.7 *
                          (P.p_outflow_etaSN / (double)9.4e-3f) * (P.p_tOB / 3.0) * (40.0 / P.p_N_SN1OB);
  const double q = rs / 3.0;
  double u = constant * (1.0 / (q * q)) * c.SFR;
  u = u * detmath::det_pow(nn, (double)(-1.f / 3.f));
  u = u * detmath::det_pow(hk / (double)0.2f, (double)(4.f / 3.f));
  u = u * detmath::det_exp(-r / rs);
  if (P.outflow_type == MAG_OUTFLOW_SUPERBUBBLE) u = (P.p_outflow_hot_gas_density / rho) * u;
  return u;
}

// ------------------------------------------------------------------ construct_profiles
// profiles.f90:33-272 (B absent).  Also prepares what set_timestep needs (min h>0, max
// etat, min tau) and, when the profiles are usable, the hoisted RHS coefficients.
__device__ inline bool construct_profiles(Gal& G) {
  const mag_params_t& P = G.a->P;
  const Units& U = G.a->U;
  const int nx = G.nx, lane = G.lane;
  double *x = G.A(A_X), *rk = G.A(A_RKPC);
  double *Om_d = G.A(A_OMD), *Om_b = G.A(A_OMB), *Om_h = G.A(A_OMH), *G_b = G.A(A_GB), *G_h = G.A(A_GH);
  double *Omk = G.A(A_OMK), *Gk = G.A(A_GK), *Om = G.A(A_OM), *Gc = G.A(A_G);
  double *h = G.A(A_H), *nn = G.A(A_N), *l = G.A(A_L), *lk = G.A(A_LKPC), *etat = G.A(A_ETAT), *tau = G.A(A_TAU);
  double *Beq = G.A(A_BEQ), *alp_k = G.A(A_ALPK), *Uz = G.A(A_UZ);
  bool result = true;
  const double r_disk_min = G.r_max_kpc * P.rmin_over_rmax;
  G.delta_r_floor = P.p_floor_kappa * P.p_ISM_turbulent_length / G.r_max_kpc;

  RotCtx c;
  c.r_disk = G.r_disk; c.v_disk = G.v_disk; c.r_bulge = G.r_bulge; c.v_bulge = G.v_bulge;
  c.r_halo = G.r_halo; c.v_halo = G.v_halo; c.nfw_cs1 = G.nfw_cs1; c.rmin = r_disk_min;
  c.f_b = (G.Mstars_disk + G.Mstars_bulge + G.Mgas_disk) / G.Mhalo;
  {
    double i0, i1, k0, k1;
    detmath::det_bessel_ik01(1.0 / MAG_DISK_SCALE_TO_HALFMASS, &i0, &i1, &k0, &k1);
    c.disk_norm = i0 * k0 - i1 * k1;
    c.halo_c = 1.0 / G.nfw_cs1;
    c.halo_A = G.v_halo * G.v_halo / (detmath::det_log(1.0 + c.halo_c) - c.halo_c / (1.0 + c.halo_c));
  }
  const bool with_halo = G.is_central || !P.p_ignore_satellite_DM_haloes;
  const double rreg = P.p_rreg_to_rdisk * G.r_disk;
  double best = CUDART_INF; int ireg = 0x7fffffff;
  double bestm = CUDART_INF; int ihalf = 0x7fffffff;
  uint32_t fl = 0;
  // pass 2 of the profile phase re-uses the rotation curves pass 1 computed for this very
  // grid and snapshot (the halo contraction -- one Brent solve per radius -- dominates the cost)
  G.chain_job = (G.chain_mode == CHAIN_MODE_IMPORT) ? chain_lookup(G) : -1;
  const double* cached = (G.chain_job >= 0) ? chain_job_array(G, 10) : nullptr;   // Om_d, Om_b, Om_h before regularisation
  if (cached && lane == 0) fl = chain_job_flags(G);
  double *raw_b = G.A(A_T1), *raw_h = G.A(A_T2);
  for (int i = lane; i < nx; i += 32) {
    const double r = rk[i];
    double oh = 0.0;
    if (cached) {
      Om_d[i] = cached[i]; Om_b[i] = cached[nx + i]; oh = cached[2 * nx + i];
    } else {
      Om_d[i] = disk_omega(c, r);
      Om_b[i] = bulge_omega(c, r);
      if (with_halo) oh = halo_velocity_contracted(c, fabs(r), Om_d[i], Om_b[i], i == MAG_NGHOST, &fl) / r;
    }
    raw_b[i] = Om_b[i]; raw_h[i] = oh;
    Om_h[i] = oh;
    Omk[i] = sqrt(Om_d[i] * Om_d[i] + Om_b[i] * Om_b[i] + oh * oh);
    const double d = fabs(r - rreg);
    if (d < best) { best = d; ireg = i; }
    const double dm = fabs(r - G.r_disk);
    if (dm < bestm) { bestm = dm; ihalf = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(FULL, fl, o);
  G.flags |= fl;
  warp_argmin_first(best, ireg);
  warp_argmin_first(bestm, ihalf);
  __syncwarp();
  {
    const double r_xi = fabs(rk[ireg]);
    const double oxi_k = Omk[ireg], oxi_h = Om_h[ireg], oxi_b = Om_b[ireg];
    __syncwarp();
    for (int i = lane; i < nx; i += 32) {
      const double rx = fabs(rk[i]);
      Omk[i] = regularize_at(rx, r_xi, oxi_k, Omk[i]);
      Om_h[i] = regularize_at(rx, r_xi, oxi_h, Om_h[i]);
      Om_b[i] = regularize_at(rx, r_xi, oxi_b, Om_b[i]);
    }
  }
  __syncwarp();
  const double fac1 = 1.0 / (60. * (x[1] - x[0]));
  const double conv = 1.0 / U.h0_km * U.t0_s;  // Om = Om_kmskpc/h0_km*h0_kpc*t0_s/t0
  for (int i = lane; i < nx; i += 32) {
    double g = xder_at(Omk, i, nx, fac1) * x[i];
    G_h[i] = xder_at(Om_h, i, nx, fac1) * x[i];
    G_b[i] = xder_at(Om_b, i, nx, fac1) * x[i];
    if (g > 0.0) g = 0.0;
    Gk[i] = g;
    Om[i] = Omk[i] / U.h0_km * U.t0_s;
    Gc[i] = g / U.h0_km * U.t0_s;
  }
  (void)conv;
  const double v_kms = P.p_ISM_sound_speed_km_s * P.p_ISM_kappa;
  G.v_code = v_kms / U.h0_km * U.t0_s;
  __syncwarp();

  // negligible-disk trap (profiles.f90:131-140)
  if (G.r_disk < r_disk_min || G.Mgas_disk < P.Mgas_disk_min) {
    for (int i = lane; i < nx; i += 32) { nn[i] = 0; l[i] = 0; h[i] = 0; etat[i] = 0; alp_k[i] = 0; Uz[i] = 0; }
    G.minh = 1e10; G.maxetat = 0.0;
    __syncwarp();
    return true;
  }

  // hydrostatic equilibrium
  double *Sigma_d = G.A(A_SIGD), *Sigma_star = G.A(A_SIGS), *Rm = G.A(A_RM), *absr = G.A(A_T0);
  {
    const double rs = MAG_DISK_SCALE_TO_HALFMASS * G.r_disk;
    const double rs_g = P.p_gasScaleRadiusToStellarScaleRadius_ratio * rs;
    for (int i = lane; i < nx; i += 32) {
      const double r = fabs(rk[i]);
      absr[i] = r;
      const double Sg = exp_surface_density(rs_g, r, G.Mgas_disk);
      const double Ss = exp_surface_density(rs, r, G.Mstars_disk);
      const double rm = molecular_to_diffuse_ratio(P, G.r_disk, Sg, Ss);
      Sigma_star[i] = Ss;
      Rm[i] = rm;
      Sigma_d[i] = Sg / (rm + 1.0);
    }
  }
  __syncwarp();
  // ---- the hydrostatic chain: exported as a job (profile pass 1), imported from the job's
  // results (pass 2), or solved inline by lane 0
  if (G.chain_mode == CHAIN_MODE_EXPORT) {
    chain_export(G, absr, fl);
    return true;   // pass 1 only walks the grid recurrence; pass 2 finishes the snapshot
  }
  int bits = -1;
  if (G.chain_mode == CHAIN_MODE_IMPORT) bits = chain_import(G);
  if (bits < 0) {
    if (lane == 0) {
      ChainIn c;
      c.n = nx; c.r_disk = G.r_disk; c.Mstars_bulge = G.Mstars_bulge; c.r = absr;
      c.Sigma_d = Sigma_d; c.Sigma_star = Sigma_star; c.Rm = Rm;
      c.Om_h = Om_h; c.G_h = G_h; c.Om_b = Om_b; c.G_b = G_b;
      c.h_d = G.A(A_HKPC); c.rho_d = G.A(A_RHO);
      bits = hydro_chain(P, c);
    }
    bits = __shfl_sync(FULL, bits, 0);
  }
  const int okc = (bits & CHAIN_OK) ? 1 : 0;
  if (bits & CHAIN_SECANT) G.flags |= MAG_FLAG_SECANT;
  if (bits & CHAIN_STATUS_P) G.status = 'P';
  __syncwarp();
  double *h_kpc = G.A(A_HKPC), *rho = G.A(A_RHO);
  if (!okc) {  // any(h_kpc<0)
    G.status = 'H';
    for (int i = lane; i < nx; i += 32) rho[i] = 0.0;
    result = false;
    __syncwarp();
  }
  if (h_kpc[ihalf] > G.r_disk) G.status = 'h';

  const double v = G.v_code;
  double mn_h = 1e10, mx_e = -CUDART_INF, mn_tau = CUDART_INF;
  const OutflowCtx oc = {G.r_disk, G.v_disk, G.SFR};
  const double* Rmol = G.A(A_RM);
  for (int i = lane; i < nx; i += 32) {
    // VERTICAL VELOCITY PROFILE (profiles.f90:225-232)
    const double Uz_kms = result ? outflow_speed(P, oc, rk[i], rho[i], h_kpc[i], v_kms, Rmol[i]) : 0.0;
    Uz[i] = Uz_kms / U.h0_km * U.t0_s;
    const double n_cm3 = rho[i] / MAG_HMASS;
    const double ni = n_cm3 / U.n0_cm3;
    nn[i] = ni;
    double lki = P.p_ISM_turbulent_length;
    if (lki > h_kpc[i]) lki = h_kpc[i];
    lk[i] = lki;
    l[i] = lki;
    Beq[i] = sqrt(4.0 * MAG_PI * ni) * v;
    const double hi = h_kpc[i];
    h[i] = hi;
    const double et = 1.0 / 3.0 * lki * v;
    etat[i] = et;
    const double ta = lki / v;
    tau[i] = ta;
    double ak;
    if (hi > 1e-70) ak = P.C_alp * (lki * lki) / hi * Om[i]; else ak = v;
    if (ak > v) ak = v;
    alp_k[i] = ak;
    if (hi > 0 && hi < mn_h) mn_h = hi;
    mx_e = fmax(mx_e, et);
    mn_tau = fmin(mn_tau, __dmul_rn(ta, U.t0_Gyr));
  }
  G.minh = warp_min(mn_h);
  G.maxetat = warp_max(mx_e);
  G.tau_min = warp_min(mn_tau);
  __syncwarp();

  if (result) {
    // hoisted RHS coefficients (reference gutsdynamo.f90:242-293,
    // floor_field.f90:34-97)
    double *cIR = G.A(C_IR), *cC1 = G.A(C_C1), *cCA = G.A(C_CA), *cCR = G.A(C_CR), *cKDC = G.A(C_KDC);
    double *cFB = G.A(C_FB), *cP3 = G.A(C_P3), *cQC = G.A(C_QC), *cW = G.A(C_W);
    const double lam = G.lambda, lam2 = lam * lam, Dr = G.delta_r_floor;
    const double pi2 = MAG_PI / 2, pi2_5 = pi2 * pi2 * pi2 * pi2 * pi2;
    const double pi32 = MAG_PI * sqrt(MAG_PI);
    for (int i = lane; i < nx; i += 32) {
      const double r = x[i], hi = h[i], et = etat[i], li = l[i], be = Beq[i], uz = Uz[i];
      cIR[i] = 1.0 / r;
      // vertical advection by the outflow (gutsdynamo.f90:243,253: -C_U*Uz*B/h; :285: -C_a*alp_m*Uz/h) joins the
      // vertical-diffusion coefficients; with 'no_outflow' uz = 0 and the terms vanish exactly
      const double c1 = MAG_PI * MAG_PI / 4 / (hi * hi) * et + MAG_C_U * uz / hi;
      cC1[i] = c1;
      cCA[i] = 2.0 / MAG_PI / hi;
      cCR[i] = et * lam2;
      const double kD = Gc[i] * (hi * hi * hi) / (et * et);
      // compute_floor_source_coefficient (floor_field.f90:83-97): R_U = Uz*h/etat,
      // Dyn_crit = -(pi/2)^5 (1 + 4 C_U R_U/pi^2)^2, A_floor = (pi/2)^2 etat/h^2 (1 + 4 C_U R_U/pi^2)(1 - Dyn_gen/Dyn_crit)
      const double tU = 1.0 + 4 * MAG_C_U * (uz * hi / et) / (MAG_PI * MAG_PI);
      cKDC[i] = kD / (pi2_5 * (tU * tU));
      const double Ncells = fabs(3.0 * r * Dr * hi / (li * li * li) / lam2);
      const double b = detmath::det_exp(-Dr / 2. / fabs(r)) * (P.fmag * be) / sqrt(Ncells) * li / Dr * lam / 3 * P.C_floor;
      cFB[i] = (pi2 * pi2) * et / (hi * hi) * tU * b;
      const double lkr = 1.0 / lk[i];
      cP3[i] = -2 * (lkr * lkr) * et / (be * be);
      cQC[i] = 3 * et / pi32 / hi * sqrt(fabs(kD));
      cW[i] = P.R_kappa * et * (-(MAG_PI * MAG_PI)) / (hi * hi) - MAG_C_A * uz / hi;
    }
    __syncwarp();
  }
  return result;
}

// set_timestep (input_parameters.f90:65-134), p_variable_timesteps=.true.
__device__ inline void set_timestep(Gal& G, bool have_profiles) {
  const mag_params_t& P = G.a->P;
  G.tsnap = G.time_between_inputs / G.a->U.t0_Gyr;
  if (have_profiles) {
    const double maxv = G.v_code, maxe = G.maxetat;
    if (fabs(maxv) > 1e-20 && fabs(maxe) > 1e-20) {
      const double d1 = __dmul_rn(P.p_courant_v, G.dx) / G.lambda / maxv;
      const double d2 = __dmul_rn(P.p_courant_eta, __dmul_rn(G.minh, G.minh)) / maxe;
      G.dt = fmin(d1, d2);
      const double nsteps_f = G.tsnap / G.dt;
      G.nsteps = (int)fmin((double)1e9f, nsteps_f);
    } else {
      G.nsteps = P.p_nsteps_max;
    }
  } else {
    G.nsteps = P.nsteps_0;  // reduce_timestep path: dt is NOT recomputed
  }
  if (G.nsteps > P.p_nsteps_max) G.nsteps = P.p_nsteps_max;
}

// adjust_grid (grid.f90:143-297), default switches -- the GRID part (x, r_kpc, dx, dr_kpc,
// r_max_kpc, lambda, nx), which depends on the r_disk history only.  Reports what the
// reference does to the field array so that the evolve phase can repeat it on f:
// *nx_mid = grid size after the optional scale-back (== nx before it if none),
// *scaled / *extended = which of the two f transformations happen.
// Returns false on nx overflow (reference: stop, grid.f90:222-226).
__device__ inline bool adjust_grid_geometry(Gal& G, int* nx_mid, bool* scaled, bool* extended) {
  const mag_params_t& P = G.a->P;
  const double GRID_ADJ_THRES = (double)0.05f;
  const int nx_def = P.p_nx_ref + 2 * MAG_NGHOST, lane = G.lane, ng = MAG_NGHOST;
  double *x = G.A(A_X), *rk = G.A(A_RKPC);
  double* tmp = G.A(A_T0);
  *scaled = false; *extended = false;
  *nx_mid = G.nx;
  if (P.p_use_fixed_physical_grid) return true;   // grid.f90:159-162: do nothing
  if (G.nx != nx_def) {  // scale back to the default resolution
    const int nx_old = G.nx;
    for (int i = lane; i < nx_old; i += 32) tmp[i] = rk[i];
    __syncwarp();
    G.nx = nx_def;
    const int n_old = nx_old - ng, n_new = nx_def - ng;
    for (int i = lane; i < n_new; i += 32) rk[ng + i] = rescale_at(tmp + ng, n_old, i, n_new);
    __syncwarp();
    G.dr_kpc = rk[ng + 1] - rk[ng];
    if (lane == 0) for (int i = ng; i >= 1; i--) rk[i - 1] = rk[i] - G.dr_kpc;
    __syncwarp();
    for (int i = lane; i < G.nx; i += 32) x[i] = rk[i] / G.r_max_kpc;
    __syncwarp();
    *scaled = true;
    // dx is NOT updated here
  }
  *nx_mid = G.nx;
  const double prev = G.r_max_kpc;
  G.r_max_kpc = P.p_rmax_over_rdisk * G.r_disk;
  const double relvar = fabs(G.r_max_kpc - prev) / prev;
  if (relvar < GRID_ADJ_THRES) { G.r_max_kpc = prev; return true; }
  if (G.r_max_kpc >= prev) {
    int nx_new = G.nx;
    double cand = prev;
    const double base = MAG_R_IN / G.dx - (double)ng;
    while (cand < G.r_max_kpc) {
      nx_new += 1;
      cand = (base + ((double)(nx_new - ng) - 1.0)) * G.dr_kpc;
      if (nx_new > P.p_nx_MAX + 1) break;
    }
    if (nx_new > P.p_nx_MAX || nx_new > G.a->nxcap) return false;
    G.r_max_kpc = cand;
    const int nx = G.nx;
    if (lane == 0) for (int i = nx; i <= nx_new; i++) rk[i - 1] = rk[i - 2] + G.dr_kpc;
    __syncwarp();
    for (int i = lane; i < nx_new; i += 32) x[i] = rk[i] / G.r_max_kpc;
    __syncwarp();
    G.dx = x[1] - x[0];
    G.nx = nx_new;
    *extended = true;
    __syncwarp();
  } else {
    const double dr_new = G.dr_kpc * G.r_max_kpc / prev;
    for (int i = lane; i < G.nx; i += 32) rk[i] = rk[i] * G.r_max_kpc / prev;
    G.dr_kpc = dr_new;
    __syncwarp();
  }
  G.lambda = 1.0 / G.r_max_kpc;
  return true;
}

// The FIELD part of adjust_grid.  Scale-back (grid.f90:167-193 -> rescale_f_array,
// interpolation.f90:170-198): the interior of every column is resampled linearly from
// nx_old-3 to nx_def-3 points; the inner ghost zone is left for impose_bc.
__device__ inline void adjust_f_scaleback(Gal& G, int nx_old, int nx_def) {
  const int lane = G.lane, ng = MAG_NGHOST;
  double *f[3] = {G.A(A_F0), G.A(A_F1), G.A(A_F2)};
  double *tmp[3] = {G.A(A_T0), G.A(A_T1), G.A(A_T2)};
  for (int i = lane; i < nx_old; i += 32) { tmp[0][i] = f[0][i]; tmp[1][i] = f[1][i]; tmp[2][i] = f[2][i]; }
  __syncwarp();
  const int n_old = nx_old - ng, n_new = nx_def - ng;
  for (int i = lane; i < n_new; i += 32) {
    f[0][ng + i] = rescale_at(tmp[0] + ng, n_old, i, n_new);
    f[1][ng + i] = rescale_at(tmp[1] + ng, n_old, i, n_new);
    f[2][ng + i] = rescale_at(tmp[2] + ng, n_old, i, n_new);
  }
  if (lane < ng) { f[0][lane] = 0.0; f[1][lane] = 0.0; f[2][lane] = 0.0; }
  __syncwarp();
}
// Extension (grid.f90:262-276): points up to nx-nxghost-1 are kept, the rest follows 1/r.
// rk = r_kpc of the EXTENDED grid (its first nx points are the old grid).
__device__ inline void adjust_f_extend(Gal& G, int nx, int nx_new) {
  const int lane = G.lane, ng = MAG_NGHOST;
  const double* rk = G.A(A_RKPC);
  const int keep = nx - ng - 1;  // f(:nx-nxghost-1,:) kept
  for (int v = 0; v < 3; v++) {
    double* f = G.A(A_F0 + v);
    const double edge = f[keep - 1];
    __syncwarp();
    for (int i = nx - ng - 1 + lane; i < nx_new; i += 32) f[i] = edge * rk[keep - 1] / rk[i];
  }
  __syncwarp();
}

// ------------------------------------------------------------------ floor field signs
// random_sign (random.f90:121-129), lane 0 draws, all lanes get it
__device__ __forceinline__ double random_sign(Gal& G) {
  double s = 0.0;
  if (G.lane == 0) s = copysign(1.0, rng_uniform_lane0(G.rng) - 0.5);
  return __shfl_sync(FULL, s, 0);
}
// the sign_array refresh inside compute_floor_target_field (floor_field.f90:55-79).  The
// stored array is the PRODUCT Bfloor_sign*sign_array(i) that multiplies the target field
// (Bfloor_sign squared cancels).
__device__ inline void refresh_signs(Gal& G) {
  const int nx = G.nx;
  double* sg = G.A(A_SIGN);
  const double* x = G.A(A_X);
  for (int i = G.lane; i < nx; i += 32) sg[i] = 1.0;
  const double Dr = G.delta_r_floor;
  double next_layer = Dr;
  for (int i = 0; i < nx; i++) {
    if (x[i] > next_layer) {
      next_layer = __dadd_rn(next_layer, Dr);
      const double s = random_sign(G);
      if (s < 0) for (int j = G.lane; j < i; j += 32) sg[j] = -sg[j];
    }
  }
  G.refresh_sign = false;
  G.sign_nx = nx;
  __syncwarp();
}

// x**n with integer n as gfortran compiles it (libgcc __powidf2: square and multiply from the low bit)
__device__ __forceinline__ double pow_int(double x, int m) {
  unsigned int n = m < 0 ? (unsigned int)(-m) : (unsigned int)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y *= x;
  }
  return m < 0 ? 1.0 / y : y;
}
// init_seed_field (seed_field.f90:27-63): 'random' (sequential draws on lane 0), 'fraction', 'decaying'
__device__ inline void init_seed_field(Gal& G) {
  const int nx = G.nx;
  double *f0 = G.A(A_F0), *f1 = G.A(A_F1), *f2 = G.A(A_F2);
  const double* Beq = G.A(A_BEQ);
  const double fs = G.a->P.frac_seed;
  for (int i = G.lane; i < nx; i += 32) f2[i] = 0.0;
  if (G.a->P.seed_choice == MAG_SEED_FRACTION) {            // :38-41
    for (int i = G.lane; i < nx; i += 32) { const double Bseed = fs * Beq[i]; f0[i] = -Bseed; f1[i] = Bseed; }
  } else if (G.a->P.seed_choice == MAG_SEED_DECAYING) {     // :42-46
    const double* x = G.A(A_X);
    const double r1 = G.seed_r1;
    const int nn = G.a->P.p_nn_seed;
    for (int i = G.lane; i < nx; i += 32) {
      const double Bseed = fs * Beq[i], r = x[i];
      f0[i] = -Bseed * r * pow_int(1.0 - r, nn) * detmath::det_exp(-r / r1);
      f1[i] = Bseed * r * pow_int(1.0 - r, nn) * detmath::det_exp(-r / r1);
    }
  } else if (G.lane == 0) {
    for (int i = 0; i < nx; i++) f0[i] = (fs * Beq[i]) * rng_normal_lane0(G.rng);
    for (int i = 0; i < nx; i++) f1[i] = (fs * Beq[i]) * rng_normal_lane0(G.rng);
  }
  __syncwarp();
}

// estimate_Bzmod (gutsdynamo.f90:101-114)
__device__ inline void estimate_Bzmod(Gal& G) {
  const int nx = G.nx;
  const double *f0 = G.A(A_F0), *x = G.A(A_X), *h = G.A(A_H);
  double* bz = G.A(A_BZ);
  const double fac1 = 1.0 / (60. * (x[1] - x[0]));
  for (int i = G.lane; i < nx; i += 32) {
    const double d = xder_at(f0, i, nx, fac1);
    double v = G.lambda * h[i] * fabs(f0[i] / x[i] + d);
    if (i == MAG_NGHOST) v = G.lambda * h[i] * fabs(2.0 * d);
    bz[i] = v;
  }
  __syncwarp();
}

// ------------------------------------------------------------------ RHS + RK3 (generic path)
// One evaluation of pde (gutsdynamo.f90:130-392, default branch) in hoisted form for the
// points of this lane; result into P0..P2, alp into A_ALP.
__device__ inline void pde_generic(Gal& G) {
  const int nx = G.nx;
  double *f0 = G.A(A_F0), *f1 = G.A(A_F1), *f2 = G.A(A_F2);
  impose_bc(G, f0, f1, f2);
  if (G.refresh_sign || G.sign_nx != nx) refresh_signs(G);
  const double* x = G.A(A_X);
  const double ddx = x[1] - x[0];
  const double fac1 = 1.0 / (60. * ddx), fac2 = 1.0 / (180. * (ddx * ddx));
  const double *cIR = G.A(C_IR), *cC1 = G.A(C_C1), *cCA = G.A(C_CA), *cCR = G.A(C_CR), *cKDC = G.A(C_KDC);
  const double *cFB = G.A(C_FB), *cP3 = G.A(C_P3), *cQC = G.A(C_QC), *cW = G.A(C_W);
  const double *Gc = G.A(A_G), *alp_k = G.A(A_ALPK), *sg = G.A(A_SIGN);
  double *p0 = G.A(A_P0), *p1 = G.A(A_P1), *p2 = G.A(A_P2), *alp = G.A(A_ALP);
  const double Rk = G.a->P.R_kappa;
  const bool alg = G.a->P.Alg_quench != 0, dyn = G.a->P.Dyn_quench != 0;
  const double *bz = G.A(A_BZ), *beq = G.A(A_BEQ);
  for (int i = G.lane; i < nx; i += 32) {
    const double Br = f0[i], Bp = f1[i], am = f2[i];
    const double dBr = xder_at(f0, i, nx, fac1), d2Br = xder2_at(f0, i, nx, fac2);
    const double dBp = xder_at(f1, i, nx, fac1), d2Bp = xder2_at(f1, i, nx, fac2);
    double dam = xder_at(f2, i, nx, fac1);
    const double d2am = xder2_at(f2, i, nx, fac2);
    if (i == MAG_NGHOST) dam = 0.0;
    const double ir = cIR[i], c1 = cC1[i], cr = cCR[i];
    // total alpha: dynamical quenching, or -- Alg_quench -- alp_k/(1 + Bsqtot/Beq^2) with Bsqtot = Br^2+Bp^2+Bzmod^2
    // (gutsdynamo.f90:199-209; Bzmod is the estimate of the start of the step, dynamo.f90:179)
    const double a = alg ? alp_k[i] / (1.0 + (Br * Br + Bp * Bp + bz[i] * bz[i]) / (beq[i] * beq[i])) : (dyn ? alp_k[i] + am : alp_k[i]);
    alp[i] = a;
    p0[i] = -c1 * Br - cCA[i] * a * Bp + cr * (d2Br + ir * (dBr - Br * ir));
    p1[i] = Gc[i] * Br - c1 * Bp + cr * (d2Bp + ir * (dBp - Bp * ir)) + sg[i] * cFB[i] * (1.0 + cKDC[i] * a);
    p2[i] = cP3[i] * (a * (Br * Br + Bp * Bp) + cQC[i] * sqrt(fabs(a)) * Br * Bp) + cW[i] * am +
            Rk * cr * (d2am + dam * ir);
    if (!dyn) p2[i] = 0.0;   // nvar = 2 (var module, gutsdynamo.f90:28-43): no alp_m equation, the third array stays zero
  }
  G.point_stages += nx;
  __syncwarp();
}

// check_f_error (gutsdynamo.f90:450-463): signed maxima over ALL points incl. ghosts
__device__ inline bool f_blew_up(const Gal& G) {
  const int nx = G.nx;
  const double *f0 = G.A(A_F0), *f1 = G.A(A_F1), *f2 = G.A(A_F2);
  bool bad = false;
  const bool dyn = G.a->P.Dyn_quench != 0;   // :460: the alp_m test only with Dyn_quench
  for (int i = G.lane; i < nx; i += 32) bad = bad || (f0[i] > 1e8) || (f1[i] > 1e8) || (dyn && f2[i] > 1000.0);
  return __any_sync(FULL, bad);
}

// rk (gutsdynamo.f90:404-448): 3-stage 2N-storage Runge-Kutta.  Returns ok.
__device__ inline bool rk_generic(Gal& G) {
  const double gam[3] = {8.0 / 15, 5.0 / 12, 3.0 / 4}, zet[2] = {-17.0 / 60, -5.0 / 12};
  const int nx = G.nx;
  double *f[3] = {G.A(A_F0), G.A(A_F1), G.A(A_F2)}, *ft[3] = {G.A(A_FT0), G.A(A_FT1), G.A(A_FT2)};
  double *p[3] = {G.A(A_P0), G.A(A_P1), G.A(A_P2)};
  double ttmp = 0.0;
  for (int s = 0; s < 3; s++) {
    pde_generic(G);
    const double dg = __dmul_rn(G.dt, gam[s]);
    const double dz = s < 2 ? __dmul_rn(G.dt, zet[s]) : 0.0;
    for (int v = 0; v < 3; v++)
      for (int i = G.lane; i < nx; i += 32) {
        const double base = (s == 0) ? f[v][i] : ft[v][i];
        const double fn = base + dg * p[v][i];
        f[v][i] = fn;
        if (s < 2) ft[v][i] = fn + dz * p[v][i];
      }
    G.t = __dadd_rn(s == 0 ? G.t : ttmp, dg);
    if (s < 2) ttmp = __dadd_rn(G.t, dz);
    __syncwarp();
    if (f_blew_up(G)) return false;
  }
  return true;
}

// ------------------------------------------------------------------ outputs
// make_ts_arrays (ts_arrays.f90:76-203): everything of snapshot `it` (1-based) that the
// caller requested, resampled to p_nx_ref points.
__device__ inline const double* profile_source(const Gal& G, int q) {
  switch (q) {
    case MAG_Q_BR: return G.A(A_F0);
    case MAG_Q_BP: return G.A(A_F1);
    case MAG_Q_ALP_M: return G.A(A_F2);
    case MAG_Q_BZMOD: return G.A(A_BZ);
    case MAG_Q_H: return G.A(A_H);
    case MAG_Q_OMEGA: return G.A(A_OM);
    case MAG_Q_OMEGA_B: return G.A(A_OMB);
    case MAG_Q_OMEGA_D: return G.A(A_OMD);
    case MAG_Q_OMEGA_H: return G.A(A_OMH);
    case MAG_Q_SHEAR: return G.A(A_G);
    case MAG_Q_L: return G.A(A_L);
    case MAG_Q_ETAT: return G.A(A_ETAT);
    case MAG_Q_TAU: return G.A(A_TAU);
    case MAG_Q_ALP_K: return G.A(A_ALPK);
    case MAG_Q_ALP: return G.A(A_ALP);
    case MAG_Q_N: return G.A(A_N);
    case MAG_Q_BEQ: return G.A(A_BEQ);
    case MAG_Q_RKPC: return G.A(A_RKPC);
    case MAG_Q_UZ: return G.A(A_UZ);
    default: return nullptr;  // v (constant), Ur (zero)
  }
}

// true for the quantities whose rows the evolve phase writes (they depend on f); all other
// profiles depend on the ISM profiles only and are written by the profile phase
__host__ __device__ __forceinline__ bool is_field_quantity(int q) {
  return q == MAG_Q_BR || q == MAG_Q_BP || q == MAG_Q_ALP_M || q == MAG_Q_BZMOD || q == MAG_Q_ALP;
}

// make_ts_arrays (ts_arrays.f90:76-203), rows of snapshot `it` that depend on the ISM profiles only
__device__ inline void write_profile_rows(Gal& G, int it) {
  const KernelArgs* a = G.a;
  const int nr = a->P.p_nx_ref, ni = G.nx - 2 * MAG_NGHOST, ng = MAG_NGHOST, lane = G.lane;
  const int nz = a->nz;
  G.w0 = min(G.w0, it); G.w1 = max(G.w1, it);
  for (int qq = 0; qq < a->n_profile_q; qq++) {
    const int q = a->profile_q[qq];
    if (is_field_quantity(q) || !a->prof[qq]) continue;
    const double* src = profile_source(G, q);
    double* dst = a->prof[qq] + ((size_t)G.g * nz + (it - 1)) * nr;
    const double cval = (q == MAG_Q_V) ? G.v_code : 0.0;
    for (int i = lane; i < nr; i += 32) dst[i] = src ? rescale_at(src + ng, ni, i, nr) : cval;
  }
  __syncwarp();
}

// make_ts_arrays, rows that depend on the field: Br, Bp, alp_m, Bzmod, alp, the scalars and
// the status character
__device__ inline void write_field_rows(Gal& G, int it) {
  const KernelArgs* a = G.a;
  const int nr = a->P.p_nx_ref, ni = G.nx - 2 * MAG_NGHOST, ng = MAG_NGHOST, lane = G.lane;
  const int nz = a->nz;
  const size_t g = (size_t)G.g;
  const bool alp_ok = (G.status == 'M' || G.status == 'm');
  G.w0 = min(G.w0, it); G.w1 = max(G.w1, it);
  if (a->o_status && lane == 0) a->o_status[g * nz + it - 1] = G.status;
  for (int qq = 0; qq < a->n_profile_q; qq++) {
    int q = a->profile_q[qq];
    if (!is_field_quantity(q) || !a->prof[qq]) continue;
    if (q == MAG_Q_ALP && !alp_ok) q = MAG_Q_H;  // ts_arrays.f90:200-201: stale tmp (= h)
    const double* src = profile_source(G, q);
    double* dst = a->prof[qq] + (g * nz + (it - 1)) * nr;
    if (q == MAG_Q_ALP_M && !a->P.Dyn_quench) {   // ts_arrays.f90:132-138: no 'alp_m' row without Dyn_quench
      for (int i = lane; i < nr; i += 32) dst[i] = MAG_INVALID;
      continue;
    }
    for (int i = lane; i < nr; i += 32) dst[i] = rescale_at(src + ng, ni, i, nr);
  }
  if (a->n_scalar_q > 0 && a->scal[0]) {
    const double *f0 = G.A(A_F0) + ng, *f1 = G.A(A_F1) + ng, *bz = G.A(A_BZ) + ng, *rk = G.A(A_RKPC) + ng, *h = G.A(A_H) + ng;
    double bmax = -CUDART_INF; int bidx = 0x7fffffff;
    double s_rh = 0, s_b2rh = 0, s_br = 0, s_r = 0;
    for (int i = lane; i < nr; i += 32) {
      const double br = rescale_at(f0, ni, i, nr), bp = rescale_at(f1, ni, i, nr), bzz = rescale_at(bz, ni, i, nr);
      const double r = rescale_at(rk, ni, i, nr), hh = rescale_at(h, ni, i, nr);
      const double bt = sqrt(br * br + bp * bp + bzz * bzz);
      if (bt > bmax) { bmax = bt; bidx = i; }
      s_rh += r * hh;
      s_b2rh += bt * bt * r * hh;
      s_br += bt * r;
      s_r += r;
    }
    warp_argmax_first(bmax, bidx);
    // an all-NaN field never passes `bt > bmax`: maxloc then stays at the first point (as the oracle's loop does) -- the
    // index below must not be the 'nothing found' sentinel
    if (bidx == 0x7fffffff) { bidx = 0; bmax = CUDART_NAN; }
    s_rh = warp_sum(s_rh); s_b2rh = warp_sum(s_b2rh); s_br = warp_sum(s_br); s_r = warp_sum(s_r);
    const double rmax = rescale_at(rk, ni, bidx, nr);
    if (lane == 0) {
      for (int qq = 0; qq < a->n_scalar_q; qq++) {
        double v = MAG_INVALID;
        switch (a->scalar_q[qq]) {
          case MAG_S_DT: v = G.t * a->U.t0_Gyr; break;
          case MAG_S_RMAX: v = rmax; break;
          case MAG_S_BMAX: v = bmax; break;
          case MAG_S_BMAX_IDX: v = (double)(bidx + 1); break;
          case MAG_S_BAVG: v = s_br / s_r; break;
          case MAG_S_BEAVG: v = (s_rh > 0) ? sqrt(s_b2rh / s_rh) : MAG_INVALID; break;
          default: break;
        }
        a->scal[qq][g * nz + it - 1] = v;
      }
    }
  }
  __syncwarp();
}

// fills a galaxy's outputs with "never written" markers (reset_ts_arrays, ts_arrays.f90:205-210)
__device__ inline void reset_outputs(const Gal& G) {
  const KernelArgs* a = G.a;
  const int nr = a->P.p_nx_ref, nz = a->nz, lane = G.lane;
  const size_t g = (size_t)G.g;
  for (int qq = 0; qq < a->n_profile_q; qq++) {
    if (!a->prof[qq]) continue;
    double* dst = a->prof[qq] + g * nz * nr;
    for (int i = lane; i < nz * nr; i += 32) dst[i] = MAG_INVALID;
  }
  for (int qq = 0; qq < a->n_scalar_q; qq++)
    if (a->scal[qq])
      for (int i = lane; i < nz; i += 32) a->scal[qq][g * nz + i] = MAG_INVALID;
  // '-' everywhere: the reference's '0' marking only happens for the first galaxy a rank
  // processes (ts_arrays.f90:99-121), an order dependence that is not reproduced
  if (a->o_status)
    for (int i = lane; i < nz; i += 32) a->o_status[g * nz + i] = '-';
}


// "Never written" markers for the rows a phase owns and did not write: snapshots outside
// [G.w0, G.w1] (a phase writes a contiguous range of snapshots).  field_part: Br, Bp, alp_m,
// Bzmod, alp, the scalars and the status (evolve phase); otherwise the ISM-profile rows
// (profile phase).  Every output row is thus stored exactly once -- this matters when the
// outputs live in mapped host memory and every store crosses PCIe.
__device__ inline void fill_unwritten(const Gal& G, bool field_part) {
  const KernelArgs* a = G.a;
  const int nr = a->P.p_nx_ref, nz = a->nz, lane = G.lane;
  const size_t g = (size_t)G.g;
  const int w0 = G.w1 >= G.w0 ? G.w0 : nz + 1, w1 = G.w1 >= G.w0 ? G.w1 : nz;   // nothing written: everything is "before"
  for (int part = 0; part < 2; part++) {
    const int s0 = part == 0 ? 1 : w1 + 1, s1 = part == 0 ? w0 - 1 : nz;   // snapshots s0..s1 (1-based)
    if (s1 < s0) continue;
    for (int qq = 0; qq < a->n_profile_q; qq++) {
      if (is_field_quantity(a->profile_q[qq]) != field_part || !a->prof[qq]) continue;
      double* dst = a->prof[qq] + (g * nz + (s0 - 1)) * nr;
      for (int i = lane; i < (s1 - s0 + 1) * nr; i += 32) dst[i] = MAG_INVALID;
    }
    if (field_part) {
      for (int qq = 0; qq < a->n_scalar_q; qq++)
        if (a->scal[qq])
          for (int i = s0 - 1 + lane; i < s1; i += 32) a->scal[qq][g * nz + i] = MAG_INVALID;
      if (a->o_status)
        for (int i = s0 - 1 + lane; i < s1; i += 32) a->o_status[g * nz + i] = '-';
    }
  }
  __syncwarp();
}

}  // namespace magdev
