// mag_rk_fast.cuh -- the time-stepping loop of one snapshot (dynamo.f90:168-212):
// nsteps x { update_floor_sign ; rk (3 stages, gutsdynamo.f90:404-448) ; impose_bc }.
//
// Register-resident fast path on a TEAM of 64 threads (2 warps, one CTA) per galaxy.
// Thread t owns PPL consecutive radial points (2t, 2t+1 for PPL = 2, nx <= 128): the state
// f = (Br, Bp, alp_m), the 2N-storage register ftmp and the per-point coefficients stay in
// registers for the whole snapshot (PPL = 2), or the 7 stencil weights come from shared
// memory (PPL = 3, 4: nx <= 192, 256).  Per stage each thread publishes its points to a
// double-buffered shared-memory line, reads the 3+3 halo points of its neighbours after ONE
// CTA barrier, and evaluates the RHS in a folded form:
//
//   L(f)_i   = sum_k w_k(i) f_{i+k}            7 per-point weights = cr*(xder2 + xder/r - 1/r^2) - c1
//   dBr/dt   = L(Br) - ca*alp*Bp
//   dBp/dt   = L(Bp) + G*Br + fb*(1 + kdc*alp)                       (floor forcing folded)
//   dalp_m/dt= p3*(alp*(Br^2+Bp^2) + qc*sqrt|alp|*Br*Bp) + Rk*sum_{k!=0} w_k am_{i+k} + w0a*am_i
//
// (gutsdynamo.f90:242-293 with the defaults; ~50 FP64 instructions per point and stage
// instead of ~235 as written).  Boundary conditions (gutsdynamo.f90:55-86) are applied
// while publishing: mirror sources also store their (negated) value into the ghost slot,
// Dirichlet points publish 0.
//
// check_f_error (gutsdynamo.f90:450-463) is evaluated by the reference after every stage on
// ALL points including ghost zones and the r = 0 point, whose values are one-sided-stencil
// garbage that impose_bc resets.  The fast path evaluates exact values only where the
// reference uses centred stencils on meaningful data; every BLK steps it reduces running
// maxima and accepts the block only if (a) no exactly evaluated value crossed a threshold and
// (b) a rigorous bound on the garbage points -- which depend on the 7 points next to each
// boundary only -- stays below the thresholds.  Otherwise the block is replayed from a
// checkpoint with the generic path (mag_galaxy.cuh), which evaluates every point exactly
// like the reference, so status codes, RNG draw order and the accumulated time stay exact.
#pragma once
#include "mag_phases.cuh"

namespace magdev {

#define MAG_FAST_BLK 16       // steps per checkpoint block
// Threads per galaxy team (template parameter TEAM of everything below).  k_evolve is
// instantiated twice in the same library (mag_solver.cu): TEAM = 32 -- one warp per galaxy, the
// register/shuffle loop of mag_rk_warp.cuh, highest throughput, used for the bulk of a batch --
// and TEAM = 128 -- four warps per galaxy on the shared-line loop of this file (1 point per
// thread on the default grid): the same work at a 4-5 times lower latency per Runge-Kutta
// stage, used for the galaxies whose history is so long that one warp would still be busy with
// it when the rest of the batch is done (the tail of the cost-ordered queue), and for grids the
// warp loop does not take (p_nx_ref = 401 of BASELINE config 5).
#define MAG_TEAM_LIGHT 32
#ifndef MAG_TEAM_HEAVY
#define MAG_TEAM_HEAVY 128
#endif
// points per thread up to which the stencil weights of a team loop live in shared memory (above: the
// streaming variant); sets the size of a team's shared memory and with it the teams per SM
#ifndef MAG_WSMEM_SLOTS
#define MAG_WSMEM_SLOTS 4
#endif
#define MAG_FAST_MAXPPL 8     // fast path up to nx = 8 * TEAM
#define MAG_LINE_S(TEAM) (4 * (TEAM) + 8)   // shared line of the PPL <= 4 variants
#define MAG_LINE_L(TEAM) (8 * (TEAM) + 8)   // shared line of the PPL = 8 variant
// coefficient placement of a fast-loop variant
#define COEF_REGS 0    // everything in registers (PPL = 1, 2)
#define COEF_WSMEM 1   // 7+1 stencil weights in shared memory, the rest in registers (PPL = 3, 4)
#define COEF_GLOBAL 2  // everything re-read from the team workspace (L1/L2) every stage (PPL = 8)

// parameters of one fast-loop call, written by the leader warp, read by both
struct FastCall {
  double* W;            // team workspace
  int32_t nxcap, nx, nsteps, ppl;
  int32_t sign_nx, refresh_sign, cmd, ok;
  double dt, t, t_Gyr, dtg, tau_min_kappa, t_last_sign_choice, Rk;
  long long point_stages;
};
#define TEAM_CMD_FAST 1
#define TEAM_CMD_EXIT 2

struct GuardK {
  double kb_lin, kb_a, fb0, fbk, akmax, ka_lin, kp, kq;
};

// per-team shared memory
template <int TEAM>
struct RkSharedT {
  RngState rng;
  RngState rng_ck;                                        // checkpoint copy
  FastCall call;
  GuardK gk[2];
  unsigned red[8];                                        // block maxima (high words)
  double stage_dg[3], stage_dz[3];                        // dt*gamma, dt*zeta of the three RK stages (looped stage body)
  // lines: [buffer][variable][4 + global point index] with row length MAG_LINE_S or MAG_LINE_L
  // (4 pad points each side, so that a thread's window -- own points +-3 -- is covered by
  // 16-byte aligned vector loads); for COEF_WSMEM the stencil weights [k][slot*64 + thread]
  // follow the short lines
  alignas(16) double pool[(2 * 3 * MAG_LINE_S(TEAM) + 8 * MAG_WSMEM_SLOTS * TEAM) > 1920 ? (2 * 3 * MAG_LINE_S(TEAM) + 8 * MAG_WSMEM_SLOTS * TEAM) : 1920];
  static_assert(2 * 3 * MAG_LINE_L(TEAM) <= 2 * 3 * MAG_LINE_S(TEAM) + 8 * MAG_WSMEM_SLOTS * TEAM || 2 * 3 * MAG_LINE_L(TEAM) <= 1920, "RkShared pool too small for the long lines");
};
using RkShared = RkSharedT<MAG_TEAM_LIGHT>;   // the one-warp loop of mag_rk_warp.cuh

// update_floor_sign (floor_field.f90:99-114)
__device__ __forceinline__ void update_floor_sign(Gal& G, double time) {
  if (time - G.t_last_sign_choice > G.a->P.p_floor_kappa * G.tau_min) {
    (void)random_sign(G);  // Bfloor_sign: its square cancels in the target field
    G.t_last_sign_choice = time;
    G.refresh_sign = true;
  }
}

// nb steps of the generic path (exact reference semantics), leader warp only.  Returns ok.
__device__ __noinline__ bool run_steps_generic(Gal& G, int jt0, int nb) {
  const double dtg = G.dt * G.a->U.t0_Gyr;
  for (int jt = jt0; jt < jt0 + nb; jt++) {
    const double this_t = G.t_Gyr + dtg * (double)jt;
    if (G.a->P.Alg_quench) estimate_Bzmod(G);   // dynamo.f90:179 -- only the algebraic quenching reads it inside pde
    update_floor_sign(G, this_t);
    const bool ok = rk_generic(G);
    impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    if (!ok) return false;
  }
  return true;
}

// ---- folded coefficients of one snapshot, written to the workspace for g in [0, npts)
// (zeros at ghost and unused points, so that those slots stay inert).  Leader warp.
__device__ __noinline__ void prepare_fast_coefs(Gal& G, int npts) {
  const int nx = G.nx, m = nx - MAG_NGHOST - 1;
  const double* x = G.A(A_X);
  const double ddx = x[1] - x[0];
  const double fac1 = 1.0 / (60. * ddx), fac2 = 1.0 / (180. * (ddx * ddx));
  const double a1[7] = {-1., 9., -45., 0., 45., -9., 1.};
  const double b2[7] = {2., -27., 270., -490., 270., -27., 2.};
  const double *cIR = G.A(C_IR), *cC1 = G.A(C_C1), *cCR = G.A(C_CR), *cW = G.A(C_W);
  const double Rk = G.a->P.R_kappa;
  for (int g = G.lane; g < npts; g += 32) {
    const bool live = (g >= MAG_NGHOST && g <= m);
    double w[7], w0a = 0.0;
    if (live) {
      const double cr = cCR[g], ir = cIR[g];
      const bool centre = (g == MAG_NGHOST);   // dalp_m/dr := 0 at r = 0 (gutsdynamo.f90:240)
      for (int k = 0; k < 7; k++) w[k] = cr * (fac2 * b2[k] + (centre ? 0.0 : ir * fac1 * a1[k]));
      w0a = Rk * cr * fac2 * b2[3] + cW[g];
      w[3] = cr * (fac2 * b2[3] - (centre ? 0.0 : ir * ir)) - cC1[g];
    } else {
      for (int k = 0; k < 7; k++) w[k] = 0.0;
    }
    for (int k = 0; k < 7; k++) G.A(C_W0 + k)[g] = w[k];
    G.A(C_W0A)[g] = w0a;
    // private copies: the generic path (replays) needs the originals at the ghost points
    G.A(F_CA)[g] = live ? G.A(C_CA)[g] : 0.0;
    G.A(F_GS)[g] = live ? G.A(A_G)[g] : 0.0;
    G.A(F_KDC)[g] = live ? G.A(C_KDC)[g] : 0.0;
    G.A(F_P3)[g] = live ? G.A(C_P3)[g] : 0.0;
    G.A(F_QC)[g] = live ? G.A(C_QC)[g] : 0.0;
    G.A(F_AK)[g] = live ? G.A(A_ALPK)[g] : 0.0;
  }
  __syncwarp();
}

// ---- rigorous bound on the values check_f_error sees at the points the fast path does not
// evaluate exactly: side 0 = the 3 inner ghost points and Br/Bp at the r = 0 point (their
// stencils read points 0..6), side 1 = the 3 outer ghost points and Br/Bp at the outer
// Dirichlet point nx-4, which the warp path does not evaluate exactly either (points nx-7..nx-1)
__device__ __noinline__ void guard_constants(Gal& G, GuardK* out /* [2], shared memory */) {
  const int nx = G.nx;
  const double* x = G.A(A_X);
  const double ddx = x[1] - x[0];
  const double S1 = 1664.0 / (60. * ddx), S2 = 18368.0 / (180. * (ddx * ddx));
  const double *cIR = G.A(C_IR), *cC1 = G.A(C_C1), *cCA = G.A(C_CA), *cCR = G.A(C_CR), *cKDC = G.A(C_KDC);
  const double *cFB = G.A(C_FB), *cP3 = G.A(C_P3), *cQC = G.A(C_QC), *cW = G.A(C_W), *Gc = G.A(A_G), *ak = G.A(A_ALPK);
  const double Rk = fabs(G.a->P.R_kappa);
  for (int side = 0; side < 2; side++) {
    GuardK k = {0, 0, 0, 0, 0, 0, 0, 0};
    // side 0: lanes 0-2 inner ghosts, lane 3 the centre; side 1: lane 0 the outer Dirichlet point, lanes 1-3 outer ghosts
    const bool active = G.lane < 4;
    if (active) {
      const int g = side == 0 ? G.lane : nx - 4 + G.lane;
      const double ir = fabs(cIR[g]), cr = fabs(cCR[g]);
      if (!(side == 0 && G.lane == 3)) {
        k.kb_lin = fabs(cC1[g]) + cr * (ir * ir + S1 * ir + S2) + fabs(Gc[g]);
        k.ka_lin = fabs(cW[g]) + Rk * cr * (S2 + S1 * ir);
        k.kp = fabs(cP3[g]);
        k.kq = fabs(cP3[g] * cQC[g]);
      } else {
        // Br = Bp = 0 at the centre when pde is evaluated: only the derivative terms survive;
        // alp_m at the centre is evaluated exactly
        k.kb_lin = cr * (110.0 / (60. * ddx) * ir + 1078.0 / (180. * (ddx * ddx))) + fabs(Gc[g]);
      }
      k.kb_a = fabs(cCA[g]);
      k.fb0 = fabs(cFB[g]);
      k.fbk = fabs(cFB[g] * cKDC[g]);
      k.akmax = fabs(ak[g]);
    }
    k.kb_lin = warp_max(k.kb_lin); k.kb_a = warp_max(k.kb_a); k.fb0 = warp_max(k.fb0); k.fbk = warp_max(k.fbk);
    k.akmax = warp_max(k.akmax); k.ka_lin = warp_max(k.ka_lin); k.kp = warp_max(k.kp); k.kq = warp_max(k.kq);
    if (G.lane == 0) out[side] = k;
  }
  __syncwarp();
}
// true when the garbage points of one side cannot have crossed the check_f_error thresholds;
// MB / MA bound |Br|,|Bp| / |alp_m| on the 7 points next to that boundary during the block
__device__ __forceinline__ bool guard_side_ok(const GuardK& k, double dt, double MB, double MA) {
  const double A = k.akmax + MA;
  const double bB = MB + 4.0 * dt * (MB * (k.kb_lin + k.kb_a * A) + k.fb0 + k.fbk * A);
  const double bA = MA + 4.0 * dt * (k.kp * A * 2.0 * MB * MB + k.kq * sqrt(A) * MB * MB + k.ka_lin * MA);
  return (bB < 1e8) && (bA < 1000.0);   // false for NaN
}

// sqrt|x| for the quenching term q*sqrt|alpha|*Br*Bp: MUFU.RSQ64H seed (rel. error ~2^-22)
// + one coupled Newton (Goldschmidt) step -> rel. error < 2e-12 (checked on the device by
// tests/test_gpu_parity.py::test_device_fast_sqrt), branch free.  The term is sub-dominant
// and errors do not amplify, so this is far inside the 1e-8 parity tolerance.  |x| is
// clamped away from 0/denormals (the term it multiplies is then < 1e-150).
#ifndef MAG_SQRT_ITERS
#define MAG_SQRT_ITERS 1
#endif
__device__ __forceinline__ double fast_sqrt_abs(double x) {
  int hi = __double2hiint(x) & 0x7fffffff;
  hi = max(hi, 0x00200000);
  x = __hiloint2double(hi, __double2loint(x));
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
#pragma unroll
  for (int it = 0; it < MAG_SQRT_ITERS; it++) {
    const double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    if (it + 1 < MAG_SQRT_ITERS) h = fma(h, e, h);
  }
  return g;
}

template <int PPL, int MODE>
struct FastState {
  double f[3][PPL], ft[3][PPL];
  double w[MODE == COEF_REGS ? PPL : 1][7], w0a[MODE == COEF_REGS ? PPL : 1];
  double ca[MODE == COEF_GLOBAL ? 1 : PPL], gs[MODE == COEF_GLOBAL ? 1 : PPL], fb[MODE == COEF_GLOBAL ? 1 : PPL];
  double kdc[MODE == COEF_GLOBAL ? 1 : PPL], p3[MODE == COEF_GLOBAL ? 1 : PPL], qc[MODE == COEF_GLOBAL ? 1 : PPL];
  double ak[MODE == COEF_GLOBAL ? 1 : PPL];
};
template <int TEAM, int PPL> struct LineLen { static constexpr int value = PPL <= 4 ? MAG_LINE_S(TEAM) : MAG_LINE_L(TEAM); };

// One Runge-Kutta stage.  STAGE: 0,1,2.  Reads the halo from line `buf`, updates f/ft in
// registers, publishes the new state (boundary conditions applied) to the other line.
template <int TEAM, int PPL, int MODE, int STAGE>
__device__ __forceinline__ void fast_stage(FastState<PPL, MODE>& S, RkSharedT<TEAM>* rs, int tid, int buf, double dg, double dz,
                                           double Rk, uint32_t pub_mask, uint32_t zero_mask, uint32_t mir_mask,
                                           uint32_t near_mask, const int* mir_addr, bool reload, unsigned& mB, unsigned& mA,
                                           unsigned& nB, unsigned& nA, double* alp_out, const double* W, int nxcap) {
  constexpr int LL = LineLen<TEAM, PPL>::value;
  const double* wts = rs->pool + 2 * 3 * MAG_LINE_S(TEAM);
  __syncthreads();
  // ---- gather the window: PPL own points + 3 halo points each side
  double win[3][PPL + 6];
  const int g0 = tid * PPL;
#pragma unroll
  for (int v = 0; v < 3; v++) {
    const double* L = rs->pool + (buf * 3 + v) * LL + 4 + g0;   // L[d] = point g0 + d
    if (PPL % 2 == 0) {  // g0 even: 16-byte aligned pairs (g0-4,g0-3) (g0-2,g0-1) | (g0+PPL,..) (g0+PPL+2,..)
      const double2 a = *reinterpret_cast<const double2*>(L - 4), b = *reinterpret_cast<const double2*>(L - 2);
      const double2 c = *reinterpret_cast<const double2*>(L + PPL), d = *reinterpret_cast<const double2*>(L + PPL + 2);
      win[v][0] = a.y; win[v][1] = b.x; win[v][2] = b.y;
      win[v][PPL + 3] = c.x; win[v][PPL + 4] = c.y; win[v][PPL + 5] = d.x;
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) { win[v][k] = L[k - 3]; win[v][PPL + 3 + k] = L[PPL + k]; }
    }
    if (reload) {  // threads that own ghost / Dirichlet points take the boundary-conditioned values
#pragma unroll
      for (int j = 0; j < PPL; j++) S.f[v][j] = L[j];
    }
#pragma unroll
    for (int j = 0; j < PPL; j++) win[v][3 + j] = S.f[v][j];
  }
  // ---- RHS and update, point by point
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    double w[7], w0a, c_ca, c_gs, c_fb, c_kdc, c_p3, c_qc, c_ak;
    if (MODE == COEF_WSMEM) {
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = wts[(k * MAG_WSMEM_SLOTS + j) * TEAM + tid];
      w0a = wts[(7 * MAG_WSMEM_SLOTS + j) * TEAM + tid];
    } else if (MODE == COEF_REGS) {
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = S.w[MODE == COEF_REGS ? j : 0][k];
      w0a = S.w0a[MODE == COEF_REGS ? j : 0];
    } else {
      const int g = g0 + j;
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = W[(size_t)(C_W0 + k) * nxcap + g];
      w0a = W[(size_t)C_W0A * nxcap + g];
    }
    if (MODE == COEF_GLOBAL) {
      const int g = g0 + j;
      c_ca = W[(size_t)F_CA * nxcap + g]; c_gs = W[(size_t)F_GS * nxcap + g]; c_fb = W[(size_t)C_FBS * nxcap + g];
      c_kdc = W[(size_t)F_KDC * nxcap + g]; c_p3 = W[(size_t)F_P3 * nxcap + g]; c_qc = W[(size_t)F_QC * nxcap + g];
      c_ak = W[(size_t)F_AK * nxcap + g];
    } else {
      constexpr int jj = 0;
      (void)jj;
      c_ca = S.ca[MODE == COEF_GLOBAL ? 0 : j]; c_gs = S.gs[MODE == COEF_GLOBAL ? 0 : j]; c_fb = S.fb[MODE == COEF_GLOBAL ? 0 : j];
      c_kdc = S.kdc[MODE == COEF_GLOBAL ? 0 : j]; c_p3 = S.p3[MODE == COEF_GLOBAL ? 0 : j]; c_qc = S.qc[MODE == COEF_GLOBAL ? 0 : j];
      c_ak = S.ak[MODE == COEF_GLOBAL ? 0 : j];
    }
    const double Br = win[0][3 + j], Bp = win[1][3 + j], am = win[2][3 + j];
    double LBr = w[0] * win[0][j], LBp = w[0] * win[1][j], Lam = w[0] * win[2][j];
#pragma unroll
    for (int k = 1; k < 7; k++) {
      LBr = fma(w[k], win[0][j + k], LBr);
      LBp = fma(w[k], win[1][j + k], LBp);
      if (k != 3) Lam = fma(w[k], win[2][j + k], Lam);
    }
    const double a = c_ak + am;
    if (alp_out) alp_out[tid * PPL + j] = a;
    const double p0 = fma(-(c_ca * a), Bp, LBr);
    const double p1 = fma(c_fb, fma(c_kdc, a, 1.0), fma(c_gs, Br, LBp));
    const double B2 = fma(Br, Br, Bp * Bp);
    const double em = fma(c_qc * fast_sqrt_abs(a) * Br, Bp, a * B2);
    const double p2 = fma(c_p3, em, fma(Rk, Lam, w0a * am));
    const double b0 = (STAGE == 0) ? Br : S.ft[0][j];
    const double b1 = (STAGE == 0) ? Bp : S.ft[1][j];
    const double b2 = (STAGE == 0) ? am : S.ft[2][j];
    const double n0 = fma(dg, p0, b0), n1 = fma(dg, p1, b1), n2 = fma(dg, p2, b2);
    S.f[0][j] = n0; S.f[1][j] = n1; S.f[2][j] = n2;
    if (STAGE < 2) {
      S.ft[0][j] = fma(dz, p0, n0); S.ft[1][j] = fma(dz, p1, n1); S.ft[2][j] = fma(dz, p2, n2);
    }
    mB = max(mB, max((unsigned)__double2hiint(n0) & 0x7fffffffu, (unsigned)__double2hiint(n1) & 0x7fffffffu));
    mA = max(mA, (unsigned)__double2hiint(n2) & 0x7fffffffu);
  }
  // ---- publish the new state with boundary conditions applied
  double* Ln = rs->pool + (buf ^ 1) * 3 * LL;   // Ln[v*LL + 4 + g]
  if (PPL % 2 == 0 && !reload) {  // plain interior thread: vector stores
#pragma unroll
    for (int j = 0; j < PPL; j += 2) {
      *reinterpret_cast<double2*>(&Ln[4 + g0 + j]) = make_double2(S.f[0][j], S.f[0][j + 1]);
      *reinterpret_cast<double2*>(&Ln[LL + 4 + g0 + j]) = make_double2(S.f[1][j], S.f[1][j + 1]);
      *reinterpret_cast<double2*>(&Ln[2 * LL + 4 + g0 + j]) = make_double2(S.f[2][j], S.f[2][j + 1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      if (pub_mask & (1u << j)) {
        const bool z = zero_mask & (1u << j);
        Ln[4 + g0 + j] = z ? 0.0 : S.f[0][j];
        Ln[LL + 4 + g0 + j] = z ? 0.0 : S.f[1][j];
        Ln[2 * LL + 4 + g0 + j] = S.f[2][j];
      }
    }
  }
  if (near_mask) {  // the few threads next to a boundary
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      if (mir_mask & (1u << j)) {
        const int a = mir_addr[j];
        Ln[a] = -S.f[0][j]; Ln[LL + a] = -S.f[1][j]; Ln[2 * LL + a] = S.f[2][j];
        nB = max(nB, max((unsigned)__double2hiint(S.f[0][j]) & 0x7fffffffu, (unsigned)__double2hiint(S.f[1][j]) & 0x7fffffffu));
      }
      if (near_mask & (1u << j)) nA = max(nA, (unsigned)__double2hiint(S.f[2][j]) & 0x7fffffffu);
    }
  }
}

// Fast time loop of one snapshot attempt, executed by all 64 threads of the team.  G is the
// leader warp's galaxy state (nullptr in the helper warp).  The result is left in rs->call
// (ok, t, t_last_sign_choice, sign_nx, refresh_sign, point_stages) and in the workspace.
template <int TEAM, int PPL, int MODE>
__device__ __noinline__ void fast_loop_team(Gal* G, RkSharedT<TEAM>* rs) {
  constexpr int LL = LineLen<TEAM, PPL>::value;
  double* const wts = rs->pool + 2 * 3 * MAG_LINE_S(TEAM);
  const int tid = threadIdx.x;
  const bool leader = tid < 32;
  const FastCall C = rs->call;
  double* const W = C.W;
  const int nxcap = C.nxcap, nx = C.nx, m = nx - MAG_NGHOST - 1, nsteps = C.nsteps;
  const int npts = TEAM * PPL;
  const double Rk = C.Rk, dt = C.dt;
  auto WA = [&](int id) { return W + (size_t)id * nxcap; };

  if (leader) { prepare_fast_coefs(*G, npts); guard_constants(*G, rs->gk); }
  __syncthreads();

  // per-thread publishing tables: which slots are published / zeroed / mirrored into a ghost slot
  uint32_t pub_mask = 0, zero_mask = 0, mir_mask = 0, near_mask = 0;
  int mir_addr[PPL];
  bool reload = false;
  int side = -1;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = tid * PPL + j;
    mir_addr[j] = 0;
    const bool ghost = (g < MAG_NGHOST) || (g > m && g < nx);
    if (!ghost) pub_mask |= 1u << j;                       // ghost slots are written by their mirror source
    if (g == MAG_NGHOST || g == m) zero_mask |= 1u << j;   // Dirichlet points of Br, Bp
    int dst = -1;
    if (g > MAG_NGHOST && g <= 2 * MAG_NGHOST) dst = 2 * MAG_NGHOST - g;          // 4,5,6 -> 2,1,0
    if (g < m && g >= m - MAG_NGHOST) dst = 2 * m - g;                            // m-1.. -> m+1..
    if (dst >= 0) { mir_mask |= 1u << j; mir_addr[j] = 4 + dst; }
    if (dst >= 0 || g == MAG_NGHOST || g == m) { near_mask |= 1u << j; side = (g <= 2 * MAG_NGHOST) ? 0 : 1; }
    if (ghost || g == MAG_NGHOST || g == m) reload = true;
  }

  FastState<PPL, MODE> S;
  auto load_state = [&]() {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = tid * PPL + j;
      const bool in = g < nx;
#pragma unroll
      for (int v = 0; v < 3; v++) { S.f[v][j] = in ? WA(A_F0 + v)[g] : 0.0; S.ft[v][j] = 0.0; }
    }
  };
  auto load_signed_floor = [&]() {
    const double *cFB = WA(C_FB), *sg = WA(A_SIGN);
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = tid * PPL + j;
      const double v = (g >= MAG_NGHOST && g <= m) ? sg[g] * cFB[g] : 0.0;
      if (MODE == COEF_GLOBAL) WA(C_FBS)[g] = v; else S.fb[MODE == COEF_GLOBAL ? 0 : j] = v;
    }
  };
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = tid * PPL + j;
    if (MODE == COEF_WSMEM) {
#pragma unroll
      for (int k = 0; k < 7; k++) wts[(k * MAG_WSMEM_SLOTS + j) * TEAM + tid] = WA(C_W0 + k)[g];
      wts[(7 * MAG_WSMEM_SLOTS + j) * TEAM + tid] = WA(C_W0A)[g];
    } else if (MODE == COEF_REGS) {
#pragma unroll
      for (int k = 0; k < 7; k++) S.w[MODE == COEF_REGS ? j : 0][k] = WA(C_W0 + k)[g];
      S.w0a[MODE == COEF_REGS ? j : 0] = WA(C_W0A)[g];
    }
    if (MODE != COEF_GLOBAL) {
      S.ca[MODE == COEF_GLOBAL ? 0 : j] = WA(F_CA)[g]; S.gs[MODE == COEF_GLOBAL ? 0 : j] = WA(F_GS)[g];
      S.kdc[MODE == COEF_GLOBAL ? 0 : j] = WA(F_KDC)[g]; S.p3[MODE == COEF_GLOBAL ? 0 : j] = WA(F_P3)[g];
      S.qc[MODE == COEF_GLOBAL ? 0 : j] = WA(F_QC)[g]; S.ak[MODE == COEF_GLOBAL ? 0 : j] = WA(F_AK)[g];
    }
  }
  load_signed_floor();
  load_state();

  auto publish = [&](int buf) {
    double* Ln = rs->pool + buf * 3 * LL;
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      if (pub_mask & (1u << j)) {
        const bool z = zero_mask & (1u << j);
        Ln[4 + tid * PPL + j] = z ? 0.0 : S.f[0][j];
        Ln[LL + 4 + tid * PPL + j] = z ? 0.0 : S.f[1][j];
        Ln[2 * LL + 4 + tid * PPL + j] = S.f[2][j];
      }
      if (mir_mask & (1u << j)) {
        const int a = mir_addr[j];
        Ln[a] = -S.f[0][j]; Ln[LL + a] = -S.f[1][j]; Ln[2 * LL + a] = S.f[2][j];
      }
    }
  };
  // unused slots of both lines must read as (finite) zero
  __syncthreads();   // (the weights above live behind the short lines; the long lines overlap them)
  for (int i = tid; i < 2 * 3 * LL; i += TEAM) rs->pool[i] = 0.0;
  __syncthreads();

  const double dg0 = dt * (8.0 / 15), dg1 = dt * (5.0 / 12), dg2 = dt * (3.0 / 4);
  const double dz0 = dt * (-17.0 / 60), dz1 = dt * (-5.0 / 12);
  const double dtg = C.dtg, t_Gyr = C.t_Gyr, tmk = C.tau_min_kappa;
  // replicated scalar state (identical in all 64 threads)
  double t = C.t, t_last = C.t_last_sign_choice;
  int sign_nx = C.sign_nx;
  bool refresh = C.refresh_sign != 0;
  long long stages = 0;
  double* ck[3] = {WA(A_T0), WA(A_T1), WA(A_T2)};
  int buf = 0;
  int jt = 1;
  bool ok = true;
  while (jt <= nsteps) {
    const int nb = min(MAG_FAST_BLK, nsteps - jt + 1);
    // ---- checkpoint
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = tid * PPL + j;
      if (g < nx) { ck[0][g] = S.f[0][j]; ck[1][g] = S.f[1][j]; ck[2][g] = S.f[2][j]; }
    }
    const double ck_t = t, ck_tl = t_last;
    const bool ck_refresh = refresh;
    const int ck_sign_nx = sign_nx;
    static_assert(sizeof(RngState) == 17 * sizeof(uint64_t), "RngState layout");
    if (tid < 17) ((uint64_t*)&rs->rng_ck)[tid] = ((const uint64_t*)&rs->rng)[tid];
    if (tid < 8) rs->red[tid] = 0u;
    bool signs_saved = false;
    // running maxima start from the block's initial state
    unsigned mB = 0, mA = 0, nB = 0, nA = 0;
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const unsigned h0 = (unsigned)__double2hiint(S.f[0][j]) & 0x7fffffffu, h1 = (unsigned)__double2hiint(S.f[1][j]) & 0x7fffffffu;
      const unsigned h2 = (unsigned)__double2hiint(S.f[2][j]) & 0x7fffffffu;
      mB = max(mB, max(h0, h1)); mA = max(mA, h2);
      if (mir_mask & (1u << j)) nB = max(nB, max(h0, h1));
      if (near_mask & (1u << j)) nA = max(nA, h2);
    }
    publish(buf);
    for (int s = 0; s < nb; s++) {
      const double this_t = t_Gyr + dtg * (double)(jt + s);
      // same order of RNG draws as the reference: update_floor_sign (dynamo.f90:196-198) may
      // draw Bfloor_sign, then the first pde of the step redraws the layer signs if flagged
      // or if the grid size changed (floor_field.f90:55-79)
      const bool due = this_t - t_last > tmk;
      if (due || refresh || sign_nx != nx) {
        __syncthreads();
        if (leader) {
          if (!signs_saved) {  // first refresh of the block: keep the old signs for a replay
            const double* sg = WA(A_SIGN); double* bk = WA(A_OMK);
            const int nsave = max(nx, ck_sign_nx);
            for (int i = tid; i < nsave; i += 32) bk[i] = sg[i];
            __syncwarp();
          }
          if (due) (void)random_sign(*G);
          refresh_signs(*G);
        }
        signs_saved = true;
        if (due) t_last = this_t;
        refresh = false; sign_nx = nx;
        __syncthreads();
        load_signed_floor();
      }
      double* alp_out = (jt + s == nsteps) ? WA(A_ALP) : nullptr;
      fast_stage<TEAM, PPL, MODE, 0>(S, rs, tid, buf, dg0, dz0, Rk, pub_mask, zero_mask, mir_mask, near_mask, mir_addr, reload, mB, mA, nB, nA, nullptr, W, nxcap);
      buf ^= 1;
      fast_stage<TEAM, PPL, MODE, 1>(S, rs, tid, buf, dg1, dz1, Rk, pub_mask, zero_mask, mir_mask, near_mask, mir_addr, reload, mB, mA, nB, nA, nullptr, W, nxcap);
      buf ^= 1;
      fast_stage<TEAM, PPL, MODE, 2>(S, rs, tid, buf, dg2, 0.0, Rk, pub_mask, zero_mask, mir_mask, near_mask, mir_addr, reload, mB, mA, nB, nA, alp_out, W, nxcap);
      buf ^= 1;
      // t = t + dt*gam1 ; ttmp = t + dt*zet1 ; t = ttmp + dt*gam2 ; ttmp = t + dt*zet2 ; t = ttmp + dt*gam3
      double tt = t + dg0;
      tt = tt + dz0; tt = tt + dg1; tt = tt + dz1; tt = tt + dg2;
      t = tt;
    }
    // ---- accept or replay the block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mB = max(mB, __shfl_xor_sync(FULL, mB, o));
      mA = max(mA, __shfl_xor_sync(FULL, mA, o));
    }
    if ((tid & 31) == 0) { atomicMax(&rs->red[0], mB); atomicMax(&rs->red[1], mA); }
    if (side >= 0) { atomicMax(&rs->red[2 + 2 * side], nB); atomicMax(&rs->red[3 + 2 * side], nA); }
    __syncthreads();
    const unsigned rB = rs->red[0], rA = rs->red[1];
    bool accept = rB < 0x7fe00000u && rA < 0x7fe00000u;
    if (accept) {
      const double MB = __hiloint2double((int)rB + 1, 0), MA = __hiloint2double((int)rA + 1, 0);
      accept = MB < 1e8 && MA < 1000.0;
      for (int sd = 0; sd < 2 && accept; sd++) {
        const double MBn = __hiloint2double((int)rs->red[2 + 2 * sd] + 1, 0), MAn = __hiloint2double((int)rs->red[3 + 2 * sd] + 1, 0);
        accept = guard_side_ok(rs->gk[sd], dt, MBn, MAn);
      }
    }
    __syncthreads();   // everyone has read rs->red before the next block clears it
    if (accept) {
      stages += (long long)nb * 3 * nx;
      jt += nb;
      continue;
    }
    // replay with the generic path from the checkpoint (leader warp), helper waits
    if (leader) {
      Gal& g = *G;
      if (g.a->replays && tid == 0) atomicAdd(g.a->replays, 1);
      for (int v = 0; v < 3; v++) {
        double* f = WA(A_F0 + v);
        for (int i = tid; i < nx; i += 32) f[i] = ck[v][i];
      }
      if (signs_saved) {
        double* sg = WA(A_SIGN); const double* bk = WA(A_OMK);
        const int nsave = max(nx, ck_sign_nx);
        for (int i = tid; i < nsave; i += 32) sg[i] = bk[i];
      }
      if (tid < 17) ((uint64_t*)&rs->rng)[tid] = ((const uint64_t*)&rs->rng_ck)[tid];
      g.t = ck_t; g.t_last_sign_choice = ck_tl; g.refresh_sign = ck_refresh; g.sign_nx = ck_sign_nx;
      const long long ps0 = g.point_stages;
      __syncwarp();
      impose_bc(g, WA(A_F0), WA(A_F1), WA(A_F2));
      const bool okg = run_steps_generic(g, jt, nb);
      if (tid == 0) {
        rs->call.ok = okg ? 1 : 0; rs->call.t = g.t; rs->call.t_last_sign_choice = g.t_last_sign_choice;
        rs->call.refresh_sign = g.refresh_sign ? 1 : 0; rs->call.sign_nx = g.sign_nx;
        rs->call.point_stages = g.point_stages - ps0;
      }
      g.point_stages = ps0;
    }
    __syncthreads();
    ok = rs->call.ok != 0;
    t = rs->call.t; t_last = rs->call.t_last_sign_choice; refresh = rs->call.refresh_sign != 0; sign_nx = rs->call.sign_nx;
    stages += rs->call.point_stages;
    __syncthreads();
    if (!ok) break;
    jt += nb;
    load_signed_floor();
    load_state();
  }
  // ---- hand the state back to the workspace (the caller applies impose_bc)
  if (ok) {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = tid * PPL + j;
      if (g < nx) { WA(A_F0)[g] = S.f[0][j]; WA(A_F1)[g] = S.f[1][j]; WA(A_F2)[g] = S.f[2][j]; }
    }
  }
  __syncthreads();
  if (tid == 0) {
    rs->call.ok = ok ? 1 : 0; rs->call.t = t; rs->call.t_last_sign_choice = t_last;
    rs->call.refresh_sign = refresh ? 1 : 0; rs->call.sign_nx = sign_nx; rs->call.point_stages = stages;
  }
  __syncthreads();
}

}  // namespace magdev
#include "mag_rk_warp.cuh"
namespace magdev {

template <int TEAM>
__device__ __forceinline__ void fast_dispatch(Gal* G, RkSharedT<TEAM>* rs) {
  switch (rs->call.ppl) {
    case 1: if (TEAM > 32) fast_loop_team<TEAM, 1, COEF_REGS>(G, rs); break;   // (one-warp teams: nx >= 64 has PPL >= 2)
    case 2: fast_loop_team<TEAM, 2, COEF_REGS>(G, rs); break;
    case 3: fast_loop_team<TEAM, 3, COEF_WSMEM>(G, rs); break;
#if MAG_WSMEM_SLOTS >= 4
    case 4: fast_loop_team<TEAM, 4, COEF_WSMEM>(G, rs); break;
#endif
    default: fast_loop_team<TEAM, 8, COEF_GLOBAL>(G, rs); break;
  }
}

// Leader warp: runs the time loop of the current snapshot attempt.  Returns ok (false: the
// field blew up, check_f_error).
template <int TEAM>
__device__ bool run_timesteps(Gal& G, RkSharedT<TEAM>* rs) {
  if (G.nsteps < 1) return true;
  if (TEAM == MAG_TEAM_LIGHT) {
    if (G.a->use_fast && warp_path_applies(G)) return warp_dispatch(G, reinterpret_cast<RkShared*>(rs));
  }
  if (G.a->use_fast && G.nx >= 32 && G.nx <= MAG_FAST_MAXPPL * TEAM) {
    if (G.lane == 0) {
      FastCall& C = rs->call;
      C.W = G.W; C.nxcap = G.a->nxcap; C.nx = G.nx; C.nsteps = G.nsteps;
      C.ppl = (TEAM > 32 && G.nx <= TEAM) ? 1 : (G.nx <= 2 * TEAM ? 2 : (G.nx <= 3 * TEAM ? 3 : ((MAG_WSMEM_SLOTS >= 4 && G.nx <= 4 * TEAM) ? 4 : 8)));
      C.sign_nx = G.sign_nx; C.refresh_sign = G.refresh_sign ? 1 : 0; C.cmd = TEAM_CMD_FAST; C.ok = 1;
      C.dt = G.dt; C.t = G.t; C.t_Gyr = G.t_Gyr; C.dtg = G.dt * G.a->U.t0_Gyr;
      C.tau_min_kappa = G.a->P.p_floor_kappa * G.tau_min; C.t_last_sign_choice = G.t_last_sign_choice;
      C.Rk = G.a->P.R_kappa; C.point_stages = 0;
    }
    __syncthreads();   // releases the helper warps (k_evolve)
    fast_dispatch<TEAM>(&G, rs);
    const FastCall& C = rs->call;
    G.t = C.t; G.t_last_sign_choice = C.t_last_sign_choice; G.refresh_sign = C.refresh_sign != 0; G.sign_nx = C.sign_nx;
    G.point_stages += C.point_stages;
    return C.ok != 0;
  }
  return run_steps_generic(G, 1, G.nsteps);
}

}  // namespace magdev
