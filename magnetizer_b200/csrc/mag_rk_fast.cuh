// mag_rk_fast.cuh -- the time-stepping loop of one snapshot (dynamo.f90:168-212):
// nsteps x { update_floor_sign ; rk (3 stages, gutsdynamo.f90:404-448) ; impose_bc }.
//
// Register-resident fast path.  Lane l of the warp owns PPL consecutive radial points
// (4l .. 4l+3 for PPL = 4, nx <= 128): the state f = (Br, Bp, alp_m), the 2N-storage
// register ftmp and -- for PPL = 4 -- all per-point coefficients stay in registers for the
// whole snapshot.  Per stage each lane publishes its points to a double-buffered
// shared-memory line, reads the 3+3 halo points of its neighbours after ONE __syncwarp,
// and evaluates the RHS in a folded form:
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
// reference uses centred stencils on meaningful data; every BLK steps it reduces a running
// maximum and accepts the block only if a rigorous bound on the garbage points stays below
// the blow-up thresholds.  Otherwise the block is replayed from a checkpoint with the
// generic path (mag_galaxy.cuh), which evaluates every point exactly like the reference --
// so status codes, RNG draw order and the accumulated time stay exact.
#pragma once
#include "mag_galaxy.cuh"

namespace magdev {

#define MAG_FAST_BLK 8        // steps per checkpoint block
#define MAG_FAST_MAXPPL 8     // shared line sized for nx <= 256

// per-warp shared memory
struct RkShared {
  RngState rng;
  RngState rng_ck;                                 // checkpoint copy
  double line[2][3][MAG_FAST_MAXPPL * 32];         // [buffer][variable][slot*32 + lane]
};

// update_floor_sign (floor_field.f90:99-114)
__device__ __forceinline__ bool floor_sign_due(const Gal& G, double time) {
  return time - G.t_last_sign_choice > G.a->P.p_floor_kappa * G.tau_min;
}
__device__ __forceinline__ void update_floor_sign(Gal& G, double time) {
  if (floor_sign_due(G, time)) {
    (void)random_sign(G);  // Bfloor_sign: its square cancels in the target field
    G.t_last_sign_choice = time;
    G.refresh_sign = true;
  }
}

// nb steps of the generic path (exact reference semantics).  Returns ok.
__device__ __noinline__ bool run_steps_generic(Gal& G, int jt0, int nb) {
  const double dtg = G.dt * G.a->U.t0_Gyr;
  for (int jt = jt0; jt < jt0 + nb; jt++) {
    const double this_t = G.t_Gyr + dtg * (double)jt;
    update_floor_sign(G, this_t);
    const bool ok = rk_generic(G);
    impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    if (!ok) return false;
  }
  return true;
}

// ---- folded coefficients of one snapshot, written to the workspace for g in [0, 32*PPL)
// (zeros at ghost and unused points, so that those slots stay inert)
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
// evaluate exactly (6 ghost points with one-sided stencils, Br/Bp at the r = 0 point)
struct GuardK {
  double kb_lin, kb_a, fb0, fbk, akmax, ka_lin, kp, kq;
};
__device__ __noinline__ GuardK guard_constants(Gal& G) {
  const int nx = G.nx;
  const double* x = G.A(A_X);
  const double ddx = x[1] - x[0];
  const double S1 = 1664.0 / (60. * ddx), S2 = 18368.0 / (180. * (ddx * ddx));
  const double *cIR = G.A(C_IR), *cC1 = G.A(C_C1), *cCA = G.A(C_CA), *cCR = G.A(C_CR), *cKDC = G.A(C_KDC);
  const double *cFB = G.A(C_FB), *cP3 = G.A(C_P3), *cQC = G.A(C_QC), *cW = G.A(C_W), *Gc = G.A(A_G), *ak = G.A(A_ALPK);
  const double Rk = fabs(G.a->P.R_kappa);
  GuardK k = {0, 0, 0, 0, 0, 0, 0, 0};
  if (G.lane < 7) {
    // lanes 0-2: left ghosts, 3-5: right ghosts, 6: the r = 0 point
    const int g = G.lane < 3 ? G.lane : (G.lane < 6 ? nx - 6 + G.lane : MAG_NGHOST);
    const double ir = fabs(cIR[g]), cr = fabs(cCR[g]);
    if (G.lane < 6) {
      k.kb_lin = fabs(cC1[g]) + cr * (ir * ir + S1 * ir + S2) + fabs(Gc[g]);
      k.ka_lin = fabs(cW[g]) + Rk * cr * (S2 + S1 * ir);
    } else {
      // Br = Bp = 0 at the centre when pde is evaluated: only the derivative terms survive
      k.kb_lin = cr * (110.0 / (60. * ddx) * ir + 1078.0 / (180. * (ddx * ddx))) + fabs(Gc[g]);
      k.ka_lin = 0.0;  // alp_m at the centre is evaluated exactly
    }
    k.kb_a = fabs(cCA[g]);
    k.fb0 = fabs(cFB[g]);
    k.fbk = fabs(cFB[g] * cKDC[g]);
    k.akmax = fabs(ak[g]);
    k.kp = G.lane < 6 ? fabs(cP3[g]) : 0.0;
    k.kq = G.lane < 6 ? fabs(cP3[g] * cQC[g]) : 0.0;
  }
  k.kb_lin = warp_max(k.kb_lin); k.kb_a = warp_max(k.kb_a); k.fb0 = warp_max(k.fb0); k.fbk = warp_max(k.fbk);
  k.akmax = warp_max(k.akmax); k.ka_lin = warp_max(k.ka_lin); k.kp = warp_max(k.kp); k.kq = warp_max(k.kq);
  return k;
}
// true when no point of the reference's f can have crossed the check_f_error thresholds
__device__ __forceinline__ bool guard_ok(const GuardK& k, double dt, double MB, double MA) {
  const double A = k.akmax + MA;
  const double bB = MB + 4.0 * dt * (MB * (k.kb_lin + k.kb_a * A) + k.fb0 + k.fbk * A);
  const double bA = MA + 4.0 * dt * (k.kp * A * 2.0 * MB * MB + k.kq * sqrt(A) * MB * MB + k.ka_lin * MA);
  return (bB < 1e8) && (bA < 1000.0);   // false for NaN
}

template <int PPL, bool COEF_REGS>
struct FastState {
  double f[3][PPL], ft[3][PPL];
  // coefficients (registers when COEF_REGS, else re-read from the workspace every stage)
  double w[COEF_REGS ? PPL : 1][7], w0a[COEF_REGS ? PPL : 1], ca[COEF_REGS ? PPL : 1], gs[COEF_REGS ? PPL : 1];
  double fb[COEF_REGS ? PPL : 1], kdc[COEF_REGS ? PPL : 1], p3[COEF_REGS ? PPL : 1], qc[COEF_REGS ? PPL : 1];
  double ak[COEF_REGS ? PPL : 1];
};

// One Runge-Kutta stage.  STAGE: 0,1,2.  Reads the halo from line `buf`, updates f/ft in
// registers, publishes the new f to the other line.  mB/mA: running maxima of the high words of
// |Br|,|Bp| and |alp_m| over this lane's live points.
template <int PPL, bool COEF_REGS, int STAGE>
__device__ __forceinline__ void fast_stage(FastState<PPL, COEF_REGS>& S, const Gal& G, double (*line)[3][MAG_FAST_MAXPPL * 32],
                                           int buf, double dg, double dz, double Rk, uint32_t pub_mask, uint32_t zero_mask,
                                           uint32_t mir_mask, const int* mir_addr, bool reload, unsigned& mB, unsigned& mA,
                                           bool write_alp) {
  const int lane = G.lane;
  const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
  __syncwarp();
  // ---- gather the window: PPL own points + 3 halo points each side
  double win[3][PPL + 6];
#pragma unroll
  for (int v = 0; v < 3; v++) {
    const double* L = line[buf][v];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      win[v][k] = (lane == 0) ? 0.0 : L[(PPL - 3 + k) * 32 + lm];
      win[v][PPL + 3 + k] = (lane == 31) ? 0.0 : L[k * 32 + lp];
    }
    if (reload) {  // lanes that own ghost / Dirichlet points take the boundary-conditioned values
#pragma unroll
      for (int j = 0; j < PPL; j++) S.f[v][j] = L[j * 32 + lane];
    }
#pragma unroll
    for (int j = 0; j < PPL; j++) win[v][3 + j] = S.f[v][j];
  }
  // ---- RHS and update, point by point
  const int g0 = lane * PPL;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    double w[7], w0a, ca, gs, fb, kdc, p3, qc, ak;
    if (COEF_REGS) {
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = S.w[COEF_REGS ? j : 0][k];
      w0a = S.w0a[COEF_REGS ? j : 0]; ca = S.ca[COEF_REGS ? j : 0]; gs = S.gs[COEF_REGS ? j : 0];
      fb = S.fb[COEF_REGS ? j : 0]; kdc = S.kdc[COEF_REGS ? j : 0]; p3 = S.p3[COEF_REGS ? j : 0];
      qc = S.qc[COEF_REGS ? j : 0]; ak = S.ak[COEF_REGS ? j : 0];
    } else {
      const int g = g0 + j;
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = G.A(C_W0 + k)[g];
      w0a = G.A(C_W0A)[g]; ca = G.A(F_CA)[g]; gs = G.A(F_GS)[g]; fb = G.A(C_FBS)[g]; kdc = G.A(F_KDC)[g];
      p3 = G.A(F_P3)[g]; qc = G.A(F_QC)[g]; ak = G.A(F_AK)[g];
    }
    const double Br = win[0][3 + j], Bp = win[1][3 + j], am = win[2][3 + j];
    double LBr = w[0] * win[0][j], LBp = w[0] * win[1][j], Lam = w[0] * win[2][j];
#pragma unroll
    for (int k = 1; k < 7; k++) {
      LBr = fma(w[k], win[0][j + k], LBr);
      LBp = fma(w[k], win[1][j + k], LBp);
      if (k != 3) Lam = fma(w[k], win[2][j + k], Lam);
    }
    const double a = ak + am;
    if (write_alp) G.A(A_ALP)[g0 + j] = a;
    const double p0 = fma(-(ca * a), Bp, LBr);
    const double p1 = fma(fb, fma(kdc, a, 1.0), fma(gs, Br, LBp));
    const double B2 = fma(Br, Br, Bp * Bp);
    const double em = fma(qc * sqrt(fabs(a)) * Br, Bp, a * B2);
    const double p2 = fma(p3, em, fma(Rk, Lam, w0a * am));
    const double b0 = (STAGE == 0) ? Br : S.ft[0][j];
    const double b1 = (STAGE == 0) ? Bp : S.ft[1][j];
    const double b2 = (STAGE == 0) ? am : S.ft[2][j];
    const double n0 = fma(dg, p0, b0), n1 = fma(dg, p1, b1), n2 = fma(dg, p2, b2);
    S.f[0][j] = n0; S.f[1][j] = n1; S.f[2][j] = n2;
    if (STAGE < 2) {
      S.ft[0][j] = fma(dz, p0, n0); S.ft[1][j] = fma(dz, p1, n1); S.ft[2][j] = fma(dz, p2, n2);
    }
    mB = max(mB, (unsigned)__double2hiint(n0) & 0x7fffffffu);
    mB = max(mB, (unsigned)__double2hiint(n1) & 0x7fffffffu);
    mA = max(mA, (unsigned)__double2hiint(n2) & 0x7fffffffu);
  }
  // ---- publish the new state with boundary conditions applied
  double (*Ln)[MAG_FAST_MAXPPL * 32] = line[buf ^ 1];
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    if (pub_mask & (1u << j)) {
      const bool z = zero_mask & (1u << j);
      Ln[0][j * 32 + lane] = z ? 0.0 : S.f[0][j];
      Ln[1][j * 32 + lane] = z ? 0.0 : S.f[1][j];
      Ln[2][j * 32 + lane] = S.f[2][j];
    }
  }
  if (mir_mask) {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      if (mir_mask & (1u << j)) {
        const int a = mir_addr[j];
        Ln[0][a] = -S.f[0][j]; Ln[1][a] = -S.f[1][j]; Ln[2][a] = S.f[2][j];
      }
    }
  }
}

// Fast time loop for one snapshot attempt.  Returns ok (false: check_f_error tripped).
template <int PPL, bool COEF_REGS>
__device__ __noinline__ bool run_timesteps_fast(Gal& G, RkShared* rs) {
  const int nx = G.nx, lane = G.lane, m = nx - MAG_NGHOST - 1;
  const int npts = 32 * PPL;
  const double Rk = G.a->P.R_kappa;
  prepare_fast_coefs(G, npts);
  const GuardK gk = guard_constants(G);

  // per-lane publishing tables: which slots are published / zeroed / mirrored into a ghost slot
  uint32_t pub_mask = 0, zero_mask = 0, mir_mask = 0;
  int mir_addr[PPL];
  bool reload = false;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = lane * PPL + j;
    mir_addr[j] = 0;
    const bool ghost = (g < MAG_NGHOST) || (g > m && g < nx);
    if (!ghost) pub_mask |= 1u << j;                       // ghost slots are written by their mirror source
    if (g == MAG_NGHOST || g == m) zero_mask |= 1u << j;   // Dirichlet points of Br, Bp
    int dst = -1;
    if (g > MAG_NGHOST && g <= 2 * MAG_NGHOST) dst = 2 * MAG_NGHOST - g;          // 4,5,6 -> 2,1,0
    if (g < m && g >= m - MAG_NGHOST) dst = 2 * m - g;                            // m-1.. -> m+1..
    if (dst >= 0) { mir_mask |= 1u << j; mir_addr[j] = (dst % PPL) * 32 + dst / PPL; }
    if (ghost || g == MAG_NGHOST || g == m) reload = true;
  }

  FastState<PPL, COEF_REGS> S;
  auto load_state = [&]() {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = lane * PPL + j;
      const bool in = g < nx;
#pragma unroll
      for (int v = 0; v < 3; v++) { S.f[v][j] = in ? G.A(A_F0 + v)[g] : 0.0; S.ft[v][j] = 0.0; }
    }
  };
  auto load_signed_floor = [&]() {
    const double *cFB = G.A(C_FB), *sg = G.A(A_SIGN);
    double* fbs = G.A(C_FBS);
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = lane * PPL + j;
      const double v = (g >= MAG_NGHOST && g <= m) ? sg[g] * cFB[g] : 0.0;
      if (COEF_REGS) S.fb[COEF_REGS ? j : 0] = v; else fbs[g] = v;
    }
    __syncwarp();
  };
  if (COEF_REGS) {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = lane * PPL + j;
#pragma unroll
      for (int k = 0; k < 7; k++) S.w[COEF_REGS ? j : 0][k] = G.A(C_W0 + k)[g];
      S.w0a[COEF_REGS ? j : 0] = G.A(C_W0A)[g]; S.ca[COEF_REGS ? j : 0] = G.A(F_CA)[g];
      S.gs[COEF_REGS ? j : 0] = G.A(F_GS)[g]; S.kdc[COEF_REGS ? j : 0] = G.A(F_KDC)[g];
      S.p3[COEF_REGS ? j : 0] = G.A(F_P3)[g]; S.qc[COEF_REGS ? j : 0] = G.A(F_QC)[g];
      S.ak[COEF_REGS ? j : 0] = G.A(F_AK)[g];
    }
  }
  load_signed_floor();
  load_state();

  // publish the initial state (boundary conditions applied) into line 0
  auto publish = [&](int buf) {
    double (*Ln)[MAG_FAST_MAXPPL * 32] = rs->line[buf];
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      if (pub_mask & (1u << j)) {
        const bool z = zero_mask & (1u << j);
        Ln[0][j * 32 + lane] = z ? 0.0 : S.f[0][j];
        Ln[1][j * 32 + lane] = z ? 0.0 : S.f[1][j];
        Ln[2][j * 32 + lane] = S.f[2][j];
      }
      if (mir_mask & (1u << j)) {
        const int a = mir_addr[j];
        Ln[0][a] = -S.f[0][j]; Ln[1][a] = -S.f[1][j]; Ln[2][a] = S.f[2][j];
      }
    }
  };
  // unused slots of both lines must read as zero
  for (int i = lane; i < 2 * 3 * MAG_FAST_MAXPPL * 32; i += 32) (&rs->line[0][0][0])[i] = 0.0;
  __syncwarp();

  const double dt = G.dt;
  const double dg0 = dt * (8.0 / 15), dg1 = dt * (5.0 / 12), dg2 = dt * (3.0 / 4);
  const double dz0 = dt * (-17.0 / 60), dz1 = dt * (-5.0 / 12);
  const double dtg = dt * G.a->U.t0_Gyr;
  const int nsteps = G.nsteps;
  double* ck[3] = {G.A(A_T0), G.A(A_T1), G.A(A_T2)};
  int buf = 0;
  int jt = 1;
  while (jt <= nsteps) {
    const int nb = min(MAG_FAST_BLK, nsteps - jt + 1);
    // ---- checkpoint
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = lane * PPL + j;
      if (g < nx) { ck[0][g] = S.f[0][j]; ck[1][g] = S.f[1][j]; ck[2][g] = S.f[2][j]; }
    }
    const double ck_t = G.t, ck_tl = G.t_last_sign_choice;
    const bool ck_refresh = G.refresh_sign;
    const int ck_sign_nx = G.sign_nx;
    static_assert(sizeof(RngState) == 17 * sizeof(uint64_t), "RngState layout");
    if (lane < 17) ((uint64_t*)&rs->rng_ck)[lane] = ((const uint64_t*)&rs->rng)[lane];
    bool signs_saved = false;
    unsigned mB = 0, mA = 0;
    publish(buf);
    for (int s = 0; s < nb; s++) {
      const double this_t = G.t_Gyr + dtg * (double)(jt + s);
      // same order of RNG draws as the reference: update_floor_sign (dynamo.f90:196-198) may
      // draw Bfloor_sign, then the first pde of the step redraws the layer signs if flagged
      // or if the grid size changed (floor_field.f90:55-79)
      update_floor_sign(G, this_t);
      if (G.refresh_sign || G.sign_nx != nx) {
        if (!signs_saved) {  // first refresh of the block: keep the old signs for a replay
          const double* sg = G.A(A_SIGN); double* bk = G.A(A_OMK);
          const int nsave = max(nx, ck_sign_nx);
          for (int i = lane; i < nsave; i += 32) bk[i] = sg[i];
          signs_saved = true;
          __syncwarp();
        }
        refresh_signs(G);
        load_signed_floor();
      }
      const bool last = (jt + s == nsteps);
      fast_stage<PPL, COEF_REGS, 0>(S, G, rs->line, buf, dg0, dz0, Rk, pub_mask, zero_mask, mir_mask, mir_addr, reload, mB, mA, false);
      buf ^= 1;
      fast_stage<PPL, COEF_REGS, 1>(S, G, rs->line, buf, dg1, dz1, Rk, pub_mask, zero_mask, mir_mask, mir_addr, reload, mB, mA, false);
      buf ^= 1;
      fast_stage<PPL, COEF_REGS, 2>(S, G, rs->line, buf, dg2, 0.0, Rk, pub_mask, zero_mask, mir_mask, mir_addr, reload, mB, mA, last);
      buf ^= 1;
      // t = t + dt*gam1 ; ttmp = t + dt*zet1 ; t = ttmp + dt*gam2 ; ttmp = t + dt*zet2 ; t = ttmp + dt*gam3
      double tt = G.t + dg0;
      tt = tt + dz0; tt = tt + dg1; tt = tt + dz1; tt = tt + dg2;
      G.t = tt;
    }
    // ---- accept or replay the block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mB = max(mB, __shfl_xor_sync(FULL, mB, o));
      mA = max(mA, __shfl_xor_sync(FULL, mA, o));
    }
    const double MB = __hiloint2double((int)mB + 1, 0), MA = __hiloint2double((int)mA + 1, 0);
    if (mB < 0x7fe00000u && mA < 0x7fe00000u && guard_ok(gk, dt, MB, MA)) {
      G.point_stages += (long long)nb * 3 * nx;
      jt += nb;
      continue;
    }
    // replay with the generic path from the checkpoint
    if (G.a->replays && lane == 0) atomicAdd(G.a->replays, 1);
    for (int v = 0; v < 3; v++) {
      double* f = G.A(A_F0 + v);
      for (int i = lane; i < nx; i += 32) f[i] = ck[v][i];
    }
    if (signs_saved) {
      double* sg = G.A(A_SIGN); const double* bk = G.A(A_OMK);
      const int nsave = max(nx, ck_sign_nx);
      for (int i = lane; i < nsave; i += 32) sg[i] = bk[i];
    }
    if (lane < 17) ((uint64_t*)&rs->rng)[lane] = ((const uint64_t*)&rs->rng_ck)[lane];
    G.t = ck_t; G.t_last_sign_choice = ck_tl; G.refresh_sign = ck_refresh; G.sign_nx = ck_sign_nx;
    __syncwarp();
    impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    if (!run_steps_generic(G, jt, nb)) return false;
    jt += nb;
    load_signed_floor();   // (a refresh still pending after the replay happens at the next step)
    load_state();
  }
  // ---- hand the state back to the workspace (the caller applies impose_bc)
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = lane * PPL + j;
    if (g < nx) { G.A(A_F0)[g] = S.f[0][j]; G.A(A_F1)[g] = S.f[1][j]; G.A(A_F2)[g] = S.f[2][j]; }
  }
  __syncwarp();
  return true;
}

// Returns ok (false: the field blew up, check_f_error).
__device__ inline bool run_timesteps(Gal& G, RkShared* rs) {
  if (G.nsteps < 1) return true;
  if (G.a->use_fast) {
    if (G.nx <= 128) return run_timesteps_fast<4, true>(G, rs);
    if (G.nx <= 256) return run_timesteps_fast<8, false>(G, rs);
  }
  return run_steps_generic(G, 1, G.nsteps);
}

}  // namespace magdev
