// mag_rk_warp.cuh -- the time-stepping loop of one snapshot (dynamo.f90:168-212) on ONE WARP
// per galaxy, state and coefficients in registers, halo exchange by warp shuffles.
//
// Layout.  The live points of the grid are g = 3 .. m (m = nx-4; gutsdynamo.f90:55-86:
// g = 3 and g = m are the Dirichlet points of Br, Bp; g < 3 and g > m are ghost zones).
// With p = g - 3 in [0, P] (P = nx-7), slot s = p + J0 belongs to lane s / PPL, register
// j = s % PPL.  J0 = PPL-1 - P % PPL puts the LAST live point into the last register of its
// lane (lane tm), so the outer ghost zone is exactly that lane's right halo; the inner ghost
// zone is the J0 leading registers of lane 0 plus its left halo.  Ghost values are never
// stored: after the shuffles the three lanes that see a ghost point (0, 1 and tm) overwrite
// those window entries with the mirror image (negated for Br, Bp) of entries they already
// hold -- a handful of predicated register moves with compile-time indices.
//
// Per stage and thread: 36 SHFL (3 variables x 6 halo values x 2 words), then for each of the
// PPL points the folded right-hand side of mag_rk_fast.cuh with two more foldings
//   fb*(1 + kdc*alp) = c0 + fk*alp_m ,  c0 = fb*(1 + kdc*alp_k)  (c0 seeds the L(Bp) chain)
//   R_kappa = 1: w0a*alp_m seeds the alp_m stencil chain
// = 43 FP64 instructions per point and stage, no shared-memory traffic, no barrier.
// PPL = 4 keeps all 15 per-point coefficients in registers (nx <= 134); PPL = 5, 6 (nx <= 198)
// read the 8 stencil weights of a point from shared memory ([j][k][lane], conflict free).
//
// The block / checkpoint / replay protocol that makes check_f_error (gutsdynamo.f90:450-463)
// exact is the one described in mag_rk_fast.cuh.
#pragma once

namespace magdev {

// Coefficient placement of a variant (MODE):
//   0  all 15 per-point coefficients in registers                       (PPL = 4)
//   1  the 8 stencil weights in shared memory, the rest in registers    (PPL = 5)
//   2  stencil weights and p3, qc in shared memory                      (PPL = 6)
//   3  everything streamed from a pair table in the team workspace (global memory, L1/L2
//      resident), only f and ftmp in registers                          (PPL = 8, 11)
// Pair tables are double2 [j*NQ + q][lane]: one conflict-free / coalesced 128-bit load per
// pair.  q = 0..3: (w0,w1) (w2,w3) (w4,w5) (w6,w0a); shared-memory tables add q = 4: (p3,qc);
// the global table adds q = 4..7: (ca,gs) (c0,fk) (ak,p3) (qc,-).
// The loads are volatile asm so that the compiler cannot hoist them out of the time loop (the
// table is loop invariant; hoisting would put it back into registers and spill).
#define WM_REGS 0
#define WM_WSM 1
#define WM_WSM2 2
#define WM_ALLG 3
__device__ __forceinline__ double2 lds_pair(unsigned saddr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(saddr));
  return r;
}
__device__ __forceinline__ double2 ldg_pair(const double2* p) {
  double2 r;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
#define WSLOT2(j, q) (((j) * 5 + (q)) * 32)
#define GSLOT2(j, q) (((j) * 8 + (q)) * 32)
#define WARP_TABLE_ARRAY A_GK   // A_GK .. A_HKPC (6 arrays) are free in the evolve phase
#define WARP_TABLE_ARRAYS 6

template <int PPL, int MODE>
struct WarpState {
  double f[3][PPL], ft[3][PPL];
  double w[MODE == WM_REGS ? PPL : 1][8];
  double ca[MODE == WM_ALLG ? 1 : PPL], gs[MODE == WM_ALLG ? 1 : PPL], c0[MODE == WM_ALLG ? 1 : PPL];
  double fk[MODE == WM_ALLG ? 1 : PPL], ak[MODE == WM_ALLG ? 1 : PPL];
  double p3[MODE >= WM_WSM2 ? 1 : PPL], qc[MODE >= WM_WSM2 ? 1 : PPL];
};

// ---- boundary conditions folded into the stencil weights.  With ghost value = sigma * mirror
// value (sigma = -1 for Br, Bp; +1 for alp_m) the weight of a ghost offset o of a point at
// distance d from the Dirichlet point moves onto its mirror offset o' = -2d - o:
//   d = 0 (Dirichlet point; only alp_m evolves there): w[o'] += w[o], w[o] = 0
//   o' = 0 (the mirror is the point itself): centre weight of B -= w[o], centre weight of
//          alp_m += R_kappa*w[o], w[o] = 0
//   else: w[o'] -= w[o] (exact for Br, Bp if their ghost entry reads 0) and w[o] *= 2, so that
//          alp_m, whose ghost entry holds the mirror value, sees w[o'] + w[o].
// Only four ghost entries keep a non-zero weight: p = -2, -1 (used by p = 1, 2) and P+1, P+2
// (used by P-2, P-1).  The values Br, Bp take AT the Dirichlet points before impose_bc resets
// them are not evaluated exactly any more; guard_constants bounds them with the ghost points.
__device__ __noinline__ void fold_boundary_weights(Gal& G) {
  const int nx = G.nx, m = nx - MAG_NGHOST - 1;
  const double Rk = G.a->P.R_kappa;
  if (G.lane < 2) {
    const bool left = G.lane == 0;
    for (int d = 0; d <= 2; d++) {
      const int g = left ? MAG_NGHOST + d : m - d;
      double w[7], wc, w0a;
      for (int k = 0; k < 7; k++) w[k] = G.A(C_W0 + k)[g];
      wc = w[3]; w0a = G.A(C_W0A)[g];
      for (int o = -3; o < -d; o++) {          // ghost offsets (inward positive)
        const int io = left ? 3 + o : 3 - o;
        const int om = -2 * d - o, im = left ? 3 + om : 3 - om;
        if (d == 0) { w[im] += w[io]; w[io] = 0.0; }
        else if (om == 0) { wc -= w[io]; w0a += Rk * w[io]; w[io] = 0.0; }
        else { w[im] -= w[io]; w[io] *= 2.0; }
      }
      w[3] = wc;
      for (int k = 0; k < 7; k++) G.A(C_W0 + k)[g] = w[k];
      G.A(C_W0A)[g] = w0a;
    }
  }
  __syncwarp();
}

// One Runge-Kutta stage.  STAGE: 0,1,2.  R_kappa = 1 (folded into the alp_m chain).
template <int PPL, int J0, int MODE, int STAGE>
__device__ __forceinline__ void warp_stage(WarpState<PPL, MODE>& S, const unsigned wsa, const double2* __restrict__ tab, const int lane,
                                           const int tm, const double dg, const double dz, unsigned& mBs, unsigned& mBu,
                                           unsigned& mAs, unsigned& mAu, double* __restrict__ alp_out, const int g0, const int m) {
  // lanes / window indices of the four alp_m ghost entries with a non-zero weight
  constexpr int L1 = (J0 + 1) / PPL, I1 = 3 + J0 - 2 - PPL * L1, S1 = 3 + J0 + 2 - PPL * L1;   // p = -2 <- p = 2, for p = 1
  constexpr int L2 = (J0 + 2) / PPL, I2 = 3 + J0 - 1 - PPL * L2, S2 = 3 + J0 + 1 - PPL * L2;   // p = -1 <- p = 1, for p = 2
  constexpr int LD = J0 / PPL, JD = J0 % PPL;                                                // inner Dirichlet point
  static_assert(I1 >= 0 && I2 >= 0 && S1 <= PPL + 5 && S2 <= PPL + 5 && J0 >= 2 && J0 <= PPL + 1, "warp_stage layout");
  const bool isM = lane == tm;
  // ---- windows: own registers + 3 halo values each side
  double win[3][PPL + 6];
#pragma unroll
  for (int v = 0; v < 3; v++) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      win[v][k] = __shfl_up_sync(FULL, S.f[v][PPL - 3 + k], 1);
      win[v][PPL + 3 + k] = __shfl_down_sync(FULL, S.f[v][k], 1);
    }
#pragma unroll
    for (int j = 0; j < PPL; j++) win[v][3 + j] = S.f[v][j];
  }
  // Neumann ghost entries of alp_m (Br, Bp read the zeros of the inert slots)
  if (lane == L1) win[2][I1] = win[2][S1];
  if (lane == L2) win[2][I2] = win[2][S2];
  if (isM) { win[2][PPL + 3] = win[2][PPL + 1]; win[2][PPL + 4] = win[2][PPL]; }
  // ---- point by point: three stencil chains, point-wise terms, Runge-Kutta update
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    double w[8], c_ca, c_gs, c_c0, c_fk, c_ak, c_p3, c_qc;
    if (MODE == WM_REGS) {
#pragma unroll
      for (int k = 0; k < 8; k++) w[k] = S.w[MODE == WM_REGS ? j : 0][k];
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double2 t = MODE == WM_ALLG ? ldg_pair(tab + GSLOT2(j, q)) : lds_pair(wsa + 16u * WSLOT2(j, q));
        w[2 * q] = t.x; w[2 * q + 1] = t.y;
      }
    }
    if (MODE == WM_ALLG) {
      const double2 t4 = ldg_pair(tab + GSLOT2(j, 4)), t5 = ldg_pair(tab + GSLOT2(j, 5)), t6 = ldg_pair(tab + GSLOT2(j, 6));
      const double2 t7 = ldg_pair(tab + GSLOT2(j, 7));
      c_ca = t4.x; c_gs = t4.y; c_c0 = t5.x; c_fk = t5.y; c_ak = t6.x; c_p3 = t6.y; c_qc = t7.x;
    } else {
      constexpr int jj = MODE == WM_ALLG ? 0 : 1;
      c_ca = S.ca[j * jj]; c_gs = S.gs[j * jj]; c_c0 = S.c0[j * jj]; c_fk = S.fk[j * jj]; c_ak = S.ak[j * jj];
      if (MODE == WM_WSM2) {
        const double2 t = lds_pair(wsa + 16u * WSLOT2(j, 4));
        c_p3 = t.x; c_qc = t.y;
      } else {
        c_p3 = S.p3[MODE >= WM_WSM2 ? 0 : j]; c_qc = S.qc[MODE >= WM_WSM2 ? 0 : j];
      }
    }
    const double Br = win[0][3 + j], Bp = win[1][3 + j], am = win[2][3 + j];
    // chains start at the centre and work outwards, so that the halo terms come last
    double a0 = w[3] * Br, a1 = fma(w[3], Bp, c_c0), a2 = w[7] * am;
#pragma unroll
    for (int d = 1; d <= 3; d++) {
      a0 = fma(w[3 - d], win[0][j + 3 - d], a0); a1 = fma(w[3 - d], win[1][j + 3 - d], a1); a2 = fma(w[3 - d], win[2][j + 3 - d], a2);
      a0 = fma(w[3 + d], win[0][j + 3 + d], a0); a1 = fma(w[3 + d], win[1][j + 3 + d], a1); a2 = fma(w[3 + d], win[2][j + 3 + d], a2);
    }
    const double a = c_ak + am;
    if (alp_out) { const int g = g0 + j; if (g >= MAG_NGHOST && g <= m) alp_out[g] = a; }
    const double p0 = fma(-(c_ca * a), Bp, a0);
    const double p1 = fma(c_fk, am, fma(c_gs, Br, a1));
    const double B2 = fma(Br, Br, Bp * Bp);
    const double em = fma(c_qc * fast_sqrt_abs(a), Br * Bp, a * B2);
    const double p2 = fma(c_p3, em, a2);
    const double b0 = (STAGE == 0) ? Br : S.ft[0][j];
    const double b1 = (STAGE == 0) ? Bp : S.ft[1][j];
    const double b2 = (STAGE == 0) ? am : S.ft[2][j];
    const double n0 = fma(dg, p0, b0), n1 = fma(dg, p1, b1), n2 = fma(dg, p2, b2);
    S.f[0][j] = n0; S.f[1][j] = n1; S.f[2][j] = n2;
    if (STAGE < 2) { S.ft[0][j] = fma(dz, p0, n0); S.ft[1][j] = fma(dz, p1, n1); S.ft[2][j] = fma(dz, p2, n2); }
    // |x| maxima on the high words: signed max catches the positive values, unsigned max the negative ones
    mBs = (unsigned)__vimax3_s32((int)mBs, __double2hiint(n0), __double2hiint(n1));
    mBu = __vimax3_u32(mBu, (unsigned)__double2hiint(n0), (unsigned)__double2hiint(n1));
    mAs = (unsigned)max((int)mAs, __double2hiint(n2));
    mAu = max(mAu, (unsigned)__double2hiint(n2));
  }
  // ---- Dirichlet points of Br, Bp: the state the next pde call sees is 0 (ftmp keeps the
  // unreset value, like the reference's ftmp)
  if (lane == LD) { S.f[0][JD] = 0.0; S.f[1][JD] = 0.0; }
  if (isM) { S.f[0][PPL - 1] = 0.0; S.f[1][PPL - 1] = 0.0; }
}

// Fast time loop of one snapshot attempt on one warp.  Returns ok (false: check_f_error).
template <int PPL, int J0, int MODE>
__device__ __noinline__ bool warp_loop(Gal& G, RkShared* rs) {
  const int lane = G.lane;
  double* const W = G.W;
  const int nxcap = G.a->nxcap, nx = G.nx, m = nx - MAG_NGHOST - 1, nsteps = G.nsteps;
  const int P = nx - 2 * MAG_NGHOST - 1;
  const int tm = (P + J0) / PPL;
  const double dt = G.dt;
  auto WA = [&](int id) { return W + (size_t)id * nxcap; };
  const int g0 = lane * PPL - J0 + MAG_NGHOST;   // grid index of register 0
  double2* const wt2 = reinterpret_cast<double2*>(rs->pool) + lane;                          // shared pair table, this lane's column
  double2* const tab = reinterpret_cast<double2*>(W + (size_t)WARP_TABLE_ARRAY * nxcap) + lane;   // global pair table
  const unsigned wsa = (unsigned)__cvta_generic_to_shared(rs->pool) + 16u * lane;

  prepare_fast_coefs(G, nx);
  guard_constants(G, rs->gk);
  fold_boundary_weights(G);

  WarpState<PPL, MODE> S;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = g0 + j;
    const bool live = g >= MAG_NGHOST && g <= m;
    const int gc = live ? g : 0;
    double w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = live ? WA(C_W0 + k)[gc] : 0.0;   // C_W0..C_W6, C_W0A are consecutive
    const double ca = live ? WA(F_CA)[gc] : 0.0, gs = live ? WA(F_GS)[gc] : 0.0, ak = live ? WA(F_AK)[gc] : 0.0;
    const double p3 = live ? WA(F_P3)[gc] : 0.0, qc = live ? WA(F_QC)[gc] : 0.0;
    if (MODE == WM_REGS) {
#pragma unroll
      for (int k = 0; k < 8; k++) S.w[MODE == WM_REGS ? j : 0][k] = w[k];
    } else if (MODE == WM_ALLG) {
#pragma unroll
      for (int q = 0; q < 4; q++) tab[GSLOT2(j, q)] = make_double2(w[2 * q], w[2 * q + 1]);
      tab[GSLOT2(j, 4)] = make_double2(ca, gs);
      tab[GSLOT2(j, 6)] = make_double2(ak, p3);
      tab[GSLOT2(j, 7)] = make_double2(qc, 0.0);
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++) wt2[WSLOT2(j, q)] = make_double2(w[2 * q], w[2 * q + 1]);
      if (MODE == WM_WSM2) wt2[WSLOT2(j, 4)] = make_double2(p3, qc);
    }
    if (MODE != WM_ALLG) { S.ca[MODE == WM_ALLG ? 0 : j] = ca; S.gs[MODE == WM_ALLG ? 0 : j] = gs; S.ak[MODE == WM_ALLG ? 0 : j] = ak; }
    if (MODE < WM_WSM2) { S.p3[MODE >= WM_WSM2 ? 0 : j] = p3; S.qc[MODE >= WM_WSM2 ? 0 : j] = qc; }
  }
  auto load_signed_floor = [&]() {
    const double *cFB = WA(C_FB), *sg = WA(A_SIGN), *kd = WA(F_KDC), *akp = WA(F_AK);
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = g0 + j;
      const bool live = g >= MAG_NGHOST && g <= m;
      const int gc = live ? g : 0;
      const double fb = live ? sg[gc] * cFB[gc] : 0.0;
      const double fk = live ? fb * kd[gc] : 0.0;
      const double c0 = fma(fk, live ? akp[gc] : 0.0, fb);
      if (MODE == WM_ALLG) tab[GSLOT2(j, 5)] = make_double2(c0, fk);
      else { S.fk[MODE == WM_ALLG ? 0 : j] = fk; S.c0[MODE == WM_ALLG ? 0 : j] = c0; }
    }
    __syncwarp();
    asm volatile("" ::: "memory");   // table stores stay above the volatile pair loads of the stages
  };
  auto load_state = [&]() {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = g0 + j;
      const bool live = g >= MAG_NGHOST && g <= m;
      const int gc = live ? g : 0;
#pragma unroll
      for (int v = 0; v < 3; v++) { S.f[v][j] = live ? WA(A_F0 + v)[gc] : 0.0; S.ft[v][j] = 0.0; }
    }
    if (lane == J0 / PPL) { S.f[0][J0 % PPL] = 0.0; S.f[1][J0 % PPL] = 0.0; }
    if (lane == tm) { S.f[0][PPL - 1] = 0.0; S.f[1][PPL - 1] = 0.0; }
  };
  load_signed_floor();
  load_state();
  __syncwarp();

  const double dg0 = dt * (8.0 / 15), dg1 = dt * (5.0 / 12), dg2 = dt * (3.0 / 4);
  const double dz0 = dt * (-17.0 / 60), dz1 = dt * (-5.0 / 12);
  const double dtg = dt * G.a->U.t0_Gyr, t_Gyr = G.t_Gyr, tmk = G.a->P.p_floor_kappa * G.tau_min;
  double t = G.t, t_last = G.t_last_sign_choice;
  int sign_nx = G.sign_nx;
  bool refresh = G.refresh_sign;
  long long stages = 0;
  double* ck[3] = {WA(A_T0), WA(A_T1), WA(A_T2)};
  int jt = 1;
  bool ok = true;
  while (jt <= nsteps) {
    const int nb = min(MAG_FAST_BLK, nsteps - jt + 1);
    // ---- checkpoint (live points; the replay applies impose_bc)
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = g0 + j;
      if (g >= MAG_NGHOST && g <= m) { ck[0][g] = S.f[0][j]; ck[1][g] = S.f[1][j]; ck[2][g] = S.f[2][j]; }
    }
    const double ck_t = t, ck_tl = t_last;
    const bool ck_refresh = refresh;
    const int ck_sign_nx = sign_nx;
    if (lane < 17) ((uint64_t*)&rs->rng_ck)[lane] = ((const uint64_t*)&rs->rng)[lane];
    __syncwarp();
    bool signs_saved = false;
    unsigned mBs = 0, mBu = 0, mAs = 0, mAu = 0;   // running maxima start from the block's initial state
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      mBs = (unsigned)__vimax3_s32((int)mBs, __double2hiint(S.f[0][j]), __double2hiint(S.f[1][j]));
      mBu = __vimax3_u32(mBu, (unsigned)__double2hiint(S.f[0][j]), (unsigned)__double2hiint(S.f[1][j]));
      mAs = (unsigned)max((int)mAs, __double2hiint(S.f[2][j]));
      mAu = max(mAu, (unsigned)__double2hiint(S.f[2][j]));
    }
    for (int s = 0; s < nb; s++) {
      const double this_t = t_Gyr + dtg * (double)(jt + s);
      // same order of RNG draws as the reference: update_floor_sign (dynamo.f90:196-198) may
      // draw Bfloor_sign, then the first pde of the step redraws the layer signs if flagged
      // or if the grid size changed (floor_field.f90:55-79)
      const bool due = this_t - t_last > tmk;
      if (due || refresh || sign_nx != nx) {
        if (!signs_saved) {  // first refresh of the block: keep the old signs for a replay
          const double* sg = WA(A_SIGN); double* bk = WA(A_OMK);
          const int nsave = max(nx, ck_sign_nx);
          for (int i = lane; i < nsave; i += 32) bk[i] = sg[i];
          __syncwarp();
        }
        if (due) (void)random_sign(G);
        refresh_signs(G);
        signs_saved = true;
        if (due) t_last = this_t;
        refresh = false; sign_nx = nx;
        load_signed_floor();
      }
      double* alp_out = (jt + s == nsteps) ? WA(A_ALP) : nullptr;
      warp_stage<PPL, J0, MODE, 0>(S, wsa, tab, lane, tm, dg0, dz0, mBs, mBu, mAs, mAu, nullptr, g0, m);
      warp_stage<PPL, J0, MODE, 1>(S, wsa, tab, lane, tm, dg1, dz1, mBs, mBu, mAs, mAu, nullptr, g0, m);
      warp_stage<PPL, J0, MODE, 2>(S, wsa, tab, lane, tm, dg2, 0.0, mBs, mBu, mAs, mAu, alp_out, g0, m);
      // t = t + dt*gam1 ; ttmp = t + dt*zet1 ; t = ttmp + dt*gam2 ; ttmp = t + dt*zet2 ; t = ttmp + dt*gam3
      double tt = t + dg0;
      tt = tt + dz0; tt = tt + dg1; tt = tt + dz1; tt = tt + dg2;
      t = tt;
    }
    // ---- accept or replay the block.  Per-lane maxima of lanes 0,1 / tm bound the 7 points
    // next to the inner / outer boundary (a superset of them: conservative).
    const unsigned mB = max(mBs, mBu & 0x7fffffffu), mA = max(mAs, mAu & 0x7fffffffu);
    const unsigned nB0 = max(__shfl_sync(FULL, mB, 0), __shfl_sync(FULL, mB, 1));
    const unsigned nA0 = max(__shfl_sync(FULL, mA, 0), __shfl_sync(FULL, mA, 1));
    const unsigned nB1 = __shfl_sync(FULL, mB, tm), nA1 = __shfl_sync(FULL, mA, tm);
    unsigned rB = mB, rA = mA;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rB = max(rB, __shfl_xor_sync(FULL, rB, o));
      rA = max(rA, __shfl_xor_sync(FULL, rA, o));
    }
    bool accept = rB < 0x7fe00000u && rA < 0x7fe00000u;
    if (accept) {
      const double MB = __hiloint2double((int)rB + 1, 0), MA = __hiloint2double((int)rA + 1, 0);
      accept = MB < 1e8 && MA < 1000.0;
      if (accept) accept = guard_side_ok(rs->gk[0], dt, __hiloint2double((int)nB0 + 1, 0), __hiloint2double((int)nA0 + 1, 0));
      if (accept) accept = guard_side_ok(rs->gk[1], dt, __hiloint2double((int)nB1 + 1, 0), __hiloint2double((int)nA1 + 1, 0));
    }
    if (accept) {
      stages += (long long)nb * 3 * nx;
      jt += nb;
      continue;
    }
    // ---- replay with the generic path from the checkpoint
    if (G.a->replays && lane == 0) atomicAdd(G.a->replays, 1);
    for (int v = 0; v < 3; v++) {
      double* f = WA(A_F0 + v);
      for (int i = lane; i < nx; i += 32) f[i] = (i >= MAG_NGHOST && i <= m) ? ck[v][i] : 0.0;
    }
    if (signs_saved) {
      double* sg = WA(A_SIGN); const double* bk = WA(A_OMK);
      const int nsave = max(nx, ck_sign_nx);
      for (int i = lane; i < nsave; i += 32) sg[i] = bk[i];
    }
    if (lane < 17) ((uint64_t*)&rs->rng)[lane] = ((const uint64_t*)&rs->rng_ck)[lane];
    G.t = ck_t; G.t_last_sign_choice = ck_tl; G.refresh_sign = ck_refresh; G.sign_nx = ck_sign_nx;
    const long long ps0 = G.point_stages;
    __syncwarp();
    impose_bc(G, WA(A_F0), WA(A_F1), WA(A_F2));
    ok = run_steps_generic(G, jt, nb);
    stages += G.point_stages - ps0;
    G.point_stages = ps0;
    t = G.t; t_last = G.t_last_sign_choice; refresh = G.refresh_sign; sign_nx = G.sign_nx;
    if (!ok) break;
    jt += nb;
    load_signed_floor();
    load_state();
  }
  // ---- hand the state back to the workspace (the caller applies impose_bc)
  if (ok) {
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const int g = g0 + j;
      if (g >= MAG_NGHOST && g <= m) { WA(A_F0)[g] = S.f[0][j]; WA(A_F1)[g] = S.f[1][j]; WA(A_F2)[g] = S.f[2][j]; }
    }
  }
  G.t = t; G.t_last_sign_choice = t_last; G.refresh_sign = refresh; G.sign_nx = sign_nx;
  G.point_stages += stages;
  __syncwarp();
  return ok;
}

// Layout of a grid of nx points: registers per lane and the offset J0 in [2, PPL+1] that puts the
// last live point into the last register of a lane <= 30 (lane 31 must stay inert: it feeds
// the zeros of the outer ghost zone); ppl = 0 if the warp path does not take this grid.
__device__ __forceinline__ void warp_layout(int nx, int& ppl, int& j0) {
  const int P = nx - 2 * MAG_NGHOST - 1;
  const int cand[5] = {4, 5, 6, 8, 11};
  for (int c = 0; c < 5; c++) {
    ppl = cand[c];
    j0 = ppl - 1 - P % ppl;
    if (j0 < 2) j0 += ppl;
    if ((P + j0) / ppl <= 30) return;
  }
  ppl = 0;
}
// (R_kappa = 1, the reference's default, is folded into the alp_m stencil chain; other values
// take the shared-line path of mag_rk_fast.cuh)
__device__ __forceinline__ bool warp_path_applies(const Gal& G) {
  int ppl, j0;
  warp_layout(G.nx, ppl, j0);
  if (ppl >= 8 && (size_t)WARP_TABLE_ARRAYS * G.a->nxcap < (size_t)16 * 32 * ppl) return false;   // pair table does not fit
  return G.a->P.R_kappa == 1.0 && G.nx >= 64 && ppl != 0;
}

template <int PPL, int MODE, int J>
struct WarpJ0 {
  static __device__ __forceinline__ bool run(Gal& G, RkShared* rs, int j0) {
    if (j0 == J) return warp_loop<PPL, J, MODE>(G, rs);
    return WarpJ0<PPL, MODE, J + 1>::run(G, rs, j0);
  }
};
template <int PPL, int MODE>
struct WarpJ0<PPL, MODE, PPL + 2> {
  static __device__ __forceinline__ bool run(Gal&, RkShared*, int) { return false; }
};
__device__ __forceinline__ bool warp_dispatch(Gal& G, RkShared* rs) {
  int ppl, j0;
  warp_layout(G.nx, ppl, j0);
  switch (ppl) {
    case 4: return WarpJ0<4, WM_REGS, 2>::run(G, rs, j0);
    case 5: return WarpJ0<5, WM_WSM, 2>::run(G, rs, j0);
    case 6: return WarpJ0<6, WM_WSM2, 2>::run(G, rs, j0);
    case 8: return WarpJ0<8, WM_ALLG, 2>::run(G, rs, j0);
    default: return WarpJ0<11, WM_ALLG, 2>::run(G, rs, j0);
  }
}

}  // namespace magdev
