// mag_rk_warp.cuh -- the time-stepping loop of one snapshot (dynamo.f90:168-212) on ONE WARP
// per galaxy: the state (Br, Bp, alp_m, ftmp) lives in registers for the whole snapshot, the
// halo is exchanged by warp shuffles, there is no barrier and no shared-memory line.
//
// Layout.  The live points of the grid are g = 3 .. m (m = nx-4; gutsdynamo.f90:55-86:
// g = 3 and g = m are the Dirichlet points of Br, Bp; g < 3 and g > m are ghost zones).
// With p = g - 3 in [0, P] (P = nx-7), slot s = p + 3 belongs to lane s / PPL, register
// j = s % PPL: the inner end is fixed (p = 0 in register 3 of lane 0, p = -3..-1 in its
// registers 0..2), the outer end falls on register jm = (P+3) % PPL of lane tm = (P+3) / PPL,
// a run-time position that one warp-uniform branch per stage looks at.  The loop is
// therefore instantiated once per PPL and not once per alignment, and the three Runge-Kutta
// stages share ONE looped body (ftmp == f at the start of a step): the two warps of an SM
// sub-partition then fit into its instruction cache even when they work on different grid
// classes.  (Measured: with one unrolled 14 KB loop per alignment ncu shows 15
// stall_no_instruction cycles per issue on the mixed catalogue and the throughput drops 4x,
// although the same loops run 40 % faster when all warps of the SM execute the same one.)
//
// Boundary conditions (impose_bc) cost nothing for Br and Bp: slots outside the live range
// are inert (all coefficients zero, value 0) and the mirror condition is folded into the
// stencil weights (fold_boundary_weights), which also keeps Br = Bp = 0 at the Dirichlet
// points without ever resetting them.  alp_m needs three mirrored values per end, copied into
// the window of the lanes that use them (warp_ghosts).
//
// Per stage and thread: 36 SHFL (3 variables x 6 halo values x 2 words) and, per point, the
// folded right-hand side of mag_rk_fast.cuh with two more foldings
//   fb*(1 + kdc*alp) = c0 + fk*alp_m ,  c0 = fb*(1 + kdc*alp_k)  (c0 seeds the L(Bp) chain)
//   R_kappa = 1: w0a*alp_m seeds the alp_m stencil chain
// = 43 FP64 instructions per point and stage; the stage is written in three phases whose
// statements are interleaved across the points of the lane (warp_stage).
//
// The block / checkpoint / replay protocol that makes check_f_error (gutsdynamo.f90:450-463)
// exact is the one described in mag_rk_fast.cuh.
#pragma once

namespace magdev {

// Coefficient placement of a variant (MODE):
//   0  all 15 per-point coefficients in registers                       (not used by default)
//   1  the 8 stencil weights in shared memory, the rest in registers    (PPL = 4, 5: nx <= 160)
//   2  stencil weights and p3, qc in shared memory                      (PPL = 6: nx <= 192)
//   3  everything streamed from a pair table in the team workspace (global memory, L1/L2
//      resident), only f and ftmp in registers                          (PPL = 8, 11: nx <= 352)
// Pair tables are double2 [j*NQ + q][lane]: one conflict-free / coalesced 128-bit load per
// pair.  q = 0..3: (w0,w1) (w2,w3) (w4,w5) (w6,w0a); shared-memory tables add q = 4: (p3,qc);
// the global table adds q = 4..7: (ca,gs) (c0,fk) (ak,p3) (qc,-).
// The loads are volatile asm so that the compiler cannot hoist them out of the time loop (the
// table is loop invariant; hoisting would put it back into registers and spill).
// Build-time experiments (see the header): MAG_WARP_UNROLL_MAXPPL unrolls the three stages up
// to that PPL, MAG_WARP_SPECIALIZE_DEFAULT adds an unrolled instantiation with a compile-time
// outer alignment for the default grid (nx = 107).  Both lose on the mixed catalogue.
#ifndef MAG_W4_MODE
#define MAG_W4_MODE WM_WSM   // same speed as WM_REGS on the looped body, smaller code, no spills
#endif
#ifndef MAG_WARP_SPECIALIZE_DEFAULT
#define MAG_WARP_SPECIALIZE_DEFAULT 0
#endif
#ifndef MAG_WARP_UNROLL_MAXPPL
#define MAG_WARP_UNROLL_MAXPPL 0
#endif
#ifndef MAG_WARP_UNROLL_JM
#define MAG_WARP_UNROLL_JM 0      // 1: the instantiations with a compile-time outer alignment unroll their three stages
#endif
#ifndef MAG_WARP_DUE_AHEAD
#define MAG_WARP_DUE_AHEAD 0     // the step of the next floor-sign refresh is found ahead of the steps of a block
#endif
#ifndef MAG_WARP_OPAQUE_ADDR
#define MAG_WARP_OPAQUE_ADDR 0   // shared addresses of the pair table / stage constants kept in registers (no S2UR + ULEA per stage)
#endif
#ifndef MAG_WARP_STAGE_LOOP2
#define MAG_WARP_STAGE_LOOP2 0   // the stage loop stops once for save_alp() instead of testing for it in every stage
#endif
#ifndef MAG_WARP_GHOST_SELECT
#define MAG_WARP_GHOST_SELECT 0  // outer ghosts of grids that fill their last lane: selects, no branch
#endif
#ifndef MAG_WARP_ALIGNED4
#define MAG_WARP_ALIGNED4 0       // own instantiation of the four-register loop for grids that fill their last lane (nx = 107: 90 % of its stages)
#endif
#define WM_REGS 0
#define WM_WSM 1
#define WM_WSM2 2
#define WM_ALLG 3
__device__ __forceinline__ double2 lds_pair(unsigned saddr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(saddr));
  return r;
}
__device__ __forceinline__ double lds_f64x(unsigned saddr) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(saddr));
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
  double hl[3][3], hr[3][3];   // halo: the last 3 registers of lane-1, the first 3 of lane+1 (filled by warp_exchange)
  double w[MODE == WM_REGS ? PPL : 1][8];
  double ca[MODE == WM_ALLG ? 1 : PPL], gs[MODE == WM_ALLG ? 1 : PPL], c0[MODE == WM_ALLG ? 1 : PPL];
  double fk[MODE == WM_ALLG ? 1 : PPL], ak[MODE == WM_ALLG ? 1 : PPL];
  double p3[MODE >= WM_WSM2 ? 1 : PPL], qc[MODE >= WM_WSM2 ? 1 : PPL];
};

// ---- boundary conditions folded into the stencil weights.  With ghost value = sigma * mirror
// value (sigma = -1 for Br, Bp; +1 for alp_m) the weight of a ghost offset o of a point at
// distance d from the Dirichlet point meets its mirror offset o' = -2d - o:
//   d = 0 (the Dirichlet point itself): w[o] += w[o'], w[o'] = 0 -- the whole weight sits on the
//          ghost side, where Br and Bp read zeros: with ca = gs = c0 = fk = 0 at that point
//          Br = Bp = 0 is then preserved exactly without ever resetting it, while alp_m, whose
//          ghost entries hold the mirror values, sees w[o] + w[o'];
//   o' = 0 (the mirror is the point itself): centre weight of B -= w[o], centre weight of
//          alp_m += R_kappa*w[o], w[o] = 0;
//   else: w[o'] -= w[o] (exact for Br, Bp, ghost entry 0) and w[o] *= 2, so that alp_m sees
//          w[o'] + w[o].
// The only boundary work left in a stage is to copy three alp_m values per end into the
// ghost entries of the window (mirror image about the Dirichlet point).  The values Br, Bp
// take AT the Dirichlet points before impose_bc resets them in the reference are not
// evaluated; guard_constants bounds them together with the ghost points.
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
        if (d == 0) { w[io] += w[im]; w[im] = 0.0; }
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

// One Runge-Kutta stage; R_kappa = 1 (folded into the alp_m chain).  The same code serves the
// three stages: ftmp == f holds at the start of every step, so stage 1 can read its base from
// ftmp like the others, and stage 3 re-establishes it (LAST, or dz = 0: fma(0, p, n) = n).
// Window entry i of alp_m (i = 0..2: left halo, 3..PPL+2: own registers, PPL+3..PPL+5: right halo)
template <int PPL, int MODE>
__device__ __forceinline__ double& wref(WarpState<PPL, MODE>& S, const int i) {
  return i < 3 ? S.hl[2][i < 3 ? i : 0] : (i < PPL + 3 ? S.f[2][(i >= 3 && i < PPL + 3) ? i - 3 : 0] : S.hr[2][i >= PPL + 3 ? i - PPL - 3 : 0]);
}
// mirror image of alp_m about a Dirichlet point that sits at window entry IP of lane `owner`
template <int PPL, int MODE, int IP>
__device__ __forceinline__ void mirror_ghosts(WarpState<PPL, MODE>& S, const bool mine, const bool outward_up) {
#pragma unroll
  for (int k = 1; k <= 3; k++) {
    const int dst = outward_up ? IP + k : IP - k, src = outward_up ? IP - k : IP + k;
    if (dst >= 0 && dst <= PPL + 5 && src >= 0 && src <= PPL + 5) {
      const double v = wref<PPL, MODE>(S, src);
      if (mine) wref<PPL, MODE>(S, dst) = v;
    }
  }
}
template <int PPL, int MODE, int JMC>
__device__ __forceinline__ void outer_ghosts(WarpState<PPL, MODE>& S, const int lane, const int tm) {
  mirror_ghosts<PPL, MODE, 3 + JMC>(S, lane == tm, true);
  if (JMC <= 1) mirror_ghosts<PPL, MODE, PPL + 3 + JMC>(S, lane == tm - 1, true);
}

// Halo exchange after a stage (or after loading the state): 36 SHFL, then the Neumann ghost
// entries of alp_m.  Br and Bp need nothing: their ghost entries are inert registers that read
// 0.  JM >= 0: the outer Dirichlet point sits in register JM of lane tm (compile time);
// JM = -1: in register jm (one warp-uniform switch).  The inner one, p = 0, sits in register
// WARP_J0 = 3 of lane 0, so that p = -3..-1 are its (inert) registers 0..2.
#define WARP_J0 3
// up to this PPL the stage applies the ghost entries after its phase 0; the large-grid variants
// apply them right after the exchange (fewer values live across the stage: no spills)
#define WARP_LATE_GHOSTS_MAXPPL 6
#ifndef WARP_T_IN_STAGE_MAXPPL
#define WARP_T_IN_STAGE_MAXPPL 0
#endif
// 64-bit shuffles as two explicit 32-bit ones (the compiler's own lowering of a double shuffle
// lands the halves in swapped registers and then swaps them back with three LOP3 each)
__device__ __forceinline__ double shfl_down1(double x) {
  const int lo = __shfl_down_sync(FULL, __double2loint(x), 1), hi = __shfl_down_sync(FULL, __double2hiint(x), 1);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up1(double x) {
  const int lo = __shfl_up_sync(FULL, __double2loint(x), 1), hi = __shfl_up_sync(FULL, __double2hiint(x), 1);
  return __hiloint2double(hi, lo);
}

template <int PPL, int MODE>
__device__ __forceinline__ void warp_exchange(WarpState<PPL, MODE>& S) {
#pragma unroll
  for (int v = 0; v < 3; v++) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      S.hr[v][k] = shfl_down1(S.f[v][k]);
      S.hl[v][k] = shfl_up1(S.f[v][PPL - 3 + k]);
    }
  }
}
// The Neumann ghost entries of alp_m, applied by the stage after its phase 0 (so that the
// shuffles of warp_exchange have landed by then).
template <int PPL, int MODE, int JM>
__device__ __forceinline__ void warp_ghosts(WarpState<PPL, MODE>& S, const int lane, const int tm, const int jm) {
  // inner end: the ghost entries of p = 0 (all three), of p = 1 (p = -2) and of p = 2 (p = -1)
  // live in the windows of the lanes that own those points
  constexpr int J0 = WARP_J0;
  constexpr int LA = J0 / PPL, LB = (J0 + 1) / PPL, LC = (J0 + 2) / PPL;
  mirror_ghosts<PPL, MODE, 3 + J0 - PPL * LA>(S, lane == LA, false);
  if (LB != LA) mirror_ghosts<PPL, MODE, 3 + J0 - PPL * LB>(S, lane == LB, false);
  if (LC != LB) mirror_ghosts<PPL, MODE, 3 + J0 - PPL * LC>(S, lane == LC, false);
  // outer end
  if (JM >= 0) {
    outer_ghosts<PPL, MODE, (JM >= 0 ? JM : 0)>(S, lane, tm);
  } else if (!MAG_WARP_GHOST_SELECT && jm == PPL - 1) {   // the default grid (nx = 107) and every other grid that fills its last lane
    outer_ghosts<PPL, MODE, PPL - 1>(S, lane, tm);
  } else {
    // the default grid (nx = 107) and every other grid that fills its last lane: selects only, no branch (90 % of the stages)
    if (MAG_WARP_GHOST_SELECT) mirror_ghosts<PPL, MODE, 3 + PPL - 1>(S, lane == tm && jm == PPL - 1, true);
    if (!MAG_WARP_GHOST_SELECT || __builtin_expect(jm != PPL - 1, 0)) switch (jm) {
      case 0: outer_ghosts<PPL, MODE, 0>(S, lane, tm); break;
      case 1: outer_ghosts<PPL, MODE, 1>(S, lane, tm); break;
      case 2: outer_ghosts<PPL, MODE, 2>(S, lane, tm); break;
      case 3: outer_ghosts<PPL, MODE, 3>(S, lane, tm); break;
      case 4: if (PPL > 4) outer_ghosts<PPL, MODE, (PPL > 4 ? 4 : 0)>(S, lane, tm); break;
      case 5: if (PPL > 5) outer_ghosts<PPL, MODE, (PPL > 5 ? 5 : 0)>(S, lane, tm); break;
      case 6: if (PPL > 6) outer_ghosts<PPL, MODE, (PPL > 6 ? 6 : 0)>(S, lane, tm); break;
      case 7: if (PPL > 7) outer_ghosts<PPL, MODE, (PPL > 7 ? 7 : 0)>(S, lane, tm); break;
      case 8: if (PPL > 8) outer_ghosts<PPL, MODE, (PPL > 8 ? 8 : 0)>(S, lane, tm); break;
      default: if (PPL > 9) outer_ghosts<PPL, MODE, (PPL > 9 ? 9 : 0)>(S, lane, tm); break;
    }
  }
}

// One Runge-Kutta stage on the window (hl, f, hr) that warp_exchange left behind, written as
// three phases whose statements are interleaved ACROSS the points of the lane, so that the
// scheduler finds 4+ independent FP64 instructions at every moment (DFMA latency is 8 cycles at
// an issue interval of 2; with two warps per scheduler the chains have to overlap inside a warp):
//   phase 0  point-wise non-linear terms (own values only; overlaps the tail of the exchange
//            and the first coefficient loads)
//   phase 1  the 3*PPL stencil chains, two points at a time, own registers before halo
//            entries, the coefficient pairs of the next two points in flight
//   phase 2  right-hand sides and the Runge-Kutta update of f and ftmp IN PLACE -- every old
//            value has been consumed by then, so nothing is copied back at the loop edge.
template <int PPL, int MODE, int JM, bool LAST>
__device__ __forceinline__ void warp_stage(WarpState<PPL, MODE>& S, const unsigned wsa, const double2* __restrict__ tab, const int lane,
                                           const int tm, const int jm, const double dg, const double dz, unsigned& mBs,
                                           unsigned& mBu, unsigned& mAs, unsigned& mAu) {
  auto W3 = [&](int v, int i) -> double {   // window entry i of variable v (compile-time i after unrolling)
    return i < 3 ? S.hl[v][i < 3 ? i : 0] : (i < PPL + 3 ? S.f[v][(i >= 3 && i < PPL + 3) ? i - 3 : 0] : S.hr[v][i >= PPL + 3 ? i - PPL - 3 : 0]);
  };
  auto load_w = [&](int j, double* w) {
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
  };
  constexpr int NG = (PPL + 1) / 2;   // groups of two points
  double wq[2][2][8];                 // [buffer][point of the group][weight]
  load_w(0, wq[0][0]);
  if (PPL > 1) load_w(1, wq[0][1]);
  // ---- phase 0: caa = ca*alp and em = alp*(Br^2+Bp^2) + qc*sqrt|alp|*Br*Bp
  double caa[PPL], em[PPL], c_gs[PPL], c_fk[PPL], c_p3[PPL], c_c0[PPL];
  {
    double al[PPL], B2[PPL], bb[PPL], g[PPL], h[PPL], c_qc[PPL];
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      double c_ca, c_ak;
      if (MODE == WM_ALLG) {
        const double2 t4 = ldg_pair(tab + GSLOT2(j, 4)), t5 = ldg_pair(tab + GSLOT2(j, 5)), t6 = ldg_pair(tab + GSLOT2(j, 6));
        const double2 t7 = ldg_pair(tab + GSLOT2(j, 7));
        c_ca = t4.x; c_gs[j] = t4.y; c_c0[j] = t5.x; c_fk[j] = t5.y; c_ak = t6.x; c_p3[j] = t6.y; c_qc[j] = t7.x;
      } else {
        constexpr int jj = MODE == WM_ALLG ? 0 : 1;
        c_ca = S.ca[j * jj]; c_gs[j] = S.gs[j * jj]; c_c0[j] = S.c0[j * jj]; c_fk[j] = S.fk[j * jj]; c_ak = S.ak[j * jj];
        if (MODE == WM_WSM2) {
          const double2 t = lds_pair(wsa + 16u * WSLOT2(j, 4));
          c_p3[j] = t.x; c_qc[j] = t.y;
        } else {
          c_p3[j] = S.p3[MODE >= WM_WSM2 ? 0 : j]; c_qc[j] = S.qc[MODE >= WM_WSM2 ? 0 : j];
        }
      }
      al[j] = c_ak + S.f[2][j];
      caa[j] = c_ca * al[j];
    }
#pragma unroll
    for (int j = 0; j < PPL; j++) {   // sqrt|alp|: MUFU.RSQ64H seed + one Goldschmidt step (fast_sqrt_abs), interleaved
      int hi = __double2hiint(al[j]) & 0x7fffffff;
      hi = max(hi, 0x00200000);
      const double x = __hiloint2double(hi, __double2loint(al[j]));
      double r;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
      g[j] = x * r; h[j] = 0.5 * r;
    }
#pragma unroll
    for (int j = 0; j < PPL; j++) { B2[j] = S.f[1][j] * S.f[1][j]; bb[j] = S.f[0][j] * S.f[1][j]; }
#pragma unroll
    for (int j = 0; j < PPL; j++) { h[j] = fma(-g[j], h[j], 0.5); B2[j] = fma(S.f[0][j], S.f[0][j], B2[j]); }
#pragma unroll
    for (int j = 0; j < PPL; j++) { g[j] = fma(g[j], h[j], g[j]); B2[j] = al[j] * B2[j]; }
#pragma unroll
    for (int j = 0; j < PPL; j++) g[j] = c_qc[j] * g[j];
#pragma unroll
    for (int j = 0; j < PPL; j++) em[j] = fma(g[j], bb[j], B2[j]);
  }
  if (PPL <= WARP_LATE_GHOSTS_MAXPPL) warp_ghosts<PPL, MODE, JM>(S, lane, tm, jm);
  // ---- phase 1: stencil chains L(Br), L(Bp) + floor constant, alp_m chain
  double acc[3][PPL];
#pragma unroll
  for (int gi = 0; gi < NG; gi++) {
    const int ja = 2 * gi, jb = 2 * gi + 1;
    const bool hasb = jb < PPL;
    const int cur = gi & 1;
    if (gi + 1 < NG) {   // coefficient pairs of the next group
      load_w(2 * gi + 2, wq[cur ^ 1][0]);
      if (2 * gi + 3 < PPL) load_w(2 * gi + 3, wq[cur ^ 1][1]);
    }
    const double* wa = wq[cur][0];
    const double* wb = wq[cur][1];
    double a0 = wa[3] * S.f[0][ja], a1 = fma(wa[3], S.f[1][ja], c_c0[ja]), a2 = wa[7] * S.f[2][ja];
    double b0 = 0, b1 = 0, b2 = 0;
    if (hasb) { b0 = wb[3] * S.f[0][hasb ? jb : 0]; b1 = fma(wb[3], S.f[1][hasb ? jb : 0], c_c0[hasb ? jb : 0]); b2 = wb[7] * S.f[2][hasb ? jb : 0]; }
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {   // pass 0: neighbours in own registers, pass 1: halo entries
#pragma unroll
      for (int k = 0; k < 7; k++) {
        if (k == 3) continue;
        const int ia = ja + k, ib = jb + k;
        if ((ia >= 3 && ia < PPL + 3) == (pass == 0)) {
          a0 = fma(wa[k], W3(0, ia), a0); a1 = fma(wa[k], W3(1, ia), a1); a2 = fma(wa[k], W3(2, ia), a2);
        }
        if (hasb && (ib >= 3 && ib < PPL + 3) == (pass == 0)) {
          b0 = fma(wb[k], W3(0, ib), b0); b1 = fma(wb[k], W3(1, ib), b1); b2 = fma(wb[k], W3(2, ib), b2);
        }
      }
    }
    acc[0][ja] = a0; acc[1][ja] = a1; acc[2][ja] = a2;
    if (hasb) { acc[0][hasb ? jb : 0] = b0; acc[1][hasb ? jb : 0] = b1; acc[2][hasb ? jb : 0] = b2; }
  }
  // ---- phase 2: right-hand sides and the Runge-Kutta update, in place
  double p0[PPL], p1[PPL], p2[PPL];
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    p0[j] = fma(-caa[j], S.f[1][j], acc[0][j]);
    p1[j] = fma(c_gs[j], S.f[0][j], acc[1][j]);
    p2[j] = fma(c_p3[j], em[j], acc[2][j]);
  }
#pragma unroll
  for (int j = 0; j < PPL; j++) p1[j] = fma(c_fk[j], S.f[2][j], p1[j]);
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    S.f[0][j] = fma(dg, p0[j], S.ft[0][j]); S.f[1][j] = fma(dg, p1[j], S.ft[1][j]); S.f[2][j] = fma(dg, p2[j], S.ft[2][j]);
  }
  // the new state is final: the halo exchange for the next stage starts now; its 36 shuffles
  // (one issue slot every ~4 cycles) are interleaved with the ftmp update and the maxima
#pragma unroll
  for (int j = 0; j < PPL; j++) {
#pragma unroll
    for (int v = 0; v < 3; v++) {
      if (j < 3) S.hr[v][j] = shfl_down1(S.f[v][j]);
      if (j >= PPL - 3) S.hl[v][j - (PPL - 3)] = shfl_up1(S.f[v][j]);
      if (LAST) S.ft[v][j] = S.f[v][j];
      else S.ft[v][j] = fma(dz, v == 0 ? p0[j] : (v == 1 ? p1[j] : p2[j]), S.f[v][j]);
    }
    // |x| maxima on the high words: signed max catches the positive values, unsigned max the negative ones
    mBs = (unsigned)__vimax3_s32((int)mBs, __double2hiint(S.f[0][j]), __double2hiint(S.f[1][j]));
    mBu = __vimax3_u32(mBu, (unsigned)__double2hiint(S.f[0][j]), (unsigned)__double2hiint(S.f[1][j]));
    mAs = (unsigned)max((int)mAs, __double2hiint(S.f[2][j]));
    mAu = max(mAu, (unsigned)__double2hiint(S.f[2][j]));
  }
}

// alp = alp_k + alp_m as the last pde call of the snapshot sees it (an output row); called once
// per snapshot, kept out of line so that the time loop only carries one predicate for it
template <int PPL>
struct AmRow { double v[PPL]; };
template <int PPL>
__device__ __noinline__ void save_alp_row(const double* __restrict__ akp, double* __restrict__ alp, const int g0, const int m, const AmRow<PPL> am) {
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = g0 + j;
    if (g >= MAG_NGHOST && g <= m) alp[g] = akp[g] + am.v[j];
  }
}

// Fast time loop of one snapshot attempt on one warp.  Returns ok (false: check_f_error).
template <int PPL, int MODE, int JM>
__device__ __noinline__ bool warp_loop(Gal& G, RkShared* rs) {
  const int lane = G.lane;
  double* const W = G.W;
  const int nxcap = G.a->nxcap, nx = G.nx, m = nx - MAG_NGHOST - 1, nsteps = G.nsteps;
  const int P = nx - 2 * MAG_NGHOST - 1;
  const int tm = (P + WARP_J0) / PPL, jm = (P + WARP_J0) % PPL;
  const double dt = G.dt;
  auto WA = [&](int id) { return W + (size_t)id * nxcap; };
  const int g0 = lane * PPL - WARP_J0 + MAG_NGHOST;   // grid index of register 0 (slot s holds p = s - WARP_J0, g = p + 3)
  double2* const wt2 = reinterpret_cast<double2*>(rs->pool) + lane;                          // shared pair table, this lane's column
  double2* const tab = reinterpret_cast<double2*>(W + (size_t)WARP_TABLE_ARRAY * nxcap) + lane;   // global pair table
  // 32-bit shared addresses of this lane's column of the pair table and of the stage constants.  They pass through an
  // empty asm so that the compiler keeps them in a register: it would otherwise rebuild the shared window base in every
  // stage (S2UR SR_CgaCtaId + ULEA: a 20-cycle special-register read in front of the first loads, 3 % of the stall samples).
  unsigned wsa = (unsigned)__cvta_generic_to_shared(rs->pool) + 16u * lane;
  unsigned sdg = (unsigned)__cvta_generic_to_shared(&rs->stage_dg[0]);
  if (MAG_WARP_OPAQUE_ADDR) asm volatile("" : "+r"(wsa), "+r"(sdg));

  PHASE_BEGIN;
  prepare_fast_coefs(G, nx);
  PHASE_END(3, lane);
  guard_constants(G, rs->gk);
  PHASE_END(4, lane);
  fold_boundary_weights(G);
  PHASE_END(5, lane);

  WarpState<PPL, MODE> S;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int g = g0 + j;
    const bool live = g >= MAG_NGHOST && g <= m;
    const int gc = live ? g : 0;
    double w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = live ? WA(C_W0 + k)[gc] : 0.0;   // C_W0..C_W6, C_W0A are consecutive
    const bool bfree = live && g != MAG_NGHOST && g != m;   // Br, Bp evolve here (not a Dirichlet point)
    const double ca = bfree ? WA(F_CA)[gc] : 0.0, gs = bfree ? WA(F_GS)[gc] : 0.0, ak = live ? WA(F_AK)[gc] : 0.0;
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
      const double fb = (live && g != MAG_NGHOST && g != m) ? sg[gc] * cFB[gc] : 0.0;
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
      for (int v = 0; v < 3; v++) S.f[v][j] = (live && (v == 2 || (g != MAG_NGHOST && g != m))) ? WA(A_F0 + v)[gc] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < PPL; j++)
#pragma unroll
      for (int v = 0; v < 3; v++) S.ft[v][j] = S.f[v][j];
    warp_exchange<PPL, MODE>(S);
    if (PPL > WARP_LATE_GHOSTS_MAXPPL) warp_ghosts<PPL, MODE, JM>(S, lane, tm, jm);
  };
  auto save_alp = [&]() {
    AmRow<PPL> r;
#pragma unroll
    for (int j = 0; j < PPL; j++) r.v[j] = S.f[2][j];
    save_alp_row<PPL>(WA(F_AK), WA(A_ALP), g0, m, r);
  };
  load_signed_floor();
  load_state();
  __syncwarp();
  PHASE_END(6, lane);

  const double dg0 = dt * (8.0 / 15), dg1 = dt * (5.0 / 12), dg2 = dt * (3.0 / 4);
  const double dz0 = dt * (-17.0 / 60), dz1 = dt * (-5.0 / 12);
  if (lane == 0) {
    rs->stage_dg[0] = dg0; rs->stage_dg[1] = dg1; rs->stage_dg[2] = dg2;
    rs->stage_dz[0] = dz0; rs->stage_dz[1] = dz1; rs->stage_dz[2] = 0.0;
  }
  __syncwarp();
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
    // The step at which the floor signs are redrawn next (this_t - t_last > tmk) is found ahead of the steps, `nb`
    // independent evaluations of the step's own expression, so that the steps do not start with that serial chain
    // (I2F, DMUL, DADD, DADD, DSETP, BRA: 2.4 % of the stall samples).
    auto first_due = [&](const int s0) {
      int r = nb;
#pragma unroll 4
      for (int s = nb - 1; s >= s0; s--)
        if ((t_Gyr + dtg * (double)(jt + s)) - t_last > tmk) r = s;
      return r;
    };
    int s_evt = !MAG_WARP_DUE_AHEAD ? -1 : (refresh || sign_nx != nx) ? 0 : first_due(0);
    for (int s = 0; s < nb; s++) {
      // same order of RNG draws as the reference: update_floor_sign (dynamo.f90:196-198) may
      // draw Bfloor_sign, then the first pde of the step redraws the layer signs if flagged
      // or if the grid size changed (floor_field.f90:55-79)
      const double this_t = t_Gyr + dtg * (double)(jt + s);
      const bool due = this_t - t_last > tmk;
      if (MAG_WARP_DUE_AHEAD ? s == s_evt : (due || refresh || sign_nx != nx)) {
#ifdef MAG_PHASE_CLOCKS
        const long long _r0 = clock64();
#endif
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
        if (MAG_WARP_DUE_AHEAD) s_evt = first_due(s + 1);
#ifdef MAG_PHASE_CLOCKS
        if (lane == 0) { atomicAdd(&g_phase_clk[8], (unsigned long long)(clock64() - _r0)); atomicAdd(&g_phase_clk[9], 1ull); }
#endif
      }
      const bool last_step = jt + s == nsteps;
      if ((JM >= 0 && MAG_WARP_UNROLL_JM) || PPL <= MAG_WARP_UNROLL_MAXPPL) {  // three stages unrolled: their shuffles and chains overlap
        warp_stage<PPL, MODE, JM, false>(S, wsa, tab, lane, tm, jm, dg0, dz0, mBs, mBu, mAs, mAu);
        warp_stage<PPL, MODE, JM, false>(S, wsa, tab, lane, tm, jm, dg1, dz1, mBs, mBu, mAs, mAu);
        if (last_step) save_alp();
        warp_stage<PPL, MODE, JM, true>(S, wsa, tab, lane, tm, jm, dg2, 0.0, mBs, mBu, mAs, mAu);
      } else {         // one stage body, looped: a three times smaller instruction footprint
        // save_alp() runs before the third stage of the snapshot's last step: the stage loop stops there once, instead of
        // testing for it (a taken branch around the call) in every stage
        int st = 0, stop = (MAG_WARP_STAGE_LOOP2 && last_step) ? 2 : 3;
        for (;;) {
#pragma unroll 1
          for (; st < stop; st++) {
            double dg, dz;
            if (MAG_WARP_OPAQUE_ADDR) { dg = lds_f64x(sdg + 8u * (unsigned)st); dz = lds_f64x(sdg + 24u + 8u * (unsigned)st); }   // stage_dz follows stage_dg
            else { dg = rs->stage_dg[st]; dz = rs->stage_dz[st]; }
            if (!MAG_WARP_STAGE_LOOP2 && st == 2 && last_step) save_alp();
            warp_stage<PPL, MODE, JM, false>(S, wsa, tab, lane, tm, jm, dg, dz, mBs, mBu, mAs, mAu);
            if (PPL > WARP_LATE_GHOSTS_MAXPPL) warp_ghosts<PPL, MODE, JM>(S, lane, tm, jm);
            // the time update of the step (below), two terms per stage, inside the stage so that its serial chain of five
            // additions hides behind the stage's arithmetic; the third stage adds stage_dz[2] = +0.0 to a positive t: exact.
            if (PPL <= WARP_T_IN_STAGE_MAXPPL) { t = t + dg; t = t + dz; }
          }
          if (st == 3) break;
          save_alp();
          stop = 3;
        }
        if (PPL <= WARP_T_IN_STAGE_MAXPPL) continue;
      }
      // t = t + dt*gam1 ; ttmp = t + dt*zet1 ; t = ttmp + dt*gam2 ; ttmp = t + dt*zet2 ; t = ttmp + dt*gam3
      double tt = t + dg0;
      tt = tt + dz0; tt = tt + dg1; tt = tt + dz1; tt = tt + dg2;
      t = tt;
    }
    // ---- accept or replay the block.  Per-lane maxima of lanes 0,1 / tm-1,tm bound the 7 points
    // next to the inner / outer boundary (a superset of them: conservative).
    const unsigned mB = max(mBs, mBu & 0x7fffffffu), mA = max(mAs, mAu & 0x7fffffffu);
    const unsigned nB0 = max(__shfl_sync(FULL, mB, 0), __shfl_sync(FULL, mB, 1));
    const unsigned nA0 = max(__shfl_sync(FULL, mA, 0), __shfl_sync(FULL, mA, 1));
    const unsigned nB1 = max(__shfl_sync(FULL, mB, tm), __shfl_sync(FULL, mB, tm - 1));
    const unsigned nA1 = max(__shfl_sync(FULL, mA, tm), __shfl_sync(FULL, mA, tm - 1));
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
  PHASE_END(7, lane);
#ifdef MAG_PHASE_CLOCKS
  if (lane == 0) { atomicAdd(&g_phase_clk[10], 1ull); atomicAdd(&g_phase_clk[11], (unsigned long long)nsteps); }
#endif
  return ok;
}

// Registers per lane for a grid of nx points: the two outer alp_m ghost entries must still
// fall into a lane (slots up to P+4 < 32*PPL); 0 if the warp path does not take this grid.
#ifndef MAG_WARP_CAND_SET
#define MAG_WARP_CAND_SET 0
#endif
#if MAG_WARP_CAND_SET == 0
#define MAG_WARP_CANDS {4, 5, 6, 8, 11}
#elif MAG_WARP_CAND_SET == 1
#define MAG_WARP_CANDS {4, 6, 8, 11}
#elif MAG_WARP_CAND_SET == 2
#define MAG_WARP_CANDS {4, 5, 8, 11}
#elif MAG_WARP_CAND_SET == 3
#define MAG_WARP_CANDS {4, 8, 11}
#else
#define MAG_WARP_CANDS {4, 6, 11}
#endif
__device__ __forceinline__ int warp_ppl(int nx) {
  const int P = nx - 2 * MAG_NGHOST - 1;
  const int cand[] = MAG_WARP_CANDS;
  const int ncand = (int)(sizeof(cand) / sizeof(cand[0]));
  for (int c = 0; c < ncand; c++) {
    // MAG_WARP_ALIGNED4 = 2: the four-register loop only takes grids that fill their last lane (one body, no alignment switch)
    if (MAG_WARP_ALIGNED4 == 2 && cand[c] == 4 && (P + WARP_J0) % 4 != 3) continue;
    if (P + WARP_J0 + 3 < 32 * cand[c]) return cand[c];
  }
  return 0;
}
// (R_kappa = 1, the reference's default, is folded into the alp_m stencil chain; other values
// take the shared-line path of mag_rk_fast.cuh)
__device__ __forceinline__ bool warp_path_applies(const Gal& G) {
  const int ppl = warp_ppl(G.nx);
  if (ppl >= 8 && (size_t)WARP_TABLE_ARRAYS * G.a->nxcap < (size_t)16 * 32 * ppl) return false;   // pair table does not fit
  return G.a->P.R_kappa == 1.0 && G.nx >= 64 && ppl != 0;
}
__device__ __forceinline__ bool warp_dispatch(Gal& G, RkShared* rs) {
  // the default grid (p_nx_ref = 101: nx = 107, most of the work) has its own instantiation with
  // a compile-time outer alignment and the three stages unrolled
  if (MAG_WARP_SPECIALIZE_DEFAULT && G.nx == 107) return warp_loop<4, MAG_W4_MODE, 3>(G, rs);
  // the outer Dirichlet point in the last register of its lane (the default grid among others): no alignment switch in the stage
  if (MAG_WARP_ALIGNED4 && warp_ppl(G.nx) == 4 && (G.nx - 2 * MAG_NGHOST - 1 + WARP_J0) % 4 == 3) return warp_loop<4, MAG_W4_MODE, 3>(G, rs);
  switch (warp_ppl(G.nx)) {
    case 4: if (MAG_WARP_ALIGNED4 == 2) return false; return warp_loop<4, MAG_W4_MODE, -1>(G, rs);
    case 5: return warp_loop<5, WM_WSM, -1>(G, rs);
    case 6: return warp_loop<6, WM_WSM2, -1>(G, rs);
    case 8: return warp_loop<8, WM_ALLG, -1>(G, rs);
    default: return warp_loop<11, WM_ALLG, -1>(G, rs);
  }
}

}  // namespace magdev
