// mag_phases.cuh -- dynamo_run (source/dynamo.f90:39-300) split into two device phases.
//
// Everything the reference does between two snapshots that does NOT depend on the magnetic
// field -- reading the inputs, the grid recurrence (grid.f90), construct_profiles
// (profiles.f90) with its root finders, set_timestep, the status codes they raise -- is a
// pure function of the galaxy's input history.  Phase A (run_profiles, one warp per
// galaxy) evaluates it for the whole history and leaves one snapshot record per output
// redshift in HBM: the control flags of that snapshot, dt and nsteps, and the per-point
// arrays the field evolution needs (hoisted RHS coefficients, x, h, Beq, r_kpc).  It also
// writes the output rows that depend on the ISM profiles only.  Because nsteps is then
// known for every snapshot, the exact cost of every galaxy is known before Phase B starts
// and the B queue is ordered longest-first.
//
// Phase B (run_evolve, one warp per galaxy; one 4-warp team in the MAG_TEAM = 128 build) owns the field f = (Br, Bp, alp_m):
// seeds (seed_field.f90), the field part of adjust_grid, the RK3 time loop with retries
// (dynamo.f90:155-243), the floor-field RNG stream, and the output rows that depend on f.
// The "profile refresh stays inside the same solve": both phases run inside one
// mag_solve_batch call, back to back on the same stream, exchanging data through HBM only.
#pragma once
#include "mag_galaxy.cuh"

namespace magdev {

// snapshot record flags
#define SR_VALID 1       // the snapshot is part of the history (init_it..last processed)
#define SR_TRAP 2        // negligible-disk trap: f = 0 (dynamo.f90:106-110)
#define SR_RESEED 4      // leaving an elliptical phase: profiles + seed field on the OLD grid (:113-118)
#define SR_SEED0 8       // seed field before the loop (:90-96), first snapshot only
#define SR_ADJUSTED 16   // adjust_grid + impose_bc + construct_profiles happened (:126-140)
#define SR_SCALEBACK 32  // ... and f is resampled back to the default resolution
#define SR_EXTEND 64     // ... and f is extended
#define SR_SOLVE 128     // the time loop runs (not test_run, not elliptical)
#define SR_FAIL 256      // construct_profiles failed: finish() and stop (:134-137 / :83-87)
#define SR_LAST 512      // last_output after this snapshot

enum RecArray { R_X = 0, R_H, R_BEQ, R_RKPC, R_IR, R_C1, R_CA, R_CR, R_KDC, R_FB, R_P3, R_QC, R_W, R_G, R_ALPK, R_COUNT };
__device__ __constant__ const int REC_WS[R_COUNT] = {A_X, A_H, A_BEQ, A_RKPC, C_IR, C_C1, C_CA, C_CR, C_KDC, C_FB, C_P3, C_QC,
                                                    C_W, A_G, A_ALPK};

struct SnapRec {       // dense [ngal][nz]
  int32_t flags, nx, nx_pre, nx_mid, nsteps, status;
  uint32_t gflags;     // MAG_FLAG_* raised by the profile phase up to and including this snapshot
  int32_t pad;
  long long off;       // arena offset (doubles) of R_COUNT arrays, each `nx` long
  long long off_pre;   // arena offset of Beq and x on the pre-adjust grid, 2 x [nx_pre] (SR_RESEED only)
  double dt, t_Gyr, lambda, delta_r_floor, tau_min, v_code;
  double seed_r1;      // p_r_seed_decay / r_max_kpc of the grid a seed field of this snapshot is drawn on ('decaying' seed)
};
struct GalHead {       // [chunk]
  int32_t init_it, max_it, error;
  uint32_t flags;
  long long cost;      // scheduling cost: sum over SR_SOLVE snapshots of nsteps*nx*(relative time per point of the one-warp loop)
  int32_t nx_max;      // largest grid of an SR_SOLVE snapshot
  int32_t pad;
};

struct ChainJob {      // one hydrostatic chain = one construct_profiles call of one galaxy
  // arena offset of CHAIN_JOB_ARRAYS arrays of nx doubles: |r|, Sigma_d, Sigma_*, Rm, Om_h, G_h, Om_b, G_b (chain
  // inputs), h, rho (chain outputs), Om_d, Om_b, Om_h before regularisation (rotation curves)
  long long off;
  int32_t nx, bits;    // bits: CHAIN_* result of k_chains
  uint32_t rot_flags;  // MAG_FLAG_* raised by the rotation curves
  int32_t pad;
  double r_disk, Mstars_bulge;
  double r_max_kpc, dr_kpc;   // with nx and r_disk: identifies the grid the job was exported on
};

__device__ inline long long arena_alloc(const KernelArgs* a, int lane, long long n) {
  long long off = -1;
  if (lane == 0) {
    off = (long long)atomicAdd(a->arena_top, (unsigned long long)n);
    if (off + n > a->arena_cap) { atomicExch(a->fail, MAG_ERR_ARENA); off = -1; }
  }
  return __shfl_sync(FULL, off, 0);
}

// ---- hydrostatic chains as jobs.  Pass 1 of the profile phase (k_profiles, CHAIN_MODE_EXPORT)
// walks the grid recurrence of every galaxy assuming that each chain succeeds and exports the
// inputs of every chain; k_chains then solves 32 chains per warp (one thread each) instead of
// one chain on lane 0 of a warp; pass 2 (CHAIN_MODE_IMPORT) repeats the walk with the chain
// results and finishes the snapshots.  If a chain fails in a way that changes the walk
// (the pre-adjust solve of a snapshot that leaves an elliptical phase), pass 2 notices that
// the job it needs does not exist or has another grid size and solves inline from there on.
#define CHAIN_JOB_ARRAYS 13
__device__ void chain_export(Gal& G, const double* absr, uint32_t rot_flags) {
  const KernelArgs* a = G.a;
  const int nx = G.nx, lane = G.lane;
  const long long off = arena_alloc(a, lane, (long long)CHAIN_JOB_ARRAYS * nx);
  int j = -1;
  if (off >= 0 && lane == 0) {
    j = atomicAdd(a->job_count, 1);
    if (j >= a->job_cap) { atomicExch(a->fail, MAG_ERR_ARENA); j = -1; }
  }
  j = __shfl_sync(FULL, j, 0);
  if (j < 0) return;
  const int src[CHAIN_JOB_ARRAYS] = {-1, A_SIGD, A_SIGS, A_RM, A_OMH, A_GH, A_OMB, A_GB, -2, -2, A_OMD, A_T1, A_T2};
  for (int k = 0; k < CHAIN_JOB_ARRAYS; k++) {
    if (src[k] == -2) continue;
    const double* s = (k == 0) ? absr : G.A(src[k]);
    double* d = a->arena + off + (long long)k * nx;
    for (int i = lane; i < nx; i += 32) d[i] = s[i];
  }
  if (lane == 0) {
    ChainJob J;
    J.off = off; J.nx = nx; J.bits = 0; J.r_disk = G.r_disk; J.Mstars_bulge = G.Mstars_bulge;
    J.r_max_kpc = G.r_max_kpc; J.dr_kpc = G.dr_kpc; J.rot_flags = rot_flags; J.pad = 0;
    a->jobs[j] = J;
    a->jobidx[((size_t)(G.g - a->g0) * a->nz + (G.chain_it - 1)) * 2 + G.chain_kind] = j;
  }
  __syncwarp();
}

__device__ int chain_lookup(Gal& G) {
  const KernelArgs* a = G.a;
  if (G.chain_diverged) return -1;
  const int j = a->jobidx[((size_t)(G.g - a->g0) * a->nz + (G.chain_it - 1)) * 2 + G.chain_kind];
  if (j < 0) { G.chain_diverged = true; return -1; }
  const ChainJob J = a->jobs[j];
  if (J.nx != G.nx || J.r_disk != G.r_disk || J.r_max_kpc != G.r_max_kpc || J.dr_kpc != G.dr_kpc) {
    G.chain_diverged = true;   // pass 1 walked another grid (it assumed a chain success that did not happen)
    return -1;
  }
  return j;
}
__device__ const double* chain_job_array(const Gal& G, int k) {
  const ChainJob& J = G.a->jobs[G.chain_job];
  return G.a->arena + J.off + (long long)k * J.nx;
}
__device__ uint32_t chain_job_flags(const Gal& G) { return G.a->jobs[G.chain_job].rot_flags; }

__device__ int chain_import(Gal& G) {
  const KernelArgs* a = G.a;
  if (G.chain_job < 0) return -1;
  const ChainJob J = a->jobs[G.chain_job];
  const double *h = a->arena + J.off + 8LL * J.nx, *rho = h + J.nx;
  double *hd = G.A(A_HKPC), *rd = G.A(A_RHO);
  for (int i = G.lane; i < J.nx; i += 32) { hd[i] = h[i]; rd[i] = rho[i]; }
  __syncwarp();
  return J.bits;
}

// one chain job, one thread (k_chains)
__device__ inline void run_chain_job(const KernelArgs& a, int j) {
  ChainJob& J = a.jobs[j];
  const int nx = J.nx;
  double* base = a.arena + J.off;
  ChainIn c;
  c.n = nx; c.r_disk = J.r_disk; c.Mstars_bulge = J.Mstars_bulge;
  c.r = base; c.Sigma_d = base + nx; c.Sigma_star = base + 2LL * nx; c.Rm = base + 3LL * nx;
  c.Om_h = base + 4LL * nx; c.G_h = base + 5LL * nx; c.Om_b = base + 6LL * nx; c.G_b = base + 7LL * nx;
  c.h_d = base + 8LL * nx; c.rho_d = base + 9LL * nx;
  J.bits = hydro_chain(a.P, c);
}

// writes the record of snapshot `it` from the warp's workspace
__device__ inline bool write_record(Gal& G, int it, int flags, int nx_pre, int nx_mid, long long off_pre, double seed_r1 = 0.0) {
  const KernelArgs* a = G.a;
  const int nx = G.nx;
  const long long off = arena_alloc(a, G.lane, (long long)R_COUNT * nx);
  if (off < 0) return false;
  for (int r = 0; r < R_COUNT; r++) {
    const double* src = G.A(REC_WS[r]);
    double* dst = a->arena + off + (long long)r * nx;
    for (int i = G.lane; i < nx; i += 32) dst[i] = src[i];
  }
  if (G.lane == 0) {
    SnapRec& R = a->recs[(size_t)(G.g - a->g0) * a->nz + it - 1];   // record tables are chunk-local
    R.flags = flags | SR_VALID; R.nx = nx; R.nx_pre = nx_pre; R.nx_mid = nx_mid; R.nsteps = G.nsteps;
    R.status = (int)G.status; R.off = off; R.off_pre = off_pre; R.gflags = G.flags; R.pad = 0;
    R.dt = G.dt; R.t_Gyr = G.t_Gyr; R.lambda = G.lambda; R.delta_r_floor = G.delta_r_floor;
    R.tau_min = G.tau_min; R.v_code = G.v_code; R.seed_r1 = seed_r1;
  }
  __syncwarp();
  return true;
}

// ------------------------------------------------------------------ Phase A
// Returns true when the reference would return error=.true. (nothing written).
__device__ static bool run_profiles(Gal& G, GalHead& H) {
  const mag_params_t& P = G.a->P;
  const bool test_run = P.p_no_magnetic_fields_test_run != 0;
  const int lane = G.lane;
  bool elliptical = false;
  const bool exporting = G.a->chain_mode == CHAIN_MODE_EXPORT;   // pass 1: no outputs, no records
  G.chain_mode = G.a->chain_mode; G.chain_diverged = false; G.chain_kind = 0;
  G.status = test_run ? 't' : 'M';
  G.flags = 0; G.point_stages = 0;
  G.last_output = false;
  G.t = 0; G.dt = 0; G.nsteps = 20; G.time_between_inputs = 0; G.minh = 1e10; G.maxetat = 0; G.tau_min = 0;
  G.t_last_sign_choice = 0; G.refresh_sign = true; G.sign_nx = -1;
  H.cost = 0; H.flags = 0; H.error = 0; H.nx_max = 0; H.pad = 0;
  G.w0 = G.a->nz + 1; G.w1 = 0;   // no output row written yet (k_profiles fills the rest afterwards)
  load_input_parameters(G);
  H.init_it = G.init_it; H.max_it = G.max_it;
  if (set_input_params(G)) return true;
  construct_grid(G);
  {
    const int ids[] = {A_F0, A_F1, A_F2, A_ALP, A_BZ, A_BEQ, A_TAU, A_LKPC, A_L, A_N, A_H, A_ETAT, A_ALPK, A_UZ,
                       C_IR, C_C1, C_CA, C_CR, C_KDC, C_FB, C_P3, C_QC, C_W, A_G};
    for (int k = 0; k < (int)(sizeof(ids) / sizeof(int)); k++) {
      double* p = G.A(ids[k]);
      for (int i = lane; i < G.nx; i += 32) p[i] = 0.0;
    }
    __syncwarp();
  }
  G.chain_it = G.init_it;
  bool able = construct_profiles(G);
  int it = G.init_it;
  if (!able) {  // finish()
    const bool okw = write_record(G, it, SR_FAIL, G.nx, G.nx, -1);
    write_profile_rows(G, it);
    H.flags = G.flags;
    return !okw;
  }
  if (G.Mgas_disk < P.Mgas_disk_min) G.flags |= MAG_FLAG_STALE_PROFILES;
  const bool seed0 = (G.r_disk > P.p_rdisk_min && G.Mgas_disk > P.Mgas_disk_min);
  for (it = G.init_it; it <= G.max_it; it++) {
    int fl = (it == G.init_it && seed0) ? SR_SEED0 : 0;
    const int nx_pre = G.nx;
    int nx_mid = G.nx;
    long long off_pre = -1;
    const double seed_r1 = P.p_r_seed_decay / G.r_max_kpc;   // seeds are drawn before adjust_grid (dynamo.f90:113-131)
    if (G.r_disk < P.p_rdisk_min || G.Mgas_disk < P.Mgas_disk_min) {
      elliptical = true;
      fl |= SR_TRAP;
    } else {
      if (elliptical) {
        G.chain_it = it; G.chain_kind = 1;
        able = construct_profiles(G);
        G.chain_kind = 0;
        if (!able) G.chain_diverged = true;   // pass 1 assumed success: its later jobs belong to another walk
        if (!exporting) {
          off_pre = arena_alloc(G.a, lane, 2 * (long long)nx_pre);
          if (off_pre < 0) return true;
          const double *beq = G.A(A_BEQ), *xg = G.A(A_X);
          for (int i = lane; i < nx_pre; i += 32) { G.a->arena[off_pre + i] = beq[i]; G.a->arena[off_pre + nx_pre + i] = xg[i]; }
          __syncwarp();
        }
        fl |= SR_RESEED;
      }
      if (able) elliptical = false;
    }
    if (it != G.init_it && !elliptical) {
      bool scaled, extended;
      if (!adjust_grid_geometry(G, &nx_mid, &scaled, &extended)) {
        if (lane == 0) atomicExch(G.a->fail, MAG_ERR_NX_OVERFLOW);
        return true;
      }
      fl |= SR_ADJUSTED | (scaled ? SR_SCALEBACK : 0) | (extended ? SR_EXTEND : 0);
      G.chain_it = it;
      able = construct_profiles(G);
      if (!able) {
        if (G.last_output) fl |= SR_LAST;
        const bool okw = write_record(G, it, fl | SR_FAIL, nx_pre, nx_mid, off_pre, seed_r1);
        write_profile_rows(G, it);
        H.flags = G.flags;
        return !okw;
      }
    }
    set_timestep(G, true);
    if (!test_run && !elliptical) {
      fl |= SR_SOLVE;
      // time of this snapshot on a one-warp team, in units of (steps x points) of the default grid: the
      // large-grid variants of the warp loop stream their coefficients and are slower per point
      // (measured, profiles/r1_bench.md: 177 / 117 / 77 G point-stages/s for nx <= 192 / 256 / 352)
      const int w16 = G.nx <= 192 ? 16 : (G.nx <= 256 ? 24 : (G.nx <= 352 ? 37 : 64));
      H.cost += (long long)G.nsteps * G.nx * w16 / 16;
      H.nx_max = max(H.nx_max, G.nx);
    }
    if (G.last_output) fl |= SR_LAST;
    if (!exporting) {
      if (!write_record(G, it, fl, nx_pre, nx_mid, off_pre, seed_r1)) return true;
      write_profile_rows(G, it);
    }
    if (G.last_output) break;
    G.status = test_run ? 't' : 'M';
    if (set_input_params(G)) break;
  }
  H.flags = G.flags;
  return false;
}

// ------------------------------------------------------------------ Phase B (leader warp)
template <int TEAM> struct RkSharedT;
template <int TEAM> __device__ bool run_timesteps(Gal& G, RkSharedT<TEAM>* rs);   // mag_rk_fast.cuh

// loads the arrays of a record into the team's workspace
__device__ inline void load_record(Gal& G, const SnapRec& R) {
  const int nx = R.nx;
  for (int r = 0; r < R_COUNT; r++) {
    const double* src = G.a->arena + R.off + (long long)r * nx;
    double* dst = G.A(REC_WS[r]);
    for (int i = G.lane; i < nx; i += 32) dst[i] = src[i];
  }
  __syncwarp();
}

// marks the rows Phase A wrote for snapshots after `it` as never written (the run stopped
// early with status 'i', dynamo.f90:245-257)
__device__ inline void invalidate_later_rows(Gal& G, int it) {
  const KernelArgs* a = G.a;
  const int nr = a->P.p_nx_ref, nz = a->nz;
  if (it >= nz) return;
  // (the host re-copies the rows of the profile phase when they were staged and copied out
  // while this kernel was running, mag_solver.cu)
  if (G.lane == 0 && a->n_invalidated) atomicAdd(a->n_invalidated, 1);
  for (int qq = 0; qq < a->n_profile_q; qq++) {
    if (is_field_quantity(a->profile_q[qq]) || !a->prof[qq]) continue;
    double* dst = a->prof[qq] + ((size_t)G.g * nz + it) * nr;
    for (int i = G.lane; i < (nz - it) * nr; i += 32) dst[i] = MAG_INVALID;
  }
  __syncwarp();
}

template <int TEAM>
__device__ static void run_evolve(Gal& G, const GalHead& H, RkSharedT<TEAM>* rs) {
  const KernelArgs* a = G.a;
  const mag_params_t& P = a->P;
  const bool test_run = P.p_no_magnetic_fields_test_run != 0;
  const int lane = G.lane;
  if (lane == 0) rng_seed(G.rng, a->gal_ids[G.g], P.p_random_seed, P.rng_kind);
  __syncwarp();
  // initialize_floor_sign (floor_field.f90:116-123)
  G.t_last_sign_choice = 0.0;
  (void)random_sign(G);
  G.refresh_sign = true;
  G.sign_nx = -1;
  // the diagnostic flags of the profile phase arrive with the snapshot records: a run that stops early
  // (status 'i') must not report what the reference would only have met in later snapshots
  G.flags = 0; G.point_stages = 0;
  G.t = 0; G.nsteps = 20;
  G.w0 = a->nz + 1; G.w1 = 0;
  G.init_it = H.init_it; G.max_it = H.max_it;
  const SnapRec* recs = a->recs + (size_t)(G.g - a->g0) * a->nz;
  G.nx = recs[H.init_it - 1].nx;
  for (int v = 0; v < 3; v++) { double* f = G.A(A_F0 + v); for (int i = lane; i < G.nx; i += 32) f[i] = 0.0; }
  { double *alp = G.A(A_ALP), *bz = G.A(A_BZ); for (int i = lane; i < G.nx; i += 32) { alp[i] = 0.0; bz[i] = 0.0; } }
  __syncwarp();
  for (int it = H.init_it; it <= H.max_it; it++) {
    const SnapRec R = recs[it - 1];
    if (!(R.flags & SR_VALID)) break;
    G.flags |= R.gflags;
    bool ok = true;
    PHASE_BEGIN;
    if (R.flags & SR_TRAP) {
      for (int v = 0; v < 3; v++) { double* f = G.A(A_F0 + v); for (int i = lane; i < G.nx; i += 32) f[i] = 0.0; }
      __syncwarp();
    } else if (R.flags & SR_RESEED) {
      double *beq = G.A(A_BEQ), *xg = G.A(A_X);
      for (int i = lane; i < R.nx_pre; i += 32) { beq[i] = a->arena[R.off_pre + i]; xg[i] = a->arena[R.off_pre + R.nx_pre + i]; }
      __syncwarp();
      G.seed_r1 = R.seed_r1;
      init_seed_field(G);
      impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    }
    if (R.flags & SR_SCALEBACK) { adjust_f_scaleback(G, R.nx_pre, R.nx_mid); G.nx = R.nx_mid; }
    load_record(G, R);
    if (R.flags & SR_EXTEND) adjust_f_extend(G, R.nx_mid, R.nx);
    G.nx = R.nx;
    G.lambda = R.lambda; G.delta_r_floor = R.delta_r_floor; G.tau_min = R.tau_min; G.v_code = R.v_code;
    G.dt = R.dt; G.nsteps = R.nsteps; G.t_Gyr = R.t_Gyr;
    G.status = (char)R.status;
    if (R.flags & SR_ADJUSTED) impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    if (!(R.flags & SR_LAST)) G.t = 0;  // set_input_params resets t except at the last snapshot (input_parameters.f90:224-233)
    if (R.flags & SR_FAIL) {  // finish() is the only writer of this snapshot
      estimate_Bzmod(G);
      write_field_rows(G, it);
      break;
    }
    if (R.flags & SR_SEED0) {
      G.seed_r1 = R.seed_r1;
      init_seed_field(G);
      impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    }
    {  // f_snapshot_beginning = f
      for (int v = 0; v < 3; v++) {
        const double* f = G.A(A_F0 + v); double* fb = G.A(A_FB0 + v);
        for (int i = lane; i < G.nx; i += 32) fb[i] = f[i];
      }
      __syncwarp();
    }
    PHASE_END(0, lane);
    for (int fail_count = 0; fail_count <= P.p_MAX_FAILS; fail_count++) {
      ok = true;
      if (!(R.flags & SR_SOLVE)) {
        double *alp = G.A(A_ALP), *bz = G.A(A_BZ);
        for (int i = lane; i < G.nx; i += 32) { alp[i] = 0.0; bz[i] = 0.0; }
        __syncwarp();
        break;
      }
      ok = run_timesteps<TEAM>(G, rs);
      if (ok) break;
      G.status = 'm';
      G.flags |= MAG_FLAG_RETRY;
      G.nsteps = P.nsteps_0;  // set_timestep(reduce_timestep): dt is NOT recomputed
      if (G.nsteps > P.p_nsteps_max) G.nsteps = P.p_nsteps_max;
      for (int v = 0; v < 3; v++) {
        double* f = G.A(A_F0 + v); const double* fb = G.A(A_FB0 + v);
        for (int i = lane; i < G.nx; i += 32) f[i] = fb[i];
      }
      __syncwarp();
    }
    PHASE_END(1, lane);
    bool last = (R.flags & SR_LAST) != 0;
    if (!ok) {
      G.status = 'i';
      last = true;
      init_seed_field(G);
      invalidate_later_rows(G, it);
    }
    impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    estimate_Bzmod(G);
    write_field_rows(G, it);
    PHASE_END(2, lane);
    if (last) break;
  }
  fill_unwritten(G, true);
  if (lane == 0) {
    if (a->o_nsteps) a->o_nsteps[G.g] = G.nsteps;
    if (a->o_point_stages) a->o_point_stages[G.g] = G.point_stages;
    if (a->o_flags) a->o_flags[G.g] = G.flags;
  }
  __syncwarp();
}

}  // namespace magdev
