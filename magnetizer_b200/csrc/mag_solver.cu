// mag_solver.cu -- persistent team-per-galaxy kernels and the single-GPU C ABI
// (include/magnetizer_b200.h).
//
// Replaces, for the per-galaxy dynamo path only, the reference's MPI master/worker farm
// (source/distributor.f90:253-384) + work routine dynamo_run (source/dynamo.f90:39): the
// farm becomes device-side work queues (atomic counters) drained by persistent teams, each
// of which runs a whole galaxy history.  No collective, no host round trip between
// snapshots; inputs are staged into HBM once per batch and read with coalesced loads
// (snapshot index fastest), outputs are written as contiguous 101-point rows.
//
// One solve = k_profiles (+ k_chains + k_profiles) -> cost sort -> k_evolve_heavy || k_evolve:
// the galaxies whose history is longer than a warp slot's fair share of the batch (the tail
// of a longest-first schedule) run on four-warp teams at a 4-5 times lower latency per
// Runge-Kutta stage, everything else on one warp per galaxy (highest throughput).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mag_galaxy.cuh"
#include "mag_phases.cuh"
#include "mag_rk_fast.cuh"

using namespace magdev;

#define WARPS_PER_CTA 4
#ifndef MAG_PROFILE_MAXREG
#define MAG_PROFILE_MAXREG 128   // 16 warps per SM: the profile phase is latency bound (serial root-finder chains)
#endif
#ifndef MAG_EVOLVE_MAXREG
#define MAG_EVOLVE_MAXREG 255    // 8 one-warp teams per SM, everything of a galaxy in registers
#endif
#ifndef MAG_HEAVY_MAXREG
#define MAG_HEAVY_MAXREG 168     // 3 four-warp teams per SM
#endif

static thread_local std::string g_last_error;
static void set_error(const std::string& s) { g_last_error = s; }
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
      return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? MAG_ERR_NO_DEVICE : MAG_ERR_CUDA; \
    }                                                                                          \
  } while (0)

// ------------------------------------------------------------------ Phase A kernel
// One warp per galaxy, persistent, any order (every history costs about the same here).
__global__ void __maxnreg__(MAG_PROFILE_MAXREG) k_profiles(KernelArgs a) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS_PER_CTA + wib;
  Gal G;
  G.a = &a;
  G.W = a.ws + (size_t)gw * A_COUNT * a.nxcap;
  G.rng = nullptr;
  G.lane = lane;
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(a.counter, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= a.g1 - a.g0) break;
    G.g = a.g0 + q;
    GalHead H;
    const bool err = run_profiles(G, H);
    H.error = err ? 1 : 0;
    if (err) H.cost = 0;
    if (a.chain_mode != CHAIN_MODE_EXPORT) {
      // rows nobody will write: everything for a galaxy the evolve phase skips (dynamo.f90:69-73),
      // else the ISM-profile rows outside the snapshots processed above
      if (err) reset_outputs(G); else fill_unwritten(G, false);
    }
    if (lane == 0 && a.chain_mode != CHAIN_MODE_EXPORT) {
      a.heads[q] = H;
      if (H.cost > 0) atomicAdd(a.cost_total, (unsigned long long)H.cost);
      if (a.o_error) a.o_error[G.g] = H.error;
      if (err) {  // nothing else is written for this galaxy (dynamo.f90:69-73)
        if (a.o_nsteps) a.o_nsteps[G.g] = G.nsteps;
        if (a.o_point_stages) a.o_point_stages[G.g] = 0;
        if (a.o_flags) a.o_flags[G.g] = G.flags;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ hydrostatic chains
// One thread per chain job exported by pass 1 of the profile phase: 32 serial root-finder
// chains per warp instead of one on lane 0.
__global__ void __launch_bounds__(128) k_chains(KernelArgs a) {
  const int n = *a.job_count < a.job_cap ? *a.job_count : a.job_cap;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) run_chain_job(a, j);
}

// ------------------------------------------------------------------ Phase B kernels
// Light class: one warp per galaxy, persistent, longest history first; the warp does
// everything, the time loop in registers with shuffles (mag_rk_warp.cuh).  The queue of this
// class starts behind the heavy entries of `order`.
__global__ void __maxnreg__(MAG_EVOLVE_MAXREG) k_evolve(KernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RkSharedT<MAG_TEAM_LIGHT>* rs = reinterpret_cast<RkSharedT<MAG_TEAM_LIGHT>*>(smem_raw);
  const int lane = threadIdx.x & 31;
  Gal G;
  G.a = &a;
  G.W = a.ws2 + (size_t)blockIdx.x * A_COUNT * a.nxcap;
  G.rng = &rs->rng;
  G.lane = lane;
  const int nh = *a.n_heavy, n = a.g1 - a.g0;
  for (;;) {
    int q = 0;
    if (lane == 0) q = nh + atomicAdd(a.counter2, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= n) break;
    const int qi = a.order[q];
    const GalHead H = a.heads[qi];
    if (H.error) continue;
    G.g = a.g0 + qi;
    run_evolve<MAG_TEAM_LIGHT>(G, H, rs);
  }
  // launched as the programmatic dependent of k_evolve_heavy (which it does not depend on): the
  // stream must not run ahead of the heavy teams, so one CTA joins them before the grid retires
  if (a.pdl_join && blockIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Heavy class: four warps per galaxy.  Warp 0 leads (seeds, grid transformations of f, outputs,
// generic path), the helper warps join for the shared-line time loop (mag_rk_fast.cuh).
__global__ void __maxnreg__(MAG_HEAVY_MAXREG) k_evolve_heavy(KernelArgs a) {
  asm volatile("griddepcontrol.launch_dependents;");   // the light kernel may start filling the free SMs at once
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RkSharedT<MAG_TEAM_HEAVY>* rs = reinterpret_cast<RkSharedT<MAG_TEAM_HEAVY>*>(smem_raw);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) {  // helper warps
    for (;;) {
      __syncthreads();
      if (rs->call.cmd == TEAM_CMD_EXIT) break;
      fast_dispatch<MAG_TEAM_HEAVY>(nullptr, rs);
    }
    return;
  }
  Gal G;
  G.a = &a;
  G.W = a.ws3 + (size_t)blockIdx.x * A_COUNT * a.nxcap;
  G.rng = &rs->rng;
  G.lane = lane;
  const int nh = *a.n_heavy;
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(a.counter3, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= nh) break;
    const int qi = a.order[q];
    const GalHead H = a.heads[qi];
    if (H.error) continue;
    G.g = a.g0 + qi;
    run_evolve<MAG_TEAM_HEAVY>(G, H, rs);
  }
  if (lane == 0) rs->call.cmd = TEAM_CMD_EXIT;
  __syncthreads();
}

// ------------------------------------------------------------------ cost-ordered queues
// The cost of every history (sum of nsteps*nx, weighted by the speed of the loop variant its grids
// take) is exact after Phase A.  `order` = heavy class first, then the
// light class, each longest-first (counting sort on 1/16-octave buckets of the cost).  A
// galaxy is heavy when its history alone would keep one warp busy for longer than `alpha`
// times the fair share of a warp slot (sum of all costs / number of slots): under a
// longest-first schedule those are the galaxies that end up as the tail.  Grids the one-warp
// loop does not take (nx_max > nx_heavy) are heavy as well.
#define COST_BUCKETS 1024
struct QueuePolicy {
  int mode;        // 0: everything light, 1: by cost, 2: everything heavy
  int nx_heavy;
  double alpha_over_slots;
};
__device__ __forceinline__ int cost_bucket(long long c) {
  if (c <= 0) return 0;
  const int b = (int)(log2f((float)c) * 16.0f) + 1;
  return b >= COST_BUCKETS ? COST_BUCKETS - 1 : b;
}
__device__ __forceinline__ int queue_slot(const GalHead& h, const QueuePolicy& p, unsigned long long total) {
  bool heavy = false;
  if (!h.error && h.cost > 0) {
    if (p.mode == 2) heavy = true;
    else if (p.mode == 1) heavy = (double)h.cost > p.alpha_over_slots * (double)total || h.nx_max > p.nx_heavy;
  }
  return (heavy ? 0 : COST_BUCKETS) + (COST_BUCKETS - 1 - cost_bucket(h.cost));   // ascending slot = descending cost
}
__global__ void k_cost_hist(const GalHead* heads, int n, int* hist, QueuePolicy p, const unsigned long long* total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(&hist[queue_slot(heads[i], p, *total)], 1);
}
__global__ void k_cost_scan(int* hist, int* n_heavy) {  // exclusive scan, in place
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < 2 * COST_BUCKETS; k++) {
      if (k == COST_BUCKETS) *n_heavy = acc;
      int c = hist[k]; hist[k] = acc; acc += c;
    }
  }
}
__global__ void k_cost_scatter(const GalHead* heads, int n, int* hist, int* order, QueuePolicy p, const unsigned long long* total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  order[atomicAdd(&hist[queue_slot(heads[i], p, *total)], 1)] = i;
}

// FP64 peak micro-benchmark: independent DFMA chains, 2 flops each
__global__ void k_dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------ host side
// device-side counters of one context (one int each)
enum Counter { CNT_A = 0, CNT_FAIL, CNT_REPLAYS, CNT_LIGHT, CNT_HEAVY, CNT_NHEAVY, CNT_INVALIDATED, CNT_NHEAVY_SUM, CNT_COUNT };

// one solve of the host-buffer API, kept until mag_wait (and for a retry with a larger arena)
struct PendingCall {
  bool active = false;
  int32_t ngal_total = 0, g0 = 0, n = 0, nz = 0;
  const int32_t* gal_ids = nullptr;
  const double* t_gyr = nullptr;
  const double* inputs = nullptr;
  int32_t n_profile_q = 0, n_scalar_q = 0;
  int32_t profile_q[MAG_NPROFILE], scalar_q[MAG_NSCALAR];
  mag_outputs_t out;
  // planes of the profile phase that were staged in HBM and copied out while k_evolve was running
  int n_early = 0;
  struct { double* dev; double* host; size_t bytes; } early[MAG_NPROFILE];
};

struct mag_ctx {
  mag_params_t P;
  int device = 0;
  int sm_count = 0;
  int ctasA_per_sm = 0, ctasB_per_sm = 0, ctasH_per_sm = 0;
  int nwarpsA = 0, nteamsB = 0, nteamsH = 0;
  int nxcap = 0;
  double* wsA = nullptr;      // Phase A: per-warp workspaces
  double* wsB = nullptr;      // Phase B: per-team workspaces, one-warp teams
  double* wsH = nullptr;      // Phase B: four-warp teams
  int* d_counter = nullptr;   // [CNT_COUNT]
  unsigned long long* d_arena_top = nullptr;   // [0] arena top, [1] cost total of the chunk
  int* d_hist = nullptr;
  int* d_order = nullptr;
  ChainJob* d_jobs = nullptr;
  int* d_jobidx = nullptr;
  int* d_job_count = nullptr;
  int split_chains = 1;       // 1: two profile passes + k_chains, 0: chains inline on lane 0
  SnapRec* d_recs = nullptr;
  GalHead* d_heads = nullptr;
  int chunk_cap = 0;          // galaxies the record tables are sized for
  int nz_cap = 0;
  double* d_arena = nullptr;
  long long arena_cap = 0;    // doubles
  int nx_budget_pct = 165;    // arena budget per snapshot in percent of the default grid
  // staging for the host-buffer API
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  int* h_status = nullptr;    // pinned: fail flag, n_invalidated, replays, heavy galaxies of the last call
  cudaStream_t stream = nullptr, stream_h = nullptr, stream_copy = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_prof = nullptr, ev_copy = nullptr;
  double last_ms[4] = {0, 0, 0, 0};
  int last_launches = 0;
  int use_fast = 1;
  int launch_pdl = 1;         // 1: light kernel as programmatic dependent of the heavy one, 0: two streams
  int stage_profile_rows = 1; // rows of the profile phase: HBM + one copy per plane during k_evolve (0: zero-copy stores)
  QueuePolicy policy;
  double heavy_alpha = 1.5;
  PendingCall pend;
};

// virtual base of an array whose first row belongs to catalogue galaxy g_first (see KernelArgs)
template <class T>
static T* vbase(T* real, long long g_first, size_t row_elems) {
  return real ? real - (ptrdiff_t)(g_first * (long long)row_elems) : nullptr;
}

extern "C" {

const char* mag_last_error(void) { return g_last_error.c_str(); }

void mag_params_default(mag_params_t* p) {
  // source/global_input_parameters.f90 defaults; literals without d0 are single precision
  // in the reference and are widened from float here
  memset(p, 0, sizeof(*p));
  p->abi_version = MAG_ABI_VERSION;
  p->nsteps_0 = 1000; p->p_nsteps_max = 20000; p->p_MAX_FAILS = 6; p->p_random_seed = 17;
  p->p_no_magnetic_fields_test_run = 0;
  p->p_courant_v = (double)0.09f; p->p_courant_eta = (double)0.09f;
  p->p_nx_ref = 101; p->p_nx_MAX = 10000; p->p_rmax_over_rdisk = 2.75;
  p->frac_seed = 0.01; p->C_alp = 1.0; p->C_floor = 1.0; p->p_floor_kappa = 10.0; p->fmag = 0.5; p->R_kappa = 1.0;
  p->p_ISM_sound_speed_km_s = 10.0; p->p_ISM_kappa = 1.0; p->p_ISM_xi = 0.25; p->p_ISM_gamma = 5.0 / 3.0;
  p->p_ISM_turbulent_length = (double)0.1f;
  p->p_stellarHeightToRadiusScale = 1.0 / (double)7.3f;
  p->p_molecularHeightToRadiusScale = (double)0.032f;
  p->p_gasScaleRadiusToStellarScaleRadius_ratio = 1.0;
  p->p_Rmol_alpha = 0.92; p->p_Rmol_P0 = 4.787e-12;
  p->p_rreg_to_rdisk = (double)0.1f;
  p->p_minimum_density = 1e-30; p->p_rdisk_min = 0.5; p->Mgas_disk_min = 1e4;
  p->rmin_over_rmax = (double)0.001f;
  p->p_ignore_satellite_DM_haloes = 0; p->p_use_Pdm = 1; p->p_use_Pbulge = 1;
  p->rng_kind = MAG_RNG_XORSHIFT1024S;
  p->p_variable_timesteps = 1; p->p_simplified_pressure = 1; p->p_halo_contraction = 1; p->p_enable_P2 = 0;
  p->Dyn_quench = 1; p->Alg_quench = 0; p->Damp = 0; p->lFloor = 1;
  p->p_space_varying_floor = 1; p->p_time_varying_floor = 1; p->p_limit_turbulent_scale = 1;
  p->p_allow_positive_shears = 0; p->seed_choice = MAG_SEED_RANDOM; p->outflow_type = MAG_OUTFLOW_NONE;
  p->p_neumann_boundary_condition_rmax = 0; p->p_use_fixed_physical_grid = 0; p->p_nn_seed = 2; p->p_r_seed_decay = 15.0;
  // outflow_parameters (global_input_parameters.f90:303-325)
  p->p_outflow_Lsn = 1e38; p->p_outflow_fOB = (double)0.7f; p->p_outflow_etaSN = 9.4e-3; p->p_tOB = 3.0; p->p_N_SN1OB = 40.0;
  p->p_outflow_hot_gas_density = 1.7e-27; p->p_outflow_nu0 = 0.5; p->p_outflow_alphahot = (double)-3.2f; p->p_outflow_Vhot = 425.0;
}

static bool params_supported(const mag_params_t& p) {
  return p.abi_version == MAG_ABI_VERSION && p.p_variable_timesteps == 1 && p.p_simplified_pressure == 1 &&
         p.p_halo_contraction == 1 && p.p_enable_P2 == 0 && (p.Dyn_quench == 0 || p.Dyn_quench == 1) && (p.Alg_quench == 0 || p.Alg_quench == 1) && p.Damp == 0 &&
         p.lFloor == 1 && p.p_space_varying_floor == 1 && p.p_time_varying_floor == 1 &&
         p.p_limit_turbulent_scale == 1 && p.p_allow_positive_shears == 0 && p.seed_choice >= MAG_SEED_RANDOM && p.seed_choice <= MAG_SEED_DECAYING &&
         (p.p_neumann_boundary_condition_rmax == 0 || p.p_neumann_boundary_condition_rmax == 1) &&
         (p.p_use_fixed_physical_grid == 0 || p.p_use_fixed_physical_grid == 1) &&
         p.outflow_type >= MAG_OUTFLOW_NONE && p.outflow_type <= MAG_OUTFLOW_SUPERBUBBLE && p.p_nx_ref >= 8 &&
         (p.rng_kind == 0 || p.rng_kind == 1) && p.p_nx_MAX >= p.p_nx_ref + 6;
}

void mag_destroy(mag_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->wsA); cudaFree(c->wsB); cudaFree(c->wsH); cudaFree(c->d_counter); cudaFree(c->d_arena_top); cudaFree(c->d_hist);
  cudaFree(c->d_order); cudaFree(c->d_recs); cudaFree(c->d_heads); cudaFree(c->d_arena);
  cudaFree(c->d_jobs); cudaFree(c->d_jobidx); cudaFree(c->d_job_count);
  cudaFree(c->d_stage);
  if (c->h_status) cudaFreeHost(c->h_status);
  for (int i = 0; i < 5; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (cudaEvent_t e : {c->ev_fork, c->ev_join, c->ev_prof, c->ev_copy}) if (e) cudaEventDestroy(e);
  for (cudaStream_t s : {c->stream, c->stream_h, c->stream_copy}) if (s) cudaStreamDestroy(s);
  delete c;
}

static int create_impl(mag_ctx* c, const cudaDeviceProp& prop) {
  const mag_params_t* params = &c->P;
  c->sm_count = prop.multiProcessorCount;
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return (e && *e) ? atoi(e) : dflt; };
  c->split_chains = env_int("MAG_INLINE_CHAINS", 0) ? 0 : 1;
  c->use_fast = env_int("MAG_DISABLE_FAST_PATH", 0) ? 0 : 1;
  c->launch_pdl = env_int("MAG_EVOLVE_PDL", 1);
  c->stage_profile_rows = env_int("MAG_STAGE_PROFILE_ROWS", 1);
  c->policy.mode = env_int("MAG_HEAVY_MODE", 1);
  c->policy.nx_heavy = env_int("MAG_HEAVY_NX", 352);
  c->nx_budget_pct = env_int("MAG_ARENA_NX_PCT", 165);
  if (const char* e = getenv("MAG_HEAVY_ALPHA")) c->heavy_alpha = atof(e);
  const size_t smemB = sizeof(RkSharedT<MAG_TEAM_LIGHT>), smemH = sizeof(RkSharedT<MAG_TEAM_HEAVY>);
  CUDA_TRY(cudaFuncSetAttribute(k_evolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
  CUDA_TRY(cudaFuncSetAttribute(k_evolve_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemH));
  int occA = 0, occB = 0, occH = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, k_profiles, WARPS_PER_CTA * 32, 0));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, k_evolve, MAG_TEAM_LIGHT, smemB));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occH, k_evolve_heavy, MAG_TEAM_HEAVY, smemH));
  if (occA < 1 || occB < 1 || occH < 1) { set_error("kernel does not fit on an SM"); return MAG_ERR_CUDA; }
  c->ctasA_per_sm = occA; c->ctasB_per_sm = occB; c->ctasH_per_sm = occH;
  c->nwarpsA = c->sm_count * occA * WARPS_PER_CTA;
  c->nteamsB = c->sm_count * occB;
  c->nteamsH = c->sm_count * occH;
  c->policy.alpha_over_slots = c->heavy_alpha / (double)c->nteamsB;
  int cap = env_int("MAG_NXCAP", 2048);
  if (cap > params->p_nx_MAX) cap = params->p_nx_MAX;
  if (cap < params->p_nx_ref + 6) cap = params->p_nx_ref + 6;
  if (cap < 1024) cap = 1024;   // the team loops address up to 8 points per thread; room for the pair table of the large-grid warp variants
  c->nxcap = (cap + 15) & ~15;
  const size_t per_team = (size_t)A_COUNT * c->nxcap * sizeof(double);
  CUDA_TRY(cudaMalloc(&c->wsA, c->nwarpsA * per_team));
  CUDA_TRY(cudaMemset(c->wsA, 0, c->nwarpsA * per_team));
  CUDA_TRY(cudaMalloc(&c->wsB, c->nteamsB * per_team));
  CUDA_TRY(cudaMemset(c->wsB, 0, c->nteamsB * per_team));
  CUDA_TRY(cudaMalloc(&c->wsH, c->nteamsH * per_team));
  CUDA_TRY(cudaMemset(c->wsH, 0, c->nteamsH * per_team));
  CUDA_TRY(cudaMalloc(&c->d_counter, CNT_COUNT * sizeof(int)));
  CUDA_TRY(cudaMemset(c->d_counter, 0, CNT_COUNT * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_arena_top, 2 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&c->d_hist, 2 * COST_BUCKETS * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_job_count, sizeof(int)));
  CUDA_TRY(cudaHostAlloc(&c->h_status, 8 * sizeof(int), cudaHostAllocDefault));
  memset(c->h_status, 0, 8 * sizeof(int));
  int prio_lo = 0, prio_hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithPriority(&c->stream_h, cudaStreamNonBlocking, prio_hi));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking));
  for (int i = 0; i < 5; i++) CUDA_TRY(cudaEventCreate(&c->ev[i]));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_prof, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
  return MAG_OK;
}

int mag_create(const mag_params_t* params, int device, mag_ctx_t** ctx_out) {
  if (!params || !ctx_out) { set_error("null argument"); return MAG_ERR_BAD_ARGUMENT; }
  *ctx_out = nullptr;
  if (!params_supported(*params)) { set_error("parameter combination outside the implemented path"); return MAG_ERR_UNSUPPORTED; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { set_error("no CUDA device (this library has no CPU fallback)"); return MAG_ERR_NO_DEVICE; }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= ndev) { set_error("device index out of range"); return MAG_ERR_BAD_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { set_error("kernels are built for sm_100a only"); return MAG_ERR_NO_DEVICE; }
  mag_ctx* c = new mag_ctx();
  c->P = *params;
  c->device = device;
  const int rc = create_impl(c, prop);
  if (rc != MAG_OK) { const std::string msg = g_last_error; mag_destroy(c); set_error(msg); return rc; }   // nothing leaks on a failed create
  *ctx_out = c;
  return MAG_OK;
}

static int check_quantities(int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q, const int32_t* scalar_q) {
  if (n_profile_q < 0 || n_profile_q > MAG_NPROFILE || n_scalar_q < 0 || n_scalar_q > MAG_NSCALAR) return MAG_ERR_BAD_ARGUMENT;
  for (int i = 0; i < n_profile_q; i++) if (profile_q[i] < 0 || profile_q[i] >= MAG_NPROFILE) return MAG_ERR_BAD_ARGUMENT;
  for (int i = 0; i < n_scalar_q; i++) if (scalar_q[i] < 0 || scalar_q[i] >= MAG_NSCALAR) return MAG_ERR_BAD_ARGUMENT;
  return MAG_OK;
}

// Record storage: SnapRec/GalHead tables for `chunk` galaxies and an arena for their arrays.
// The arena budget per snapshot is ARENA_ARRAYS arrays of nx_budget points (1.65 x the default
// grid; grid extensions reach 3 x nx in the example input but only for ~1 % of the snapshots).
// A range that does not fit is solved in several chunks; a chunk that overflows the arena
// raises MAG_ERR_ARENA on the device and mag_wait repeats the solve with twice the budget.
#define ARENA_ARRAYS (R_COUNT + CHAIN_JOB_ARRAYS + 2)   // record arrays + the arrays of a chain job (+ Beq and x of reseeded snapshots)
static long long arena_need(const mag_ctx* c, int chunk, int nz) {
  const long long nx_budget = ((long long)(c->P.p_nx_ref + 2 * MAG_NGHOST) * c->nx_budget_pct + 99) / 100;
  return (long long)chunk * nz * ARENA_ARRAYS * nx_budget;
}

static void free_records(mag_ctx* c) {
  cudaFree(c->d_order); cudaFree(c->d_recs); cudaFree(c->d_heads); cudaFree(c->d_arena); cudaFree(c->d_jobs); cudaFree(c->d_jobidx);
  c->d_order = nullptr; c->d_recs = nullptr; c->d_heads = nullptr; c->d_arena = nullptr; c->d_jobs = nullptr; c->d_jobidx = nullptr;
  c->chunk_cap = 0; c->arena_cap = 0;
}

static int ensure_records(mag_ctx* c, int ngal, int nz, int* chunk_out) {
  if (c->chunk_cap > 0 && nz <= c->nz_cap && (ngal <= c->chunk_cap)) { *chunk_out = ngal; return MAG_OK; }
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  free_b += (size_t)c->arena_cap * sizeof(double) +
            (size_t)c->chunk_cap * ((size_t)c->nz_cap * (sizeof(SnapRec) + 2 * sizeof(ChainJob) + 2 * sizeof(int)) + sizeof(GalHead) + sizeof(int));
  const char* env = getenv("MAG_ARENA_FRACTION");
  const double frac = env ? atof(env) : 0.5;
  const size_t per_gal = (size_t)nz * (sizeof(SnapRec) + 2 * sizeof(ChainJob) + 2 * sizeof(int)) + sizeof(GalHead) + sizeof(int) +
                         (size_t)arena_need(c, 1, nz) * sizeof(double);
  long long chunk = (long long)((double)free_b * frac / (double)per_gal);
  if (const char* e = getenv("MAG_MAX_CHUNK")) { const long long m = atoll(e); if (m > 0 && chunk > m) chunk = m; }
  if (chunk > ngal) chunk = ngal;
  if (chunk < 1) { set_error("not enough device memory for the snapshot records"); return MAG_ERR_CUDA; }
  if (chunk <= c->chunk_cap && nz <= c->nz_cap) { *chunk_out = (int)chunk; return MAG_OK; }
  free_records(c);
  CUDA_TRY(cudaMalloc(&c->d_jobs, (size_t)chunk * nz * 2 * sizeof(ChainJob)));
  CUDA_TRY(cudaMalloc(&c->d_jobidx, (size_t)chunk * nz * 2 * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_order, (size_t)chunk * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_recs, (size_t)chunk * nz * sizeof(SnapRec)));
  CUDA_TRY(cudaMalloc(&c->d_heads, (size_t)chunk * sizeof(GalHead)));
  CUDA_TRY(cudaMalloc(&c->d_arena, (size_t)arena_need(c, (int)chunk, nz) * sizeof(double)));
  c->chunk_cap = (int)chunk; c->nz_cap = nz; c->arena_cap = arena_need(c, (int)chunk, nz);
  *chunk_out = (int)chunk;
  return MAG_OK;
}

// Enqueues the kernels for catalogue galaxies [g0, g0+n) on `st`.  `a` arrives with the array
// bindings (virtual base pointers), nz, t_gyr and the quantity lists filled in.
static int launch_solve(mag_ctx* c, KernelArgs& a, int32_t g0, int32_t n, cudaStream_t st, bool early_copy_event) {
  const int nz = a.nz;
  if (nz < 1 || nz > 4000) { set_error("nz out of range"); return MAG_ERR_BAD_ARGUMENT; }
  int chunk = 0;
  int rc = ensure_records(c, n, nz, &chunk);
  if (rc) return rc;
  int launches = 0;
  CUDA_TRY(cudaMemsetAsync(c->d_counter + CNT_FAIL, 0, 2 * sizeof(int), st));  // fail flag, replays
  CUDA_TRY(cudaMemsetAsync(c->d_counter + CNT_INVALIDATED, 0, 2 * sizeof(int), st));  // ... and the heavy-galaxy count
  a.P = c->P;
  a.U = make_units();
  a.order = c->d_order;
  a.ws = c->wsA; a.ws2 = c->wsB; a.ws3 = c->wsH; a.nxcap = c->nxcap;
  a.counter = c->d_counter + CNT_A; a.fail = c->d_counter + CNT_FAIL; a.replays = c->d_counter + CNT_REPLAYS;
  a.counter2 = c->d_counter + CNT_LIGHT; a.counter3 = c->d_counter + CNT_HEAVY; a.n_heavy = c->d_counter + CNT_NHEAVY;
  a.n_invalidated = c->d_counter + CNT_INVALIDATED;
  a.cost_total = c->d_arena_top + 1;
  // the register loops hard-wire alp = alp_k + alp_m; algebraic quenching (gutsdynamo.f90:202-203) needs Bzmod in the
  // RHS at every step and runs on the any-nx generic path
  // ... and so do the Neumann outer boundary (the register loops fold the Dirichlet one into their stencil weights) and the
  // runs without the alp_m equation
  a.use_fast = (c->P.Alg_quench || c->P.p_neumann_boundary_condition_rmax || !c->P.Dyn_quench) ? 0 : c->use_fast;
  a.recs = c->d_recs; a.heads = c->d_heads; a.arena = c->d_arena; a.arena_top = c->d_arena_top; a.arena_cap = c->arena_cap;
  a.jobs = c->d_jobs; a.job_count = c->d_job_count; a.jobidx = c->d_jobidx;
  a.pdl_join = c->launch_pdl;
  for (int s0 = 0; s0 < n; s0 += chunk) {
    const int m = (n - s0 < chunk) ? n - s0 : chunk;
    a.g0 = g0 + s0; a.g1 = g0 + s0 + m;
    CUDA_TRY(cudaMemsetAsync(c->d_counter + CNT_A, 0, sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(c->d_counter + CNT_LIGHT, 0, 3 * sizeof(int), st));   // light head, heavy head, n_heavy
    CUDA_TRY(cudaMemsetAsync(c->d_arena_top, 0, 2 * sizeof(unsigned long long), st));
    CUDA_TRY(cudaMemsetAsync(c->d_hist, 0, 2 * COST_BUCKETS * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(c->d_recs, 0, (size_t)m * nz * sizeof(SnapRec), st));
    int ctasA = c->sm_count * c->ctasA_per_sm;
    const int needA = (m + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (needA < ctasA) ctasA = needA;
    a.job_cap = m * nz * 2;
    if (c->split_chains) {
      // pass 1: walk the grid recurrence, export the hydrostatic chains; then 32 chains per warp
      CUDA_TRY(cudaMemsetAsync(c->d_job_count, 0, sizeof(int), st));
      CUDA_TRY(cudaMemsetAsync(c->d_jobidx, 0xFF, (size_t)m * nz * 2 * sizeof(int), st));
      a.chain_mode = CHAIN_MODE_EXPORT;
      k_profiles<<<ctasA, WARPS_PER_CTA * 32, 0, st>>>(a);
      k_chains<<<c->sm_count * 8, 128, 0, st>>>(a);
      CUDA_TRY(cudaMemsetAsync(c->d_counter + CNT_A, 0, sizeof(int), st));
      a.chain_mode = CHAIN_MODE_IMPORT;
      launches += 2;
    } else {
      a.chain_mode = CHAIN_MODE_INLINE;
    }
    k_profiles<<<ctasA, WARPS_PER_CTA * 32, 0, st>>>(a);
    if (early_copy_event && s0 + m >= n) CUDA_TRY(cudaEventRecord(c->ev_prof, st));   // the rows of the profile phase are complete
    const int tb = 256, gb = (m + tb - 1) / tb;
    k_cost_hist<<<gb, tb, 0, st>>>(c->d_heads, m, c->d_hist, c->policy, a.cost_total);
    k_cost_scan<<<1, 32, 0, st>>>(c->d_hist, c->d_counter + CNT_NHEAVY);
    k_cost_scatter<<<gb, tb, 0, st>>>(c->d_heads, m, c->d_hist, c->d_order, c->policy, a.cost_total);
    if (s0 == 0) CUDA_TRY(cudaEventRecord(c->ev[4], st));
    launches += 4;
    int teams = c->nteamsB, teamsH = c->nteamsH;
    if (m < teams) teams = m;
    if (m < teamsH) teamsH = m;
    const bool heavy_possible = c->policy.mode != 0;
    if (heavy_possible && c->launch_pdl) {
      // one stream: heavy teams first; the light kernel is their programmatic dependent and
      // starts as soon as every heavy CTA is resident (it never reads what they write)
      k_evolve_heavy<<<teamsH, MAG_TEAM_HEAVY, sizeof(RkSharedT<MAG_TEAM_HEAVY>), st>>>(a);
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(teams); cfg.blockDim = dim3(MAG_TEAM_LIGHT);
      cfg.dynamicSmemBytes = sizeof(RkSharedT<MAG_TEAM_LIGHT>); cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CUDA_TRY(cudaLaunchKernelEx(&cfg, k_evolve, a));
      launches += 2;
    } else if (heavy_possible) {
      // two streams: the heavy teams on a high-priority side stream, forked and joined by events
      a.pdl_join = 0;
      CUDA_TRY(cudaEventRecord(c->ev_fork, st));
      CUDA_TRY(cudaStreamWaitEvent(c->stream_h, c->ev_fork, 0));
      k_evolve_heavy<<<teamsH, MAG_TEAM_HEAVY, sizeof(RkSharedT<MAG_TEAM_HEAVY>), c->stream_h>>>(a);
      CUDA_TRY(cudaEventRecord(c->ev_join, c->stream_h));
      k_evolve<<<teams, MAG_TEAM_LIGHT, sizeof(RkSharedT<MAG_TEAM_LIGHT>), st>>>(a);
      CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
      launches += 2;
    } else {
      a.pdl_join = 0;
      k_evolve<<<teams, MAG_TEAM_LIGHT, sizeof(RkSharedT<MAG_TEAM_LIGHT>), st>>>(a);
      launches += 1;
    }
    // diagnostics: heavy galaxies of the call (summed over chunks on the device would need a kernel; keep the last chunk's)
    CUDA_TRY(cudaMemcpyAsync(c->d_counter + CNT_NHEAVY_SUM, c->d_counter + CNT_NHEAVY, sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaGetLastError());
  }
  c->last_launches = launches;
  return MAG_OK;
}

int mag_solve_batch_device(mag_ctx_t* c, int32_t ngal, const int32_t* d_gal_ids, int32_t nz, const double* d_t_gyr,
                           const double* d_inputs, int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q,
                           const int32_t* scalar_q, const mag_outputs_t* d_out, void* cuda_stream) {
  return mag_solve_range_device(c, ngal, 0, ngal, d_gal_ids, nz, d_t_gyr, d_inputs, n_profile_q, profile_q, n_scalar_q, scalar_q,
                                d_out, cuda_stream);
}

int mag_solve_range_device(mag_ctx_t* c, int32_t ngal_total, int32_t g0, int32_t n, const int32_t* d_gal_ids, int32_t nz,
                           const double* d_t_gyr, const double* d_inputs, int32_t n_profile_q, const int32_t* profile_q,
                           int32_t n_scalar_q, const int32_t* scalar_q, const mag_outputs_t* d_out, void* cuda_stream) {
  if (!c || !d_out || ngal_total < 0 || g0 < 0 || n < 0 || g0 + n > ngal_total) { set_error("bad argument"); return MAG_ERR_BAD_ARGUMENT; }
  if (check_quantities(n_profile_q, profile_q, n_scalar_q, scalar_q)) { set_error("bad quantity id"); return MAG_ERR_BAD_ARGUMENT; }
  if (n == 0) return MAG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  const size_t nr = c->P.p_nx_ref;
  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.nz = nz; a.t_gyr = d_t_gyr; a.gal_ids = d_gal_ids;
  for (int col = 0; col < MAG_NINPUT; col++) a.in[col] = d_inputs + (size_t)col * ngal_total * nz;
  a.n_profile_q = n_profile_q; a.n_scalar_q = n_scalar_q;
  for (int i = 0; i < n_profile_q; i++) { a.profile_q[i] = profile_q[i]; a.prof[i] = d_out->profiles ? d_out->profiles + (size_t)i * ngal_total * nz * nr : nullptr; }
  for (int i = 0; i < n_scalar_q; i++) { a.scalar_q[i] = scalar_q[i]; a.scal[i] = d_out->scalars ? d_out->scalars + (size_t)i * ngal_total * nz : nullptr; }
  a.o_status = d_out->status; a.o_nsteps = d_out->nsteps; a.o_error = d_out->error; a.o_flags = d_out->flags;
  a.o_point_stages = d_out->point_stages;
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  int rc = launch_solve(c, a, g0, n, st, false);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(c->ev[2], st));
  return MAG_OK;
}

int mag_last_device_status(mag_ctx_t* c) {
  if (!c) return MAG_ERR_BAD_ARGUMENT;
  CUDA_TRY(cudaSetDevice(c->device));
  int fail = 0;
  CUDA_TRY(cudaMemcpy(&fail, c->d_counter + CNT_FAIL, sizeof(int), cudaMemcpyDeviceToHost));
  if (fail == MAG_ERR_ARENA) {
    set_error("snapshot-record arena exhausted (grids grew beyond the budget): the results of this call are incomplete; "
              "repeat it through mag_solve_batch, which retries with a larger arena, or set MAG_ARENA_NX_PCT");
    return MAG_ERR_CUDA;
  }
  if (fail) { set_error("grid extension beyond p_nx_MAX / workspace capacity (reference: stop, grid.f90:222-226)"); return fail; }
  return MAG_OK;
}

// ---- host-buffer API: enqueue (H2D, kernels, D2H) for catalogue galaxies [g0, g0+n) of full-catalogue
// host arrays; mag_wait completes the call.
static int enqueue_range(mag_ctx* c) {
  PendingCall& pc = c->pend;
  const int32_t ngal_total = pc.ngal_total, g0 = pc.g0, n = pc.n, nz = pc.nz;
  const int32_t n_profile_q = pc.n_profile_q, n_scalar_q = pc.n_scalar_q;
  const mag_outputs_t* out = &pc.out;
  const size_t nr = c->P.p_nx_ref;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  // Output arrays in pinned (mapped) host memory are written by the kernels directly: every row
  // is written once, by one warp, as contiguous 8-byte stores, so the transfer overlaps the
  // solve instead of following it.  Pageable buffers are staged in HBM and copied afterwards.
  auto mapped = [](const void* h) -> void* {
    if (!h) return nullptr;
    const char* env = getenv("MAG_ZERO_COPY");
    if (env && env[0] == '0') return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
  };
  double* const m_prof = (double*)mapped(out->profiles);
  double* const m_scal = (double*)mapped(out->scalars);
  char* const m_stat = (char*)mapped(out->status);
  // planes of the profile phase (h, n, Omega ...: everything that does not depend on the field) are
  // produced in a burst at the start of the solve; staged in HBM and copied out by the copy engine
  // while k_evolve runs, they do not compete with the other GPUs' bursts for the host's write bandwidth
  const size_t plane = sizeof(double) * (size_t)n * nz * nr;
  int n_staged_planes = 0;
  bool plane_staged[MAG_NPROFILE];
  for (int i = 0; i < n_profile_q; i++) {
    const bool field = is_field_quantity(pc.profile_q[i]);
    plane_staged[i] = out->profiles && (!m_prof || (c->stage_profile_rows && !field));
    n_staged_planes += plane_staged[i] ? 1 : 0;
  }
  const size_t b_in = al(sizeof(double) * MAG_NINPUT * n * nz), b_t = al(sizeof(double) * nz), b_id = al(sizeof(int32_t) * n);
  const size_t b_plane = al(plane);
  const size_t b_scal = (out->scalars && !m_scal) ? al(sizeof(double) * n_scalar_q * n * nz) : 0;
  const size_t b_stat = (out->status && !m_stat) ? al((size_t)n * nz) : 0;
  const size_t b_i32 = al(sizeof(int32_t) * n), b_i64 = al(sizeof(int64_t) * n);
  const size_t total = b_in + b_t + b_id + n_staged_planes * b_plane + b_scal + b_stat + 3 * b_i32 + b_i64;
  if (total > c->stage_bytes) {
    cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0;
    CUDA_TRY(cudaMalloc(&c->d_stage, total));
    c->stage_bytes = total;
  }
  unsigned char* p = (unsigned char*)c->d_stage;
  double* d_in = (double*)p; p += b_in;
  double* d_t = (double*)p; p += b_t;
  int32_t* d_id = (int32_t*)p; p += b_id;
  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.nz = nz; a.t_gyr = d_t;
  a.gal_ids = vbase(d_id, g0, 1);
  for (int col = 0; col < MAG_NINPUT; col++) a.in[col] = vbase(d_in + (size_t)col * n * nz, g0, nz);
  a.n_profile_q = n_profile_q; a.n_scalar_q = n_scalar_q;
  pc.n_early = 0;
  double* staged_plane[MAG_NPROFILE];
  for (int i = 0; i < n_profile_q; i++) {
    a.profile_q[i] = pc.profile_q[i];
    staged_plane[i] = nullptr;
    if (!out->profiles) { a.prof[i] = nullptr; continue; }
    if (plane_staged[i]) {
      staged_plane[i] = (double*)p; p += b_plane;
      a.prof[i] = vbase(staged_plane[i], g0, (size_t)nz * nr);
      if (m_prof && !is_field_quantity(pc.profile_q[i])) {
        auto& e = pc.early[pc.n_early++];
        e.dev = staged_plane[i]; e.host = out->profiles + ((size_t)i * ngal_total + g0) * nz * nr; e.bytes = plane;
      }
    } else {
      a.prof[i] = m_prof + (size_t)i * ngal_total * nz * nr;
    }
  }
  double* d_scal = nullptr;
  if (out->scalars && !m_scal) { d_scal = (double*)p; p += b_scal; }
  for (int i = 0; i < n_scalar_q; i++) {
    a.scalar_q[i] = pc.scalar_q[i];
    a.scal[i] = !out->scalars ? nullptr : (m_scal ? m_scal + (size_t)i * ngal_total * nz : vbase(d_scal + (size_t)i * n * nz, g0, nz));
  }
  char* d_stat = nullptr;
  if (out->status && !m_stat) { d_stat = (char*)p; p += b_stat; }
  a.o_status = !out->status ? nullptr : (m_stat ? m_stat : vbase(d_stat, g0, nz));
  int32_t* d_nsteps = (int32_t*)p; p += b_i32;
  int32_t* d_error = (int32_t*)p; p += b_i32;
  uint32_t* d_flags = (uint32_t*)p; p += b_i32;
  int64_t* d_ps = (int64_t*)p; p += b_i64;
  a.o_nsteps = vbase(d_nsteps, g0, 1); a.o_error = vbase(d_error, g0, 1); a.o_flags = vbase(d_flags, g0, 1);
  a.o_point_stages = vbase(d_ps, g0, 1);

  cudaStream_t st = c->stream;
  CUDA_TRY(cudaEventRecord(c->ev[0], st));
  // the range's slice of every input column: [13][ngal_total][nz] -> [13][n][nz]
  CUDA_TRY(cudaMemcpy2DAsync(d_in, sizeof(double) * n * nz, pc.inputs + (size_t)g0 * nz, sizeof(double) * ngal_total * nz,
                             sizeof(double) * n * nz, MAG_NINPUT, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_t, pc.t_gyr, sizeof(double) * nz, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_id, pc.gal_ids + g0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  int rc = launch_solve(c, a, g0, n, st, pc.n_early > 0);
  if (rc) return rc;
  if (pc.n_early > 0) {
    CUDA_TRY(cudaStreamWaitEvent(c->stream_copy, c->ev_prof, 0));
    for (int i = 0; i < pc.n_early; i++)
      CUDA_TRY(cudaMemcpyAsync(pc.early[i].host, pc.early[i].dev, pc.early[i].bytes, cudaMemcpyDeviceToHost, c->stream_copy));
    CUDA_TRY(cudaEventRecord(c->ev_copy, c->stream_copy));
    CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copy, 0));
  }
  CUDA_TRY(cudaEventRecord(c->ev[2], st));
  if (out->profiles && !m_prof)
    for (int i = 0; i < n_profile_q; i++)
      CUDA_TRY(cudaMemcpyAsync(out->profiles + ((size_t)i * ngal_total + g0) * nz * nr, staged_plane[i], plane, cudaMemcpyDeviceToHost, st));
  if (d_scal)
    CUDA_TRY(cudaMemcpy2DAsync(out->scalars + (size_t)g0 * nz, sizeof(double) * ngal_total * nz, d_scal, sizeof(double) * n * nz,
                               sizeof(double) * n * nz, n_scalar_q, cudaMemcpyDeviceToHost, st));
  if (d_stat) CUDA_TRY(cudaMemcpyAsync(out->status + (size_t)g0 * nz, d_stat, (size_t)n * nz, cudaMemcpyDeviceToHost, st));
  if (out->nsteps) CUDA_TRY(cudaMemcpyAsync(out->nsteps + g0, d_nsteps, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  if (out->error) CUDA_TRY(cudaMemcpyAsync(out->error + g0, d_error, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  if (out->flags) CUDA_TRY(cudaMemcpyAsync(out->flags + g0, d_flags, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
  if (out->point_stages) CUDA_TRY(cudaMemcpyAsync(out->point_stages + g0, d_ps, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(c->h_status, c->d_counter + CNT_FAIL, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(c->h_status + 1, c->d_counter + CNT_INVALIDATED, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(c->h_status + 2, c->d_counter + CNT_REPLAYS, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(c->h_status + 3, c->d_counter + CNT_NHEAVY_SUM, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(c->ev[3], st));
  return MAG_OK;
}

int mag_solve_range_async(mag_ctx_t* c, int32_t ngal_total, int32_t g0, int32_t n, const int32_t* gal_ids, int32_t nz,
                          const double* t_gyr, const double* inputs, int32_t n_profile_q, const int32_t* profile_q,
                          int32_t n_scalar_q, const int32_t* scalar_q, const mag_outputs_t* out) {
  if (!c || !out || ngal_total < 0 || g0 < 0 || n < 0 || g0 + n > ngal_total || !gal_ids || !t_gyr || !inputs) {
    set_error("bad argument"); return MAG_ERR_BAD_ARGUMENT;
  }
  if (check_quantities(n_profile_q, profile_q, n_scalar_q, scalar_q)) { set_error("bad quantity id"); return MAG_ERR_BAD_ARGUMENT; }
  if (c->pend.active) { set_error("a solve is still pending on this context: call mag_wait first"); return MAG_ERR_BAD_ARGUMENT; }
  if (n == 0) return MAG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  PendingCall& pc = c->pend;
  pc.ngal_total = ngal_total; pc.g0 = g0; pc.n = n; pc.nz = nz; pc.gal_ids = gal_ids; pc.t_gyr = t_gyr; pc.inputs = inputs;
  pc.n_profile_q = n_profile_q; pc.n_scalar_q = n_scalar_q;
  for (int i = 0; i < n_profile_q; i++) pc.profile_q[i] = profile_q[i];
  for (int i = 0; i < n_scalar_q; i++) pc.scalar_q[i] = scalar_q[i];
  pc.out = *out;
  const int rc = enqueue_range(c);
  if (rc == MAG_OK) pc.active = true;
  return rc;
}

int mag_wait(mag_ctx_t* c) {
  if (!c) { set_error("null context"); return MAG_ERR_BAD_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(c->device));
  if (!c->pend.active) {   // device-array calls on the context's own stream: just drain it
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return MAG_OK;
  }
  for (int attempt = 0;; attempt++) {
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->pend.active = false; set_error(std::string("solve failed: ") + cudaGetErrorString(e)); return MAG_ERR_CUDA; }
    float ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->last_ms[2] = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->last_ms[1] = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->last_ms[3] = ms;
    const int fail = c->h_status[0];
    if (fail == MAG_ERR_ARENA && attempt < 3) {
      // grids grew beyond the arena budget: repeat the range with twice the budget (smaller chunks)
      c->nx_budget_pct *= 2;
      free_records(c);
      const int rc = enqueue_range(c);
      if (rc != MAG_OK) { c->pend.active = false; return rc; }
      continue;
    }
    c->pend.active = false;
    if (fail == MAG_ERR_ARENA) { set_error("snapshot-record arena exhausted after three retries"); return MAG_ERR_CUDA; }
    if (fail) { set_error("grid extension beyond p_nx_MAX / workspace capacity (reference: stop, grid.f90:222-226)"); return fail; }
    if (c->h_status[1] > 0 && c->pend.n_early > 0) {
      // a run stopped early (status 'i') and rewrote rows of the profile phase after their planes had
      // been copied out: copy them again (rare)
      for (int i = 0; i < c->pend.n_early; i++)
        CUDA_TRY(cudaMemcpy(c->pend.early[i].host, c->pend.early[i].dev, c->pend.early[i].bytes, cudaMemcpyDeviceToHost));
    }
    return MAG_OK;
  }
}

int mag_solve_batch(mag_ctx_t* c, int32_t ngal, const int32_t* gal_ids, int32_t nz, const double* t_gyr,
                    const double* inputs, int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q,
                    const int32_t* scalar_q, const mag_outputs_t* out) {
  const int rc = mag_solve_range_async(c, ngal, 0, ngal, gal_ids, nz, t_gyr, inputs, n_profile_q, profile_q, n_scalar_q, scalar_q, out);
  if (rc != MAG_OK) return rc;
  return mag_wait(c);
}

int mag_last_timing(mag_ctx_t* c, double ms_out[4]) {
  if (!c) return 0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) c->last_ms[1] = ms;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[4]) == cudaSuccess) c->last_ms[0] = ms;  // profile phase (first chunk)
  for (int i = 0; i < 4; i++) ms_out[i] = c->last_ms[i];
  return c->last_launches;
}

double mag_measure_fp64_peak(mag_ctx_t* c, int repeats) {
  if (!c) return 0.0;
  cudaSetDevice(c->device);
  const int threads = 256, blocks = c->sm_count * 8, iters = 1 << 14;
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * threads * blocks) != cudaSuccess) return 0.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < repeats + 1; r++) {
    cudaEventRecord(e0, c->stream);
    k_dfma_peak<<<blocks, threads, 0, c->stream>>>(d, iters);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3);
    if (r > 0 && flops > best) best = flops;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return best;
}

void* mag_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void mag_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- test hooks (exported, not part of the reference-facing ABI)
__global__ void k_test_bessel(const double* x, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double i0, i1, k0, k1;
  detmath::det_bessel_ik01(x[i], &i0, &i1, &k0, &k1);
  out[i] = i0; out[n + i] = i1; out[2 * n + i] = k0; out[3 * n + i] = k1;
}
int mag_test_bessel(const double* x, int n, double* out /* [4][n] */) {
  double *dx = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMalloc(&dx, sizeof(double) * n));
  CUDA_TRY(cudaMalloc(&dout, sizeof(double) * 4 * n));
  CUDA_TRY(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  k_test_bessel<<<(n + 127) / 128, 128>>>(dx, n, dout);
  CUDA_TRY(cudaMemcpy(out, dout, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dout);
  return MAG_OK;
}
__global__ void k_test_fast_sqrt(int n, const double* x, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fast_sqrt_abs(x[i]);
}
// the hot loop's sqrt|x| (mag_rk_fast.cuh)
int mag_test_fast_sqrt(int32_t n, const double* x, double* out) {
  double *dx = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMalloc(&dx, sizeof(double) * n));
  CUDA_TRY(cudaMalloc(&dout, sizeof(double) * n));
  CUDA_TRY(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  k_test_fast_sqrt<<<(n + 127) / 128, 128>>>(n, dx, dout);
  CUDA_TRY(cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dout);
  return MAG_OK;
}
__global__ void k_test_det(int which, int n, const double* x, const double* y, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = which == 0 ? detmath::det_log(x[i]) : which == 1 ? detmath::det_exp(x[i]) : detmath::det_pow(x[i], y[i]);
}
// elementary functions of mag_det_math.cuh: 0 log, 1 exp, 2 pow(x,y)
int mag_test_det_math(int32_t which, int32_t n, const double* x, const double* y, double* out) {
  double *dx = nullptr, *dy = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMalloc(&dx, sizeof(double) * n));
  CUDA_TRY(cudaMalloc(&dy, sizeof(double) * n));
  CUDA_TRY(cudaMalloc(&dout, sizeof(double) * n));
  CUDA_TRY(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dy, y, sizeof(double) * n, cudaMemcpyHostToDevice));
  k_test_det<<<(n + 127) / 128, 128>>>(which, n, dx, dy, dout);
  CUDA_TRY(cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dy); cudaFree(dout);
  return MAG_OK;
}
__global__ void k_test_rng(int32_t igal, int32_t seed, int kind, int n, double* out) {
  __shared__ RngState r;
  if (threadIdx.x == 0) {
    rng_seed(&r, igal, seed, kind);
    for (int i = 0; i < n; i++) out[i] = rng_uniform_lane0(&r);
  }
}
int mag_test_rng(int32_t igal, int32_t seed, int32_t kind, int32_t n, double* out) {
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * n));
  k_test_rng<<<1, 32>>>(igal, seed, kind, n, d);
  CUDA_TRY(cudaMemcpy(out, d, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return MAG_OK;
}
int mag_ctx_info(mag_ctx_t* c, int32_t* out /* [16]: sm_count, profile CTAs/SM, profile warps, nxcap, use_fast, fast-path replays and
                                               heavy-class galaxies of the last solve (last chunk), one-warp teams/SM, one-warp teams,
                                               four-warp teams/SM, four-warp teams, heavy mode, galaxies the snapshot-record arena
                                               holds (a larger call runs as several sub-launches), 3 spare */) {
  if (!c) return MAG_ERR_BAD_ARGUMENT;
  out[0] = c->sm_count; out[1] = c->ctasA_per_sm; out[2] = c->nwarpsA; out[3] = c->nxcap; out[4] = c->use_fast;
  out[5] = 0; out[6] = 0;
  cudaSetDevice(c->device);
  cudaMemcpy(&out[5], c->d_counter + CNT_REPLAYS, sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(&out[6], c->d_counter + CNT_NHEAVY_SUM, sizeof(int), cudaMemcpyDeviceToHost);
  out[7] = c->ctasB_per_sm; out[8] = c->nteamsB; out[9] = c->ctasH_per_sm; out[10] = c->nteamsH; out[11] = c->policy.mode;
  out[12] = c->chunk_cap; out[13] = out[14] = out[15] = 0;
  return MAG_OK;
}
int mag_set_fast_path(mag_ctx_t* c, int32_t on) {
  if (!c) return MAG_ERR_BAD_ARGUMENT;
  c->use_fast = on ? 1 : 0;
  return MAG_OK;
}
// which galaxies run on four-warp teams: mode 0 none, 1 by cost (alpha x the fair share of a warp slot;
// alpha <= 0 keeps the current value), 2 all; nx_heavy > 0: grids above it are heavy
int mag_set_heavy_policy(mag_ctx_t* c, int32_t mode, double alpha, int32_t nx_heavy) {
  if (!c || mode < 0 || mode > 2) return MAG_ERR_BAD_ARGUMENT;
  c->policy.mode = mode;
  if (alpha > 0) { c->heavy_alpha = alpha; c->policy.alpha_over_slots = alpha / (double)c->nteamsB; }
  if (nx_heavy > 0) c->policy.nx_heavy = nx_heavy;
  return MAG_OK;
}

}  // extern "C"

#ifdef MAG_PHASE_CLOCKS
extern "C" int mag_debug_phase_clocks(unsigned long long* out, int reset) {
  if (out && cudaMemcpyFromSymbol(out, magdev::g_phase_clk, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(magdev::g_phase_clk, z, sizeof(z)) != cudaSuccess) return -1; }
  return 0;
}
#endif
