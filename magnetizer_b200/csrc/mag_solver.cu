// mag_solver.cu -- persistent warp-per-galaxy kernel and the C ABI (include/magnetizer_b200.h).
//
// Replaces, for the per-galaxy dynamo path only, the reference's MPI master/worker farm
// (source/distributor.f90:253-384) + work routine dynamo_run (source/dynamo.f90:39): the
// farm becomes a device-side work queue (one atomic counter) drained by persistent warps,
// each of which runs a whole galaxy history.  No collective, no host round trip between
// snapshots; inputs are staged into HBM once per batch and read with coalesced loads
// (snapshot index fastest), outputs are written as contiguous 101-point rows.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "mag_galaxy.cuh"
#include "mag_phases.cuh"
#include "mag_rk_fast.cuh"

using namespace magdev;

#define WARPS_PER_CTA 4
#ifndef MAG_PROFILE_MAXREG
#define MAG_PROFILE_MAXREG 128   // 16 warps per SM: the profile phase is latency bound (serial root-finder chains)
#endif
#ifndef MAG_EVOLVE_MAXREG
#if MAG_TEAM == 32
#define MAG_EVOLVE_MAXREG 255   // 8 one-warp teams per SM, everything of a galaxy in registers
#else
#define MAG_EVOLVE_MAXREG 168
#endif
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
      if (a.out.error) a.out.error[G.g] = H.error;
      if (err) {  // nothing else is written for this galaxy (dynamo.f90:69-73)
        if (a.out.nsteps) a.out.nsteps[G.g] = G.nsteps;
        if (a.out.point_stages) a.out.point_stages[G.g] = 0;
        if (a.out.flags) a.out.flags[G.g] = G.flags;
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

// ------------------------------------------------------------------ Phase B kernel
// One team of MAG_TEAM threads per galaxy, persistent, longest history first.  MAG_TEAM = 32
// (default build): the team is ONE warp that does everything, the time loop in registers with
// shuffles (mag_rk_warp.cuh).  MAG_TEAM = 128 (grids of more than 256 points): warp 0 leads
// (seeds, grid transformations of f, outputs, generic path), the helper warps join for the
// shared-line time loop (mag_rk_fast.cuh).
__global__ void __maxnreg__(MAG_EVOLVE_MAXREG) k_evolve(KernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RkShared* rs = reinterpret_cast<RkShared*>(smem_raw);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) {  // helper warps (MAG_TEAM > 32 only)
    for (;;) {
      __syncthreads();
      if (rs->call.cmd == TEAM_CMD_EXIT) break;
      fast_dispatch(nullptr, rs);
    }
    return;
  }
  Gal G;
  G.a = &a;
  G.W = a.ws2 + (size_t)blockIdx.x * A_COUNT * a.nxcap;
  G.rng = &rs->rng;
  G.lane = lane;
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(a.counter2, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= a.g1 - a.g0) break;
    const int qi = a.order[q];
    const GalHead H = a.heads[qi];
    if (H.error) continue;
    G.g = a.g0 + qi;
    run_evolve(G, H, rs);
  }
  if (lane == 0) rs->call.cmd = TEAM_CMD_EXIT;
  __syncthreads();
}

// ------------------------------------------------------------------ cost-ordered queue
// Longest histories first (exact cost = sum of nsteps*nx known after Phase A): a counting
// sort on 1/16-octave buckets of the cost.
#define COST_BUCKETS 1024
__device__ __forceinline__ int cost_bucket(long long c) {
  if (c <= 0) return 0;
  const int b = (int)(log2f((float)c) * 16.0f) + 1;
  return b >= COST_BUCKETS ? COST_BUCKETS - 1 : b;
}
__global__ void k_cost_hist(const GalHead* heads, int n, int* hist) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(&hist[cost_bucket(heads[i].cost)], 1);
}
__global__ void k_cost_scan(int* hist) {  // descending exclusive scan, in place
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int k = COST_BUCKETS - 1; k >= 0; k--) { int c = hist[k]; hist[k] = acc; acc += c; }
  }
}
__global__ void k_cost_scatter(const GalHead* heads, int n, int* hist, int* order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  order[atomicAdd(&hist[cost_bucket(heads[i].cost)], 1)] = i;
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
struct mag_ctx {
  mag_params_t P;
  int device = 0;
  int sm_count = 0;
  int ctasA_per_sm = 0, ctasB_per_sm = 0;
  int nwarpsA = 0, nteamsB = 0;
  int nxcap = 0;
  double* wsA = nullptr;      // Phase A: per-warp workspaces
  double* wsB = nullptr;      // Phase B: per-team workspaces
  int* d_counter = nullptr;   // [0] A queue head, [1] fail flag, [2] fast-path replays, [3] B queue head
  unsigned long long* d_arena_top = nullptr;
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
  // staging for the host-buffer API
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double last_ms[4] = {0, 0, 0, 0};
  int last_launches = 0;
  int last_replays = 0;
  int use_fast = 1;
  // forwarding to the MAG_TEAM = 128 build of this library for grids of more than 256 points
  void* fwd_lib = nullptr;
  mag_ctx* fwd_ctx = nullptr;
};

#if MAG_TEAM == 32
namespace {
struct FwdApi {
  int (*create)(const mag_params_t*, int, mag_ctx_t**);
  void (*destroy)(mag_ctx_t*);
  int (*solve)(mag_ctx_t*, int32_t, const int32_t*, int32_t, const double*, const double*, int32_t, const int32_t*, int32_t,
               const int32_t*, const mag_outputs_t*);
  int (*solve_dev)(mag_ctx_t*, int32_t, const int32_t*, int32_t, const double*, const double*, int32_t, const int32_t*, int32_t,
                   const int32_t*, const mag_outputs_t*, void*);
  int (*timing)(mag_ctx_t*, double*);
  double (*peak)(mag_ctx_t*, int);
  int (*info)(mag_ctx_t*, int32_t*);
  int (*set_fast)(mag_ctx_t*, int32_t);
  const char* (*last_error)(void);
};
FwdApi g_fwd;
void* open_big_team_library() {
  Dl_info di;
  if (!dladdr((void*)&open_big_team_library, &di) || !di.dli_fname) return nullptr;
  std::string path(di.dli_fname);
  const size_t slash = path.rfind('/');
  path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/libmagnetizer_b200_team128.so";
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) return nullptr;
  g_fwd.create = (decltype(g_fwd.create))dlsym(h, "mag_create");
  g_fwd.destroy = (decltype(g_fwd.destroy))dlsym(h, "mag_destroy");
  g_fwd.solve = (decltype(g_fwd.solve))dlsym(h, "mag_solve_batch");
  g_fwd.solve_dev = (decltype(g_fwd.solve_dev))dlsym(h, "mag_solve_batch_device");
  g_fwd.timing = (decltype(g_fwd.timing))dlsym(h, "mag_last_timing");
  g_fwd.peak = (decltype(g_fwd.peak))dlsym(h, "mag_measure_fp64_peak");
  g_fwd.info = (decltype(g_fwd.info))dlsym(h, "mag_ctx_info");
  g_fwd.set_fast = (decltype(g_fwd.set_fast))dlsym(h, "mag_set_fast_path");
  g_fwd.last_error = (decltype(g_fwd.last_error))dlsym(h, "mag_last_error");
  if (!g_fwd.create || !g_fwd.destroy || !g_fwd.solve || !g_fwd.solve_dev || !g_fwd.timing || !g_fwd.peak || !g_fwd.info ||
      !g_fwd.set_fast || !g_fwd.last_error) { dlclose(h); return nullptr; }
  return h;
}
}  // namespace
#define MAG_FWD(c, call) if ((c) && (c)->fwd_ctx) return call
#else
#define MAG_FWD(c, call)
#endif

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
  p->p_allow_positive_shears = 0; p->seed_choice = 0; p->outflow_type = 0;
}

static bool params_supported(const mag_params_t& p) {
  return p.abi_version == MAG_ABI_VERSION && p.p_variable_timesteps == 1 && p.p_simplified_pressure == 1 &&
         p.p_halo_contraction == 1 && p.p_enable_P2 == 0 && p.Dyn_quench == 1 && p.Alg_quench == 0 && p.Damp == 0 &&
         p.lFloor == 1 && p.p_space_varying_floor == 1 && p.p_time_varying_floor == 1 &&
         p.p_limit_turbulent_scale == 1 && p.p_allow_positive_shears == 0 && p.seed_choice == 0 &&
         p.outflow_type == 0 && p.p_nx_ref >= 8 && (p.rng_kind == 0 || p.rng_kind == 1) &&
         p.p_nx_MAX >= p.p_nx_ref + 6;
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
#if MAG_TEAM == 32
  if (params->p_nx_ref + 2 * MAG_NGHOST > 256) {
    // more than 256 grid points: use the 128-thread-team build if it is there (same ABI)
    if (void* h = open_big_team_library()) {
      mag_ctx* w = new mag_ctx();
      w->P = *params; w->device = device; w->fwd_lib = h;
      const int rc = g_fwd.create(params, device, &w->fwd_ctx);
      if (rc != MAG_OK) { set_error(g_fwd.last_error()); delete w; return rc; }
      *ctx_out = w;
      return MAG_OK;
    }
  }
#endif
  mag_ctx* c = new mag_ctx();
  c->P = *params;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  const char* env_split = getenv("MAG_INLINE_CHAINS");
  c->split_chains = (env_split && env_split[0] == '1') ? 0 : 1;
  const char* env_fast = getenv("MAG_DISABLE_FAST_PATH");
  c->use_fast = (env_fast && env_fast[0] == '1') ? 0 : 1;
  const size_t smemB = sizeof(RkShared);
  CUDA_TRY(cudaFuncSetAttribute(k_evolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
  int occA = 0, occB = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, k_profiles, WARPS_PER_CTA * 32, 0));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, k_evolve, MAG_TEAM, smemB));
  if (occA < 1 || occB < 1) { set_error("kernel does not fit on an SM"); delete c; return MAG_ERR_CUDA; }
  c->ctasA_per_sm = occA; c->ctasB_per_sm = occB;
  c->nwarpsA = c->sm_count * occA * WARPS_PER_CTA;
  c->nteamsB = c->sm_count * occB;
  const char* env_cap = getenv("MAG_NXCAP");
  int cap = env_cap ? atoi(env_cap) : 2048;
  if (cap > params->p_nx_MAX) cap = params->p_nx_MAX;
  if (cap < params->p_nx_ref + 6) cap = params->p_nx_ref + 6;
  if (cap < MAG_FAST_MAXPPL * MAG_TEAM) cap = MAG_FAST_MAXPPL * MAG_TEAM;  // the fast path addresses up to 256 points
#if MAG_TEAM == 32
  if (cap < 1024) cap = 1024;   // room for the coefficient pair table of the large-grid warp variants (mag_rk_warp.cuh)
#endif
  c->nxcap = (cap + 15) & ~15;
  const size_t wsA = (size_t)c->nwarpsA * A_COUNT * c->nxcap * sizeof(double);
  const size_t wsB = (size_t)c->nteamsB * A_COUNT * c->nxcap * sizeof(double);
  CUDA_TRY(cudaMalloc(&c->wsA, wsA));
  CUDA_TRY(cudaMemset(c->wsA, 0, wsA));
  CUDA_TRY(cudaMalloc(&c->wsB, wsB));
  CUDA_TRY(cudaMemset(c->wsB, 0, wsB));
  CUDA_TRY(cudaMalloc(&c->d_counter, 4 * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_arena_top, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&c->d_hist, COST_BUCKETS * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_job_count, sizeof(int)));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 6; i++) CUDA_TRY(cudaEventCreate(&c->ev[i]));
  *ctx_out = c;
  return MAG_OK;
}

void mag_destroy(mag_ctx_t* c) {
  if (!c) return;
#if MAG_TEAM == 32
  if (c->fwd_ctx) { g_fwd.destroy(c->fwd_ctx); delete c; return; }
#endif
  cudaSetDevice(c->device);
  cudaFree(c->wsA); cudaFree(c->wsB); cudaFree(c->d_counter); cudaFree(c->d_arena_top); cudaFree(c->d_hist);
  cudaFree(c->d_order); cudaFree(c->d_recs); cudaFree(c->d_heads); cudaFree(c->d_arena);
  cudaFree(c->d_jobs); cudaFree(c->d_jobidx); cudaFree(c->d_job_count);
  cudaFree(c->d_stage);
  for (int i = 0; i < 6; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

static int check_quantities(int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q, const int32_t* scalar_q) {
  if (n_profile_q < 0 || n_profile_q > MAG_NPROFILE || n_scalar_q < 0 || n_scalar_q > MAG_NSCALAR) return MAG_ERR_BAD_ARGUMENT;
  for (int i = 0; i < n_profile_q; i++) if (profile_q[i] < 0 || profile_q[i] >= MAG_NPROFILE) return MAG_ERR_BAD_ARGUMENT;
  for (int i = 0; i < n_scalar_q; i++) if (scalar_q[i] < 0 || scalar_q[i] >= MAG_NSCALAR) return MAG_ERR_BAD_ARGUMENT;
  return MAG_OK;
}

// Record storage: SnapRec/GalHead tables for `chunk` galaxies and an arena for their arrays.
// The arena budget per snapshot is R_COUNT arrays of ARENA_NX_BUDGET points (default grid
// 107; grid extensions reach ~330 in the example input but are rare); a batch that does not
// fit is solved in several chunks, and a chunk that overflows the arena is reported as
// MAG_ERR_CUDA with an explanatory message.
#define ARENA_ARRAYS (R_COUNT + CHAIN_JOB_ARRAYS + 1)   // record arrays + the arrays of a chain job (+ seeds of reseeded snapshots)
static long long arena_need(const mag_ctx* c, int chunk, int nz) {
  // average points per snapshot budgeted: 1.65 x the default grid (176 for nx = 107; grid
  // extensions reach 3 x nx in the example input but only for ~1 % of the snapshots)
  const long long nx_budget = ((long long)(c->P.p_nx_ref + 2 * MAG_NGHOST) * 165 + 99) / 100;
  return (long long)chunk * nz * ARENA_ARRAYS * nx_budget;
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
  if (chunk > ngal) chunk = ngal;
  if (chunk < 1) { set_error("not enough device memory for the snapshot records"); return MAG_ERR_CUDA; }
  if (chunk <= c->chunk_cap && nz <= c->nz_cap) { *chunk_out = (int)chunk; return MAG_OK; }
  cudaFree(c->d_order); cudaFree(c->d_recs); cudaFree(c->d_heads); cudaFree(c->d_arena); cudaFree(c->d_jobs); cudaFree(c->d_jobidx);
  c->d_order = nullptr; c->d_recs = nullptr; c->d_heads = nullptr; c->d_arena = nullptr; c->d_jobs = nullptr; c->d_jobidx = nullptr;
  c->chunk_cap = 0; c->arena_cap = 0;
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

static int launch_solve(mag_ctx* c, int32_t ngal, const int32_t* d_gal_ids, int32_t nz, const double* d_t,
                        const double* d_inputs, int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q,
                        const int32_t* scalar_q, const mag_outputs_t* d_out, cudaStream_t st) {
  if (nz < 1 || nz > 4000) { set_error("nz out of range"); return MAG_ERR_BAD_ARGUMENT; }
  int chunk = 0;
  int rc = ensure_records(c, ngal, nz, &chunk);
  if (rc) return rc;
  int launches = 0;
  CUDA_TRY(cudaMemsetAsync(c->d_counter + 1, 0, 2 * sizeof(int), st));  // fail flag, replays
  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.P = c->P;
  a.U = make_units();
  a.ngal = ngal; a.nz = nz;
  a.gal_ids = d_gal_ids; a.t_gyr = d_t; a.inputs = d_inputs; a.order = c->d_order;
  a.n_profile_q = n_profile_q; a.n_scalar_q = n_scalar_q;
  for (int i = 0; i < n_profile_q; i++) a.profile_q[i] = profile_q[i];
  for (int i = 0; i < n_scalar_q; i++) a.scalar_q[i] = scalar_q[i];
  a.out = *d_out;
  a.ws = c->wsA; a.ws2 = c->wsB; a.nxcap = c->nxcap;
  a.counter = c->d_counter; a.fail = c->d_counter + 1; a.replays = c->d_counter + 2; a.counter2 = c->d_counter + 3;
  a.use_fast = c->use_fast;
  a.recs = c->d_recs; a.heads = c->d_heads; a.arena = c->d_arena; a.arena_top = c->d_arena_top; a.arena_cap = c->arena_cap;
  a.jobs = c->d_jobs; a.job_count = c->d_job_count; a.jobidx = c->d_jobidx;
  for (int g0 = 0; g0 < ngal; g0 += chunk) {
    const int n = (ngal - g0 < chunk) ? ngal - g0 : chunk;
    a.g0 = g0; a.g1 = g0 + n;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(c->d_counter + 3, 0, sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(c->d_arena_top, 0, sizeof(unsigned long long), st));
    CUDA_TRY(cudaMemsetAsync(c->d_hist, 0, COST_BUCKETS * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(c->d_recs, 0, (size_t)n * nz * sizeof(SnapRec), st));
    int ctasA = c->sm_count * c->ctasA_per_sm;
    const int needA = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (needA < ctasA) ctasA = needA;
    a.job_cap = n * nz * 2;
    if (c->split_chains) {
      // pass 1: walk the grid recurrence, export the hydrostatic chains; then 32 chains per warp
      CUDA_TRY(cudaMemsetAsync(c->d_job_count, 0, sizeof(int), st));
      CUDA_TRY(cudaMemsetAsync(c->d_jobidx, 0xFF, (size_t)n * nz * 2 * sizeof(int), st));
      a.chain_mode = CHAIN_MODE_EXPORT;
      k_profiles<<<ctasA, WARPS_PER_CTA * 32, 0, st>>>(a);
      k_chains<<<c->sm_count * 8, 128, 0, st>>>(a);
      CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, sizeof(int), st));
      a.chain_mode = CHAIN_MODE_IMPORT;
      launches += 2;
    } else {
      a.chain_mode = CHAIN_MODE_INLINE;
    }
    k_profiles<<<ctasA, WARPS_PER_CTA * 32, 0, st>>>(a);
    const int tb = 256, gb = (n + tb - 1) / tb;
    k_cost_hist<<<gb, tb, 0, st>>>(c->d_heads, n, c->d_hist);
    k_cost_scan<<<1, 32, 0, st>>>(c->d_hist);
    k_cost_scatter<<<gb, tb, 0, st>>>(c->d_heads, n, c->d_hist, c->d_order);
    if (g0 == 0) CUDA_TRY(cudaEventRecord(c->ev[4], st));
    int teams = c->nteamsB;
    if (n < teams) teams = n;
    k_evolve<<<teams, MAG_TEAM, sizeof(RkShared), st>>>(a);
    launches += 5;
    CUDA_TRY(cudaGetLastError());
  }
  c->last_launches = launches;
  return MAG_OK;
}

int mag_solve_batch_device(mag_ctx_t* c, int32_t ngal, const int32_t* d_gal_ids, int32_t nz, const double* d_t_gyr,
                           const double* d_inputs, int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q,
                           const int32_t* scalar_q, const mag_outputs_t* d_out, void* cuda_stream) {
  MAG_FWD(c, g_fwd.solve_dev(c->fwd_ctx, ngal, d_gal_ids, nz, d_t_gyr, d_inputs, n_profile_q, profile_q, n_scalar_q, scalar_q, d_out, cuda_stream));
  if (!c || !d_out || ngal < 0) { set_error("bad argument"); return MAG_ERR_BAD_ARGUMENT; }
  if (check_quantities(n_profile_q, profile_q, n_scalar_q, scalar_q)) { set_error("bad quantity id"); return MAG_ERR_BAD_ARGUMENT; }
  if (ngal == 0) return MAG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  int rc = launch_solve(c, ngal, d_gal_ids, nz, d_t_gyr, d_inputs, n_profile_q, profile_q, n_scalar_q, scalar_q, d_out, st);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(c->ev[2], st));
  return MAG_OK;
}

int mag_solve_batch(mag_ctx_t* c, int32_t ngal, const int32_t* gal_ids, int32_t nz, const double* t_gyr,
                    const double* inputs, int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q,
                    const int32_t* scalar_q, const mag_outputs_t* out) {
#if MAG_TEAM == 32
  if (c && c->fwd_ctx) {
    const int rc = g_fwd.solve(c->fwd_ctx, ngal, gal_ids, nz, t_gyr, inputs, n_profile_q, profile_q, n_scalar_q, scalar_q, out);
    if (rc != MAG_OK) set_error(g_fwd.last_error());
    return rc;
  }
#endif
  if (!c || !out || ngal < 0 || !gal_ids || !t_gyr || !inputs) { set_error("bad argument"); return MAG_ERR_BAD_ARGUMENT; }
  if (check_quantities(n_profile_q, profile_q, n_scalar_q, scalar_q)) { set_error("bad quantity id"); return MAG_ERR_BAD_ARGUMENT; }
  if (ngal == 0) return MAG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
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
  void* const m_prof = mapped(out->profiles);
  void* const m_scal = mapped(out->scalars);
  void* const m_stat = mapped(out->status);
  const size_t b_in = al(sizeof(double) * MAG_NINPUT * ngal * nz), b_t = al(sizeof(double) * nz), b_id = al(sizeof(int32_t) * ngal);
  const size_t b_prof = (out->profiles && !m_prof) ? al(sizeof(double) * n_profile_q * ngal * nz * nr) : 0;
  const size_t b_scal = (out->scalars && !m_scal) ? al(sizeof(double) * n_scalar_q * ngal * nz) : 0;
  const size_t b_stat = (out->status && !m_stat) ? al((size_t)ngal * nz) : 0;
  const size_t b_i32 = al(sizeof(int32_t) * ngal), b_i64 = al(sizeof(int64_t) * ngal);
  const size_t total = b_in + b_t + b_id + b_prof + b_scal + b_stat + 3 * b_i32 + b_i64;
  if (total > c->stage_bytes) {
    cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0;
    CUDA_TRY(cudaMalloc(&c->d_stage, total));
    c->stage_bytes = total;
  }
  unsigned char* p = (unsigned char*)c->d_stage;
  double* d_in = (double*)p; p += b_in;
  double* d_t = (double*)p; p += b_t;
  int32_t* d_id = (int32_t*)p; p += b_id;
  mag_outputs_t d;
  d.profiles = out->profiles ? (m_prof ? (double*)m_prof : (double*)p) : nullptr; p += b_prof;
  d.scalars = out->scalars ? (m_scal ? (double*)m_scal : (double*)p) : nullptr; p += b_scal;
  d.status = out->status ? (m_stat ? (char*)m_stat : (char*)p) : nullptr; p += b_stat;
  d.nsteps = (int32_t*)p; p += b_i32;
  d.error = (int32_t*)p; p += b_i32;
  d.flags = (uint32_t*)p; p += b_i32;
  d.point_stages = (int64_t*)p; p += b_i64;
  cudaStream_t st = c->stream;
  CUDA_TRY(cudaEventRecord(c->ev[0], st));
  CUDA_TRY(cudaMemcpyAsync(d_in, inputs, sizeof(double) * MAG_NINPUT * ngal * nz, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_t, t_gyr, sizeof(double) * nz, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_id, gal_ids, sizeof(int32_t) * ngal, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  int rc = launch_solve(c, ngal, d_id, nz, d_t, d_in, n_profile_q, profile_q, n_scalar_q, scalar_q, &d, st);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(c->ev[2], st));
  if (out->profiles && !m_prof) CUDA_TRY(cudaMemcpyAsync(out->profiles, d.profiles, sizeof(double) * n_profile_q * ngal * nz * nr, cudaMemcpyDeviceToHost, st));
  if (out->scalars && !m_scal) CUDA_TRY(cudaMemcpyAsync(out->scalars, d.scalars, sizeof(double) * n_scalar_q * ngal * nz, cudaMemcpyDeviceToHost, st));
  if (out->status && !m_stat) CUDA_TRY(cudaMemcpyAsync(out->status, d.status, (size_t)ngal * nz, cudaMemcpyDeviceToHost, st));
  if (out->nsteps) CUDA_TRY(cudaMemcpyAsync(out->nsteps, d.nsteps, sizeof(int32_t) * ngal, cudaMemcpyDeviceToHost, st));
  if (out->error) CUDA_TRY(cudaMemcpyAsync(out->error, d.error, sizeof(int32_t) * ngal, cudaMemcpyDeviceToHost, st));
  if (out->flags) CUDA_TRY(cudaMemcpyAsync(out->flags, d.flags, sizeof(uint32_t) * ngal, cudaMemcpyDeviceToHost, st));
  if (out->point_stages) CUDA_TRY(cudaMemcpyAsync(out->point_stages, d.point_stages, sizeof(int64_t) * ngal, cudaMemcpyDeviceToHost, st));
  int fail = 0;
  CUDA_TRY(cudaMemcpyAsync(&fail, c->d_counter + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(c->ev[3], st));
  CUDA_TRY(cudaStreamSynchronize(st));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->last_ms[2] = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->last_ms[1] = ms;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->last_ms[3] = ms;
  if (fail == MAG_ERR_ARENA) { set_error("snapshot-record arena exhausted (grids grew beyond the budget): lower MAG_ARENA_FRACTION chunking or split the batch"); return MAG_ERR_CUDA; }
  if (fail) { set_error("grid extension beyond p_nx_MAX / workspace capacity (reference: stop, grid.f90:222-226)"); return fail; }
  return MAG_OK;
}

int mag_last_timing(mag_ctx_t* c, double ms_out[4]) {
  MAG_FWD(c, g_fwd.timing(c->fwd_ctx, ms_out));
  if (!c) return 0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) c->last_ms[1] = ms;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[4]) == cudaSuccess) c->last_ms[0] = ms;  // profile phase (first chunk)
  for (int i = 0; i < 4; i++) ms_out[i] = c->last_ms[i];
  return c->last_launches;
}

double mag_measure_fp64_peak(mag_ctx_t* c, int repeats) {
  MAG_FWD(c, g_fwd.peak(c->fwd_ctx, repeats));
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
int mag_ctx_info(mag_ctx_t* c, int32_t* out /* sm_count, ctas_per_sm, nwarps, nxcap, use_fast, replays of the last solve, ..., team */) {
  MAG_FWD(c, g_fwd.info(c->fwd_ctx, out));
  if (!c) return MAG_ERR_BAD_ARGUMENT;
  out[9] = MAG_TEAM;
  out[0] = c->sm_count; out[1] = c->ctasA_per_sm; out[2] = c->nwarpsA; out[3] = c->nxcap; out[4] = c->use_fast;
  out[5] = 0;
  cudaSetDevice(c->device);
  cudaMemcpy(&out[5], c->d_counter + 2, sizeof(int), cudaMemcpyDeviceToHost);
  out[6] = c->ctasA_per_sm; out[7] = c->ctasB_per_sm; out[8] = c->nteamsB;
  return MAG_OK;
}
int mag_set_fast_path(mag_ctx_t* c, int32_t on) {
  MAG_FWD(c, g_fwd.set_fast(c->fwd_ctx, on));
  if (!c) return MAG_ERR_BAD_ARGUMENT;
  c->use_fast = on ? 1 : 0;
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
