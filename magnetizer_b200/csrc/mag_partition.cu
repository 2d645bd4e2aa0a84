// mag_partition.cu -- static + work-stealing partition of a galaxy catalogue over the GPUs
// of one box, host-side gather of the outputs.
//
// Replaces the reference's MPI master/worker farm (source/distributor.f90:253-384: the
// master hands out one galaxy id at a time with MPI_Send tag 17 and receives the finished
// datasets, data_transfer.f90:80-131) for the per-galaxy dynamo path.  Galaxies are
// independent, so there is no collective and no NCCL on the solve path: the
// catalogue is cut into chunks, every GPU owns a static round-robin share of the chunks and
// steals unclaimed chunks from the other GPUs' shares once its own is exhausted; each GPU
// is driven by one host thread through the single-GPU C ABI (mag_solve_batch) with pinned
// staging buffers, and its results are gathered into the caller's arrays in galaxy order.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/magnetizer_b200.h"

namespace {

struct Pinned {
  void* p = nullptr;
  size_t bytes = 0;
  bool ensure(size_t n) {
    if (n <= bytes) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    if (cudaHostAlloc(&p, n, cudaHostAllocDefault) != cudaSuccess) return false;
    bytes = n;
    return true;
  }
  ~Pinned() { if (p) cudaFreeHost(p); }
};

struct Job {
  const mag_params_t* params;
  int32_t ngal, nz, chunk, nchunks, ndev;
  const int32_t* devices;
  const int32_t* gal_ids;
  const double* t_gyr;
  const double* inputs;
  int32_t n_profile_q, n_scalar_q;
  const int32_t *profile_q, *scalar_q;
  const mag_outputs_t* out;
  std::vector<std::atomic<int>>* claimed;  // [nchunks]
  std::atomic<int>* rc;
  int32_t* chunks_done;                    // [ndev]
  int32_t* chunks_stolen;                  // [ndev]
  std::string* err;                        // [ndev]
};

// chunk k belongs statically to device k % ndev
void worker(const Job& J, int d) {
  const int dev = J.devices[d];
  mag_ctx_t* ctx = nullptr;
  int rc = mag_create(J.params, dev, &ctx);
  if (rc != MAG_OK) { J.rc->store(rc); J.err[d] = mag_last_error(); return; }
  const size_t nr = (size_t)J.params->p_nx_ref, nz = (size_t)J.nz;
  Pinned s_in, s_prof, s_scal, s_stat, s_i32[3], s_i64;
  auto run_chunk = [&](int k) -> int {
    const int g0 = k * J.chunk;
    const int n = (J.ngal - g0 < J.chunk) ? J.ngal - g0 : J.chunk;
    if (!s_in.ensure(sizeof(double) * MAG_NINPUT * n * nz)) return MAG_ERR_CUDA;
    double* in = (double*)s_in.p;
    for (int c = 0; c < MAG_NINPUT; c++)
      std::memcpy(in + (size_t)c * n * nz, J.inputs + ((size_t)c * J.ngal + g0) * nz, sizeof(double) * n * nz);
    mag_outputs_t o;
    std::memset(&o, 0, sizeof(o));
    if (J.out->profiles) { if (!s_prof.ensure(sizeof(double) * J.n_profile_q * n * nz * nr)) return MAG_ERR_CUDA; o.profiles = (double*)s_prof.p; }
    if (J.out->scalars) { if (!s_scal.ensure(sizeof(double) * J.n_scalar_q * n * nz)) return MAG_ERR_CUDA; o.scalars = (double*)s_scal.p; }
    if (J.out->status) { if (!s_stat.ensure((size_t)n * nz)) return MAG_ERR_CUDA; o.status = (char*)s_stat.p; }
    if (J.out->nsteps) { if (!s_i32[0].ensure(sizeof(int32_t) * n)) return MAG_ERR_CUDA; o.nsteps = (int32_t*)s_i32[0].p; }
    if (J.out->error) { if (!s_i32[1].ensure(sizeof(int32_t) * n)) return MAG_ERR_CUDA; o.error = (int32_t*)s_i32[1].p; }
    if (J.out->flags) { if (!s_i32[2].ensure(sizeof(int32_t) * n)) return MAG_ERR_CUDA; o.flags = (uint32_t*)s_i32[2].p; }
    if (J.out->point_stages) { if (!s_i64.ensure(sizeof(int64_t) * n)) return MAG_ERR_CUDA; o.point_stages = (int64_t*)s_i64.p; }
    const int r = mag_solve_batch(ctx, n, J.gal_ids + g0, J.nz, J.t_gyr, in, J.n_profile_q, J.profile_q, J.n_scalar_q,
                                  J.scalar_q, &o);
    if (r != MAG_OK) return r;
    // host-side gather into the caller's arrays, galaxy order preserved
    for (int q = 0; q < J.n_profile_q && o.profiles; q++)
      std::memcpy(J.out->profiles + ((size_t)q * J.ngal + g0) * nz * nr, o.profiles + (size_t)q * n * nz * nr, sizeof(double) * n * nz * nr);
    for (int q = 0; q < J.n_scalar_q && o.scalars; q++)
      std::memcpy(J.out->scalars + ((size_t)q * J.ngal + g0) * nz, o.scalars + (size_t)q * n * nz, sizeof(double) * n * nz);
    if (o.status) std::memcpy(J.out->status + (size_t)g0 * nz, o.status, (size_t)n * nz);
    if (o.nsteps) std::memcpy(J.out->nsteps + g0, o.nsteps, sizeof(int32_t) * n);
    if (o.error) std::memcpy(J.out->error + g0, o.error, sizeof(int32_t) * n);
    if (o.flags) std::memcpy(J.out->flags + g0, o.flags, sizeof(uint32_t) * n);
    if (o.point_stages) std::memcpy(J.out->point_stages + g0, o.point_stages, sizeof(int64_t) * n);
    return MAG_OK;
  };
  auto try_claim = [&](int k) { int exp = 0; return (*J.claimed)[k].compare_exchange_strong(exp, 1); };
  // 1. the static share, in order
  for (int k = d; k < J.nchunks && J.rc->load() == MAG_OK; k += J.ndev) {
    if (!try_claim(k)) continue;
    rc = run_chunk(k);
    if (rc != MAG_OK) { J.rc->store(rc); J.err[d] = mag_last_error(); break; }
    J.chunks_done[d]++;
  }
  // 2. steal from the tails of the other devices' shares
  for (int v = 1; v < J.ndev && J.rc->load() == MAG_OK; v++) {
    const int victim = (d + v) % J.ndev;
    int last = victim + ((J.nchunks - 1 - victim) / J.ndev) * J.ndev;
    for (int k = last; k >= 0 && J.rc->load() == MAG_OK; k -= J.ndev) {
      if (k >= J.nchunks || !try_claim(k)) continue;
      rc = run_chunk(k);
      if (rc != MAG_OK) { J.rc->store(rc); J.err[d] = mag_last_error(); break; }
      J.chunks_done[d]++;
      J.chunks_stolen[d]++;
    }
  }
  mag_destroy(ctx);
}

thread_local std::string g_part_error;

}  // namespace

extern "C" {

const char* mag_partition_last_error(void) { return g_part_error.c_str(); }

int mag_solve_partitioned(const mag_params_t* params, int32_t ndev, const int32_t* devices, int32_t ngal,
                          const int32_t* gal_ids, int32_t nz, const double* t_gyr, const double* inputs,
                          int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q, const int32_t* scalar_q,
                          const mag_outputs_t* out, int32_t chunk, int32_t* chunks_done, int32_t* chunks_stolen) {
  if (!params || ndev < 1 || !devices || ngal < 0 || !out || chunk < 1) { g_part_error = "bad argument"; return MAG_ERR_BAD_ARGUMENT; }
  if (ngal == 0) return MAG_OK;
  const int nchunks = (ngal + chunk - 1) / chunk;
  std::vector<std::atomic<int>> claimed(nchunks);
  for (auto& c : claimed) c.store(0);
  std::atomic<int> rc(MAG_OK);
  std::vector<int32_t> done(ndev, 0), stolen(ndev, 0);
  std::vector<std::string> errs(ndev);
  Job J{params, ngal, nz, chunk, nchunks, ndev, devices, gal_ids, t_gyr, inputs, n_profile_q, n_scalar_q, profile_q, scalar_q,
        out, &claimed, &rc, done.data(), stolen.data(), errs.data()};
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; d++) th.emplace_back(worker, std::cref(J), d);
  for (auto& t : th) t.join();
  for (int d = 0; d < ndev; d++) {
    if (chunks_done) chunks_done[d] = done[d];
    if (chunks_stolen) chunks_stolen[d] = stolen[d];
    if (!errs[d].empty()) g_part_error = errs[d];
  }
  return rc.load();
}

}  // extern "C"
