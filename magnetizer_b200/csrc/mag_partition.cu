// mag_partition.cu -- static + work-stealing partition of a galaxy catalogue over the GPUs
// of one box, results gathered in place in the caller's arrays.
//
// Replaces the reference's MPI master/worker farm (source/distributor.f90:253-384: the
// master hands out one galaxy id at a time with MPI_Send tag 17 and receives the finished
// datasets, data_transfer.f90:80-131) for the per-galaxy dynamo path.  Galaxies are
// independent, so there is no collective and no NCCL on the solve path: the
// catalogue is cut into chunks, every GPU owns a static round-robin share of the chunks and
// steals unclaimed chunks from the tails of the other GPUs' shares once its own is exhausted.
//
// Each GPU is driven by two persistent host threads, each with its own solver context and
// stream (mag_solve_range_async / mag_wait of the single-GPU ABI): while one chunk drains --
// its last, longest histories -- the other thread's chunk already runs its profile phase and
// the head of its evolve queue on the SMs that became free, so chunk boundaries cost nothing
// and a chunk's tail is never exposed except at the very end of the catalogue.  Inputs are
// copied straight from the caller's arrays (one strided H2D per chunk) and results are stored
// straight into them -- by the kernels themselves when the arrays are page-locked -- in galaxy
// order: the "host-side gather" of the reference's master is the address computation.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/magnetizer_b200.h"

namespace {

struct Job {
  bool device_mode = false;
  int32_t ngal = 0, nz = 0, chunk = 0, nchunks = 0;
  // host mode: full-catalogue host arrays
  const int32_t* gal_ids = nullptr;
  const double* t_gyr = nullptr;
  const double* inputs = nullptr;
  const mag_outputs_t* out = nullptr;
  // device mode: per-device copies
  const int32_t* const* d_gal_ids = nullptr;
  const double* const* d_t_gyr = nullptr;
  const double* const* d_inputs = nullptr;
  const mag_outputs_t* d_out = nullptr;
  int32_t* owner = nullptr;
  int32_t n_profile_q = 0, n_scalar_q = 0;
  const int32_t *profile_q = nullptr, *scalar_q = nullptr;
  std::unique_ptr<std::atomic<int>[]> claimed;   // [nchunks]
};

}  // namespace

struct mag_partition {
  mag_params_t params;
  int ndev = 0, per_dev = 2;
  std::vector<int32_t> devices;
  struct Worker {
    std::thread th;
    mag_ctx_t* ctx = nullptr;
    int d = 0;           // index into devices
  };
  std::vector<Worker> workers;
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  long generation = 0;
  int running = 0;
  bool shutdown = false;
  Job job;
  std::atomic<int> rc{MAG_OK};
  std::vector<std::atomic<int>> done, stolen;   // [ndev]
  std::vector<std::atomic<int>> active;         // [ndev] threads of the device that are inside a chunk right now
  std::atomic<int> unclaimed{0};                // chunks nobody has taken yet
  std::vector<std::atomic<long long>> launches, busy_us;   // [ndev] kernel launches / device time inside the kernels of the last solve
  std::string err;

  mag_partition(int nd) : done(nd), stolen(nd), active(nd), launches(nd), busy_us(nd) {}
};

namespace {

thread_local std::string g_part_error;

void fail(mag_partition* P, int rc, const char* msg) {
  int expect = MAG_OK;
  if (P->rc.compare_exchange_strong(expect, rc)) {
    std::lock_guard<std::mutex> lk(P->mu);
    P->err = msg ? msg : "";
  }
}

int run_chunk(mag_partition* P, mag_partition::Worker& w, int k) {
  const Job& J = P->job;
  const int g0 = k * J.chunk;
  const int n = (J.ngal - g0 < J.chunk) ? J.ngal - g0 : J.chunk;
  if (J.device_mode) {
    int rc = mag_solve_range_device(w.ctx, J.ngal, g0, n, J.d_gal_ids[w.d], J.nz, J.d_t_gyr[w.d], J.d_inputs[w.d], J.n_profile_q,
                                    J.profile_q, J.n_scalar_q, J.scalar_q, &J.d_out[w.d], nullptr);
    if (rc != MAG_OK) return rc;
    rc = mag_wait(w.ctx);
    if (rc != MAG_OK) return rc;
    rc = mag_last_device_status(w.ctx);
    if (rc != MAG_OK) return rc;
    if (J.owner) for (int g = g0; g < g0 + n; g++) J.owner[g] = w.d;
    return MAG_OK;
  }
  const int rc = mag_solve_range_async(w.ctx, J.ngal, g0, n, J.gal_ids, J.nz, J.t_gyr, J.inputs, J.n_profile_q, J.profile_q,
                                       J.n_scalar_q, J.scalar_q, J.out);
  if (rc != MAG_OK) return rc;
  return mag_wait(w.ctx);
}

// chunk k belongs statically to device k % ndev
void run_job(mag_partition* P, mag_partition::Worker& w) {
  const Job& J = P->job;
  const int d = w.d, ndev = P->ndev;
  auto try_claim = [&](int k) {
    int exp = 0;
    if (!J.claimed[k].compare_exchange_strong(exp, 1)) return false;
    P->unclaimed--;
    return true;
  };
  auto one = [&](int k, bool stolen) {
    P->active[d]++;
    const int rc = run_chunk(P, w, k);
    P->active[d]--;
    if (rc != MAG_OK) { fail(P, rc, mag_last_error()); return false; }
    double ms[4] = {0, 0, 0, 0};
    P->launches[d] += mag_last_timing(w.ctx, ms);
    P->busy_us[d] += (long long)(ms[1] * 1e3);
    P->done[d]++;
    if (stolen) P->stolen[d]++;
    return true;
  };
  // 1. the static share, in order (the two threads of a device take alternate chunks of it)
  for (int k = d; k < J.nchunks && P->rc.load() == MAG_OK; k += ndev)
    if (try_claim(k) && !one(k, false)) return;
  // 2. steal from the tails of the other devices' shares -- but only from a device whose threads are ALL inside a
  // chunk: an unclaimed chunk of a device with a free thread is about to be taken by its owner (at start-up every
  // share is unclaimed; a thread that loses the race for its own device's only chunk must not grab the neighbour's)
  while (P->unclaimed.load() > 0 && P->rc.load() == MAG_OK) {
    bool took = false;
    for (int v = 1; v < ndev && !took; v++) {
      const int victim = (d + v) % ndev;
      if (victim >= J.nchunks) continue;
      const bool same_gpu = P->devices[victim] == P->devices[d];   // a GPU listed twice: its threads share the work freely
      if (!same_gpu && P->active[victim].load() < P->per_dev) continue;
      const int last = victim + ((J.nchunks - 1 - victim) / ndev) * ndev;
      for (int k = last; k >= 0 && !took; k -= ndev)
        if (try_claim(k)) { took = true; if (!one(k, true)) return; }
    }
    if (!took) std::this_thread::sleep_for(std::chrono::microseconds(200));
  }
}

void worker_main(mag_partition* P, int wi) {
  mag_partition::Worker& w = P->workers[wi];
  long seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(P->mu);
      P->cv_start.wait(lk, [&] { return P->shutdown || P->generation != seen; });
      if (P->shutdown) return;
      seen = P->generation;
    }
    run_job(P, w);
    {
      std::lock_guard<std::mutex> lk(P->mu);
      if (--P->running == 0) P->cv_done.notify_all();
    }
  }
}

int run_partition(mag_partition* P, int32_t* chunks_done, int32_t* chunks_stolen) {
  Job& J = P->job;
  if (J.chunk <= 0) {
    // automatic: one chunk per device, cut into equal pieces of at most 25 000 galaxies.  Measured on one
    // B200 (profiles/r2_bench.md): a chunk must be large for its own work to cover the latency of its
    // longest histories (12 500 galaxies: 9.9 k galaxies/s; 2 x 6 250: 8.2 k), and beyond ~25 000 the
    // cost-ordered queue of one solve runs its large-grid histories all at once and loses 10 %
    // (4 x 25 000: 11.5 k galaxies/s, 1 x 100 000: 10.7 k)
    long long c = ((long long)J.ngal + P->ndev - 1) / P->ndev;
    if (c > 25000) { const long long pieces = (c + 24999) / 25000; c = (c + pieces - 1) / pieces; }
    if (c < 256) c = 256;
    J.chunk = (int32_t)c;
  }
  J.nchunks = (J.ngal + J.chunk - 1) / J.chunk;
  J.claimed.reset(new std::atomic<int>[J.nchunks]);
  for (int k = 0; k < J.nchunks; k++) J.claimed[k].store(0);
  P->unclaimed.store(J.nchunks);
  P->rc.store(MAG_OK);
  for (int d = 0; d < P->ndev; d++) { P->done[d].store(0); P->stolen[d].store(0); P->active[d].store(0); P->launches[d].store(0); P->busy_us[d].store(0); }
  {
    std::lock_guard<std::mutex> lk(P->mu);
    P->err.clear();
    P->running = (int)P->workers.size();
    P->generation++;
  }
  P->cv_start.notify_all();
  {
    std::unique_lock<std::mutex> lk(P->mu);
    P->cv_done.wait(lk, [&] { return P->running == 0; });
  }
  for (int d = 0; d < P->ndev; d++) {
    if (chunks_done) chunks_done[d] = P->done[d].load();
    if (chunks_stolen) chunks_stolen[d] = P->stolen[d].load();
  }
  if (P->rc.load() != MAG_OK) g_part_error = P->err;
  return P->rc.load();
}

}  // namespace

extern "C" {

const char* mag_partition_last_error(void) { return g_part_error.c_str(); }

int mag_partition_create(const mag_params_t* params, int32_t ndev, const int32_t* devices, mag_partition_t** part_out) {
  if (!params || ndev < 1 || !devices || !part_out) { g_part_error = "bad argument"; return MAG_ERR_BAD_ARGUMENT; }
  *part_out = nullptr;
  mag_partition* P = new mag_partition(ndev);
  P->params = *params;
  P->ndev = ndev;
  P->devices.assign(devices, devices + ndev);
  if (const char* e = getenv("MAG_PARTITION_THREADS_PER_GPU")) { const int v = atoi(e); if (v >= 1 && v <= 4) P->per_dev = v; }
  P->workers.resize((size_t)ndev * P->per_dev);
  // contexts are created here, on the calling thread, one after the other (nothing concurrent in mag_create)
  for (size_t i = 0; i < P->workers.size(); i++) {
    P->workers[i].d = (int)(i % ndev);
    const int rc = mag_create(params, devices[P->workers[i].d], &P->workers[i].ctx);
    if (rc != MAG_OK) {
      g_part_error = mag_last_error();
      for (auto& w : P->workers) mag_destroy(w.ctx);
      delete P;
      return rc;
    }
  }
  for (size_t i = 0; i < P->workers.size(); i++) P->workers[i].th = std::thread(worker_main, P, (int)i);
  *part_out = P;
  return MAG_OK;
}

void mag_partition_destroy(mag_partition_t* P) {
  if (!P) return;
  {
    std::lock_guard<std::mutex> lk(P->mu);
    P->shutdown = true;
  }
  P->cv_start.notify_all();
  for (auto& w : P->workers) if (w.th.joinable()) w.th.join();
  for (auto& w : P->workers) mag_destroy(w.ctx);
  delete P;
}

int mag_partition_solve(mag_partition_t* P, int32_t ngal, const int32_t* gal_ids, int32_t nz, const double* t_gyr,
                        const double* inputs, int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q,
                        const int32_t* scalar_q, const mag_outputs_t* out, int32_t chunk, int32_t* chunks_done,
                        int32_t* chunks_stolen) {
  if (!P || ngal < 0 || !out || !gal_ids || !t_gyr || !inputs) { g_part_error = "bad argument"; return MAG_ERR_BAD_ARGUMENT; }
  if (ngal == 0) return MAG_OK;
  Job& J = P->job;
  J = Job();
  J.device_mode = false;
  J.ngal = ngal; J.nz = nz; J.chunk = chunk;
  J.gal_ids = gal_ids; J.t_gyr = t_gyr; J.inputs = inputs; J.out = out;
  J.n_profile_q = n_profile_q; J.n_scalar_q = n_scalar_q; J.profile_q = profile_q; J.scalar_q = scalar_q;
  return run_partition(P, chunks_done, chunks_stolen);
}

int mag_partition_solve_device(mag_partition_t* P, int32_t ngal, const int32_t* const* d_gal_ids, int32_t nz,
                               const double* const* d_t_gyr, const double* const* d_inputs, int32_t n_profile_q,
                               const int32_t* profile_q, int32_t n_scalar_q, const int32_t* scalar_q,
                               const mag_outputs_t* d_out, int32_t chunk, int32_t* chunks_done, int32_t* chunks_stolen,
                               int32_t* owner) {
  if (!P || ngal < 0 || !d_out || !d_gal_ids || !d_t_gyr || !d_inputs) { g_part_error = "bad argument"; return MAG_ERR_BAD_ARGUMENT; }
  if (ngal == 0) return MAG_OK;
  Job& J = P->job;
  J = Job();
  J.device_mode = true;
  J.ngal = ngal; J.nz = nz; J.chunk = chunk;
  J.d_gal_ids = d_gal_ids; J.d_t_gyr = d_t_gyr; J.d_inputs = d_inputs; J.d_out = d_out; J.owner = owner;
  J.n_profile_q = n_profile_q; J.n_scalar_q = n_scalar_q; J.profile_q = profile_q; J.scalar_q = scalar_q;
  return run_partition(P, chunks_done, chunks_stolen);
}

int mag_partition_stats(mag_partition_t* P, int32_t* chunk_used, int64_t* launches, double* busy_ms) {
  if (!P) return MAG_ERR_BAD_ARGUMENT;
  if (chunk_used) *chunk_used = P->job.chunk;
  for (int d = 0; d < P->ndev; d++) {
    if (launches) launches[d] = P->launches[d].load();
    if (busy_ms) busy_ms[d] = (double)P->busy_us[d].load() * 1e-3;
  }
  return MAG_OK;
}

int mag_solve_partitioned(const mag_params_t* params, int32_t ndev, const int32_t* devices, int32_t ngal,
                          const int32_t* gal_ids, int32_t nz, const double* t_gyr, const double* inputs,
                          int32_t n_profile_q, const int32_t* profile_q, int32_t n_scalar_q, const int32_t* scalar_q,
                          const mag_outputs_t* out, int32_t chunk, int32_t* chunks_done, int32_t* chunks_stolen) {
  mag_partition_t* P = nullptr;
  int rc = mag_partition_create(params, ndev, devices, &P);
  if (rc != MAG_OK) return rc;
  rc = mag_partition_solve(P, ngal, gal_ids, nz, t_gyr, inputs, n_profile_q, profile_q, n_scalar_q, scalar_q, out, chunk,
                           chunks_done, chunks_stolen);
  const std::string msg = g_part_error;
  mag_partition_destroy(P);
  g_part_error = msg;
  return rc;
}

}  // extern "C"
