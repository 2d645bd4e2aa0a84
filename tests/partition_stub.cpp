// partition_stub.cpp -- TEST INFRASTRUCTURE: a CPU fake of the single-GPU entry points of
// include/magnetizer_b200.h, linked with the real csrc/mag_partition.cu so that the multi-GPU
// partition logic (chunking, static share, stealing, in-place gather, resident form, error
// propagation) can be exercised without a GPU (tests/test_partition_host.py).  It solves
// nothing: every output element is a function of the galaxy's id and inputs, and "device"
// pointers are host pointers.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "../include/magnetizer_b200.h"

struct mag_ctx { int device; int launches; };
static thread_local std::string g_err;

static void fill(int device, int32_t ngal_total, int32_t g0, int32_t n, const int32_t* ids, int32_t nz, const double* inputs,
                 int32_t nq, const int32_t* pq, int32_t ns, const int32_t* sq, const mag_outputs_t* out, int nr) {
  for (int g = g0; g < g0 + n; g++) {
    for (int q = 0; q < nq && out->profiles; q++)
      for (int iz = 0; iz < nz; iz++)
        for (int i = 0; i < nr; i++)
          out->profiles[(((size_t)q * ngal_total + g) * nz + iz) * nr + i] = ids[g] * 1000.0 + pq[q] * 10.0 + inputs[(size_t)g * nz + iz] + i * 1e-3;
    for (int q = 0; q < ns && out->scalars; q++)
      for (int iz = 0; iz < nz; iz++) out->scalars[((size_t)q * ngal_total + g) * nz + iz] = ids[g] + sq[q] * 0.5 + iz * 0.01;
    if (out->status) for (int iz = 0; iz < nz; iz++) out->status[(size_t)g * nz + iz] = 'M';
    if (out->nsteps) out->nsteps[g] = ids[g];
    if (out->error) out->error[g] = 0;
    if (out->point_stages) out->point_stages[g] = (int64_t)ids[g] * 10;
    if (out->flags) out->flags[g] = (uint32_t)device;
  }
  std::this_thread::sleep_for(std::chrono::milliseconds(2));
}

extern "C" {
const char* mag_last_error(void) { return g_err.c_str(); }
void mag_params_default(mag_params_t* p) { std::memset(p, 0, sizeof(*p)); p->abi_version = MAG_ABI_VERSION; p->p_nx_ref = 101; }
int mag_create(const mag_params_t* params, int device, mag_ctx_t** out) {
  if (device >= 100) { g_err = "stub: no such device"; return MAG_ERR_NO_DEVICE; }
  (void)params;
  *out = new mag_ctx{device, 0};
  return MAG_OK;
}
void mag_destroy(mag_ctx_t* c) { delete c; }
int mag_solve_range_async(mag_ctx_t* c, int32_t ngal_total, int32_t g0, int32_t n, const int32_t* ids, int32_t nz, const double* t,
                          const double* inputs, int32_t nq, const int32_t* pq, int32_t ns, const int32_t* sq, const mag_outputs_t* out) {
  (void)t;
  if (g0 < 0 || n < 0 || g0 + n > ngal_total) { g_err = "stub: bad range"; return MAG_ERR_BAD_ARGUMENT; }
  if (ids[g0] < 0) { g_err = "stub: poisoned galaxy"; return MAG_ERR_CUDA; }   // lets a test inject a failure
  fill(c->device, ngal_total, g0, n, ids, nz, inputs, nq, pq, ns, sq, out, 101);
  c->launches = 8;
  return MAG_OK;
}
int mag_wait(mag_ctx_t*) { return MAG_OK; }
int mag_solve_range_device(mag_ctx_t* c, int32_t ngal_total, int32_t g0, int32_t n, const int32_t* ids, int32_t nz, const double* t,
                           const double* inputs, int32_t nq, const int32_t* pq, int32_t ns, const int32_t* sq, const mag_outputs_t* out,
                           void*) {
  return mag_solve_range_async(c, ngal_total, g0, n, ids, nz, t, inputs, nq, pq, ns, sq, out);
}
int mag_last_device_status(mag_ctx_t*) { return MAG_OK; }
int mag_last_timing(mag_ctx_t* c, double ms[4]) { ms[0] = 0.1; ms[1] = 1.0; ms[2] = ms[3] = 0; return c->launches; }
double mag_measure_fp64_peak(mag_ctx_t*, int) { return 0.0; }
int mag_set_heavy_policy(mag_ctx_t*, int32_t, double, int32_t) { return MAG_OK; }
void* mag_host_alloc(size_t n) { return std::malloc(n); }
void mag_host_free(void* p) { std::free(p); }
}
