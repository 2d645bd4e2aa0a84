// FP64 issue probe (GPU box): how many DFMA per clock one SM sub-partition sustains as a
// function of resident warps, independent chains per warp (ILP) and the share of non-FP64
// instructions interleaved with them.  Guides the register/occupancy trade of k_evolve.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_issue_probe fp64_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OPF /* integer ops per 4 DFMA */>
__global__ void k(double* out, int iters, long long* cyc) {
  extern __shared__ double sm[];
  double a[ILP];
  unsigned x[4] = {threadIdx.x, threadIdx.x + 1u, threadIdx.x + 2u, threadIdx.x + 3u};
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
  const double m = 1.0000001, c = 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        a[i] = fma(a[i], m, c);
        if (OPF > 0 && ((r * ILP + i) % 4) < OPF) { const int q = (r * ILP + i) & 3; x[q] += x[(q + 1) & 3]; }
        if (OPF > 4 && ((r * ILP + i) % 4) < OPF - 4) { const int q = (r * ILP + i + 2) & 3; x[q] ^= x[(q + 1) & 3]; }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(x[0] ^ x[1] ^ x[2] ^ x[3]);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int OPF>
void run(int warps_per_sm, int nsm) {
  const int iters = 4000;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 32 * nsm * warps_per_sm);
  cudaMalloc(&cyc, 8);
  const int smem = (200 * 1024 / warps_per_sm) & ~1023;
  cudaFuncSetAttribute(k<ILP, OPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<ILP, OPF>, 32, smem);
  k<ILP, OPF><<<nsm * warps_per_sm, 32, smem>>>(out, 10, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, OPF><<<nsm * warps_per_sm, 32, smem>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double dfma_per_warp = (double)iters * 8 * ILP;
  const double warps_per_smsp = warps_per_sm / 4.0;
  // DFMA warp-instructions per cycle and sub-partition (peak 0.5)
  printf("ILP %2d  int ops per 4 DFMA %d  warps/SM %2d (occ %d): %.3f DFMA/clk/SMSP (peak 0.5), %.2f cycles per DFMA and warp, %.1f TFLOP/s\n",
         ILP, OPF, warps_per_sm, occ, dfma_per_warp * warps_per_smsp / (double)h, (double)h / dfma_per_warp,
         2.0 * 32 * dfma_per_warp * nsm * warps_per_sm / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, nsm);
  for (int w : {4, 8, 12, 16}) {
    run<1, 0>(w, nsm); run<2, 0>(w, nsm); run<3, 0>(w, nsm); run<4, 0>(w, nsm); run<6, 0>(w, nsm); run<8, 0>(w, nsm); run<12, 0>(w, nsm);
    run<4, 2>(w, nsm); run<4, 3>(w, nsm); run<4, 4>(w, nsm); run<4, 6>(w, nsm); run<8, 2>(w, nsm); run<8, 3>(w, nsm); run<8, 4>(w, nsm);
    run<8, 6>(w, nsm); run<8, 8>(w, nsm); run<12, 3>(w, nsm); run<12, 4>(w, nsm);
  }
  return 0;
}
