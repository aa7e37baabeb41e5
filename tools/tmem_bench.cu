// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) and its interaction with MUFU.EX2.
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace maru;

// mode 0: LDTM only; 1: MUFU only (32 ex2 per iteration); 2: both interleaved
template <int MODE>
__global__ void __launch_bounds__(256) bench(int iters, int nwarps, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = threadIdx.x * 1e-9f;
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; i++) {
      uint32_t v[32];
      if (MODE != 1) { tmem_ld32(tb + (i & 7) * 32, v); tmem_ld_wait(); }
      else { for (int e = 0; e < 32; e++) v[e] = __float_as_uint(acc - e); }
      if (MODE != 0) {
#pragma unroll
        for (int e = 0; e < 32; e++) acc += ex2_approx(__uint_as_float(v[e]) * 1e-30f - acc);
      } else {
#pragma unroll
        for (int e = 0; e < 32; e++) acc += __uint_as_float(v[e]) * 1e-30f;
      }
    }
  }
  long long t1 = clock64();
  if (sink) sink[blockIdx.x * 256 + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE> void run(int nwarps, long long* d_out, float* sink) {
  const int iters = 2000;
  bench<MODE><<<148, 256>>>(iters, nwarps, d_out, sink);
  cudaDeviceSynchronize();
  bench<MODE><<<148, 256>>>(iters, nwarps, d_out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  const double cyc = (double)mx / iters;
  printf("mode=%s warps=%d  cycles/iter=%7.1f  TMEM B/clk/SM=%6.1f  MUFU/clk/SM=%5.1f %s\n",
         MODE == 0 ? "ldtm " : MODE == 1 ? "mufu " : "both ", nwarps, cyc,
         MODE != 1 ? nwarps * 4096.0 / cyc : 0.0, MODE != 0 ? nwarps * 32 * 32.0 / cyc : 0.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  float* sink; cudaMalloc(&sink, 148 * 256 * 4);
  for (int nw : {1, 4, 8}) { run<0>(nw, d_out, sink); run<1>(nw, d_out, sink); run<2>(nw, d_out, sink); }
  return 0;
}
