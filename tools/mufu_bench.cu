// Microbenchmark: MUFU.EX2 throughput per SM (independent chains), f32 vs packed bf16x2 / f16x2.
#include <cstdio>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void __launch_bounds__(1024) bench(int iters, long long* out, float* sink) {
  float a[16];
  unsigned u[16];
  for (int e = 0; e < 16; e++) { a[e] = -1.0f - 0.01f * e - threadIdx.x * 1e-4f; u[e] = 0xBC00BC00u + e; }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int e = 0; e < 16; e++) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[e]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[e]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[e]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int e = 0; e < 16; e++) s += a[e] + __uint_as_float(u[e]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int threads, long long* d_out, float* sink) {
  const int iters = 2000;
  bench<MODE><<<148, threads>>>(iters, d_out, sink); cudaDeviceSynchronize();
  bench<MODE><<<148, threads>>>(iters, d_out, sink); cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  const double per_clk = (double)threads * 16 * iters / mx;
  printf("mode=%s threads=%4d  MUFU instr-lanes/clk/SM=%6.2f  (elements/clk/SM=%6.2f)\n",
         MODE == 0 ? "f32   " : MODE == 1 ? "bf16x2" : "f16x2 ", threads, per_clk, MODE == 0 ? per_clk : 2 * per_clk);
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  float* sink; cudaMalloc(&sink, 148 * 1024 * 4);
  for (int t : {128, 256, 512, 1024}) { run<0>(t, d_out, sink); run<1>(t, d_out, sink); run<2>(t, d_out, sink); }
  return 0;
}
