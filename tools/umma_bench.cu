// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, of the number of
// independent TMEM accumulators the issue loop rotates over, and of the smem operand layout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I maru_b200/csrc -I include tools/umma_bench.cu -o /tmp/umma_bench
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace maru;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;    // SBO = 8 rows * 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}

template <int N, int NACC, int LAYOUT>
__global__ void __launch_bounds__(128) bench(int iters, long long* out, int aoff = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {                      // warp-uniform role, one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 24 * 1024;
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (LAYOUT == 0) { da[k] = umma_desc(a0 + aoff * 16 + k * 2 * 2816, 2816, 128); db[k] = umma_desc(b0 + k * 2 * N * 16, N * 16, 128); }
      else { da[k] = desc_sw128(a0 + k * 32); db[k] = desc_sw128(b0 + k * 32); }
    }
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) umma_bf16(tb + (k % NACC) * N, da[k], db[k], idesc, 1);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (out && (threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int NACC, int LAYOUT> void run1(int blocks, long long* d_out) {
  auto k = bench<N, NACC, LAYOUT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  k<<<blocks, 128, 64 * 1024>>>(iters, d_out, 0);
  cudaDeviceSynchronize();
  k<<<blocks, 128, 64 * 1024>>>(iters, d_out, 0);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(blocks);
  cudaMemcpy(h.data(), d_out, blocks * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("N=%3d layout=%s nacc=%d blocks=%3d  cycles/mma=%7.1f  (ideal %d)  %s\n", N, LAYOUT ? "sw128" : "none ",
         NACC, blocks, (double)mx / iters, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
template <int N> void run_off(int aoff, long long* d_out) {
  auto k = bench<N, 1, 0>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  k<<<148, 128, 64 * 1024>>>(iters, d_out, aoff); cudaDeviceSynchronize();
  k<<<148, 128, 64 * 1024>>>(iters, d_out, aoff); cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("N=%3d layout=none A start offset = %d rows (x16 B): cycles/mma=%7.1f\n", N, aoff, (double)mx / iters);
}
template <int N> void run(int blocks, long long* d_out) {
  run1<N, 1, 0>(blocks, d_out); run1<N, 2, 0>(blocks, d_out);
  if (4 * N <= 512) run1<N, (4 * N <= 512 ? 4 : 1), 0>(blocks, d_out);
  run1<N, 1, 1>(blocks, d_out); run1<N, 2, 1>(blocks, d_out);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  for (int aoff : {0, 1, 3, 4, 5, 7, 8}) { run_off<64>(aoff, d_out); }
  for (int aoff : {0, 1, 4}) { run_off<128>(aoff, d_out); }
  for (int blocks : {148}) { run<32>(blocks, d_out); run<64>(blocks, d_out); run<128>(blocks, d_out); run<256>(blocks, d_out); }
  return 0;
}
