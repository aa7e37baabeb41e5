// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (CTA pair, M=256 over two SMs, K=16, bf16,
// SWIZZLE_NONE K-major operands) as a function of N.  In pair mode each CTA supplies its own 128 rows
// of A and HALF of the N rows of B from its own shared memory, so the per-SM operand traffic of an
// N=64 conv tile drops from (128+64)*32 B to (128+32)*32 B per K=16 step.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I maru_b200/csrc -I include tools/umma2_bench.cu -o /tmp/umma2_bench
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace maru;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void umma2_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int N, int NACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) bench2(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_rank();
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc2(&slot, 512); tmem_relinquish2(); }
  fence_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 24 * 1024;
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      da[k] = umma_desc(a0 + k * 2 * 2816, 2816, 128);
      db[k] = umma_desc(b0 + k * 2 * (N / 2) * 16, (N / 2) * 16, 128);   // this CTA's half of B
    }
    long long t0 = clock64();
    if (rank == 0) {
      if (elect_one()) {
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; k++) umma2_bf16(tb + (k % NACC) * N, da[k], db[k], idesc, 1);
        }
        umma2_commit_mc(&bar, 3);
      }
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (out && (threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tb, 512); }
}

template <int N, int NACC> void run2(int blocks, long long* d_out) {
  auto k = bench2<N, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  k<<<blocks, 128, 64 * 1024>>>(iters, d_out);
  cudaDeviceSynchronize();
  k<<<blocks, 128, 64 * 1024>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(blocks);
  cudaMemcpy(h.data(), d_out, blocks * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  // one pair MMA = 256 x N x 16 MACs over 2 SMs = the work of one M=128 MMA per SM
  printf("cta_group::2 M=256 N=%3d nacc=%d ctas=%3d  cycles/mma=%7.1f  (tensor-bound ideal %d; 1-SM M=128 measured max(N/2,(128+N)/4)=%d)  %s\n",
         N, NACC, blocks, (double)mx / iters, 128 * N / 256, (128 * N / 256 > (128 + N) / 4) ? 128 * N / 256 : (128 + N) / 4,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 1024 * 8);
  run2<64, 1>(2, d_out);
  run2<64, 2>(2, d_out);
  run2<64, 2>(148, d_out);
  run2<128, 2>(148, d_out);
  run2<32, 2>(148, d_out);
  run2<256, 2>(148, d_out);
  return 0;
}
