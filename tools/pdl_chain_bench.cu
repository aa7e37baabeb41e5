// Micro-benchmark: steady-state gap between consecutive kernels of a PDL chain shaped like the conv layers
// (148 CTAs x 384 threads, ~190 KB shared memory, ~6 us of work), with and without a TMEM allocation, launched
// from a stream or replayed from a CUDA graph.  Reports, per boundary: first / last CTA entry of kernel k+1 and
// the return of griddepcontrol.wait, relative to the last CTA exit of kernel k.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I maru_b200/csrc -I include tools/pdl_chain_bench.cu -o tools/_pdl_chain_bench
#include <cstdio>
#include <vector>
#include <algorithm>
#include "common.cuh"
using namespace maru;

struct P { long long* stamps; int k; long long work_ns; int tmem; int mma; int ld; uint4* out; int store_kb; const uint4* in; int load_kb; int param_pad[32]; };

template <int VARIANT>
__global__ void __launch_bounds__(384, 1) layer(const P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t mbar;
  pdl_launch_dependents();
  long long* s = p.stamps + ((size_t)p.k * gridDim.x + blockIdx.x) * 4;
  if (threadIdx.x == 0) s[0] = globaltimer_ns();
  if (p.tmem && threadIdx.x < 32) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  if (threadIdx.x == 0) s[1] = globaltimer_ns();
  const long long t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < p.work_ns + (long long)(blockIdx.x % 61) * 15 + VARIANT) {}
  smem[threadIdx.x] = (uint8_t)threadIdx.x;
  // optional burst of global stores right before the CTA exits (the epilogue of a conv tile), and of loads
  if (p.load_kb > 0) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint4* src = p.in + (size_t)blockIdx.x * (p.load_kb * 64);
    for (int i = threadIdx.x; i < p.load_kb * 64; i += 384) { uint4 v = src[i]; acc.x ^= v.x; acc.y ^= v.y; }
    if (acc.x == 0x12345678u) smem[1] = 1;
  }
  if (p.store_kb > 0) {
    uint4* dst = p.out + (size_t)blockIdx.x * (p.store_kb * 64);
    for (int i = threadIdx.x; i < p.store_kb * 64; i += 384) dst[i] = make_uint4(i, p.k, 0, 0);
  }
  if (p.tmem && p.mma > 0) {
    // a conv tile's worth of tensor work on the allocation: MMAs from one elected lane, commit, then tcgen05.ld
    if (threadIdx.x < 32) {
      if (elect_one()) {
        const uint64_t ad = umma_desc(smem_u32(smem), 128 * 16, 128), bd = umma_desc(smem_u32(smem) + 65536, 64 * 16, 128);
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        for (int i = 0; i < p.mma; i++) umma_bf16(slot, ad, bd, idesc, i > 0);
        umma_commit(&mbar);
      }
      __syncwarp();
    }
    mbar_wait(&mbar, 0);
    tc_fence_after();
    if (p.ld && threadIdx.x < 128) {
      uint32_t v[16];
      tmem_ld16(slot + ((uint32_t)(threadIdx.x & ~31) << 16), v);
      tmem_ld_wait();
      if (v[3] == 0x12345u) smem[2] = 1;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) s[3] = globaltimer_ns();
  if (p.tmem && threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(slot, 128); }
  if (threadIdx.x == 0) s[2] = globaltimer_ns();
}

static void report(const char* what, const std::vector<long long>& h, int K, int n) {
  double e_first = 0, e_last = 0, d_min = 0, d_max = 0, period = 0, dealloc = 0; int cnt = 0;
  for (int k = 4; k + 1 < K; k++) {
    long long last_exit = 0, first_exit = 1LL << 62;
    for (int b = 0; b < n; b++) { last_exit = std::max(last_exit, h[((size_t)k * n + b) * 4 + 2]); first_exit = std::min(first_exit, h[((size_t)k * n + b) * 4 + 2]); }
    long long ef = 1LL << 62, el = 0, dmin = 1LL << 62, dmax = 0, nexit = 0;
    for (int b = 0; b < n; b++) {
      const long long* s = &h[((size_t)(k + 1) * n + b) * 4];
      ef = std::min(ef, s[0]); el = std::max(el, s[0]); dmin = std::min(dmin, s[1]); dmax = std::max(dmax, s[1]); nexit = std::max(nexit, s[2]);
    }
    e_first += (ef - last_exit) / 1e3; e_last += (el - last_exit) / 1e3; d_min += (dmin - last_exit) / 1e3; d_max += (dmax - last_exit) / 1e3;
    period += (nexit - last_exit) / 1e3; cnt++;
    for (int b = 0; b < n; b++) { const long long* s = &h[((size_t)k * n + b) * 4]; dealloc += (s[2] - s[3]) / 1e3 / n; }
  }
  printf("%-34s next kernel: first CTA enters %+5.2f us, last %+5.2f us, dependency resolved %+5.2f .. %+5.2f us after the previous last exit; period %.2f us; syncthreads->after dealloc %.2f us\n",
         what, e_first / cnt, e_last / cnt, d_min / cnt, d_max / cnt, period / cnt, dealloc / cnt);
}

int main() {
  const int n = 148, K = 24;
  long long* d; cudaMalloc(&d, (size_t)K * n * 4 * 8);
  uint4 *buf_a, *buf_b; cudaMalloc(&buf_a, (size_t)n * 512 * 1024); cudaMalloc(&buf_b, (size_t)n * 512 * 1024);
  cudaMemset(buf_a, 0, (size_t)n * 512 * 1024); cudaMemset(buf_b, 0, (size_t)n * 512 * 1024);
  cudaStream_t st; cudaStreamCreate(&st);
  std::vector<long long> h((size_t)K * n * 4);
  cudaFuncSetAttribute(layer<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  struct Mode { const char* name; int store_kb, load_kb, mma, ld, tmem; } modes[] = {
      {"no TMEM:", 0, 0, 0, 0, 0}, {"TMEM, no MMA:", 0, 0, 0, 0, 1}, {"TMEM, 1 MMA, no ld:", 0, 0, 1, 0, 1}, {"TMEM, 1 MMA + ld:", 0, 0, 1, 1, 1},
      {"TMEM, 37 MMA + ld:", 0, 0, 37, 1, 1}, {"TMEM, 222 MMA + ld:", 0, 0, 222, 1, 1}, {"TMEM, 222 MMA + ld + 128 KB st:", 128, 0, 222, 1, 1},
      {"no memory traffic:", 0, 0}, {"128 KB stores per CTA at the end:", 128, 0}, {"512 KB stores per CTA at the end:", 512, 0},
      {"128 KB loads + 128 KB stores:", 128, 128}, {"512 KB loads + 128 KB stores:", 128, 512}};
  for (auto& m : modes) {
    cudaGraphExec_t ex = nullptr;
    auto enqueue = [&]() {
      for (int k = 0; k < K; k++) {
        P p{}; p.stamps = d; p.k = k; p.work_ns = 5000; p.tmem = m.name[0]=='n' && m.name[3]=='T' ? 0 : 1; if (m.name[0]=='T') p.tmem = m.tmem; p.mma = m.mma; p.ld = m.ld; p.store_kb = m.store_kb; p.load_kb = m.load_kb;
        p.out = (k & 1) ? buf_a : buf_b; p.in = (k & 1) ? buf_b : buf_a;       // reads what the previous kernel wrote
        launch_kernel(layer<0>, dim3(n), dim3(384), (size_t)190 * 1024, st, true, p);
      }
    };
    cudaGraph_t g; cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal); enqueue(); cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ex, g, 0); cudaGraphDestroy(g);
    for (int rep = 0; rep < 3; rep++) { cudaGraphLaunch(ex, st); cudaStreamSynchronize(st); }
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    report(m.name, h, K, n);
    cudaGraphExecDestroy(ex);
  }
  return 0;
}
