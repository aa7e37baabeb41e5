// Micro-benchmark: how long after the LAST CTA of a primary grid has exited does griddepcontrol.wait return in a
// dependent grid whose CTAs are ALREADY RESIDENT (programmatic dependent launch, both grids small enough to be
// co-resident)?  Decides whether conv CTAs that fit two per SM could hide the layer boundary.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I maru_b200/csrc -I include tools/pdl_bench.cu -o tools/_pdl_bench
#include <cstdio>
#include <vector>
#include <algorithm>
#include "common.cuh"
using namespace maru;

__global__ void primary(long long spin_ns, long long* exit_ns, int early_trigger) {
  if (early_trigger) pdl_launch_dependents();
  const long long t0 = globaltimer_ns();
  // CTAs finish staggered: CTA i spins spin_ns + i * 20 ns
  while (globaltimer_ns() - t0 < spin_ns + (long long)blockIdx.x * 20) {}
  __syncthreads();
  if (threadIdx.x == 0) exit_ns[blockIdx.x] = globaltimer_ns();
}
__global__ void dependent(long long* entry_ns, long long* dep_ns, const long long* exit_ns, long long* seen) {
  if (threadIdx.x == 0) entry_ns[blockIdx.x] = globaltimer_ns();
  pdl_wait();
  if (threadIdx.x == 0) {
    dep_ns[blockIdx.x] = globaltimer_ns();
    seen[blockIdx.x] = exit_ns[(blockIdx.x * 7) % gridDim.x];     // visibility check: must be this round's stamp
  }
}

int main() {
  const int n = 148;
  long long *d_exit, *d_entry, *d_dep, *d_seen;
  cudaMalloc(&d_exit, n * 8); cudaMalloc(&d_entry, n * 8); cudaMalloc(&d_dep, n * 8); cudaMalloc(&d_seen, n * 8);
  cudaStream_t st; cudaStreamCreate(&st);
  for (int smem_kb : {16, 100, 200}) {
    for (int early : {1, 0}) {
      cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      cudaFuncSetAttribute(dependent, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      double gap_first = 0, gap_dep_min = 0, gap_dep_max = 0, entry_first = 0; int reps = 20;
      for (int r = 0; r < reps + 3; r++) {
        launch_kernel(primary, dim3(n), dim3(128), (size_t)smem_kb * 1024, st, false, (long long)20000, d_exit, early);
        launch_kernel(dependent, dim3(n), dim3(128), (size_t)smem_kb * 1024, st, true, d_entry, d_dep, (const long long*)d_exit, d_seen);
        cudaStreamSynchronize(st);
        std::vector<long long> ex(n), en(n), de(n);
        cudaMemcpy(ex.data(), d_exit, n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(en.data(), d_entry, n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(de.data(), d_dep, n * 8, cudaMemcpyDeviceToHost);
        if (r < 3) continue;
        const long long last_exit = *std::max_element(ex.begin(), ex.end());
        entry_first += (*std::min_element(en.begin(), en.end()) - last_exit) / 1e3;
        gap_first += (*std::max_element(en.begin(), en.end()) - last_exit) / 1e3;
        gap_dep_min += (*std::min_element(de.begin(), de.end()) - last_exit) / 1e3;
        gap_dep_max += (*std::max_element(de.begin(), de.end()) - last_exit) / 1e3;
      }
      printf("smem %3d KB/CTA, launch_dependents %s: first dependent CTA enters %+6.2f us, last enters %+6.2f us, "
             "griddepcontrol.wait returns %+5.2f .. %+5.2f us after the primary's last CTA exit\n",
             smem_kb, early ? "at kernel start" : "never (implicit)", entry_first / reps, gap_first / reps,
             gap_dep_min / reps, gap_dep_max / reps);
    }
  }
  return 0;
}
