// conv_pair.cu — 3x3 convolution 128 -> 128 channels on CTA PAIRS (tcgen05 cta_group::2), sm_100a.
//
// The inner layers of the C = 256 nets.  conv_ksplit.cu runs them as two groups of 64 output channels: two N = 64
// MMAs of 48 cycles (shared-memory operand bound) per K step and tile.  Here two CTAs of a cluster work on two
// neighbouring M tiles at once: every CTA stages its OWN tile's half-slabs and HALF of the weights (the 64 output
// channels of "its" group, 147 KB — exactly the per-group blob conv_ksplit uses), and the pair's leader issues
// M = 256, N = 128 pair-MMAs: each SM multiplies its 128 rows by all 128 output channels in 64 cycles per K step
// (tensor bound, tools/umma2_bench.cu) instead of 2 x 48.
//
// Per CTA (352 threads, persistent over tile pairs):
//   warp 0   : producer — weight blob once, then per tile two half-slabs (64 input channels each) into a 3-stage ring
//   warp 1   : rank 0: MMA issue for the pair (bias, 9 taps x 4 K-steps per half), commits multicast to both CTAs
//              rank 1: relay — forwards "my half-slab has landed" / "my weights have landed" to the leader's barriers
//   warps 2-9: epilogue of the CTA's own tile (all 128 output channels; residual added here from global memory);
//              accumulator hand-back arrives on the LEADER's barrier (16 warps of the pair)
// Accumulation order: channels 0-63 of all taps, then 64-127 (as conv_ksplit); residual in fp32 in the epilogue.
#include "conv_cfg.cuh"

namespace maru {

constexpr int CP_THREADS = 320;
constexpr int CP_TAPS = 9;
constexpr int CP_LEAD = 24, CP_SLAB_ROWS = TILE_M + 2 * CP_LEAD;      // 176
constexpr int CP_HALF_PLANES = 8;                                     // a half-item is 64 input channels
constexpr int CP_HALF_BYTES = CP_HALF_PLANES * CP_SLAB_ROWS * 16;     // 22 528
constexpr int CP_ONES_BYTES = 2 * TILE_M * 16;
constexpr int CP_BAR_BYTES = 256;
// <CIN, N>: <128, 128> the inner layers of the C = 256 nets (two half-items per tile, 147 KB of taps per CTA, 3 stages);
//           <64, 64>   the inner layers of the C = 128 nets (one item per tile, 37 KB of taps per CTA, 4 stages)
template <int CIN, int N>
struct PairCfg {
  static constexpr int NH = N / 2;                                    // B rows (output channels) held per CTA
  static constexpr int NHALF = CIN / 64;
  static constexpr int KCH = CIN / 8;
  static constexpr int W_BYTES = CP_TAPS * CIN * NH * 2;
  static constexpr int BIASB_BYTES = 2 * NH * 16;
  static constexpr int STAGES = (CIN == 128) ? 3 : 4;
  static constexpr int SMEM = W_BYTES + BIASB_BYTES + CP_ONES_BYTES + CP_BAR_BYTES + STAGES * CP_HALF_BYTES;
  static_assert(SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t cp_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cp_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cp_mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void cp_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cp_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > MARU_SPIN_LIMIT) {
      printf("maru: pair-kernel barrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void cp_tmem_alloc2(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void cp_tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int ACCUMULATE>
__device__ __forceinline__ void cp_umma2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACCUMULATE)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all prior pair-MMAs have completed
__device__ __forceinline__ void cp_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int CIN, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CP_THREADS, 1) conv_pair_kernel(const ConvParams p) {
  using Cfg = PairCfg<CIN, N>;
  constexpr int CP_W_BYTES = Cfg::W_BYTES, CP_BIASB_BYTES = Cfg::BIASB_BYTES, CP_NH = Cfg::NH, CP_N = N, CP_KCH = Cfg::KCH;
  constexpr int NHALF = Cfg::NHALF;
  constexpr uint32_t CP_STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_biasb = s_w + CP_W_BYTES;                    // [W | bias image] is one blob
  uint8_t* s_ones = s_biasb + CP_BIASB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + CP_ONES_BYTES);
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(bars) + CP_BAR_BYTES;
  uint64_t* full = bars;            // [4] this CTA's half-slab landed
  uint64_t* pfull = bars + 4;       // [4] leader only: the peer's half-slab landed (relay arrives)
  uint64_t* empty = bars + 8;       // [4] half-slab consumed (pair-MMA commit, multicast to both CTAs)
  uint64_t* tfull = bars + 12;      // [2] accumulator ready (commit, multicast)
  uint64_t* tempty = bars + 14;     // [2] leader only: accumulator drained by the 16 epilogue warps of the pair
  uint64_t* wbar = bars + 16;       // this CTA's weights + bias operand landed
  uint64_t* pwbar = bars + 17;      // leader only: the peer's weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cp_cluster_rank();
  const int m_rows = (p.n_dev != nullptr) ? total_rows_g(*p.n_dev, p.pos_rows) : p.m_rows;
  const int n_tiles = (m_rows + TILE_M - 1) / TILE_M;
  const int n_pairs = (n_tiles + 1) / 2;
  const int n_clusters = (int)gridDim.x / 2;
  const int cluster = (int)blockIdx.x / 2;

  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    fence_mbar_init();
    mbar_expect_tx(wbar, CP_W_BYTES + CP_BIASB_BYTES);
    // this CTA's half of B: the 64 output channels of group `rank`
    bulk_load_1d(s_w, p.blob + (size_t)rank * (CP_W_BYTES + CP_BIASB_BYTES), CP_W_BYTES + CP_BIASB_BYTES, wbar);
    for (int i = 0; i < (int)CP_STAGES; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&pfull[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);
    }
    mbar_init(pwbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    cp_tmem_alloc2(tmem_slot, 2 * CP_N);     // two accumulators of N columns (in each CTA of the pair)
    cp_tmem_relinquish2();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;                               // ones: (1, 1, 0, ...) in chunk 0
    const uint32_t w0 = (t < TILE_M) ? 0x3F803F80u : 0u;
    *reinterpret_cast<uint4*>(s_ones + t * 16) = make_uint4(w0, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cp_cluster_sync();                          // both CTAs' barriers are initialised before anybody arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // the tile of this CTA in pair q: 2q + rank; the odd tile of the last pair may not exist (it is loaded from the last
  // real tile and never stored)
  if (warp == 0) {
    // =========================== producer ===========================
    pdl_wait();
    uint32_t h = 0;
    for (int q = cluster; q < n_pairs; q += n_clusters) {
      int tile = 2 * q + (int)rank;
      if (tile >= n_tiles) tile = n_tiles - 1;
      const size_t row0 = (size_t)GUARD_ROWS + (size_t)tile * TILE_M;
#pragma unroll 1
      for (int half = 0; half < NHALF; half++, h++) {
        const uint32_t s = h % CP_STAGES;
        cp_wait_cluster(&empty[s], ((h / CP_STAGES) & 1u) ^ 1u);   // arrived by the leader's multicast commit
        if (elect_one()) {
          mbar_expect_tx(&full[s], CP_HALF_BYTES);
          uint8_t* dst = s_stage + s * CP_HALF_BYTES;
#pragma unroll 1
          for (int c = 0; c < CP_HALF_PLANES; c++) {
            const __nv_bfloat16* src =
                p.in + ((size_t)(p.in_chunk0 + half * CP_HALF_PLANES + c) * p.rows_alloc + row0 - CP_LEAD) * 8;
            bulk_load_1d(dst + c * CP_SLAB_ROWS * 16, src, CP_SLAB_ROWS * 16, &full[s]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && rank == 1) {
    // =========================== relay (peer CTA) ===========================
    mbar_wait(wbar, 0);
    if (lane == 0) cp_arrive_cluster(cp_mapa(smem_u32(pwbar), 0));
    __syncwarp();
    uint32_t h = 0;
    for (int q = cluster; q < n_pairs; q += n_clusters) {
#pragma unroll 1
      for (int half = 0; half < NHALF; half++, h++) {
        const uint32_t s = h % CP_STAGES;
        mbar_wait(&full[s], (h / CP_STAGES) & 1u);
        if (lane == 0) cp_arrive_cluster(cp_mapa(smem_u32(&pfull[s]), 0));
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issue (leader CTA) ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(2 * TILE_M, N, 0, 0);
    const uint64_t a_d = umma_desc(0, CP_SLAB_ROWS * 16, 128);
    const uint64_t r_d = umma_desc(0, TILE_M * 16, 128);
    const uint64_t b_d = umma_desc(0, CP_NH * 16, 128);            // each CTA holds N/2 = 64 rows of B per K chunk
    const uint32_t a_hi = (uint32_t)(a_d >> 32), r_hi = (uint32_t)(r_d >> 32), b_hi = (uint32_t)(b_d >> 32);
    const uint32_t w_lo = (uint32_t)b_d + (smem_u32(s_w) >> 4);
    const uint32_t ones_lo = (uint32_t)r_d + (smem_u32(s_ones) >> 4);
    const uint32_t biasb_lo = (uint32_t)b_d + (smem_u32(s_biasb) >> 4);
    mbar_wait(wbar, 0);
    cp_wait_cluster(pwbar, 0);
    uint32_t h = 0;
    int it = 0;
    for (int q = cluster; q < n_pairs; q += n_clusters, it++) {
      const int acc = it & 1;
      const uint32_t d_tmem = tmem_base + acc * CP_N;
      cp_wait_cluster(&tempty[acc], (uint32_t)((it >> 1) & 1) ^ 1u);
#pragma unroll 1
      for (int half = 0; half < NHALF; half++, h++) {
        const uint32_t s = h % CP_STAGES;
        mbar_wait(&full[s], (h / CP_STAGES) & 1u);
        cp_wait_cluster(&pfull[s], (h / CP_STAGES) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (uint32_t)a_d + (smem_u32(s_stage + s * CP_HALF_BYTES) >> 4);
          if (half == 0) cp_umma2<0>(d_tmem, ones_lo, r_hi, biasb_lo, b_hi, idesc);           // D = bias
          const uint32_t wh = w_lo + (uint32_t)half * CP_HALF_PLANES * CP_NH;                  // this half's K chunks
#pragma unroll
          for (int t = 0; t < CP_TAPS; t++) {
            const int off = (t / 3 - 1) * p.pitch + (t % 3 - 1);
#pragma unroll
            for (int j = 0; j < CP_HALF_PLANES / 2; j++) {
              const uint32_t a16 = (2 * j) * CP_SLAB_ROWS + CP_LEAD + off;                     // in 16-byte units
              const uint32_t b16 = (t * CP_KCH + 2 * j) * CP_NH;
              cp_umma2<1>(d_tmem, a_lo + a16, a_hi, wh + b16, b_hi, idesc);
            }
          }
          cp_commit_both(&empty[s]);
          if (half == NHALF - 1) cp_commit_both(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (8 warps, this CTA's tile, all 128 output channels) ===========================
    constexpr int CW = CP_N / 2;                           // 64 columns per warp
    const int lg = warp & 3;
    const int hcol = (warp - 2) >> 2;
    const int trow = lg * 32 + lane;
    const bool has_res = p.res != nullptr;
    const uint32_t tempty_leader = cp_mapa(smem_u32(tempty), 0);
    pdl_wait();                                            // the mask and the residual are read from global memory
    int it = 0;
    for (int q = cluster; q < n_pairs; q += n_clusters, it++) {
      const int acc = it & 1;
      const int tile = 2 * q + (int)rank;
      const int row = tile * TILE_M + trow;
      const bool valid = tile < n_tiles && row < m_rows;
      const size_t grow = (size_t)GUARD_ROWS + row;
      const bool on = valid && (p.mask[row] != 0);
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * CP_N + hcol * CW;
      const int chunk0 = p.out_chunk0 + hcol * (CW / 8);
      const size_t pstride = p.out_blocked ? (size_t)TILE_M * 8 : (size_t)p.rows_alloc * 8;
      __nv_bfloat16* out_row = p.out_blocked
                                   ? p.out + (((size_t)tile * p.out_planes + chunk0) * TILE_M + trow) * 8
                                   : p.out + ((size_t)chunk0 * p.rows_alloc + grow) * 8;
      uint4 rv[CW / 8];
      if (has_res) {
        const size_t rstride = p.res_blocked ? (size_t)TILE_M * 8 : (size_t)p.rows_alloc * 8;
        const int rc0 = p.res_chunk0 + hcol * (CW / 8);
        const __nv_bfloat16* res_row = p.res_blocked
                                           ? p.res + (((size_t)tile * p.res_planes + rc0) * TILE_M + trow) * 8
                                           : p.res + ((size_t)rc0 * p.rows_alloc + grow) * 8;
#pragma unroll
        for (int k = 0; k < CW / 8; k++)
          rv[k] = on ? *reinterpret_cast<const uint4*>(res_row + (size_t)k * rstride) : make_uint4(0u, 0u, 0u, 0u);
      }
      cp_wait_cluster(&tfull[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      uint32_t v[CW];
      tmem_ld32(t_addr, v);
      if constexpr (CW == 64) tmem_ld32(t_addr + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) cp_arrive_cluster(tempty_leader + 8 * acc);
#pragma unroll
      for (int k = 0; k < CW / 8; k++) {
        uint32_t o[4];
        uint32_t rw[4] = {0u, 0u, 0u, 0u};
        if (has_res) {
          rw[0] = rv[k].x;
          rw[1] = rv[k].y;
          rw[2] = rv[k].z;
          rw[3] = rv[k].w;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float lo = __uint_as_float(v[k * 8 + 2 * e]), hi = __uint_as_float(v[k * 8 + 2 * e + 1]);
          if (has_res) {
            lo += bf16_lo(rw[e]);
            hi += bf16_hi(rw[e]);
          }
          uint32_t pk = pack_bf16x2(lo, hi);
          if (p.relu) pk = relu_bf16x2(pk);
          o[e] = on ? pk : 0u;
        }
        if (valid) *reinterpret_cast<uint4*>(out_row + (size_t)k * pstride) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cp_cluster_sync();                          // nobody frees TMEM or exits while the peer may still signal / be read
  if (warp == 1) {
    tc_fence_after();
    cp_tmem_dealloc2(tmem_base, 2 * CP_N);
  }
}

template <int CIN, int N>
static cudaError_t configure_pair_cfg() {
  cudaError_t e = cudaFuncSetAttribute(conv_pair_kernel<CIN, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       PairCfg<CIN, N>::SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_pair_kernel<CIN, N>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}
cudaError_t configure_conv_pair() {
  cudaError_t e = configure_pair_cfg<128, 128>();
#ifdef MARU_EXPERIMENTS
  if (e != cudaSuccess) return e;
  e = configure_pair_cfg<64, 64>();
#endif
  return e;
}

bool conv_pair_supports(int cin, int cout, int taps) {
#ifdef MARU_EXPERIMENTS
  if (taps == CP_TAPS && cin == 64 && cout == 64) return true;
#endif
  return taps == CP_TAPS && cin == 128 && cout == 128;
}
int conv_pair_blob_nt(int cout) { return cout / 2; }      // output channels per CTA of the pair = blob group width

// p.blob: [2 groups of cout/2 output channels][W image | bias operand image]
cudaError_t launch_conv_pair(ConvParams p, int cin, int sm_count, cudaStream_t stream) {
  if (p.blob == nullptr || p.in_blocked) return cudaErrorInvalidValue;
  const int n_tiles = (p.m_rows + TILE_M - 1) / TILE_M;
  const int n_pairs = (n_tiles + 1) / 2;
  int clusters = sm_count / 2;
  if (clusters > n_pairs) clusters = n_pairs;
  if (clusters < 1) clusters = 1;
  if (cin == 128 && p.cout_total == 128)
    return launch_kernel(conv_pair_kernel<128, 128>, dim3(2 * clusters), dim3(CP_THREADS), (size_t)PairCfg<128, 128>::SMEM,
                         stream, p.pdl != 0, p);
#ifdef MARU_EXPERIMENTS   // the 64 -> 64 instantiation is correct but slower than conv_gemm<64,64,9> at every batch size
  if (cin == 64 && p.cout_total == 64)
    return launch_kernel(conv_pair_kernel<64, 64>, dim3(2 * clusters), dim3(CP_THREADS), (size_t)PairCfg<64, 64>::SMEM,
                         stream, p.pdl != 0, p);
#endif
  return cudaErrorInvalidValue;
}

}  // namespace maru
