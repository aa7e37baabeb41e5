// conv_chain.cu — persistent multi-layer conv kernel (sm_100a).
//
// Runs a CHAIN of consecutive conv layers (everything between two attention layers of the trunk) in
// ONE kernel: one CTA per SM, all co-resident, a grid-wide barrier between layers.  The layer math is
// the one of conv_gemm.cu (tcgen05 implicit GEMM, bias / residual as extra MMAs, fused ReLU + mask
// epilogue); what changes is everything around it:
//   * no kernel boundary between layers.  tools/timeline.py: with one launch per layer the first
//     dependent load of layer l+1 is issued ~4.6 us after the last CTA of layer l has exited (grid
//     completion + launch + CTA ramp), and a batch-256 layer itself only runs 3-5 us;
//   * TMEM and the mbarriers are set up ONCE.  The barriers are never re-initialised: every role
//     keeps the use-count parity of each pipeline slot / accumulator in registers across layers;
//   * the weight blob of layer l+1 (weights + bias image, ONE bulk copy) is requested as soon as the
//     CTA's last MMA of layer l has retired, i.e. it travels during the last epilogue and the barrier;
//   * shared memory: [weight blob region, max over the chain][barriers][per layer: ones | identity |
//     slab / mask / residual stages].
// Co-residency: the grid is one CTA per SM at > 113 KB of shared memory, launched without the
// cooperative attribute; griddepcontrol.launch_dependents is only issued after the first grid barrier,
// so no later kernel can occupy an SM before every CTA of the chain is resident.
#include <stdio.h>

#include "conv_cfg.cuh"

namespace maru {

constexpr int CHAIN_THREADS = 352;   // warp 0 producer, warps 1 and 10 MMA issue, warps 2..9 epilogue
constexpr int CHAIN_SMEM = 232448;   // 227 KB
constexpr int CHAIN_BAR_BYTES = 256;
constexpr int CHAIN_MAXS = 4;

struct ChainBars {
  uint64_t* full;     // [2][4] slab + mask (+ residual) of an even / odd tile landed: one set per MMA warp, so
                      // that neither warp ever skips a phase of a barrier it waits on (see conv_gemm.cu)
  uint64_t* empty;    // [4] stage reusable (MMA commit + 8 epilogue warps)
  uint64_t* tfull;    // [2] accumulator ready
  uint64_t* tempty;   // [2] accumulator drained (8 epilogue warps)
  uint64_t* wbar;     // weight blob landed
  uint64_t* mdone;    // both MMA warps' last MMAs of the layer retired
};

// per-thread pipeline state that survives layer boundaries (the mbarriers are never re-initialised)
struct ChainState {
  uint32_t slot_par;   // bit s: parity of the number of uses of stage slot s so far (empty[] barriers)
  uint32_t full_par;   // bit 4*(tile parity)+s: parity of the uses of full[tile parity][s] so far
  uint32_t acc_par;    // bit a: parity of the number of uses of accumulator a so far
  uint32_t n_active;   // layers in which this CTA had a role so far (weight blob / mdone phases)
};

__device__ __forceinline__ void chain_grid_sync(unsigned* ctr, unsigned n_ctas) {
  fence_proxy_async_all();   // this thread's global stores will be read through the async proxy (bulk copies)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    uint32_t spins = 0;
    while (ld_acquire_gpu(ctr) < n_ctas) {
      if (++spins > MARU_SPIN_LIMIT) {
        printf("maru: chain grid barrier timed out (block %d)\n", (int)blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// request the weight blob of `L` for this CTA's channel group (one bulk copy)
template <int CIN, int NT, int TAPS>
__device__ __forceinline__ void chain_load_blob(const ChainLayer& L, int group, uint8_t* s_w, uint64_t* wbar) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  constexpr uint32_t BLOB = Cfg::W_BYTES + Cfg::BIASB_BYTES;
  mbar_expect_tx(wbar, BLOB);
  bulk_load_1d(s_w, L.blob + (size_t)group * BLOB, BLOB, wbar);
}

template <int CIN, int NT, int TAPS>
__device__ __forceinline__ void chain_layer(const ChainParams& cp, const ChainLayer& L, uint8_t* smem,
                                            const ChainBars& B, const uint32_t tmem_base, const int bx,
                                            const int gx, const int group, ChainState& st, long long* stamp) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  const bool has_res = L.res != nullptr;
  const int STAGES = L.stages;
  const int STAGE_BYTES = Cfg::SLAB_BYTES + Cfg::MASK_BYTES + (has_res ? Cfg::RES_BYTES : 0);
  uint8_t* s_w = smem;
  uint8_t* s_biasb = s_w + Cfg::W_BYTES;                     // second part of the blob
  uint8_t* s_ones = smem + cp.w_region + CHAIN_BAR_BYTES;
  uint8_t* s_ident = s_ones + Cfg::ONES_BYTES;
  uint8_t* s_stage = s_ident + (has_res ? Cfg::IDENT_BYTES : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_rows = (cp.n_dev != nullptr) ? total_rows(*cp.n_dev) : cp.m_rows;
  const int n_tiles = (m_rows + TILE_M - 1) / TILE_M;
  // bring-up: per-tile %globaltimer stamps of CTA 0 for the layer selected by the host (stamp[3] slot reuse):
  // tr[it][0] producer issue, [1] MMA operands ready, [2] MMAs issued, [3] accumulator ready (epilogue), [4] epilogue done
  long long* tr = (stamp != nullptr && cp.trace != nullptr && stamp == cp.stamps + 4 * cp.trace_layer) ? cp.trace : nullptr;

  if (warp >= 2 && warp <= 9) {
    // constant MMA operands, K-major SWIZZLE_NONE images: [k-chunk][row][8 bf16]
    const int t = threadIdx.x - 64;
    for (int i = t; i < 2 * TILE_M; i += 256) {               // ones: (1, 1, 0, ...) in chunk 0
      const uint32_t w0 = (i < TILE_M) ? 0x3F803F80u : 0u;
      *reinterpret_cast<uint4*>(s_ones + i * 16) = make_uint4(w0, 0u, 0u, 0u);
    }
    if (has_res) {
      for (int i = t; i < (NT / 8) * NT; i += 256) {          // identity: [k-chunk c][row n]
        const int c = i / NT, n = i - c * NT;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if ((n >> 3) == c) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;
        *reinterpret_cast<uint4*>(s_ident + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_async_smem();                                       // generic stores -> visible to UMMA
  }
  __syncthreads();

  if (warp == 0) {
    // =========================== producer ===========================
    uint32_t par = st.slot_par;
    int stage = 0, it = 0;
    for (int tile = bx; tile < n_tiles; tile += gx, it++) {
      mbar_wait(&B.empty[stage], ((par >> stage) & 1u) ^ 1u);
      par ^= 1u << stage;
      uint64_t* fbar = &B.full[(it & 1) * CHAIN_MAXS + stage];
      if (tr != nullptr && lane == 0) tr[((tile - bx) / gx) * 8 + 0] = globaltimer_ns();
      if (elect_one()) {
        mbar_expect_tx(fbar, STAGE_BYTES);
        const size_t row0 = (size_t)GUARD_ROWS + (size_t)tile * TILE_M;
        uint8_t* dst = s_stage + stage * STAGE_BYTES;
        if (TAPS == 1 && L.in_blocked) {
          bulk_load_1d(dst, L.in + (size_t)tile * L.in_planes * (TILE_M * 8), Cfg::SLAB_BYTES, fbar);
        } else {
#pragma unroll 1
          for (int c = 0; c < Cfg::KCH; c++) {
            const __nv_bfloat16* src = L.in + ((size_t)c * cp.rows_alloc + row0 - Cfg::LEAD) * 8;
            bulk_load_1d(dst + c * Cfg::SLAB_ROWS * 16, src, Cfg::SLAB_ROWS * 16, fbar);
          }
        }
        bulk_load_1d(dst + Cfg::SLAB_BYTES, cp.mask + (size_t)tile * TILE_M, Cfg::MASK_BYTES, fbar);
        if (has_res) {
          uint8_t* rdst = dst + Cfg::SLAB_BYTES + Cfg::MASK_BYTES;
          const int rc0 = group * (NT / 8);
          if (L.res_blocked) {
            bulk_load_1d(rdst, L.res + ((size_t)tile * L.res_planes + rc0) * (TILE_M * 8), Cfg::RES_BYTES,
                         fbar);
          } else {
#pragma unroll 1
            for (int c = 0; c < NT / 8; c++)
              bulk_load_1d(rdst + c * TILE_M * 16, L.res + ((size_t)(rc0 + c) * cp.rows_alloc + row0) * 8,
                           TILE_M * 16, fbar);
          }
        }
      }
      __syncwarp();
      if (++stage == STAGES) stage = 0;
    }
  } else if (warp == 1 || warp == 10) {
    // =========================== MMA issuers (alternate tiles / accumulators) ===========================
    const int mw = (warp == 1) ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, NT, 0, 0);
    const uint64_t a_d = umma_desc(0, Cfg::SLAB_ROWS * 16, 128);   // slab:      LBO = slab plane
    const uint64_t r_d = umma_desc(0, TILE_M * 16, 128);           // res, ones: LBO = 128 rows
    const uint64_t b_d = umma_desc(0, NT * 16, 128);               // W, biasB, I: LBO = NT rows
    const uint32_t a_hi = (uint32_t)(a_d >> 32), r_hi = (uint32_t)(r_d >> 32), b_hi = (uint32_t)(b_d >> 32);
    const uint32_t w_lo = (uint32_t)b_d + (smem_u32(s_w) >> 4);
    const uint32_t ones_lo = (uint32_t)r_d + (smem_u32(s_ones) >> 4);
    const uint32_t biasb_lo = (uint32_t)b_d + (smem_u32(s_biasb) >> 4);
    const uint32_t ident_lo = (uint32_t)b_d + (smem_u32(s_ident) >> 4);
    mbar_wait(B.wbar, st.n_active & 1u);                           // this layer's weights + bias image
    if (stamp != nullptr && warp == 1 && lane == 0) stamp[3] = globaltimer_ns();   // weights landed
    uint32_t fpar = st.full_par, apar = st.acc_par;
    int stage = 0, it = 0;
    for (int tile = bx; tile < n_tiles; tile += gx, it++) {
      const int acc = it & 1;
      const int fslot = acc * CHAIN_MAXS + stage;
      const uint32_t fph = (fpar >> fslot) & 1u, eph = ((apar >> acc) & 1u) ^ 1u;
      fpar ^= 1u << fslot;
      apar ^= 1u << acc;
      const int my_stage = stage;
      if (++stage == STAGES) stage = 0;
      if (acc != mw) continue;
      mbar_wait(&B.tempty[acc], eph);
      mbar_wait(&B.full[mw * CHAIN_MAXS + my_stage], fph);
      if (stamp != nullptr && it == 0 && lane == 0) stamp[1] = globaltimer_ns();   // first operands landed
      if (tr != nullptr && lane == 0) tr[it * 8 + 1] = globaltimer_ns();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st_addr = smem_u32(s_stage + my_stage * STAGE_BYTES);
        const uint32_t a_lo = (uint32_t)a_d + (st_addr >> 4);
        const uint32_t res_lo = (uint32_t)r_d + ((st_addr + Cfg::SLAB_BYTES + Cfg::MASK_BYTES) >> 4);
        const uint32_t d_tmem = tmem_base + acc * NT;
        umma_bf16_lohi<0>(d_tmem, ones_lo, r_hi, biasb_lo, b_hi, idesc);           // D = bias
        if (has_res) {
#pragma unroll
          for (int j = 0; j < NT / 16; j++)                                         // D += residual
            umma_bf16_lohi<1>(d_tmem, res_lo + 2 * j * TILE_M, r_hi, ident_lo + 2 * j * NT, b_hi, idesc);
        }
#pragma unroll
        for (int t = 0; t < TAPS; t++) {
          const int off = (TAPS == 9) ? ((t / 3 - 1) * PITCH + (t % 3 - 1)) : 0;
#pragma unroll
          for (int j = 0; j < CIN / 16; j++) {
            const uint32_t a16 = (2 * j) * Cfg::SLAB_ROWS + Cfg::LEAD + off;   // in 16-byte units
            const uint32_t b16 = (t * Cfg::KCH + 2 * j) * NT;
            umma_bf16_lohi<1>(d_tmem, a_lo + a16, a_hi, w_lo + b16, b_hi, idesc);
          }
        }
        umma_commit(&B.empty[my_stage]);
        umma_commit(&B.tfull[acc]);
      }
      __syncwarp();
      if (tr != nullptr && lane == 0) tr[it * 8 + 2] = globaltimer_ns();
    }
    // both issuers' last MMAs retired => the weight region may be overwritten
    if (elect_one()) umma_commit(B.mdone);
    __syncwarp();
  } else {
    // =========================== epilogue (8 warps) ===========================
    constexpr int CW = NT / 2;               // columns per warp
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;        // which half of the NT columns
    uint32_t fpar = st.full_par, apar = st.acc_par;
    int stage = 0, it = 0;
    for (int tile = bx; tile < n_tiles; tile += gx, it++) {
      const int acc = it & 1;
      const int fslot = acc * CHAIN_MAXS + stage;
      const uint32_t fph = (fpar >> fslot) & 1u, tph = (apar >> acc) & 1u;
      fpar ^= 1u << fslot;
      apar ^= 1u << acc;
      const int trow = lg * 32 + lane;
      const int row = tile * TILE_M + trow;
      const bool valid = row < m_rows;
      const size_t grow = (size_t)GUARD_ROWS + row;
      const uint8_t* s_mask = s_stage + stage * STAGE_BYTES + Cfg::SLAB_BYTES;
      mbar_wait(&B.full[fslot], fph);                        // mask landed with the slab
      const bool on = valid && (s_mask[trow] != 0);
      mbar_wait(&B.tfull[acc], tph);
      if (tr != nullptr && warp == 2 && lane == 0) tr[it * 8 + 3] = globaltimer_ns();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * NT + half * CW;
      const int chunk0 = group * (NT / 8) + half * (CW / 8);
      const size_t pstride = L.out_blocked ? (size_t)TILE_M * 8 : (size_t)cp.rows_alloc * 8;
      __nv_bfloat16* out_row = L.out_blocked
                                   ? L.out + (((size_t)tile * L.out_planes + chunk0) * TILE_M + trow) * 8
                                   : L.out + ((size_t)chunk0 * cp.rows_alloc + grow) * 8;
      if constexpr (CW == 16) {
        uint32_t v[16];
        tmem_ld16(t_addr, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; q++) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            uint32_t pk = pack_bf16x2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
            if (L.relu) pk = relu_bf16x2(pk);
            o[e] = on ? pk : 0u;
          }
          if (valid) *reinterpret_cast<uint4*>(out_row + (size_t)q * pstride) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < CW; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; q++) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              uint32_t pk = pack_bf16x2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
              if (L.relu) pk = relu_bf16x2(pk);
              o[e] = on ? pk : 0u;
            }
            if (valid)
              *reinterpret_cast<uint4*>(out_row + (size_t)(c0 / 8 + q) * pstride) =
                  make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&B.tempty[acc]);
        mbar_arrive(&B.empty[stage]);
      }
      if (tr != nullptr && warp == 2 && lane == 0) tr[it * 8 + 4] = globaltimer_ns();
      if (++stage == STAGES) stage = 0;
    }
  }

  // every role advances the shared use-count parities by the same tile sequence
  {
    int stage = 0, it = 0;
    for (int tile = bx; tile < n_tiles; tile += gx, it++) {
      st.slot_par ^= 1u << stage;
      st.full_par ^= 1u << ((it & 1) * CHAIN_MAXS + stage);
      st.acc_par ^= 1u << (it & 1);
      if (++stage == STAGES) stage = 0;
    }
  }
}

// (cin, channels per tile, taps) the chain can run: the weight blob must leave room for the stages
#define MARU_CHAIN_CASES(X)                                                          \
  X(64, 64, 9) X(32, 32, 9)                                                          \
  X(256, 128, 1) X(256, 32, 1) X(128, 128, 1) X(128, 64, 1) X(128, 32, 1)            \
  X(64, 128, 1) X(64, 64, 1) X(64, 32, 1) X(32, 64, 1) X(32, 32, 1)

__global__ void __launch_bounds__(CHAIN_THREADS, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cp.w_region);
  ChainBars B;
  B.full = bars;
  B.empty = bars + 2 * CHAIN_MAXS;
  B.tfull = bars + 3 * CHAIN_MAXS;
  B.tempty = B.tfull + 2;
  B.wbar = B.tempty + 2;
  B.mdone = B.wbar + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 31);   // last slot of the 256-byte barrier block

  const int warp = threadIdx.x >> 5;
  const int n_ctas = (int)gridDim.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < CHAIN_MAXS; i++) {
      mbar_init(&B.full[i], 1);
      mbar_init(&B.full[CHAIN_MAXS + i], 1);
      mbar_init(&B.empty[i], 9);     // MMA commit + 8 epilogue warps
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&B.tfull[i], 1);
      mbar_init(&B.tempty[i], 8);
    }
    mbar_init(B.wbar, 1);
    mbar_init(B.mdone, 2);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem, 256);         // two accumulators of up to 128 columns for the whole chain
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  auto geometry = [&](int l, int& groups, int& gx, int& group, int& bx) {
    groups = cp.layer[l].cout_total / cp.layer[l].nt;
    gx = n_ctas / groups;
    group = (int)blockIdx.x % groups;
    bx = (int)blockIdx.x / groups;
  };
  // weight blob of layer l for this CTA (one elected lane of MMA warp 1)
  auto request_blob = [&](int l) {
    int groups, gx, group, bx;
    geometry(l, groups, gx, group, bx);
    if (bx >= gx) return;
    const ChainLayer& L = cp.layer[l];
#define X(CIN_, NT_, TAPS_) \
    if (L.cin == CIN_ && L.nt == NT_ && L.taps == TAPS_) chain_load_blob<CIN_, NT_, TAPS_>(L, group, smem, B.wbar);
    MARU_CHAIN_CASES(X)
#undef X
  };
  if (warp == 1 && elect_one()) request_blob(0);   // weights do not depend on the previous kernel
  __syncwarp();
  pdl_wait();                                      // the first layer reads the previous kernel's output

  ChainState st;
  st.slot_par = 0u;
  st.full_par = 0u;
  st.acc_par = 0u;
  st.n_active = 0u;
  for (int l = 0; l < cp.n_layers; l++) {
    const ChainLayer& L = cp.layer[l];
    int groups, gx, group, bx;
    geometry(l, groups, gx, group, bx);
    long long* stamp = (cp.stamps != nullptr && blockIdx.x == 0) ? cp.stamps + 4 * l : nullptr;
    if (stamp != nullptr && threadIdx.x == 0) stamp[0] = globaltimer_ns();
    if (bx < gx) {
#define X(CIN_, NT_, TAPS_)                                                   \
      if (L.cin == CIN_ && L.nt == NT_ && L.taps == TAPS_)                    \
        chain_layer<CIN_, NT_, TAPS_>(cp, L, smem, B, tmem_base, bx, gx, group, st, stamp);
      MARU_CHAIN_CASES(X)
#undef X
      if (warp == 1) {
        // weight region free once both issuers' MMAs have retired: fetch the next layer's blob now, it
        // travels while the last epilogue drains and the grid barrier is crossed
        mbar_wait(B.mdone, st.n_active & 1u);
        if (l + 1 < cp.n_layers && elect_one()) request_blob(l + 1);
        __syncwarp();
      }
      st.n_active++;
    } else if (l + 1 < cp.n_layers && warp == 1) {
      if (elect_one()) request_blob(l + 1);         // idle in this layer (grid not divisible by the groups)
      __syncwarp();
    }
    if (stamp != nullptr && threadIdx.x == 64) stamp[2] = globaltimer_ns();        // an epilogue thread: CTA 0 done
    if (l + 1 < cp.n_layers) {
      chain_grid_sync(cp.sync + l, (unsigned)n_ctas);
      if (l == 0) pdl_launch_dependents();          // every CTA of the chain is resident by now
    }
  }
  if (cp.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) cp.stamps[4 * cp.n_layers] = globaltimer_ns();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

cudaError_t configure_conv_chain() {
  cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       CHAIN_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

bool conv_chain_supports(int cin, int nt, int taps) {
#define X(CIN_, NT_, TAPS_) \
  if (cin == CIN_ && nt == NT_ && taps == TAPS_) return true;
  MARU_CHAIN_CASES(X)
#undef X
  return false;
}

int conv_chain_blob_bytes(int cin, int nt, int taps) { return taps * cin * nt * 2 + 2 * nt * 16; }

template <int CIN, int NT, int TAPS>
static int chain_stage_count(bool has_res, int budget) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  const int fixed = Cfg::ONES_BYTES + (has_res ? Cfg::IDENT_BYTES : 0);
  int n = (budget - fixed) / Cfg::stage_bytes(has_res);
  return n > CHAIN_MAXS ? CHAIN_MAXS : n;
}

int conv_chain_stages(int cin, int nt, int taps, bool has_res, int w_region) {
  const int budget = CHAIN_SMEM - w_region - CHAIN_BAR_BYTES;
#define X(CIN_, NT_, TAPS_) \
  if (cin == CIN_ && nt == NT_ && taps == TAPS_) return chain_stage_count<CIN_, NT_, TAPS_>(has_res, budget);
  MARU_CHAIN_CASES(X)
#undef X
  return 0;
}

cudaError_t launch_conv_chain(ChainParams& cp, int sm_count, cudaStream_t stream) {
  const int smem_dyn = CHAIN_SMEM;
  int w_region = 0;
  for (int l = 0; l < cp.n_layers; l++) {
    const ChainLayer& L = cp.layer[l];
    if (!conv_chain_supports(L.cin, L.nt, L.taps)) return cudaErrorInvalidConfiguration;
    const int b = conv_chain_blob_bytes(L.cin, L.nt, L.taps);
    if (b > w_region) w_region = b;
  }
  w_region = (w_region + 1023) & ~1023;                  // keeps the per-layer region 1 KB aligned
  cp.w_region = w_region;
  const int budget = smem_dyn - w_region - CHAIN_BAR_BYTES;
  for (int l = 0; l < cp.n_layers; l++) {
    ChainLayer& L = cp.layer[l];
    const bool has_res = L.res != nullptr;
    int st = 0;
#define X(CIN_, NT_, TAPS_) \
    if (L.cin == CIN_ && L.nt == NT_ && L.taps == TAPS_) st = chain_stage_count<CIN_, NT_, TAPS_>(has_res, budget);
    MARU_CHAIN_CASES(X)
#undef X
    if (st < 2) return cudaErrorInvalidConfiguration;   // a single stage cannot overlap loads with MMAs
    L.stages = st;
  }
  return launch_kernel(conv_chain_kernel, dim3(sm_count), dim3(CHAIN_THREADS), (size_t)smem_dyn, stream,
                       cp.pdl != 0, cp);
}

}  // namespace maru
