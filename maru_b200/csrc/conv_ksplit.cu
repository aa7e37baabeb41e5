// conv_ksplit.cu — 3x3 convolution with 128 input channels (the inner layers of the C = 256 nets), sm_100a.
//
// Same math and tensor layout as conv_gemm.cu (tcgen05 implicit GEMM over the padded-flat rows, bias and residual
// as extra MMAs, fused ReLU + board mask).  What differs is the shared-memory budget: the taps of 64 output
// channels are 147 KB, a whole input slab (176 rows x 128 channels) is 45 KB, so conv_gemm<128, 64, 9> has room for
// ONE stage and every tile pays its slab load (~2 us) before its 1.8 us of MMAs (bench at batch 512: 4.5 us per
// tile, 34 % of tensor peak).  Here a stage is HALF a slab — 64 of the 128 input channels — and a tile is two
// half-items: while the 36 MMAs of one half run, the next half (of this tile or the next) lands.  Three half-stages
// fit next to the weights (two when the layer has a residual, whose tile gets a buffer of its own).
//
// Roles (320 threads, persistent over M tiles, 1 CTA per SM, one CTA column per group of 64 output channels):
//   warp 0   : producer — weights + bias operand (one blob), then per tile two half-slabs (8 planes each) and, for
//              residual layers, the residual tile
//   warp 1   : MMA issue (one elected lane): bias, residual x I, then 9 taps x 4 K-steps per half
//   warps 2-9: epilogue — tcgen05.ld, bf16 pack, ReLU, mask (read from global memory), 16-byte stores
// The accumulation order differs from conv_gemm (channels 0-63 of all taps, then 64-127), so results agree with it
// to fp32 rounding, not bit for bit; tests compare both against the torch port of the network.
#include "conv_cfg.cuh"

namespace maru {

constexpr int KS_THREADS = 320;
constexpr int KS_CIN = 128, KS_NT = 64, KS_TAPS = 9;
constexpr int KS_LEAD = 24, KS_SLAB_ROWS = TILE_M + 2 * KS_LEAD;        // 176
constexpr int KS_KCH = KS_CIN / 8;                                     // 16 planes
constexpr int KS_HALF_BYTES = (KS_KCH / 2) * KS_SLAB_ROWS * 16;        // 22 528
constexpr int KS_W_BYTES = KS_TAPS * KS_CIN * KS_NT * 2;               // 147 456
constexpr int KS_BIASB_BYTES = 2 * KS_NT * 16;
constexpr int KS_ONES_BYTES = 2 * TILE_M * 16;
constexpr int KS_IDENT_BYTES = KS_NT * KS_NT * 2;
constexpr int KS_RES_BYTES = TILE_M * KS_NT * 2;
constexpr int KS_BAR_BYTES = 256;
constexpr int KS_SMEM = 232448;

static int ks_stages(bool has_res) {
  const int fixed = KS_W_BYTES + KS_BIASB_BYTES + KS_ONES_BYTES + KS_BAR_BYTES + (has_res ? KS_IDENT_BYTES + KS_RES_BYTES : 0);
  int n = (KS_SMEM - fixed) / KS_HALF_BYTES;
  return n > 4 ? 4 : n;
}
static int ks_smem(bool has_res) {
  return KS_W_BYTES + KS_BIASB_BYTES + KS_ONES_BYTES + KS_BAR_BYTES + (has_res ? KS_IDENT_BYTES + KS_RES_BYTES : 0) +
         ks_stages(has_res) * KS_HALF_BYTES;
}

__global__ void __launch_bounds__(KS_THREADS, 1) conv_ksplit_kernel(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const bool has_res = p.res != nullptr;
  const int HS = p.stages;                                // half-stages in the ring
  uint8_t* s_w = smem;
  uint8_t* s_biasb = s_w + KS_W_BYTES;                    // [W | bias image] is one blob
  uint8_t* s_ones = s_biasb + KS_BIASB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + KS_ONES_BYTES);
  uint8_t* s_ident = reinterpret_cast<uint8_t*>(bars) + KS_BAR_BYTES;
  uint8_t* s_res = s_ident + (has_res ? KS_IDENT_BYTES : 0);
  uint8_t* s_stage = s_res + (has_res ? KS_RES_BYTES : 0);
  uint64_t* full = bars;            // [4] half-slab landed
  uint64_t* empty = bars + 4;       // [4] half-slab consumed (MMA commit)
  uint64_t* tfull = bars + 8;       // [2] accumulator ready
  uint64_t* tempty = bars + 10;     // [2] accumulator drained (8 epilogue warps)
  uint64_t* wbar = bars + 12;       // weights + bias operand landed
  uint64_t* rfull = bars + 13;      // residual tile landed
  uint64_t* rempty = bars + 14;     // residual MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.y;
  const int m_rows = (p.n_dev != nullptr) ? total_rows_g(*p.n_dev, p.pos_rows) : p.m_rows;
  const int n_tiles = (m_rows + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    fence_mbar_init();
    mbar_expect_tx(wbar, KS_W_BYTES + KS_BIASB_BYTES);
    bulk_load_1d(s_w, p.blob + (size_t)group * (KS_W_BYTES + KS_BIASB_BYTES), KS_W_BYTES + KS_BIASB_BYTES, wbar);
    for (int i = 0; i < 4; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(rfull, 1);
    mbar_init(rempty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);          // two accumulators of 64 columns
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    {                                                           // ones: (1, 1, 0, ...) in chunk 0
      const uint32_t w0 = (t < TILE_M) ? 0x3F803F80u : 0u;
      *reinterpret_cast<uint4*>(s_ones + t * 16) = make_uint4(w0, 0u, 0u, 0u);
    }
    if (has_res) {
      for (int i = t; i < (KS_NT / 8) * KS_NT; i += 256) {      // identity: [k-chunk c][row n]
        const int c = i / KS_NT, n = i - c * KS_NT;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if ((n >> 3) == c) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;
        *reinterpret_cast<uint4*>(s_ident + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== producer ===========================
    pdl_wait();
    uint32_t h = 0;                                      // half-items issued so far
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const size_t row0 = (size_t)GUARD_ROWS + (size_t)tile * TILE_M;
      if (has_res) {
        mbar_wait(rempty, (uint32_t)(it & 1) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(rfull, KS_RES_BYTES);
          const int rc0 = p.res_chunk0 + group * (KS_NT / 8);
          if (p.res_blocked) {
            bulk_load_1d(s_res, p.res + ((size_t)tile * p.res_planes + rc0) * (TILE_M * 8), KS_RES_BYTES, rfull);
          } else {
#pragma unroll 1
            for (int c = 0; c < KS_NT / 8; c++)
              bulk_load_1d(s_res + c * TILE_M * 16, p.res + ((size_t)(rc0 + c) * p.rows_alloc + row0) * 8, TILE_M * 16,
                           rfull);
          }
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int half = 0; half < 2; half++, h++) {
        const uint32_t s = h % (uint32_t)HS;
        mbar_wait(&empty[s], ((h / (uint32_t)HS) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[s], KS_HALF_BYTES);
          uint8_t* dst = s_stage + s * KS_HALF_BYTES;
#pragma unroll 1
          for (int c = 0; c < KS_KCH / 2; c++) {
            const __nv_bfloat16* src =
                p.in + ((size_t)(p.in_chunk0 + half * (KS_KCH / 2) + c) * p.rows_alloc + row0 - KS_LEAD) * 8;
            bulk_load_1d(dst + c * KS_SLAB_ROWS * 16, src, KS_SLAB_ROWS * 16, &full[s]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issue ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, KS_NT, 0, 0);
    const uint64_t a_d = umma_desc(0, KS_SLAB_ROWS * 16, 128);
    const uint64_t r_d = umma_desc(0, TILE_M * 16, 128);
    const uint64_t b_d = umma_desc(0, KS_NT * 16, 128);
    const uint32_t a_hi = (uint32_t)(a_d >> 32), r_hi = (uint32_t)(r_d >> 32), b_hi = (uint32_t)(b_d >> 32);
    const uint32_t w_lo = (uint32_t)b_d + (smem_u32(s_w) >> 4);
    const uint32_t ones_lo = (uint32_t)r_d + (smem_u32(s_ones) >> 4);
    const uint32_t biasb_lo = (uint32_t)b_d + (smem_u32(s_biasb) >> 4);
    const uint32_t ident_lo = (uint32_t)b_d + (smem_u32(s_ident) >> 4);
    const uint32_t res_lo = (uint32_t)r_d + (smem_u32(s_res) >> 4);
    mbar_wait(wbar, 0);
    uint32_t h = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const uint32_t d_tmem = tmem_base + acc * KS_NT;
      mbar_wait(&tempty[acc], (uint32_t)((it >> 1) & 1) ^ 1u);
      if (has_res) mbar_wait(rfull, (uint32_t)(it & 1));
#pragma unroll 1
      for (int half = 0; half < 2; half++, h++) {
        const uint32_t s = h % (uint32_t)HS;
        mbar_wait(&full[s], (h / (uint32_t)HS) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (uint32_t)a_d + (smem_u32(s_stage + s * KS_HALF_BYTES) >> 4);
          if (half == 0) {
            umma_bf16_lohi<0>(d_tmem, ones_lo, r_hi, biasb_lo, b_hi, idesc);           // D = bias
            if (has_res) {
#pragma unroll
              for (int j = 0; j < KS_NT / 16; j++)                                      // D += residual
                umma_bf16_lohi<1>(d_tmem, res_lo + 2 * j * TILE_M, r_hi, ident_lo + 2 * j * KS_NT, b_hi, idesc);
              umma_commit(rempty);
            }
          }
          const uint32_t wh = w_lo + (uint32_t)half * (KS_KCH / 2) * KS_NT;            // this half's K chunks
#pragma unroll
          for (int t = 0; t < KS_TAPS; t++) {
            const int off = (t / 3 - 1) * p.pitch + (t % 3 - 1);
#pragma unroll
            for (int j = 0; j < KS_CIN / 32; j++) {
              const uint32_t a16 = (2 * j) * KS_SLAB_ROWS + KS_LEAD + off;              // in 16-byte units
              const uint32_t b16 = (t * KS_KCH + 2 * j) * KS_NT;
              umma_bf16_lohi<1>(d_tmem, a_lo + a16, a_hi, wh + b16, b_hi, idesc);
            }
          }
          umma_commit(&empty[s]);
          if (half == 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    constexpr int CW = KS_NT / 2;
    const int lg = warp & 3;
    const int hcol = (warp - 2) >> 2;
    const int trow = lg * 32 + lane;
    pdl_wait();                                            // the mask is read from global memory
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const int row = tile * TILE_M + trow;
      const bool valid = row < m_rows;
      const size_t grow = (size_t)GUARD_ROWS + row;
      const bool on = valid && (p.mask[row] != 0);
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * KS_NT + hcol * CW;
      const int chunk0 = p.out_chunk0 + group * (KS_NT / 8) + hcol * (CW / 8);
      const size_t pstride = p.out_blocked ? (size_t)TILE_M * 8 : (size_t)p.rows_alloc * 8;
      __nv_bfloat16* out_row = p.out_blocked
                                   ? p.out + (((size_t)tile * p.out_planes + chunk0) * TILE_M + trow) * 8
                                   : p.out + ((size_t)chunk0 * p.rows_alloc + grow) * 8;
      mbar_wait(&tfull[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(t_addr, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          uint32_t pk = pack_bf16x2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
          if (p.relu) pk = relu_bf16x2(pk);
          o[e] = on ? pk : 0u;
        }
        if (valid) *reinterpret_cast<uint4*>(out_row + (size_t)q * pstride) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

cudaError_t configure_conv_ksplit() {
  cudaError_t e = cudaFuncSetAttribute(conv_ksplit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KS_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_ksplit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

bool conv_ksplit_supports(int cin, int nt, int taps) { return cin == KS_CIN && nt == KS_NT && taps == KS_TAPS; }

cudaError_t launch_conv_ksplit(ConvParams p, int sm_count, cudaStream_t stream) {
  if (p.blob == nullptr || p.in_blocked || (p.cout_total % KS_NT) != 0) return cudaErrorInvalidValue;
  const bool has_res = p.res != nullptr;
  p.stages = ks_stages(has_res);
  if (p.stages < 2) return cudaErrorInvalidConfiguration;
  const int groups = p.cout_total / KS_NT;
  const int n_tiles = (p.m_rows + TILE_M - 1) / TILE_M;
  int gx = sm_count / groups;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  return launch_kernel(conv_ksplit_kernel, dim3(gx, groups, 1), dim3(KS_THREADS), (size_t)ks_smem(has_res), stream,
                       p.pdl != 0, p);
}

}  // namespace maru
