// conv_wide.cu — 1x1 convolution with 256 input channels and 256 output channels per CTA (the C = 256 nets: the
// 256 -> 768 qkv projection = three groups, and the 256 -> 256 output projection with its residual), sm_100a.
//
// conv_gemm<256, 128, 1> runs this layer as six groups of 128 output channels; every group loads the same 64 KB
// input tile again, 616 MB of L2 -> SM traffic per layer at batch 512 (~100 us at 6 TB/s; measured 120 us).  Here a
// CTA owns 256 output channels: N = 256 MMAs (128 cycles, tensor-bound: operand fetch is 96), the 131 KB weight image
// resident, and the input tile streamed as two 32 KB halves of 128 input channels each (a tile-blocked tensor: one
// bulk copy per half) through a two-stage ring.  Half as many redundant tile loads, and the next half lands while
// this one multiplies.
//
// Roles (320 threads, persistent over M tiles, 1 CTA per SM, one CTA column per group of 256 output channels):
//   warp 0   : producer — weights + bias operand (one blob), then two half-tiles per tile
//   warp 1   : MMA issue (one elected lane): bias, then 8 K-steps per half into one of two 256-column accumulators
//   warps 2-9: epilogue — 128 columns per thread in four tcgen05.ld of 32, bf16 pack, ReLU, mask, 16-byte stores
// Same accumulation order as conv_gemm (K ascending); a residual is added by the epilogue in fp32 (conv_gemm adds it
// with an identity MMA before the K loop), so residual layers agree with conv_gemm to one bf16 rounding.
#include "conv_cfg.cuh"

namespace maru {

constexpr int CW_THREADS = 320;
constexpr int CW_CIN = 256, CW_NT = 256;
constexpr int CW_KCH = CW_CIN / 8;                                   // 32 planes
constexpr int CW_HALF_BYTES = (CW_KCH / 2) * TILE_M * 16;            // 32 768
constexpr int CW_W_BYTES = CW_CIN * CW_NT * 2;                       // 131 072
constexpr int CW_BIASB_BYTES = 2 * CW_NT * 16;                       // 8 192
constexpr int CW_ONES_BYTES = 2 * TILE_M * 16;
constexpr int CW_BAR_BYTES = 256;
constexpr int CW_STAGES = 2;
constexpr int CW_SMEM = CW_W_BYTES + CW_BIASB_BYTES + CW_ONES_BYTES + CW_BAR_BYTES + CW_STAGES * CW_HALF_BYTES;
static_assert(CW_SMEM <= 232448, "shared memory budget");

__global__ void __launch_bounds__(CW_THREADS, 1) conv_wide_kernel(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_biasb = s_w + CW_W_BYTES;                    // [W | bias image] is one blob
  uint8_t* s_ones = s_biasb + CW_BIASB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + CW_ONES_BYTES);
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(bars) + CW_BAR_BYTES;
  uint64_t* full = bars;            // [2] half-tile landed
  uint64_t* empty = bars + 2;       // [2] half-tile consumed (MMA commit)
  uint64_t* tfull = bars + 4;       // [2] accumulator ready
  uint64_t* tempty = bars + 6;      // [2] accumulator drained (8 epilogue warps)
  uint64_t* wbar = bars + 8;        // weights + bias operand landed
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
    mbar_expect_tx(wbar, CW_W_BYTES + CW_BIASB_BYTES);
    bulk_load_1d(s_w, p.blob + (size_t)group * (CW_W_BYTES + CW_BIASB_BYTES), CW_W_BYTES + CW_BIASB_BYTES, wbar);
    for (int i = 0; i < 2; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);          // two accumulators of 256 columns
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;                               // ones: (1, 1, 0, ...) in chunk 0
    const uint32_t w0 = (t < TILE_M) ? 0x3F803F80u : 0u;
    *reinterpret_cast<uint4*>(s_ones + t * 16) = make_uint4(w0, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== producer ===========================
    pdl_wait();
    uint32_t h = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int half = 0; half < 2; half++, h++) {
        const uint32_t s = h & 1u;
        mbar_wait(&empty[s], ((h >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[s], CW_HALF_BYTES);
          // tile-blocked input: [tile][plane][128 rows][8 ch]: 16 consecutive planes are one contiguous 32 KB
          const __nv_bfloat16* src =
              p.in + ((size_t)tile * p.in_planes + p.in_chunk0 + half * (CW_KCH / 2)) * (TILE_M * 8);
          bulk_load_1d(s_stage + s * CW_HALF_BYTES, src, CW_HALF_BYTES, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issue ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, CW_NT, 0, 0);
    const uint64_t a_d = umma_desc(0, TILE_M * 16, 128);           // half-tile, ones: LBO = 128 rows
    const uint64_t b_d = umma_desc(0, CW_NT * 16, 128);            // W, biasB: LBO = NT rows
    const uint32_t a_hi = (uint32_t)(a_d >> 32), b_hi = (uint32_t)(b_d >> 32);
    const uint32_t w_lo = (uint32_t)b_d + (smem_u32(s_w) >> 4);
    const uint32_t ones_lo = (uint32_t)a_d + (smem_u32(s_ones) >> 4);
    const uint32_t biasb_lo = (uint32_t)b_d + (smem_u32(s_biasb) >> 4);
    mbar_wait(wbar, 0);
    uint32_t h = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const uint32_t d_tmem = tmem_base + acc * CW_NT;
      mbar_wait(&tempty[acc], (uint32_t)((it >> 1) & 1) ^ 1u);
#pragma unroll 1
      for (int half = 0; half < 2; half++, h++) {
        const uint32_t s = h & 1u;
        mbar_wait(&full[s], (h >> 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (uint32_t)a_d + (smem_u32(s_stage + s * CW_HALF_BYTES) >> 4);
          if (half == 0) umma_bf16_lohi<0>(d_tmem, ones_lo, a_hi, biasb_lo, b_hi, idesc);      // D = bias
          const uint32_t wh = w_lo + (uint32_t)half * (CW_KCH / 2) * CW_NT;                    // this half's K chunks
#pragma unroll
          for (int j = 0; j < CW_CIN / 32; j++)
            umma_bf16_lohi<1>(d_tmem, a_lo + 2 * j * TILE_M, a_hi, wh + 2 * j * CW_NT, b_hi, idesc);
          umma_commit(&empty[s]);
          if (half == 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    constexpr int CW = CW_NT / 2;                          // 128 columns per warp
    const int lg = warp & 3;
    const int hcol = (warp - 2) >> 2;
    const int trow = lg * 32 + lane;
    pdl_wait();                                            // the mask (and the residual) are read from global memory
    const bool has_res = p.res != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const int row = tile * TILE_M + trow;
      const bool valid = row < m_rows;
      const size_t grow = (size_t)GUARD_ROWS + row;
      const bool on = valid && (p.mask[row] != 0);
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * CW_NT + hcol * CW;
      const int chunk0 = p.out_chunk0 + group * (CW_NT / 8) + hcol * (CW / 8);
      const size_t pstride = p.out_blocked ? (size_t)TILE_M * 8 : (size_t)p.rows_alloc * 8;
      __nv_bfloat16* out_row = p.out_blocked
                                   ? p.out + (((size_t)tile * p.out_planes + chunk0) * TILE_M + trow) * 8
                                   : p.out + ((size_t)chunk0 * p.rows_alloc + grow) * 8;
      // residual (same tensor geometry as the output, e.g. h = h + proj(o) in place): all 128 columns of this
      // thread's row are requested before the accumulator is waited for
      uint4 rv[CW / 8];
      if (has_res) {
        const size_t rstride = p.res_blocked ? (size_t)TILE_M * 8 : (size_t)p.rows_alloc * 8;
        const int rc0 = p.res_chunk0 + group * (CW_NT / 8) + hcol * (CW / 8);
        const __nv_bfloat16* res_row = p.res_blocked
                                           ? p.res + (((size_t)tile * p.res_planes + rc0) * TILE_M + trow) * 8
                                           : p.res + ((size_t)rc0 * p.rows_alloc + grow) * 8;
#pragma unroll
        for (int q = 0; q < CW / 8; q++)
          rv[q] = on ? *reinterpret_cast<const uint4*>(res_row + (size_t)q * rstride) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&tfull[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_addr + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == CW) {                               // last read: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          uint32_t o[4];
          uint32_t rw[4] = {0u, 0u, 0u, 0u};
          if (has_res) {
            rw[0] = rv[c0 / 8 + q].x;
            rw[1] = rv[c0 / 8 + q].y;
            rw[2] = rv[c0 / 8 + q].z;
            rw[3] = rv[c0 / 8 + q].w;
          }
#pragma unroll
          for (int e = 0; e < 4; e++) {
            float lo = __uint_as_float(v[q * 8 + 2 * e]), hi = __uint_as_float(v[q * 8 + 2 * e + 1]);
            if (has_res) {
              lo += bf16_lo(rw[e]);
              hi += bf16_hi(rw[e]);
            }
            uint32_t pk = pack_bf16x2(lo, hi);
            if (p.relu) pk = relu_bf16x2(pk);
            o[e] = on ? pk : 0u;
          }
          if (valid)
            *reinterpret_cast<uint4*>(out_row + (size_t)(c0 / 8 + q) * pstride) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

cudaError_t configure_conv_wide() {
  cudaError_t e = cudaFuncSetAttribute(conv_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_wide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

bool conv_wide_supports(int cin, int cout, int taps) { return cin == CW_CIN && taps == 1 && cout % CW_NT == 0; }
int conv_wide_blob_bytes() { return CW_W_BYTES + CW_BIASB_BYTES; }

// p.blob: [groups of 256 output channels][W image [32 k-chunks][256][8] | bias operand image]
cudaError_t launch_conv_wide(ConvParams p, int sm_count, cudaStream_t stream) {
  if (p.blob == nullptr || !p.in_blocked || (p.cout_total % CW_NT) != 0) return cudaErrorInvalidValue;
  const int groups = p.cout_total / CW_NT;
  const int n_tiles = (p.m_rows + TILE_M - 1) / TILE_M;
  int gx = sm_count / groups;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  p.stages = CW_STAGES;
  return launch_kernel(conv_wide_kernel, dim3(gx, groups, 1), dim3(CW_THREADS), (size_t)CW_SMEM, stream, p.pdl != 0, p);
}

}  // namespace maru
