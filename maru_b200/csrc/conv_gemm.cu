// conv_gemm.cu — K2/K3: 3x3 and 1x1 convolutions as tcgen05 implicit GEMMs (sm_100a).
//
//   out[m, :] = select(mask[m], act(sum_t  in[m + off_t, :] . W_t  + bias + res[m, :]), 0)
//
// Replaces the cuDNN/cuBLAS convolutions the reference reaches through TorchScript
// (reference src/deepgo/native/cpp/Model.cpp:96 `_model.forward`), with BN folded into W/bias
// (maru_b200/convert.py) and bias + residual + ReLU + board mask fused into the epilogue.
//
// Layout: activations are "chunk-planar" bf16  [C/8][rows_alloc][8]  over the padded-flat row
// geometry of common.cuh.  One plane holds 8 channels (16 bytes) per row, so
//   * a slab of consecutive rows of one plane is ONE contiguous bulk copy (cp.async.bulk, UBLKCP),
//   * it lands in shared memory already in the UMMA SWIZZLE_NONE K-major canonical layout
//     (core matrix = 8 rows x 16 B contiguous; SBO = 128 B; LBO = slab_rows*16 B),
//   * a 3x3 tap is the SAME slab with the descriptor start address moved by off_t*16 bytes,
//     so each input row is fetched from L2 once per tile, not nine times,
//   * the epilogue (one TMEM lane = one row per thread) reads residuals and writes outputs as
//     16-byte row-consecutive accesses: a warp store is one contiguous 512-byte segment.
//
// Roles (192 threads, persistent over M tiles of 128 rows, 1 CTA per SM):
//   warp 0   : producer  — weights once, then per tile one input slab (+ the residual tile, so the
//              epilogue never waits on a global load) into an mbarrier ring of STAGES stages
//   warp 1   : TMEM alloc + elected-lane tcgen05.mma issue, accumulators double-buffered in TMEM
//   warps 2-5: epilogue  — tcgen05.ld, bias/residual/ReLU/mask, bf16 pack, global store
#include "kernels.h"

namespace maru {

template <int CIN, int NT, int TAPS>
struct ConvCfg {
  static constexpr int LEAD = (TAPS == 9) ? 24 : 0;            // slab rows before the tile
  static constexpr int SLAB_ROWS = TILE_M + 2 * LEAD;          // 176 or 128 (multiple of 8)
  static constexpr int KCH = CIN / 8;                          // 16-byte channel chunks
  static constexpr int W_BYTES = TAPS * CIN * NT * 2;
  static constexpr int SLAB_BYTES = SLAB_ROWS * CIN * 2;
  static constexpr int RES_BYTES = TILE_M * NT * 2;            // residual tile [NT/8][128][16 B]
  static constexpr int MAX_STAGES = 4;
  static constexpr int BIAS_BYTES = NT * 4;
  static constexpr int BAR_BYTES = 128;
  static constexpr int FIXED_BYTES = W_BYTES + BIAS_BYTES + BAR_BYTES + 16;
  static constexpr int SMEM_LIMIT = 232448;                    // 227 KB per CTA
  // a stage = input slab (+ residual tile when the layer has one); as many stages as fit, <= 4
  static int stage_bytes(bool has_res) { return SLAB_BYTES + (has_res ? RES_BYTES : 0); }
  static int stages(bool has_res) {
    int n = (SMEM_LIMIT - FIXED_BYTES) / stage_bytes(has_res);
    return n > MAX_STAGES ? MAX_STAGES : n;
  }
  static int smem_bytes(bool has_res) { return FIXED_BYTES + stages(has_res) * stage_bytes(has_res); }
  static constexpr int TMEM_COLS = (2 * NT <= 32) ? 32 : (2 * NT <= 64) ? 64 : (2 * NT <= 128) ? 128
                                   : (2 * NT <= 256) ? 256 : 512;
  static_assert(CIN % 16 == 0 && NT % 16 == 0 && NT >= 16 && NT <= 256, "UMMA shape");
  static_assert(TAPS == 1 || LEAD >= HALO, "slab lead must cover the 3x3 halo");
  static_assert(FIXED_BYTES + SLAB_BYTES + RES_BYTES <= SMEM_LIMIT, "one stage must fit in 227 KB");
};

template <int CIN, int NT, int TAPS>
__global__ void __launch_bounds__(192, 1) conv_gemm_kernel(const ConvParams p) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int MAXS = Cfg::MAX_STAGES;
  const bool has_res = p.res != nullptr;
  const int STAGES = p.stages;
  const int STAGE_BYTES = Cfg::SLAB_BYTES + (has_res ? Cfg::RES_BYTES : 0);
  uint8_t* s_w = smem;
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::W_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::W_BYTES + Cfg::BIAS_BYTES);
  uint8_t* s_stage = smem + Cfg::W_BYTES + Cfg::BIAS_BYTES + Cfg::BAR_BYTES + 16;
  uint64_t* full = bars;                   // [MAXS] input slab (+ residual tile) landed
  uint64_t* empty = bars + MAXS;           // [MAXS] stage reusable
  uint64_t* tfull = bars + 2 * MAXS;       // [2] accumulator ready
  uint64_t* tempty = tfull + 2;            // [2] accumulator drained
  uint64_t* wbar = tempty + 2;             // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.y;                 // output-channel group of NT channels
  const int m_rows = (p.n_dev != nullptr) ? total_rows(*p.n_dev) : p.m_rows;
  const int n_tiles = (m_rows + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) {
      mbar_init(&full[i], 1);
      // the MMA commit frees the slab; with a residual the 4 epilogue warps must also be done
      mbar_init(&empty[i], has_res ? 5 : 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < NT; i += 128) s_bias[i] = p.bias[group * NT + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Roles are WARP-uniform and a single elected lane issues the async instructions: in a
  // thread-divergent branch ptxas wraps every UTCHMMA / UBLKCP in an ELECT + BRA.U.ANY loop
  // (measured ~115 cycles per MMA instead of 48, profiles/r1_umma_microbench.txt).
  if (warp == 0) {
    // =========================== producer ===========================
    if (elect_one()) {
      // weights of this group: [TAPS][KCH][cout_total][8] in HBM -> [TAPS][KCH][NT][8] in smem
      mbar_expect_tx(wbar, Cfg::W_BYTES);
      for (int tc = 0; tc < TAPS * Cfg::KCH; tc++) {
        const __nv_bfloat16* src = p.w + ((size_t)tc * p.cout_total + (size_t)group * NT) * 8;
        bulk_load_1d(s_w + (size_t)tc * NT * 16, src, NT * 16, wbar);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[stage], Cfg::SLAB_BYTES + (has_res ? Cfg::RES_BYTES : 0));
        const size_t row0 = (size_t)GUARD_ROWS + (size_t)tile * TILE_M;
        uint8_t* dst = s_stage + stage * STAGE_BYTES;
#pragma unroll 1
        for (int c = 0; c < Cfg::KCH; c++) {
          const __nv_bfloat16* src = p.in + ((size_t)(p.in_chunk0 + c) * p.rows_alloc + row0 - Cfg::LEAD) * 8;
          bulk_load_1d(dst + c * Cfg::SLAB_ROWS * 16, src, Cfg::SLAB_ROWS * 16, &full[stage]);
        }
        if (has_res) {
#pragma unroll 1
          for (int c = 0; c < NT / 8; c++) {
            const __nv_bfloat16* src =
                p.res + ((size_t)(p.res_chunk0 + group * (NT / 8) + c) * p.rows_alloc + row0) * 8;
            bulk_load_1d(dst + Cfg::SLAB_BYTES + c * TILE_M * 16, src, TILE_M * 16, &full[stage]);
          }
        }
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, NT, 0, 0);
    // descriptors differ only in the 14-bit start-address field: precompute the rest once
    const uint64_t a_hi = umma_desc(0, Cfg::SLAB_ROWS * 16, 128);
    const uint64_t b_hi = umma_desc(0, NT * 16, 128);
    const uint32_t w_addr = smem_u32(s_w);
    mbar_wait(wbar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        // smem addresses are < 256 KB, so (addr >> 4) never carries out of the 14-bit field:
        // per-MMA descriptor = base + compile-time constant (one uniform add)
        const uint64_t a_base = a_hi + (smem_u32(s_stage + stage * STAGE_BYTES) >> 4);
        const uint64_t b_base = b_hi + (w_addr >> 4);
        const uint32_t d_tmem = tmem_base + acc * NT;
#pragma unroll
        for (int t = 0; t < TAPS; t++) {
          const int off = (TAPS == 9) ? ((t / 3 - 1) * PITCH + (t % 3 - 1)) : 0;
#pragma unroll
          for (int j = 0; j < CIN / 16; j++) {
            const uint32_t a16 = (2 * j) * Cfg::SLAB_ROWS + Cfg::LEAD + off;   // in 16-byte units
            const uint32_t b16 = (t * Cfg::KCH + 2 * j) * NT;
            umma_bf16(d_tmem, a_base + a16, b_base + b16, idesc, (t | j) != 0);
          }
        }
        umma_commit(&empty[stage]);  // slab reusable once these MMAs retire
        umma_commit(&tfull[acc]);    // accumulator complete
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    int it = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int trow = lg * 32 + lane;
      const int row = tile * TILE_M + trow;
      const bool valid = row < m_rows;
      const uint8_t mk = valid ? p.mask[row] : (uint8_t)0;   // issued before the wait: latency hidden
      const size_t grow = (size_t)GUARD_ROWS + row;
      const uint8_t* s_res = s_stage + stage * STAGE_BYTES + Cfg::SLAB_BYTES;
      if (has_res) mbar_wait(&full[stage], phase);           // residual tile landed with the slab
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const bool on = mk != 0;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * NT + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int lchunk = c0 / 8 + q;                      // chunk within this group's NT channels
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; e++) x[e] = __uint_as_float(v[q * 8 + e]) + s_bias[c0 + q * 8 + e];
          if (has_res) {
            const uint4 r = *reinterpret_cast<const uint4*>(s_res + (lchunk * TILE_M + trow) * 16);
            x[0] += bf16_lo(r.x); x[1] += bf16_hi(r.x);
            x[2] += bf16_lo(r.y); x[3] += bf16_hi(r.y);
            x[4] += bf16_lo(r.z); x[5] += bf16_hi(r.z);
            x[6] += bf16_lo(r.w); x[7] += bf16_hi(r.w);
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; e++) x[e] = fmaxf(x[e], 0.0f);
          }
          uint4 o;
          o.x = on ? pack_bf16x2(x[0], x[1]) : 0u;
          o.y = on ? pack_bf16x2(x[2], x[3]) : 0u;
          o.z = on ? pack_bf16x2(x[4], x[5]) : 0u;
          o.w = on ? pack_bf16x2(x[6], x[7]) : 0u;
          if (valid) {
            const int chunk = p.out_chunk0 + group * (NT / 8) + lchunk;
            *reinterpret_cast<uint4*>(p.out + ((size_t)chunk * p.rows_alloc + grow) * 8) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty[acc]);
        if (has_res) mbar_arrive(&empty[stage]);
      }
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
template <int CIN, int NT, int TAPS>
static cudaError_t configure_cfg() {
  return cudaFuncSetAttribute(conv_gemm_kernel<CIN, NT, TAPS>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                              ConvCfg<CIN, NT, TAPS>::SMEM_LIMIT);
}

template <int CIN, int NT, int TAPS>
static cudaError_t launch_cfg(ConvParams p, int sm_count, cudaStream_t stream) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  const bool has_res = p.res != nullptr;
  p.stages = Cfg::stages(has_res);
  const int groups = p.cout_total / NT;
  const int n_tiles = (p.m_rows + TILE_M - 1) / TILE_M;
  int gx = sm_count / groups;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  dim3 grid(gx, groups, 1);
  conv_gemm_kernel<CIN, NT, TAPS><<<grid, 192, Cfg::smem_bytes(has_res), stream>>>(p);
  return cudaGetLastError();
}

// (cin, channels per tile, taps) instantiated for C = 64 / 128 / 256 networks
#define MARU_CONV_CASES(X)                                                                   \
  X(64, 128, 9) /* stem C>=128 */ X(64, 64, 9) /* stem C=64, inner C=128 */                  \
  X(32, 32, 9)  /* inner C=64 */  X(128, 64, 9) /* inner C=256 (weights 147 KB) */           \
  X(256, 128, 1) X(256, 32, 1) X(128, 128, 1) X(128, 64, 1) X(128, 32, 1)                    \
  X(64, 128, 1) X(64, 64, 1) X(64, 32, 1) X(32, 64, 1) X(32, 32, 1)

// per-device: raise the dynamic shared-memory limit of every instantiation
cudaError_t configure_conv_gemm() {
#define X(CIN_, NT_, TAPS_)                                        \
  {                                                                \
    cudaError_t e = configure_cfg<CIN_, NT_, TAPS_>();             \
    if (e != cudaSuccess) return e;                                \
  }
  MARU_CONV_CASES(X)
#undef X
  return cudaSuccess;
}

// cin: input channels, nt: output channels per tile (cout_total % nt == 0), taps: 1 or 9
cudaError_t launch_conv_gemm(const ConvParams& p, int cin, int nt, int taps, int sm_count,
                             cudaStream_t stream) {
#define X(CIN_, NT_, TAPS_) \
  if (cin == CIN_ && nt == NT_ && taps == TAPS_) return launch_cfg<CIN_, NT_, TAPS_>(p, sm_count, stream);
  MARU_CONV_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace maru
