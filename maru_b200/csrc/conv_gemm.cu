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
// Roles (352 threads, persistent over M tiles of 128 rows, 1 CTA per SM):
//   warp 0    : producer  — weights once (one bulk copy), then per tile the input slab, the mask and
//               (if any) the residual tile into an mbarrier ring of <= 4 stages: one bulk copy per
//               8-channel plane for chunk-planar tensors, ONE copy for tile-blocked tensors
//   warps 1,10: elected-lane tcgen05.mma issue, alternating tiles / TMEM accumulators (bias and
//               residual are MMAs too); warp 1 also owns the TMEM allocation
//   warps 2-9 : epilogue  — tcgen05.ld, bf16 pack, packed ReLU, mask select, 16-byte stores
#include "conv_cfg.cuh"

namespace maru {

// Per-tile cross-layer flags (a conv waits only for the producer tiles it reads instead of the grid dependency):
// correct but measured ~20 % slower, kept for A/B in -DMARU_EXPERIMENTS builds.  In the product build
// both pointers fold to nullptr: the polling code and the signaller warp's loop compile away.
#ifdef MARU_EXPERIMENTS
#define GEMM_IN_FLAGS(p) ((p).in_flags)
#define GEMM_OUT_FLAGS(p) ((p).out_flags)
#else
#define GEMM_IN_FLAGS(p) (static_cast<const unsigned*>(nullptr))
#define GEMM_OUT_FLAGS(p) (static_cast<unsigned*>(nullptr))
#endif


// Epilogue work is pushed onto the (otherwise ~70 % idle) tensor pipe:
//   * bias     : D  = ones[128 x 16] . biasB[NT x 16]^T   (bias split into bf16 hi + lo, K cols 0/1)
//   * residual : D += res[128 x NT]  . I[NT x NT]^T       (bf16 x 1.0 is exact in the fp32 accumulator)
//   * conv     : D += sum_t slab_t[128 x CIN] . W_t[NT x CIN]^T
// so that the 8 epilogue warps only do  tcgen05.ld -> cvt.bf16x2 -> max.bf16x2 -> mask select -> st.
template <int CIN, int NT, int TAPS>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_gemm_kernel(const ConvParams p) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int MAXS = Cfg::MAX_STAGES;
  const bool has_res = p.res != nullptr;
  const int STAGES = p.stages;
  const int STAGE_BYTES = Cfg::SLAB_BYTES + Cfg::MASK_BYTES + (has_res ? Cfg::RES_BYTES : 0);
  uint8_t* s_w = smem;
  uint8_t* s_biasb = s_w + Cfg::W_BYTES;               // directly behind the weights: [W | bias image] is one blob
  uint8_t* s_ones = s_biasb + Cfg::BIASB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + Cfg::ONES_BYTES);
  uint8_t* s_ident = reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES;   // keeps stages 128-B aligned (TMA)
  uint8_t* s_stage = s_ident + (has_res ? Cfg::IDENT_BYTES : 0);
  // "stage landed" barriers come in TWO sets, one per MMA-issuing warp (tiles alternate between the two
  // warps): a warp that only sees every second tile would otherwise skip phases of a shared barrier
  // whenever the stage count is odd — its parity wait for tile it+2 is already satisfied by the phase
  // of tile it, before the data of it+2 (or even it+1) has landed. Found by the full-size determinism
  // test on the 1-stage 128->64 3x3 layers of the C=256 net.
  uint64_t* full = bars;                   // [2][MAXS] input slab + mask (+ residual tile) of an even / odd tile landed
  uint64_t* empty = bars + 2 * MAXS;       // [MAXS] stage reusable
  uint64_t* tfull = bars + 3 * MAXS;       // [2] accumulator ready
  uint64_t* tempty = tfull + 2;            // [2] accumulator drained
  uint64_t* wbar = tempty + 2;             // weights landed
  // number of (tile, epilogue warp) store completions so far: a counter, not an mbarrier, because the
  // flag signaller may fall more than one phase behind the epilogue
  volatile uint32_t* stored = reinterpret_cast<volatile uint32_t*>(wbar + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);   // last slot of the 256-byte barrier block
  // uses of full[tile parity][stage] repeat every STAGES tiles (even count) or 2*STAGES tiles (odd count)
  const int full_period = (STAGES & 1) ? 2 * STAGES : STAGES;

  pdl_launch_dependents();                      // let the next layer's CTAs start their prologue
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // timeline (bring-up): first and last CTA of group 0 stamp entry / prologue / dependency / first
  // operands / first accumulator / exit with the global nanosecond timer
  long long* tl = nullptr;
  if (p.timeline != nullptr && blockIdx.y == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    tl = p.timeline + ((size_t)p.seq * 2 + (blockIdx.x == 0 ? 0 : 1)) * 8;
  if (tl && threadIdx.x == 0) tl[0] = globaltimer_ns();
  const int group = blockIdx.y;                 // output-channel group of NT channels
  const int m_rows = (p.n_dev != nullptr) ? total_rows_g(*p.n_dev, p.pos_rows) : p.m_rows;
  const int n_tiles = (m_rows + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    // The weights do not depend on the previous layer: request them before anything else so the
    // ~2 us L2 round trip of the 74-147 KB image overlaps the rest of the prologue, the dependency
    // wait and the first slab load (tools/timeline.py).
    mbar_init(wbar, 1);
    fence_mbar_init();
    if (p.blob != nullptr) {
      // weights AND the bias operand of the bias MMA, prepared at load time (engine.cu): one request, and no
      // dependent global load of the bias (~0.7 us) on the prologue's critical path. The CTAs of a layer only start
      // when the previous layer's CTA has left the SM, so the prologue is fully exposed at every layer boundary.
      mbar_expect_tx(wbar, Cfg::W_BYTES + Cfg::BIASB_BYTES);
      bulk_load_1d(s_w, p.blob + (size_t)group * (Cfg::W_BYTES + Cfg::BIASB_BYTES), Cfg::W_BYTES + Cfg::BIASB_BYTES, wbar);
    } else {
    mbar_expect_tx(wbar, Cfg::W_BYTES);
    if (p.cout_total == NT) {
      bulk_load_1d(s_w, p.w, Cfg::W_BYTES, wbar);            // whole image is contiguous: one request
    } else {
      // weights of this group: [TAPS][KCH][cout_total][8] in HBM -> [TAPS][KCH][NT][8] in smem
      for (int tc = 0; tc < TAPS * Cfg::KCH; tc++) {
        const __nv_bfloat16* src = p.w + ((size_t)tc * p.cout_total + (size_t)group * NT) * 8;
        bulk_load_1d(s_w + (size_t)tc * NT * 16, src, NT * 16, wbar);
      }
    }
    }
    for (int i = 0; i < MAXS; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&full[MAXS + i], 1);
      // a stage is free once the MMA commit retired (slab, residual) and the 8 epilogue warps are
      // done with its mask
      mbar_init(&empty[i], 9);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    *stored = 0;
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    // constant MMA operands, K-major SWIZZLE_NONE images: [k-chunk][row][8 bf16]
    const int t = threadIdx.x - 64;
    for (int i = t; i < 2 * TILE_M; i += 256) {               // ones: (1, 1, 0, ...) in chunk 0
      const uint32_t w0 = (i < TILE_M) ? 0x3F803F80u : 0u;
      *reinterpret_cast<uint4*>(s_ones + i * 16) = make_uint4(w0, 0u, 0u, 0u);
    }
    for (int i = t; p.blob == nullptr && i < 2 * NT; i += 256) {   // bias: (hi, lo, 0, ...) in chunk 0
      uint32_t w0 = 0u;
      if (i < NT) {
        const float bv = p.bias[group * NT + i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
        const __nv_bfloat16 lo = __float2bfloat16_rn(bv - __bfloat162float(hi));
        w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
      }
      *reinterpret_cast<uint4*>(s_biasb + i * 16) = make_uint4(w0, 0u, 0u, 0u);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tl && threadIdx.x == 0) tl[1] = globaltimer_ns();

  // Roles are WARP-uniform and a single elected lane issues the async instructions: in a
  // thread-divergent branch ptxas wraps every UTCHMMA / UBLKCP in an ELECT + BRA.U.ANY loop
  // (measured ~115 cycles per MMA instead of 48, profiles/r1_umma_microbench.txt).
  if (warp == 0) {
    // =========================== producer ===========================
    // Without tile flags: wait for the whole previous grid. With flags the mask (written by the
    // encode kernel, several full dependencies upstream) and the residual are
    // already complete by transitivity; only the input tiles are polled below.
    if (GEMM_IN_FLAGS(p) == nullptr) pdl_wait();
    if (tl && lane == 0) tl[2] = globaltimer_ns();
    int stage = 0;
    uint32_t phase = 0;
    int tix = 0;
    unsigned fnext1 = 0u, fnext2 = 0u;   // flag values prefetched for the next / the one after next tile
    if (GEMM_IN_FLAGS(p) != nullptr) {
      const int nl = (TAPS == 9) ? 3 : 1;
      const int t0 = ((TAPS == 9) ? (int)blockIdx.x - 1 + lane : (int)blockIdx.x);
      const int t1 = t0 + (int)gridDim.x;
      fnext1 = (lane < nl && t0 >= 0 && t0 < n_tiles) ? ld_acquire_gpu(GEMM_IN_FLAGS(p) + t0) : 0u;
      fnext2 = (lane < nl && t1 >= 0 && t1 < n_tiles) ? ld_acquire_gpu(GEMM_IN_FLAGS(p) + t1) : 0u;
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tix++) {
      long long* tr = (p.trace && blockIdx.y == 0 && tix < TRACE_TILES)
                          ? p.trace + ((size_t)blockIdx.x * TRACE_TILES + tix) * TRACE_SLOTS : nullptr;
      if (tr && lane == 0) tr[0] = clock64();
      // Tiles this slab reads: tile-1 .. tile+1 for a 3x3 (24-row halo), the tile itself for a 1x1.
      // The flag loads are software-pipelined two tiles ahead (an acquire load is an L2 round trip
      // of ~1 us, longer than a tile period): fnext2 holds the value loaded two iterations ago.
      unsigned fval = 0u;
      bool fpoll = false;
      int ft = 0;
      if (GEMM_IN_FLAGS(p) != nullptr) {
        const int nl = (TAPS == 9) ? 3 : 1;
        ft = (TAPS == 9) ? tile - 1 + lane : tile;
        fpoll = lane < nl && ft >= 0 && ft < n_tiles;
        fval = fnext1;                                   // loaded for this tile two iterations ago
        fnext1 = fnext2;
        const int t2 = ft + 2 * (int)gridDim.x;          // prefetch for the tile after next
        fnext2 = (lane < nl && t2 >= 0 && t2 < n_tiles) ? ld_acquire_gpu(GEMM_IN_FLAGS(p) + t2) : 0u;
      }
      mbar_wait(&empty[stage], phase ^ 1);
      if (GEMM_IN_FLAGS(p) != nullptr) {
        if (fpoll) {
          uint32_t spins = 0;
          while (fval < p.in_expect) {
            __nanosleep(128);
            fval = ld_acquire_gpu(GEMM_IN_FLAGS(p) + ft);
            if (++spins > MARU_SPIN_LIMIT) {
              printf("maru: tile flag wait timed out (block %d tile %d)\n", (int)blockIdx.x, ft);
              __trap();
            }
          }
        }
        __syncwarp();
        // no reader-side proxy fence: it would drain the bulk copies already in flight and serialise
        // the stage ring; the writer publishes with fence.proxy.async + gpu fence before the flag
      }
      if (tr && lane == 0) tr[1] = clock64();
      if (elect_one()) {
        uint64_t* fbar = &full[(tix & 1) * MAXS + stage];   // the set of the MMA warp that owns this tile
        mbar_expect_tx(fbar, STAGE_BYTES);
        const size_t row0 = (size_t)GUARD_ROWS + (size_t)tile * TILE_M;
        uint8_t* dst = s_stage + stage * STAGE_BYTES;
        if (TAPS == 1 && p.in_blocked) {
          // tile-blocked input: the whole [KCH][128][16 B] slab is contiguous
          bulk_load_1d(dst, p.in + ((size_t)tile * p.in_planes + p.in_chunk0) * (TILE_M * 8), Cfg::SLAB_BYTES,
                       fbar);
        } else {
#pragma unroll 1
          for (int c = 0; c < Cfg::KCH; c++) {
            const __nv_bfloat16* src = p.in + ((size_t)(p.in_chunk0 + c) * p.rows_alloc + row0 - Cfg::LEAD) * 8;
            bulk_load_1d(dst + c * Cfg::SLAB_ROWS * 16, src, Cfg::SLAB_ROWS * 16, fbar);
          }
        }
        bulk_load_1d(dst + Cfg::SLAB_BYTES, p.mask + (size_t)tile * TILE_M, Cfg::MASK_BYTES, fbar);
        if (has_res) {
          uint8_t* rdst = dst + Cfg::SLAB_BYTES + Cfg::MASK_BYTES;
          const int rc0 = p.res_chunk0 + group * (NT / 8);
          if (p.res_blocked) {
            bulk_load_1d(rdst, p.res + ((size_t)tile * p.res_planes + rc0) * (TILE_M * 8), Cfg::RES_BYTES,
                         fbar);
          } else {
#pragma unroll 1
            for (int c = 0; c < NT / 8; c++)
              bulk_load_1d(rdst + c * TILE_M * 16, p.res + ((size_t)(rc0 + c) * p.rows_alloc + row0) * 8,
                           TILE_M * 16, fbar);
          }
        }
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 || warp == 10) {
    // =========================== MMA issuers ===========================
    // Two warps alternate tiles (and TMEM accumulators): the ~270 cycles of mbarrier polling before a
    // tile and the blocking issue of its 37 MMAs (tools/trace_conv.py) overlap across the two.
    const int mw = (warp == 1) ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, NT, 0, 0);
    // descriptors differ only in the 14-bit start-address field (smem < 256 KB: no carry)
    const uint64_t a_d = umma_desc(0, Cfg::SLAB_ROWS * 16, 128);   // slab:      LBO = slab plane
    const uint64_t r_d = umma_desc(0, TILE_M * 16, 128);           // res, ones: LBO = 128 rows
    const uint64_t b_d = umma_desc(0, NT * 16, 128);               // W, biasB, I: LBO = NT rows
    const uint32_t a_hi = (uint32_t)(a_d >> 32), r_hi = (uint32_t)(r_d >> 32), b_hi = (uint32_t)(b_d >> 32);
    const uint32_t w_lo = (uint32_t)b_d + (smem_u32(s_w) >> 4);
    const uint32_t ones_lo = (uint32_t)r_d + (smem_u32(s_ones) >> 4);
    const uint32_t biasb_lo = (uint32_t)b_d + (smem_u32(s_biasb) >> 4);
    const uint32_t ident_lo = (uint32_t)b_d + (smem_u32(s_ident) >> 4);
    mbar_wait(wbar, 0);
    for (int it = mw, tile = blockIdx.x + mw * gridDim.x; tile < n_tiles; tile += 2 * gridDim.x, it += 2) {
      const int acc = it & 1;                       // == mw
      const uint32_t acc_phase = (it >> 1) & 1;
      const int stage = it % STAGES;
      const uint32_t phase = (it / full_period) & 1;
      long long* tr = (p.trace && blockIdx.y == 0 && it < TRACE_TILES)
                          ? p.trace + ((size_t)blockIdx.x * TRACE_TILES + it) * TRACE_SLOTS : nullptr;
      if (tr && lane == 0) tr[2] = clock64();
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      mbar_wait(&full[mw * MAXS + stage], phase);
      if (tl && lane == 0 && it == 0) tl[3] = globaltimer_ns();
      if (tr && lane == 0) tr[3] = clock64();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st_addr = smem_u32(s_stage + stage * STAGE_BYTES);
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
          const int off = (TAPS == 9) ? ((t / 3 - 1) * p.pitch + (t % 3 - 1)) : 0;
#pragma unroll
          for (int j = 0; j < CIN / 16; j++) {
            const uint32_t a16 = (2 * j) * Cfg::SLAB_ROWS + Cfg::LEAD + off;   // in 16-byte units
            const uint32_t b16 = (t * Cfg::KCH + 2 * j) * NT;
            umma_bf16_lohi<1>(d_tmem, a_lo + a16, a_hi, w_lo + b16, b_hi, idesc);
          }
        }
        umma_commit(&empty[stage]);  // slab / residual reusable once these MMAs retire
        umma_commit(&tfull[acc]);    // accumulator complete
      }
      __syncwarp();
      if (tr && lane == 0) tr[4] = clock64();
    }
  } else if (warp == 11) {
    // =========================== tile-flag signaller ===========================
    // The gpu-scope fence that publishes a finished tile to the next layer's CTAs costs hundreds of
    // cycles; it runs here, off the epilogue warps' critical path (CTA barrier -> fence -> atomic).
    if (GEMM_OUT_FLAGS(p) != nullptr) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        if (lane == 0) {
          uint32_t spins = 0;
          while (*stored < 8u * (uint32_t)(it + 1)) {
            __nanosleep(256);   // a tight LDS loop would steal shared-memory cycles from the UMMA operand fetch
            if (++spins > MARU_SPIN_LIMIT) {
              printf("maru: store counter wait timed out (block %d)\n", (int)blockIdx.x);
              __trap();
            }
          }
          fence_proxy_async_all();             // the consumers read these rows through the async proxy
          __threadfence();                     // acquire the epilogue warps' stores, publish gpu-wide
          atomicAdd(GEMM_OUT_FLAGS(p) + tile, 8u);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    constexpr int CW = NT / 2;               // columns per warp
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;        // which half of the NT columns
    int it = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int trow = lg * 32 + lane;
      const int row = tile * TILE_M + trow;
      const bool valid = row < m_rows;
      const size_t grow = (size_t)GUARD_ROWS + row;
      const uint8_t* s_mask = s_stage + stage * STAGE_BYTES + Cfg::SLAB_BYTES;
      long long* tr = (p.trace && blockIdx.y == 0 && it < TRACE_TILES && warp == 2 && lane == 0)
                          ? p.trace + ((size_t)blockIdx.x * TRACE_TILES + it) * TRACE_SLOTS : nullptr;
      if (tr) tr[5] = clock64();
      mbar_wait(&full[acc * MAXS + stage], (it / full_period) & 1);   // mask landed with the slab
      const bool on = valid && (s_mask[trow] != 0);
      mbar_wait(&tfull[acc], acc_phase);
      if (tl && it == 0 && warp == 2 && lane == 0) tl[4] = globaltimer_ns();
      if (tr) tr[6] = clock64();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * NT + half * CW;
      const int chunk0 = p.out_chunk0 + group * (NT / 8) + half * (CW / 8);
      // consecutive planes of this row are `pstride` elements apart in either layout
      const size_t pstride = p.out_blocked ? (size_t)TILE_M * 8 : (size_t)p.rows_alloc * 8;
      __nv_bfloat16* out_row = p.out_blocked
                                   ? p.out + (((size_t)tile * p.out_planes + chunk0) * TILE_M + trow) * 8
                                   : p.out + ((size_t)chunk0 * p.rows_alloc + grow) * 8;
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
            if (p.relu) pk = relu_bf16x2(pk);
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
              if (p.relu) pk = relu_bf16x2(pk);
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
        mbar_arrive(&tempty[acc]);
        mbar_arrive(&empty[stage]);
        if (GEMM_OUT_FLAGS(p) != nullptr) {
          __threadfence_block();                                   // this warp's stores before the count
          atomicAdd_block(const_cast<uint32_t*>(stored), 1u);
        }
      }
      if (tr) tr[7] = clock64();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tl && threadIdx.x == 0) tl[5] = globaltimer_ns();
  if (warp == 1) {
    tc_fence_after();
    if (tl && threadIdx.x == 32) tl[7] = globaltimer_ns();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    if (tl && threadIdx.x == 32) tl[6] = globaltimer_ns();     // after the TMEM release: the CTA's last instruction
  }
}

// ------------------------------------------------------------------------------------------
template <int CIN, int NT, int TAPS>
static cudaError_t configure_cfg() {
  cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<CIN, NT, TAPS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ConvCfg<CIN, NT, TAPS>::SMEM_LIMIT);
  if (e != cudaSuccess) return e;
  // every kernel of the forward asks for the same (maximum) shared-memory carveout so that the SMs
  // never have to be drained and re-partitioned between consecutive launches
  return cudaFuncSetAttribute(conv_gemm_kernel<CIN, NT, TAPS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

template <int CIN, int NT, int TAPS>
static cudaError_t launch_cfg(ConvParams p, int sm_count, cudaStream_t stream) {
  using Cfg = ConvCfg<CIN, NT, TAPS>;
  const bool has_res = p.res != nullptr;
  p.stages = Cfg::stages(has_res);
  if (p.stages < 1) return cudaErrorInvalidConfiguration;   // e.g. stem shape with a residual
  const int groups = p.cout_total / NT;
  const int n_tiles = (p.m_rows + TILE_M - 1) / TILE_M;
  int gx = sm_count / groups;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  dim3 grid(gx, groups, 1);
  return launch_kernel(conv_gemm_kernel<CIN, NT, TAPS>, grid, dim3(CONV_THREADS), Cfg::smem_bytes(has_res),
                       stream, p.pdl != 0, p);
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
