// conv_flow.cu — persistent multi-layer conv kernel without grid barriers (sm_100a).
//
// Runs a CHAIN of consecutive conv layers (everything between two attention layers of the trunk) in ONE
// kernel.  The layer math is the one of conv_gemm.cu (tcgen05 implicit GEMM over the padded-flat row
// geometry, bias as an extra MMA, fused ReLU + board-mask epilogue).  What changes is the dataflow:
//
//   * every CTA owns a CONTIGUOUS range of M tiles [t0, t0+k) for all layers.  A 3x3 tile of layer l+1 reads
//     tiles t-1, t, t+1 of layer l, so all dependencies but the two range ends are CTA-local: the epilogue
//     warps bump a shared-memory counter per tile, the producer warp polls it before requesting the slab.
//     The range ends depend on the neighbouring CTA's edge tile: two gpu-scope flags per CTA and layer.
//   * tiles are processed OUTSIDE-IN (first, last, second, second-to-last, ...): the edge tiles the
//     neighbours wait for are finished first, and the tile finished last (the middle one) is only needed by
//     the third item of the next layer — the MMA pipe never drains at a layer boundary.
//   * weights are double-buffered (two regions, layer l uses region l & 1): the blob of layer l+1 is requested
//     as soon as the last MMA of layer l-1 has retired and lands while layer l computes.
//   * input slabs live in a byte ring (different layers have different slab sizes); a slab is released by the
//     MMA commit alone because the board mask of the CTA's tiles is loaded once, and the residual is added
//     by the epilogue warps (prefetched from global memory before the accumulator is ready).
//   * TMEM, mbarriers and the constant bias operand are set up once per kernel.
//
// tools/timeline.py on the one-launch-per-layer path: of the ~10.2 us a 3x3 64->64 layer takes at batch
// 256, 6.5 us are MMAs; the rest is the launch boundary (prologue, grid dependency, first slab, last
// epilogue) that programmatic dependent launch cannot hide because two ~190 KB CTAs are never co-resident.
//
// Co-residency: one CTA per SM, launched without the cooperative attribute.  griddepcontrol.launch_dependents
// is only issued when a CTA starts its last layer, so no later kernel can occupy an SM before every CTA of
// the chain is resident.  All waits are bounded (MARU_SPIN_LIMIT) and trap instead of hanging.
#include <stdio.h>
#include <stdlib.h>

#include "conv_cfg.cuh"

namespace maru {

// Timing experiments inside the kernel (FlowParams::dbg, tools/flow_ab.py: no dependency waits, no MMAs, no
// stores, ... — results are WRONG with any of them) exist only in -DMARU_EXPERIMENTS builds; the product
// build folds the bits to 0 and the branches disappear.
#ifdef MARU_EXPERIMENTS
#define FLOW_DBG(fp) ((fp).dbg)
#else
#define FLOW_DBG(fp) 0
#endif

constexpr int FLOW_THREADS = 448;     // warp 0 producer, 1 and 10 MMA issue, 2..9 epilogue, 11 weights, 12 and 13 publishers
constexpr int FLOW_SMEM = 232448;     // 227 KB
constexpr int FLOW_SLOTS = 4;         // input slabs in flight (barrier slots of the byte ring)
// Epilogue organisation.  1: the eight epilogue warps share every tile (two warps per TMEM lane quarter, half of the
// columns each) and tiles drain one after the other.  2: two groups of four warps (warps 2-5 even items, 6-9 odd items),
// a thread owns all columns of its row: two tiles drain CONCURRENTLY, which hides the latency chain of a drain
// (residual load -> accumulator wait -> tcgen05.ld -> stores -> count) behind the other group's.
#ifndef FLOW_EPI_GROUPS
#define FLOW_EPI_GROUPS 2
#endif
constexpr uint32_t FLOW_EPI_WARPS = 8 / FLOW_EPI_GROUPS;   // warps that drain one tile (= arrivals per accumulator / item)
constexpr int FLOW_CTRL_BYTES = 1024;
constexpr int FLOW_ONES_BYTES = 2 * TILE_M * 16;
constexpr int FLOW_MASK_BYTES = FLOW_MAX_TILES * TILE_M;
constexpr int FLOW_FIXED_BYTES = FLOW_CTRL_BYTES + FLOW_ONES_BYTES + FLOW_MASK_BYTES;

// (cin, channels per tile, taps) the flow kernel can run: flow_mma dispatches on the triple,
// flow_epilogue on nt
#define MARU_FLOW_CASES(X)                                                           \
  X(64, 64, 9) X(32, 32, 9)                                                          \
  X(256, 128, 1) X(256, 32, 1) X(128, 128, 1) X(128, 64, 1) X(128, 32, 1)            \
  X(64, 128, 1) X(64, 64, 1) X(64, 32, 1) X(32, 64, 1) X(32, 32, 1)

// control block (FLOW_CTRL_BYTES at smem + 2 * w_region): byte offsets of the mbarriers and the small arrays
constexpr uint32_t CB_FULL = 0;        // [4] slab of item g (slot g % 4) landed
constexpr uint32_t CB_EMPTY = 32;      // [4] slab of item g consumed (MMA commit)
constexpr uint32_t CB_TFULL = 64;      // [4] accumulator ready
constexpr uint32_t CB_TEMPTY = 96;     // [4] accumulator drained (8 epilogue warps)
constexpr uint32_t CB_WBAR = 128;      // [2] weight blob landed in region r
constexpr uint32_t CB_MDONE = 144;     // [2] both MMA warps' last MMAs of layer l retired (barrier l & 1: alternating,
                                       //     so that the waiter can never fall two phases behind however short a layer is)
constexpr uint32_t CB_SLOT_OFF = 160;  // [4] u32 byte offset of the slab of the item in slot s
constexpr uint32_t CB_TMEM = 176;      // u32 TMEM base
constexpr uint32_t CB_RAW = 256;       // [FLOW_MAX_TILES] u32 epilogue-warp completions per item index (FLOW_EPI_WARPS per layer): stored
constexpr uint32_t CB_DONE = 448;      // [FLOW_MAX_TILES] u32 the same count once a publisher warp has fenced: readable by bulk copies

// shared-memory helpers on 32-bit shared addresses (no generic 64-bit pointers in the hot loops)
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.volatile.shared::cta.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_inc(uint32_t a) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_init_a(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}
__device__ __noinline__ void flow_timeout(int what, int l, int i, uint32_t have, uint32_t need) {
  if ((threadIdx.x & 31) == 0)
    printf("maru: flow kernel wait timed out (block %d warp %d: %s, layer %d item %d, %u / %u)\n", (int)blockIdx.x,
           (int)(threadIdx.x >> 5), what == 0 ? "mbarrier" : what == 1 ? "tile counter" : "neighbour flag", l, i, have, need);
  __trap();
}
__device__ __forceinline__ void mbar_wait_a(uint32_t a, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > MARU_SPIN_LIMIT) flow_timeout(0, -1, (int)a, parity, 0);
  }
}
__device__ __forceinline__ void umma_commit_a(uint32_t a) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void bulk_load_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// item index (processing order) -> local tile, outside-in; and back
__device__ __forceinline__ int flow_tile_of(int i, int k) { return (i & 1) ? k - 1 - (i >> 1) : (i >> 1); }
__device__ __forceinline__ int flow_item_of(int j, int k) { return (j < (k + 1) / 2) ? 2 * j : 2 * (k - 1 - j) + 1; }

// relaxed polling; the caller issues ONE acquire fence after all the counters it needs have been seen
__device__ __forceinline__ void flow_wait_done(uint32_t ctr, uint32_t need, int l, int i) {
  uint32_t spins = 0;
  while (lds_u32(ctr) < need) {
    __nanosleep(32);
    if (++spins > MARU_SPIN_LIMIT) flow_timeout(1, l, i, lds_u32(ctr), need);
  }
}
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;\n" ::: "memory"); }
__device__ __forceinline__ void flow_wait_flag(const unsigned* flag, unsigned need, int l, int i) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(flag) < need) {
    __nanosleep(128);
    if (++spins > MARU_SPIN_LIMIT) flow_timeout(2, l, i, ld_acquire_gpu(flag), need);
  }
}
// The blob of a 3x3 layer is 74 KB: as ONE bulk copy it occupies the SM's copy engine for ~0.9 us and the slab requests
// of the tiles being computed queue behind it; in 4 KB pieces the slab requests issued meanwhile slip in between.
__device__ __forceinline__ void flow_request_blob(const FlowLayer& N, uint32_t dst, uint32_t bar) {
  const uint32_t blob = (uint32_t)(N.taps * N.cin * N.nt * 2 + 2 * N.nt * 16);
  const uint8_t* src = N.blob;
  mbar_expect_tx_a(bar, blob);
#pragma unroll 1
  for (uint32_t off = 0; off < blob; off += 4096u) {
    const uint32_t n = (blob - off < 4096u) ? blob - off : 4096u;
    bulk_load_a(dst + off, src + off, n, bar);
  }
}

// what every role needs to know about the CTA
struct FlowCta {
  uint32_t smem;       // shared address of the dynamic shared memory (weight region 0)
  uint32_t ctrl;       // shared address of the control block
  uint32_t area;       // shared address of the slab ring
  uint32_t tmem_base;
  int cta, t0, k, n_tiles, m_rows;
  long long* tr;       // bring-up: per-item %globaltimer stamps of one CTA: [item g][8] = producer {dependencies met,
                       // slab requested}, MMA warp {operands landed, MMAs issued}, epilogue warp 2 {accumulator
                       // ready, stored, fenced + counted}
};

// ------------------------------------------------------------------------------------------ producer (warp 0)
__device__ __noinline__ void flow_producer(const FlowParams& fp, uint32_t c_smem, uint32_t c_ctrl, uint32_t c_area, uint32_t c_tmem, int c_cta, int c_t0, int c_k, int c_n_tiles, int c_m_rows, long long* c_tr) {
  // scalars, not a struct: a by-value struct argument of a non-inlined function lives on the local-memory stack
  FlowCta c;
  c.smem = c_smem; c.ctrl = c_ctrl; c.area = c_area; c.tmem_base = c_tmem; c.cta = c_cta; c.t0 = c_t0; c.k = c_k;
  c.n_tiles = c_n_tiles; c.m_rows = c_m_rows; c.tr = c_tr;
  const int lane = threadIdx.x & 31;
  const int k = c.k;
  // Everything read from the kernel parameters is copied into locals first: the asm statements below clobber
  // "memory", and a parameter reached through a reference is re-loaded (generic load, ~1 us per item in the
  // first version of this kernel) after every one of them.
  const int n_layers = fp.n_layers;
  const unsigned* const flags = fp.flags;
  const unsigned flag_base = fp.flag_base;
  const uint32_t area_bytes = (uint32_t)fp.area;
  const size_t pstride = (size_t)fp.rows_alloc * 8;
  const int dbg = FLOW_DBG(fp);
  long long* const stamps = (c.cta == 0) ? fp.stamps : nullptr;
  uint32_t g = 0, head = 0;
  uint32_t o1 = 0, e1 = 0, o2 = 0, e2 = 0, o3 = 0, e3 = 0;
  // the fields of layer l+1 are fetched while layer l is being requested (see flow_mma)
  int nx_taps = fp.layer[0].taps, nx_cin = fp.layer[0].cin, nx_planes = fp.layer[0].in_planes;
  int nx_blocked = fp.layer[0].in_blocked, nx_dep = fp.layer[0].dep_back;
  const __nv_bfloat16* nx_in = fp.layer[0].in;
  for (int l = 0; l < n_layers; l++) {
    const int taps = nx_taps, cin = nx_cin;
    const __nv_bfloat16* const in = nx_in;
    const size_t in_tile_stride = (size_t)nx_planes * (TILE_M * 8);
    const bool blocked = (taps == 1) && nx_blocked;
    const int dep_back = nx_dep;
    if (l + 1 < n_layers) {
      nx_taps = fp.layer[l + 1].taps;
      nx_cin = fp.layer[l + 1].cin;
      nx_planes = fp.layer[l + 1].in_planes;
      nx_blocked = fp.layer[l + 1].in_blocked;
      nx_dep = fp.layer[l + 1].dep_back;
      nx_in = fp.layer[l + 1].in;
    }
    const int lead = (taps == 9) ? 24 : 0;
    const uint32_t slab_rows = TILE_M + 2 * lead;
    const uint32_t slab_bytes = slab_rows * cin * 2;
    const int kch = cin / 8;
    if (stamps != nullptr && lane == 0) stamps[4 * l] = globaltimer_ns();
    for (int i = 0; i < k; i++, g++) {
      const int j = flow_tile_of(i, k);
      const int tile = c.t0 + j;
      if (dep_back > 0 && !(dbg & 4)) {
        // this tile (and for a 3x3 its two neighbours) of the layer that wrote the input must be stored and
        // published. Usually that is the previous layer; the channel groups of one convolution (q, k, v; the two
        // halves of the stem) all read the same input and do not wait for each other.
        const uint32_t need = 8u * (uint32_t)(l - dep_back + 1);
        const uint32_t done = c.ctrl + CB_DONE;
        if (c.tr != nullptr && lane == 0 && g < 256) c.tr[g * 8 + 7] = globaltimer_ns();
        flow_wait_done(done + 4 * i, need, l, i);
        if (taps == 9) {
          if (j > 0) flow_wait_done(done + 4 * flow_item_of(j - 1, k), need, l, i);
          else if (tile > 0 && !(dbg & 16)) flow_wait_flag(flags + 2 * (c.cta - 1) + 1, flag_base + (unsigned)(l - dep_back + 1), l, i);
          if (j < k - 1) flow_wait_done(done + 4 * flow_item_of(j + 1, k), need, l, i);
          else if (tile + 1 < c.n_tiles && !(dbg & 16)) flow_wait_flag(flags + 2 * (c.cta + 1), flag_base + (unsigned)(l - dep_back + 1), l, i);
        }
        if (c.tr != nullptr && lane == 0 && g < 192) c.tr[(320 + g) * 8 + 3] = globaltimer_ns();
        fence_acq_rel_cta();
      }
      if (c.tr != nullptr && lane == 0 && g < 256) c.tr[g * 8 + 0] = globaltimer_ns();
      // ring allocation: the slab must not overlap a slab still in flight, and its barrier slot must be free.
      // (o1,e1) .. (o3,e3) are the byte ranges of the previous three items.
      uint32_t start = head;
      if (start + slab_bytes > area_bytes) start = 0;
      const uint32_t end = start + slab_bytes;
      if (g >= FLOW_SLOTS) mbar_wait_a(c.ctrl + CB_EMPTY + 8 * (g & 3), ((g - 4) >> 2) & 1u);
      if (g >= 3 && o3 < end && start < e3) mbar_wait_a(c.ctrl + CB_EMPTY + 8 * ((g - 3) & 3), ((g - 3) >> 2) & 1u);
      if (g >= 2 && o2 < end && start < e2) mbar_wait_a(c.ctrl + CB_EMPTY + 8 * ((g - 2) & 3), ((g - 2) >> 2) & 1u);
      if (g >= 1 && o1 < end && start < e1) mbar_wait_a(c.ctrl + CB_EMPTY + 8 * ((g - 1) & 3), ((g - 1) >> 2) & 1u);
      o3 = o2; e3 = e2; o2 = o1; e2 = e1; o1 = start; e1 = end;
      if (c.tr != nullptr && lane == 0 && g < 192) c.tr[(320 + g) * 8 + 4] = globaltimer_ns();
      if (lane == 0) sts_u32(c.ctrl + CB_SLOT_OFF + 4 * (g & 3), start);
      __syncwarp();
      if (c.tr != nullptr && lane == 0 && g < 192) c.tr[(320 + g) * 8 + 5] = globaltimer_ns();
      head = end;
      if (dbg & 8) {                 // timing experiment: no slab traffic at all, MMAs run on stale operands
        if (elect_one()) mbar_arrive_a(c.ctrl + CB_FULL + 8 * (g & 3));
      } else if (elect_one()) {
        const uint32_t fbar = c.ctrl + CB_FULL + 8 * (g & 3);
        mbar_expect_tx_a(fbar, slab_bytes);
        const uint32_t dst = c.area + start;
        if (blocked) {
          bulk_load_a(dst, in + (size_t)tile * in_tile_stride, slab_bytes, fbar);
        } else {
          const __nv_bfloat16* src = in + ((size_t)GUARD_ROWS + (size_t)tile * TILE_M - lead) * 8;
#pragma unroll 1
          for (int ch = 0; ch < kch; ch++) bulk_load_a(dst + ch * slab_rows * 16, src + ch * pstride, slab_rows * 16, fbar);
        }
      }
      __syncwarp();
      if (c.tr != nullptr && lane == 0 && g < 256) c.tr[g * 8 + 1] = globaltimer_ns();
    }
  }
}

// ------------------------------------------------------------------------------------------ MMA issue (warps 1, 10)
// All MMAs of one tile from the elected lane: D = bias (ones x biasB), then taps x K-steps; the A offsets are
// immediates, the B descriptor advances by 2*NT 16-byte units per MMA ([tap][k-chunk][nt][8] weight image).
template <int CIN, int NT, int TAPS>
__device__ __forceinline__ void flow_issue(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t w_lo, uint32_t b_hi,
                                           uint32_t ones_lo, uint32_t r_hi, uint32_t biasb_lo, int pitch) {
  constexpr int LEAD = (TAPS == 9) ? 24 : 0;
  constexpr int SLAB_ROWS = TILE_M + 2 * LEAD;
  constexpr int KCH = CIN / 8;
  constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, NT, 0, 0);
  umma_bf16_lohi<0>(d_tmem, ones_lo, r_hi, biasb_lo, b_hi, idesc);           // D = bias
#pragma unroll
  for (int t = 0; t < TAPS; t++) {
    const int toff = (TAPS == 9) ? ((t / 3 - 1) * pitch + (t % 3 - 1)) : 0;
#pragma unroll
    for (int jj = 0; jj < CIN / 16; jj++) {
      const uint32_t a16 = (2 * jj) * SLAB_ROWS + LEAD + toff;   // in 16-byte units
      const uint32_t b16 = (t * KCH + 2 * jj) * NT;
      umma_bf16_lohi<1>(d_tmem, a_lo + a16, a_hi, w_lo + b16, b_hi, idesc);
    }
  }
}

__device__ __noinline__ void flow_mma(const FlowParams& fp, uint32_t c_smem, uint32_t c_ctrl, uint32_t c_area, uint32_t c_tmem, int c_cta, int c_t0, int c_k, int c_n_tiles, int c_m_rows, long long* c_tr, const uint32_t mw) {
  // scalars, not a struct: a by-value struct argument of a non-inlined function lives on the local-memory stack
  FlowCta c;
  c.smem = c_smem; c.ctrl = c_ctrl; c.area = c_area; c.tmem_base = c_tmem; c.cta = c_cta; c.t0 = c_t0; c.k = c_k;
  c.n_tiles = c_n_tiles; c.m_rows = c_m_rows; c.tr = c_tr;
  const int lane = threadIdx.x & 31;
  const int k = c.k;
  const uint32_t ones = c.ctrl + FLOW_CTRL_BYTES;
  const uint64_t r_d = umma_desc(0, TILE_M * 16, 128);             // ones:  LBO = 128 rows
  const uint32_t r_hi = (uint32_t)(r_d >> 32);
  const uint32_t ones_lo = (uint32_t)r_d + (ones >> 4);
  const int n_layers = fp.n_layers;
  const int dbg = FLOW_DBG(fp);
  const int pitch = fp.pitch;
  const uint32_t w_region = (uint32_t)fp.w_region;
  long long* const stamps = (c.cta == 0 && mw == 0) ? fp.stamps : nullptr;
  uint32_t g = 0;
  // the shape of layer l+1 is fetched while layer l runs: reached through the parameter reference these are generic
  // loads (~0.7 us), and at a layer boundary they would sit between two tiles' MMAs
  int nx_taps = fp.layer[0].taps, nx_cin = fp.layer[0].cin, nx_nt = fp.layer[0].nt;
  for (int l = 0; l < n_layers; l++) {
    const int taps = nx_taps, cin = nx_cin, nt = nx_nt;
    if (l + 1 < n_layers) {
      nx_taps = fp.layer[l + 1].taps;
      nx_cin = fp.layer[l + 1].cin;
      nx_nt = fp.layer[l + 1].nt;
    }
    const uint32_t slab_rows = (taps == 9) ? TILE_M + 48 : TILE_M;
    const uint64_t a_d = umma_desc(0, slab_rows * 16, 128);        // slab:  LBO = slab plane
    const uint64_t b_d = umma_desc(0, nt * 16, 128);               // W, biasB: LBO = NT rows
    const uint32_t a_hi = (uint32_t)(a_d >> 32), b_hi = (uint32_t)(b_d >> 32);
    const uint32_t s_w = c.smem + (l & 1) * w_region;
    const uint32_t w_lo = (uint32_t)b_d + (s_w >> 4);
    const uint32_t biasb_lo = (uint32_t)b_d + ((s_w + taps * cin * nt * 2) >> 4);
    mbar_wait_a(c.ctrl + CB_WBAR + 8 * (l & 1), ((uint32_t)l >> 1) & 1u);   // this layer's weights + bias image
    if (stamps != nullptr && lane == 0) stamps[4 * l + 1] = globaltimer_ns();
    for (int i = 0; i < k; i++, g++) {
      if ((g & 1u) != mw) continue;
      // FOUR accumulators (all 512 TMEM columns): with two, an issuing warp waited for the epilogue of its
      // previous tile before every tile (issue 1.3 us + drain 0.8 us per accumulator = 1.25 us per tile)
      const uint32_t acc = g & 3u;
      mbar_wait_a(c.ctrl + CB_TEMPTY + 8 * acc, ((g >> 2) & 1u) ^ 1u);
      mbar_wait_a(c.ctrl + CB_FULL + 8 * (g & 3), (g >> 2) & 1u);
      const uint32_t off = lds_u32(c.ctrl + CB_SLOT_OFF + 4 * (g & 3));
      if (c.tr != nullptr && lane == 0 && g < 256) c.tr[g * 8 + 2] = globaltimer_ns();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = (uint32_t)a_d + ((c.area + off) >> 4);
        const uint32_t d_tmem = c.tmem_base + acc * 128u;
        // The weight descriptors are the same for every tile of a layer; left alone, the compiler hoists all 37
        // of them out of the tile loop into vector registers and feeds every UTCHMMA through R2UR moves and
        // uniform-register spills (~65 cycles per MMA). Made opaque here, they are one base register plus an
        // immediate per MMA, like the slab descriptors.
        uint32_t w_lo_i = w_lo, biasb_lo_i = biasb_lo, ones_lo_i = ones_lo;
        asm volatile("" : "+r"(w_lo_i), "+r"(biasb_lo_i), "+r"(ones_lo_i));
#define X(CIN_, NT_, TAPS_)                                \
        else if (cin == CIN_ && nt == NT_ && taps == TAPS_) \
          flow_issue<CIN_, NT_, TAPS_>(d_tmem, a_lo, a_hi, w_lo_i, b_hi, ones_lo_i, r_hi, biasb_lo_i, pitch);
        if (dbg & 32) {}      // timing experiment: no MMAs at all
        MARU_FLOW_CASES(X)
#undef X
        umma_commit_a(c.ctrl + CB_EMPTY + 8 * (g & 3));   // slab reusable once these MMAs retire
        umma_commit_a(c.ctrl + CB_TFULL + 8 * acc);       // accumulator complete
      }
      __syncwarp();
      if (c.tr != nullptr && lane == 0 && g < 256) c.tr[g * 8 + 3] = globaltimer_ns();
    }
    // both issuers' last MMAs of the layer retired => the weight region may be overwritten
    if (elect_one()) umma_commit_a(c.ctrl + CB_MDONE + 8 * (l & 1));
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ weights + neighbour flags (warp 11)
__device__ __noinline__ void flow_aux(const FlowParams& fp, uint32_t c_smem, uint32_t c_ctrl, uint32_t c_area, uint32_t c_tmem, int c_cta, int c_t0, int c_k, int c_n_tiles, int c_m_rows, long long* c_tr) {
  // scalars, not a struct: a by-value struct argument of a non-inlined function lives on the local-memory stack
  FlowCta c;
  c.smem = c_smem; c.ctrl = c_ctrl; c.area = c_area; c.tmem_base = c_tmem; c.cta = c_cta; c.t0 = c_t0; c.k = c_k;
  c.n_tiles = c_n_tiles; c.m_rows = c_m_rows; c.tr = c_tr;
  const int lane = threadIdx.x & 31;
  const int k = c.k;
  const int n_layers = fp.n_layers;
  const uint32_t w_region = (uint32_t)fp.w_region;
  unsigned* const flags = fp.flags;
  const unsigned flag_base = fp.flag_base;
  for (int l = 0; l + 1 < n_layers; l++) {
    if (l >= 1) {
      // region (l+1)&1 was read by layer l-1: fetch the blob of layer l+1 once those MMAs have retired
      mbar_wait_a(c.ctrl + CB_MDONE + 8 * ((l - 1) & 1), ((uint32_t)(l - 1) >> 1) & 1u);
      if (elect_one())
        flow_request_blob(fp.layer[l + 1], c.smem + ((l + 1) & 1) * w_region, c.ctrl + CB_WBAR + 8 * ((l + 1) & 1));
      __syncwarp();
    }
    // The two edge tiles of the layer (item 0 = first tile, item 1 = last tile of the CTA's range), once
    // published inside the CTA, go to the neighbouring CTAs' producers at gpu scope. The gpu-scope release
    // takes ~2 us: it lives here, off the publishers' path (the neighbours need it k-2 items later).
    const uint32_t need = 8u * (uint32_t)(l + 1);
    const unsigned val = flag_base + (unsigned)(l + 1);
    flow_wait_done(c.ctrl + CB_DONE, need, l, -1);
    if (lane == 0) {
      __threadfence();
      st_release_gpu(flags + 2 * c.cta, val);
      if (k == 1) st_release_gpu(flags + 2 * c.cta + 1, val);
      if (c.tr != nullptr && l < 64) c.tr[(256 + l) * 8 + 1] = globaltimer_ns();
    }
    if (k > 1) {
      flow_wait_done(c.ctrl + CB_DONE + 4, need, l, -2);
      if (lane == 0) {
        __threadfence();
        st_release_gpu(flags + 2 * c.cta + 1, val);
        if (c.tr != nullptr && l < 64) c.tr[(256 + l) * 8 + 2] = globaltimer_ns();
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ publishers (warps 12, 13)
// The epilogue warps only count their stores (cta-scope release). A publisher (one per item parity) waits for the
// 8 counts of an item, runs the generic -> async proxy fence ONCE on behalf of the 256 storing threads (a proxy
// fence anywhere on the causality path from the writes to the bulk-copy reads is sufficient; it costs ~0.7 us,
// which used to stall every epilogue warp after every tile), and re-publishes the count for the producer warp.
__device__ __noinline__ void flow_publisher(const FlowParams& fp, uint32_t c_smem, uint32_t c_ctrl, uint32_t c_area, uint32_t c_tmem, int c_cta, int c_t0, int c_k, int c_n_tiles, int c_m_rows, long long* c_tr, const uint32_t pw) {
  // scalars, not a struct: a by-value struct argument of a non-inlined function lives on the local-memory stack
  FlowCta c;
  c.smem = c_smem; c.ctrl = c_ctrl; c.area = c_area; c.tmem_base = c_tmem; c.cta = c_cta; c.t0 = c_t0; c.k = c_k;
  c.n_tiles = c_n_tiles; c.m_rows = c_m_rows; c.tr = c_tr;
  const int lane = threadIdx.x & 31;
  const int k = c.k;
  const int n_layers = fp.n_layers;
  uint32_t g = 0;
  for (int l = 0; l < n_layers; l++) {
    const uint32_t need = FLOW_EPI_WARPS * (uint32_t)(l + 1);
    for (int i = 0; i < k; i++, g++) {
      if ((g & 1u) != pw) continue;
      flow_wait_done(c.ctrl + CB_RAW + 4 * i, need, l, -1 - i);
      if (lane == 0) {
        if (c.tr != nullptr && g < 192) c.tr[(320 + g) * 8 + 0] = globaltimer_ns();
        fence_acq_rel_cta();                 // acquire the epilogue warps' stores
        if (c.tr != nullptr && g < 192) c.tr[(320 + g) * 8 + 1] = globaltimer_ns();
        fence_proxy_async_all();             // ... and order them before later async-proxy reads
        if (c.tr != nullptr && g < 192) c.tr[(320 + g) * 8 + 2] = globaltimer_ns();
        asm volatile("red.release.cta.shared::cta.add.u32 [%0], 8;\n" ::"r"(c.ctrl + CB_DONE + 4 * i) : "memory");
        if (c.tr != nullptr && g < 256) c.tr[g * 8 + 6] = globaltimer_ns();
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------ epilogue (warps 2..9)
// one tile: CW accumulator columns of this thread's row -> (+ residual) -> bf16 -> ReLU -> board mask -> 16-byte stores
// (a 128-column row is drained as two calls of 64: FIRST waits for the accumulator, LAST hands it back)
template <int CW, bool FIRST = true, bool LAST = true>
__device__ __forceinline__ void flow_epilogue_tile(const int relu, const uint32_t t_addr, __nv_bfloat16* out_row,
                                                   const size_t pstride, const __nv_bfloat16* res_row,
                                                   const size_t rstride, const bool valid, const bool on,
                                                   const uint32_t tfull, const uint32_t tempty, const uint32_t parity,
                                                   long long* trs) {
  const bool has_res = res_row != nullptr;
  // residual (same channel range as the output): requested before the accumulator is ready
  uint4 rv[CW / 8];
  if (has_res) {
#pragma unroll
    for (int q = 0; q < CW / 8; q++)
      rv[q] = on ? *reinterpret_cast<const uint4*>(res_row + (size_t)q * rstride) : make_uint4(0u, 0u, 0u, 0u);
  }
  if constexpr (FIRST) {
    mbar_wait_a(tfull, parity);
    if (trs != nullptr) trs[4] = globaltimer_ns();
    tc_fence_after();
  }
  // all of this call's accumulator columns in one go (one wait); after the last read the accumulator is handed back
  // to the MMA warps, before the conversion and the stores
  uint32_t v[CW];
  if constexpr (CW == 16) {
    tmem_ld16(t_addr, v);
  } else if constexpr (CW == 32) {
    tmem_ld32(t_addr, v);
  } else {
    tmem_ld32(t_addr, v);
    tmem_ld32(t_addr + 32, v + 32);
  }
  tmem_ld_wait();
  if constexpr (LAST) {
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_a(tempty);
  }
#pragma unroll
  for (int q = 0; q < CW / 8; q++) {
    uint32_t o[4];
    uint32_t rw[4] = {0u, 0u, 0u, 0u};
    if (has_res) {
      rw[0] = rv[q].x;
      rw[1] = rv[q].y;
      rw[2] = rv[q].z;
      rw[3] = rv[q].w;
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float lo = __uint_as_float(v[q * 8 + 2 * e]), hi = __uint_as_float(v[q * 8 + 2 * e + 1]);
      if (has_res) {
        lo += bf16_lo(rw[e]);
        hi += bf16_hi(rw[e]);
      }
      uint32_t pk = pack_bf16x2(lo, hi);
      if (relu) pk = relu_bf16x2(pk);
      o[e] = on ? pk : 0u;
    }
    if (valid) *reinterpret_cast<uint4*>(out_row + (size_t)q * pstride) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __noinline__ void flow_epilogue(const FlowParams& fp, uint32_t c_smem, uint32_t c_ctrl, uint32_t c_area, uint32_t c_tmem, int c_cta, int c_t0, int c_k, int c_n_tiles, int c_m_rows, long long* c_tr) {
  // scalars, not a struct: a by-value struct argument of a non-inlined function lives on the local-memory stack
  FlowCta c;
  c.smem = c_smem; c.ctrl = c_ctrl; c.area = c_area; c.tmem_base = c_tmem; c.cta = c_cta; c.t0 = c_t0; c.k = c_k;
  c.n_tiles = c_n_tiles; c.m_rows = c_m_rows; c.tr = c_tr;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k = c.k;
  const int lg = warp & 3;                 // TMEM lane group this warp may access
#if FLOW_EPI_GROUPS == 2
  const uint32_t eg = (uint32_t)(warp - 2) >> 2;   // epilogue group: the items of this parity
  const int half = 0;
#else
  const int half = (warp - 2) >> 2;        // which half of the NT columns
#endif
  const int trow = lg * 32 + lane;
  const uint32_t mask = c.ctrl + FLOW_CTRL_BYTES + FLOW_ONES_BYTES;
  const int n_layers = fp.n_layers;
  const size_t rows_alloc = (size_t)fp.rows_alloc;
  const int dbg = FLOW_DBG(fp);
  long long* const stamps = (c.cta == 0 && warp == 2 && lane == 0) ? fp.stamps : nullptr;
  uint32_t g = 0;
  for (int l = 0; l < n_layers; l++) {
    __nv_bfloat16* const out = fp.layer[l].out;
    const __nv_bfloat16* const res = fp.layer[l].res;
    const int out_blocked = fp.layer[l].out_blocked, res_blocked = fp.layer[l].res_blocked;
    const size_t out_planes = (size_t)fp.layer[l].out_planes, res_planes = (size_t)fp.layer[l].res_planes;
    const int relu = fp.layer[l].relu;
    const int cw = fp.layer[l].nt * FLOW_EPI_GROUPS / 2;   // columns per thread
    const int chunk0 = fp.layer[l].out_chunk0 + half * (cw / 8);
    const size_t pstride = out_blocked ? (size_t)TILE_M * 8 : rows_alloc * 8;
    const size_t rstride = res_blocked ? (size_t)TILE_M * 8 : rows_alloc * 8;
    for (int i = 0; i < k; i++, g++) {
#if FLOW_EPI_GROUPS == 2
      if ((g & 1u) != eg) continue;
#endif
      const uint32_t acc = g & 3u;
      const int j = flow_tile_of(i, k);
      const int tile = c.t0 + j;
      const int row = tile * TILE_M + trow;
      const bool valid = row < c.m_rows && !(dbg & 64);
      const size_t grow = (size_t)GUARD_ROWS + row;
      uint32_t mv;
      asm volatile("ld.shared::cta.u8 %0, [%1];\n" : "=r"(mv) : "r"(mask + j * TILE_M + trow));
      const bool on = valid && (mv != 0);
      __nv_bfloat16* out_row = out_blocked ? out + (((size_t)tile * out_planes + chunk0) * TILE_M + trow) * 8
                                           : out + ((size_t)chunk0 * rows_alloc + grow) * 8;
      const __nv_bfloat16* res_row = nullptr;
      if (res != nullptr)
        res_row = res_blocked ? res + (((size_t)tile * res_planes + chunk0) * TILE_M + trow) * 8
                              : res + ((size_t)chunk0 * rows_alloc + grow) * 8;
      const uint32_t t_addr = c.tmem_base + ((uint32_t)(lg * 32) << 16) + acc * 128u + half * cw;
      const uint32_t tfull = c.ctrl + CB_TFULL + 8 * acc, tempty = c.ctrl + CB_TEMPTY + 8 * acc;
      long long* trs = (c.tr != nullptr && (warp == 2 || (FLOW_EPI_GROUPS == 2 && warp == 6)) && lane == 0 && g < 256) ? c.tr + g * 8 : nullptr;
      if (dbg & 128) {            // timing experiment: the epilogue only hands the accumulator back
        mbar_wait_a(tfull, (g >> 2) & 1u);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(tempty);
      } else if (cw == 128) {
        flow_epilogue_tile<64, true, false>(relu, t_addr, out_row, pstride, res_row, rstride, valid, on, tfull, tempty, (g >> 2) & 1u, trs);
        flow_epilogue_tile<64, false, true>(relu, t_addr + 64, out_row + 8 * pstride, pstride,
                                            res_row != nullptr ? res_row + 8 * rstride : nullptr, rstride, valid, on, tfull,
                                            tempty, (g >> 2) & 1u, trs);
      } else if (cw == 64)
        flow_epilogue_tile<64>(relu, t_addr, out_row, pstride, res_row, rstride, valid, on, tfull, tempty, (g >> 2) & 1u, trs);
      else if (cw == 32)
        flow_epilogue_tile<32>(relu, t_addr, out_row, pstride, res_row, rstride, valid, on, tfull, tempty, (g >> 2) & 1u, trs);
      else
        flow_epilogue_tile<16>(relu, t_addr, out_row, pstride, res_row, rstride, valid, on, tfull, tempty, (g >> 2) & 1u, trs);
      if (trs != nullptr) trs[5] = globaltimer_ns();
      __syncwarp();
      if (lane == 0) {
        if (dbg & 1) {                         // A/B: fence in the storing warp (slow, the first version)
          fence_proxy_async_all();
          __threadfence();
        }
        if (dbg & 2) asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], 1;\n" ::"r"(c.ctrl + CB_RAW + 4 * i) : "memory");
        else red_release_inc(c.ctrl + CB_RAW + 4 * i);
      }
    }
    if (stamps != nullptr) stamps[4 * l + 2] = globaltimer_ns();
  }
}

__global__ void __launch_bounds__(FLOW_THREADS, 1) conv_flow_kernel(const __grid_constant__ FlowParams fp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  FlowCta c;
  c.smem = smem_u32(smem);
  c.ctrl = c.smem + 2 * fp.w_region;
  c.area = c.ctrl + FLOW_FIXED_BYTES;
  uint8_t* ctrl_p = smem + 2 * fp.w_region;

  const int warp = threadIdx.x >> 5;
  // bring-up: [cta][8] %globaltimer stamps of every CTA: entry, dependency resolved + mask loaded, roles done, exit
  long long* const tc = (fp.trace != nullptr) ? fp.trace + 8192 + (size_t)blockIdx.x * 8 : nullptr;
  if (tc != nullptr && threadIdx.x == 0) tc[0] = globaltimer_ns();
  if (threadIdx.x == 0) {
    for (int i = 0; i < FLOW_SLOTS; i++) {
      mbar_init_a(c.ctrl + CB_FULL + 8 * i, 1);
      mbar_init_a(c.ctrl + CB_EMPTY + 8 * i, 1);
    }
    for (int i = 0; i < 4; i++) {
      mbar_init_a(c.ctrl + CB_TFULL + 8 * i, 1);
      mbar_init_a(c.ctrl + CB_TEMPTY + 8 * i, FLOW_EPI_WARPS);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init_a(c.ctrl + CB_WBAR + 8 * i, 1);
      mbar_init_a(c.ctrl + CB_MDONE + 8 * i, 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(reinterpret_cast<uint32_t*>(ctrl_p + CB_TMEM), 512);   // four accumulators of up to 128 columns
    tmem_relinquish();
  }
  if (threadIdx.x < 2 * FLOW_MAX_TILES) sts_u32(c.ctrl + CB_RAW + 4 * threadIdx.x, 0u);   // raw + published counters
  if (warp >= 2 && warp <= 9) {
    // A operand of the bias MMA, K-major SWIZZLE_NONE image [2 k-chunks][128 rows][8 bf16]: (1, 1, 0, ...)
    const int t = threadIdx.x - 64;
    const uint32_t w0 = (t < TILE_M) ? 0x3F803F80u : 0u;
    *reinterpret_cast<uint4*>(ctrl_p + FLOW_CTRL_BYTES + t * 16) = make_uint4(w0, 0u, 0u, 0u);
    fence_async_smem();              // generic stores -> visible to UMMA
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *reinterpret_cast<volatile uint32_t*>(ctrl_p + CB_TMEM);

  // tile range of this CTA (the live batch size is written before the graph starts, not by the previous kernel)
  c.m_rows = (fp.n_dev != nullptr) ? total_rows_g(*fp.n_dev, fp.pos_rows) : fp.m_rows;
  c.n_tiles = (c.m_rows + TILE_M - 1) / TILE_M;
  const int n_ctas = ((int)gridDim.x < c.n_tiles) ? (int)gridDim.x : c.n_tiles;
  c.cta = (int)blockIdx.x;
  c.t0 = 0;
  c.k = 0;
  if (c.cta < n_ctas) {
    c.t0 = (int)(((long long)c.cta * c.n_tiles) / n_ctas);
    c.k = (int)(((long long)(c.cta + 1) * c.n_tiles) / n_ctas) - c.t0;
  }
  if (c.k > FLOW_MAX_TILES) {
    if (threadIdx.x == 0) printf("maru: flow kernel launched with %d tiles per CTA (max %d)\n", c.k, FLOW_MAX_TILES);
    __trap();
  }
  c.tr = (fp.trace != nullptr && c.cta == fp.trace_cta) ? fp.trace : nullptr;
  // the weights do not depend on the previous kernel: request the first two blobs right away
  if (c.k > 0 && warp == 11 && elect_one()) {
    for (int l = 0; l < 2 && l < fp.n_layers; l++)
      flow_request_blob(fp.layer[l], c.smem + l * fp.w_region, c.ctrl + CB_WBAR + 8 * l);
  }
  __syncwarp();
  pdl_wait();                        // everything below reads what the previous kernel wrote

  // board mask of the CTA's tiles
  for (int i = threadIdx.x; i < c.k * (TILE_M / 16); i += FLOW_THREADS)
    reinterpret_cast<uint4*>(ctrl_p + FLOW_CTRL_BYTES + FLOW_ONES_BYTES)[i] =
        reinterpret_cast<const uint4*>(fp.mask + (size_t)c.t0 * TILE_M)[i];
  __syncthreads();
  if (tc != nullptr && threadIdx.x == 0) tc[1] = globaltimer_ns();

  if (c.k > 0) {
    if (warp == 0) flow_producer(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr);
    else if (warp == 1) flow_mma(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr, 0u);
    else if (warp == 10) flow_mma(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr, 1u);
    else if (warp == 11) flow_aux(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr);
    else if (warp == 12) flow_publisher(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr, 0u);
    else if (warp == 13) flow_publisher(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr, 1u);
    else flow_epilogue(fp, c.smem, c.ctrl, c.area, c.tmem_base, c.cta, c.t0, c.k, c.n_tiles, c.m_rows, c.tr);
  }
  // Every role is past its last layer: the next kernel's CTAs may become resident (they still wait for this
  // grid to complete). Triggering any earlier could let them occupy an SM that a CTA of this grid needs.
  pdl_launch_dependents();
  if (fp.stamps != nullptr && c.cta == 0 && threadIdx.x == 64) fp.stamps[4 * fp.n_layers] = globaltimer_ns();
  if (tc != nullptr && threadIdx.x == 64) tc[2] = globaltimer_ns();     // an epilogue thread: last stores issued
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(c.tmem_base, 512);
    if (tc != nullptr && threadIdx.x == 32) {
      tc[3] = globaltimer_ns();
      tc[4] = c.k;
    }
  }
}

cudaError_t configure_conv_flow() {
  cudaError_t e = cudaFuncSetAttribute(conv_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOW_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_flow_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

bool conv_flow_supports(int cin, int nt, int taps) {
#define X(CIN_, NT_, TAPS_) \
  if (cin == CIN_ && nt == NT_ && taps == TAPS_) return true;
  MARU_FLOW_CASES(X)
#undef X
  return false;
}

int conv_flow_blob_bytes(int cin, int nt, int taps) { return taps * cin * nt * 2 + 2 * nt * 16; }

static int flow_slab_bytes(int cin, int taps) { return (taps == 9 ? TILE_M + 48 : TILE_M) * cin * 2; }

// slab ring left for a chain whose largest weight blob is w_region bytes
int conv_flow_area(int w_region) { return FLOW_SMEM - 2 * w_region - FLOW_FIXED_BYTES; }

// a layer fits a chain if two of its slabs fit the ring (one computing, one in flight)
bool conv_flow_fits(int cin, int nt, int taps, int w_region) {
  if (!conv_flow_supports(cin, nt, taps)) return false;
  if (conv_flow_blob_bytes(cin, nt, taps) > w_region) return false;
  return 2 * flow_slab_bytes(cin, taps) <= conv_flow_area(w_region);
}

cudaError_t launch_conv_flow(FlowParams& fp, int sm_count, cudaStream_t stream) {
  int w_region = 0;
  for (int l = 0; l < fp.n_layers; l++) {
    const FlowLayer& L = fp.layer[l];
    if (!conv_flow_supports(L.cin, L.nt, L.taps)) return cudaErrorInvalidConfiguration;
    const int b = conv_flow_blob_bytes(L.cin, L.nt, L.taps);
    if (b > w_region) w_region = b;
  }
  w_region = (w_region + 1023) & ~1023;
  fp.w_region = w_region;
  fp.area = conv_flow_area(w_region);
  for (int l = 0; l < fp.n_layers; l++) {
    const FlowLayer& L = fp.layer[l];
    if (!conv_flow_fits(L.cin, L.nt, L.taps, w_region)) return cudaErrorInvalidConfiguration;
  }
  const int n_tiles = (fp.m_rows + TILE_M - 1) / TILE_M;
  if ((n_tiles + sm_count - 1) / sm_count > FLOW_MAX_TILES) return cudaErrorInvalidConfiguration;
  // The CTAs wait for each other, so the grid is launched COOPERATIVELY: the runtime makes all of it resident
  // at once or fails the launch (another process through MPS, an SM-limited context, a long kernel of another
  // stream holding SMs) and the engine falls back to one launch per layer — instead of a partially resident grid
  // spinning until its bounded waits give up.  MARU_B200_FLOW_COOP: 2 = cooperative + programmatic dependent launch
  // (default), 1 = cooperative only, 0 = plain launch (A/B).
  static const int coop = [] {
    const char* v = getenv("MARU_B200_FLOW_COOP");
    return v != nullptr ? atoi(v) : 2;
  }();
  if (coop == 0)
    return launch_kernel(conv_flow_kernel, dim3(sm_count), dim3(FLOW_THREADS), (size_t)FLOW_SMEM, stream, fp.pdl != 0, fp);
  return launch_kernel_coop(conv_flow_kernel, dim3(sm_count), dim3(FLOW_THREADS), (size_t)FLOW_SMEM, stream,
                            fp.pdl != 0 && coop == 2, fp);
}

// how many CTAs of the flow kernel fit one SM right now (0: the kernel cannot run on this device / context)
int conv_flow_ctas_per_sm() {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_flow_kernel, FLOW_THREADS, (size_t)FLOW_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace maru
