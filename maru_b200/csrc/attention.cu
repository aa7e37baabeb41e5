// attention.cu — K4: multi-head self-attention core over the board tokens of one position.
//
//   O = softmax(Q K^T  with off-board KEYS masked to -inf) V        per (position, head)
//
// Replaces the ATen scaled-dot-product attention inside the TorchScript network (reference
// src/deepgo/native/cpp/Model.cpp:96).  Q is pre-scaled by log2(e)/sqrt(d) in the folded QKV
// weights (maru_b200/convert.py) so the softmax is a bare exp2.
//
// The product kernel is attention_tc_kernel below (tcgen05 / TMEM).  The v1 implementation it replaced is kept
// for A/B measurements in -DMARU_EXPERIMENTS builds only: warp-level online softmax with QK^T / PV through
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  One CTA = (position, head): the head's K and V
// planes (400 padded rows x d) are bulk-copied to shared memory once — in the chunk-planar layout a
// plane slice of one position is a single contiguous 6 400-byte run — and 5 warps sweep the 25
// query tiles of 16 rows.  K fragments come from ldmatrix, V fragments from ldmatrix.trans, P never
// leaves registers (S accumulator fragment == A fragment of the PV MMA).
// Tokens are the 400 padded rows; halo / off-board rows are masked as keys and written as zeros as
// queries, which makes the result identical to attention over the w x h on-board cells.
#include "kernels.h"

namespace maru {

#ifdef MARU_EXPERIMENTS   // v1 (mma.sync) kernel: A/B reference for the tcgen05 kernel below, not in the product build
constexpr int ATT_WARPS = 5;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int ATT_KB = 80;  // keys per online-softmax block (10 n8 tiles)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnParams p) {
  constexpr int DCH = D / 8;                       // 16-byte chunks per head
  constexpr int PLANE_BYTES = POS_ROWS * 16;       // 6400
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_k = smem;                             // [DCH][400][16 B]
  uint8_t* s_v = smem + DCH * PLANE_BYTES;
  uint8_t* s_mask = s_v + DCH * PLANE_BYTES;       // [400]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_mask + 512);

  pdl_launch_dependents();
  const int heads = p.channels / D;
  const int pos = blockIdx.x / heads;
  const int head = blockIdx.x - pos * heads;
  if (pos >= ((p.n_dev != nullptr) ? *p.n_dev : p.n)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const size_t prow0 = (size_t)GUARD_ROWS + (size_t)pos * POS_ROWS;
  const int cch = p.channels / 8;
  const int q_chunk0 = head * DCH, k_chunk0 = cch + head * DCH, v_chunk0 = 2 * cch + head * DCH;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 2 * DCH * PLANE_BYTES);
    for (int c = 0; c < DCH; c++) {
      bulk_load_1d(s_k + c * PLANE_BYTES, p.qkv + ((size_t)(k_chunk0 + c) * p.rows_alloc + prow0) * 8,
                   PLANE_BYTES, bar);
      bulk_load_1d(s_v + c * PLANE_BYTES, p.qkv + ((size_t)(v_chunk0 + c) * p.rows_alloc + prow0) * 8,
                   PLANE_BYTES, bar);
    }
  }
  for (int r = threadIdx.x; r < POS_ROWS; r += ATT_THREADS)
    s_mask[r] = p.mask[(size_t)pos * POS_ROWS + r];
  __syncthreads();
  mbar_wait(bar, 0);

  const uint32_t k_addr = smem_u32(s_k), v_addr = smem_u32(s_v);
  const int lm_i = lane >> 3, lm_r = lane & 7;  // ldmatrix: matrix index, row within matrix

  for (int tile = warp; tile < POS_ROWS / 16; tile += ATT_WARPS) {
    const int r0 = tile * 16;
    // ---- Q fragments straight from HBM/L2 (each element is used by exactly one warp)
    uint32_t qa[D / 16][4];
#pragma unroll
    for (int j = 0; j < D / 16; j++) {
      const __nv_bfloat16* c0 = p.qkv + ((size_t)(q_chunk0 + 2 * j) * p.rows_alloc + prow0 + r0) * 8;
      const __nv_bfloat16* c1 = p.qkv + ((size_t)(q_chunk0 + 2 * j + 1) * p.rows_alloc + prow0 + r0) * 8;
      qa[j][0] = *reinterpret_cast<const uint32_t*>(c0 + g * 8 + tq * 2);
      qa[j][1] = *reinterpret_cast<const uint32_t*>(c0 + (g + 8) * 8 + tq * 2);
      qa[j][2] = *reinterpret_cast<const uint32_t*>(c1 + g * 8 + tq * 2);
      qa[j][3] = *reinterpret_cast<const uint32_t*>(c1 + (g + 8) * 8 + tq * 2);
    }
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

#pragma unroll 1
    for (int kb = 0; kb < POS_ROWS / ATT_KB; kb++) {
      const int key0 = kb * ATT_KB;
      float s[ATT_KB / 8][4];
#pragma unroll
      for (int i = 0; i < ATT_KB / 8; i++) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
      // ---- S = Q K^T
#pragma unroll
      for (int n2 = 0; n2 < ATT_KB / 16; n2++) {
#pragma unroll
        for (int j = 0; j < D / 16; j++) {
          // matrices: (keys +0..7, chunk 2j) (keys +0..7, chunk 2j+1) (keys +8..15, 2j) (keys +8..15, 2j+1)
          const uint32_t addr = k_addr + (2 * j + (lm_i & 1)) * PLANE_BYTES +
                                (key0 + n2 * 16 + (lm_i >> 1) * 8 + lm_r) * 16;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(addr, b0, b1, b2, b3);
          mma_bf16_16816(s[2 * n2], qa[j], b0, b1);
          mma_bf16_16816(s[2 * n2 + 1], qa[j], b2, b3);
        }
      }
      // ---- mask off-board keys, block row max
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < ATT_KB / 8; i++) {
        const int c = key0 + i * 8 + tq * 2;
        const bool k0 = s_mask[c] != 0, k1 = s_mask[c + 1] != 0;
        s[i][0] = k0 ? s[i][0] : -INFINITY;
        s[i][1] = k1 ? s[i][1] : -INFINITY;
        s[i][2] = k0 ? s[i][2] : -INFINITY;
        s[i][3] = k1 ? s[i][3] : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
        mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float ms0 = (mn0 == -INFINITY) ? 0.0f : mn0, ms1 = (mn1 == -INFINITY) ? 0.0f : mn1;
      const float corr0 = exp2f(m0 - ms0), corr1 = exp2f(m1 - ms1);
      m0 = mn0;
      m1 = mn1;
      l0 *= corr0;
      l1 *= corr1;
#pragma unroll
      for (int i = 0; i < D / 8; i++) {
        o[i][0] *= corr0;
        o[i][1] *= corr0;
        o[i][2] *= corr1;
        o[i][3] *= corr1;
      }
      // ---- P = exp2(S - m) as bf16 A fragments; row sums from the rounded values
      uint32_t pa[ATT_KB / 8][2];
#pragma unroll
      for (int i = 0; i < ATT_KB / 8; i++) {
        pa[i][0] = pack_bf16x2(exp2f(s[i][0] - ms0), exp2f(s[i][1] - ms0));
        pa[i][1] = pack_bf16x2(exp2f(s[i][2] - ms1), exp2f(s[i][3] - ms1));
        l0 += bf16_lo(pa[i][0]) + bf16_hi(pa[i][0]);
        l1 += bf16_lo(pa[i][1]) + bf16_hi(pa[i][1]);
      }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < ATT_KB / 16; kk++) {
        const uint32_t a[4] = {pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1]};
#pragma unroll
        for (int d2 = 0; d2 < D / 16; d2++) {
          // matrices: (keys +0..7, chunk 2d2) (keys +8..15, chunk 2d2) (keys +0..7, 2d2+1) (keys +8..15, 2d2+1)
          const uint32_t addr = v_addr + (2 * d2 + (lm_i >> 1)) * PLANE_BYTES +
                                (key0 + kk * 16 + (lm_i & 1) * 8 + lm_r) * 16;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(addr, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * d2], a, b0, b1);
          mma_bf16_16816(o[2 * d2 + 1], a, b2, b3);
        }
      }
    }
    // ---- normalise, mask off-board queries, store (quad writes 16 contiguous bytes per row)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const bool on0 = s_mask[r0 + g] != 0, on1 = s_mask[r0 + g + 8] != 0;
    const float inv0 = on0 ? 1.0f / l0 : 0.0f, inv1 = on1 ? 1.0f / l1 : 0.0f;
#pragma unroll
    for (int i = 0; i < D / 8; i++) {
      const int m0 = pos * POS_ROWS + r0 + g, m1 = m0 + 8;
      const size_t off0 = p.out_blocked ? blocked_offset(cch, q_chunk0 + i, m0)
                                        : ((size_t)(q_chunk0 + i) * p.rows_alloc + GUARD_ROWS + m0) * 8;
      const size_t off1 = p.out_blocked ? blocked_offset(cch, q_chunk0 + i, m1)
                                        : ((size_t)(q_chunk0 + i) * p.rows_alloc + GUARD_ROWS + m1) * 8;
      *reinterpret_cast<uint32_t*>(p.out + off0 + tq * 2) = on0 ? pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0) : 0u;
      *reinterpret_cast<uint32_t*>(p.out + off1 + tq * 2) = on1 ? pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1) : 0u;
    }
  }
}

#endif  // MARU_EXPERIMENTS

// =============================================================================================
// v2: tcgen05 / TMEM attention for head_dim 32.
//
// One CTA = (position, head), two CTAs co-resident per SM (256 TMEM columns, ~113 KB smem each).
// The 3 query tiles x 5 key parts (400 padded keys = 5 x 80) form one linear sequence of 15 steps:
//     S  = Q K^T          tcgen05.mma  M=128, N=80, K=32    (Q, K: K-major chunk-planar images)
//     P  = exp2(S - c_i)  one TMEM lane (= query row) per thread, bf16 to shared memory
//     O += P [V | 1]      tcgen05.mma  M=128, N=48, K=80    (V read as an MN-major B operand; the
//                         extra "ones" column, zero for masked keys, yields the row sums l_i)
// S, P and O are double buffered (TMEM: 2x80 + 2x48 columns): two softmax warp groups alternate
// steps, the control warp keeps the tensor pipe two steps ahead, and the O / l epilogue of a tile
// runs while the next tile is already in flight.
// c_i = min(|q_i| * max_j |k_j|, 96) >= max_j S_ij (Cauchy-Schwarz) replaces the running row max:
// softmax is shift invariant, p <= 1 cannot overflow, and because V rows and the ones column of
// off-board keys are zero, the key mask costs nothing in the inner loop and the key parts are
// independent (no online rescaling of O).
// Query tiles start at padded rows 20, 148 and 276: row 20 is the first real board row (rows 0..19 are
// the position's top halo, whose output stays at the zeros the buffer was created with), and the 380
// rows 20..399 fit 3 tiles of 128 (rows 400..403 of the last tile belong to the next position's halo
// or to the tail rows: they are computed and dropped).
// =============================================================================================
// per-step clock64 stamps of one CTA (debug key 6, tools/att_trace.py), experiment builds only: the product kernel
// carries none of it
#ifdef MARU_EXPERIMENTS
#define ATT_STAMP(cond, step, slot) do { if (tr != nullptr && (cond) && (step) < 64) tr[(step) * 16 + (slot)] = clock64(); } while (0)
#else
#define ATT_STAMP(cond, step, slot) do { } while (0)
#endif

constexpr int A2_SM_WARPS = 8;              // warps 0..7 softmax/epilogue: lane group = w & 3, group = w >> 2
constexpr int A2_WARPS = A2_SM_WARPS + 1;   // warp 8: control (bulk loads + MMA issue)
constexpr int A2_THREADS = A2_WARPS * 32;
constexpr int A2_D = 32, A2_DCH = 4;
constexpr int A2_PLANE = POS_ROWS * 16;     // 6400 B: one 8-channel plane of one position
constexpr int A2_K_BYTES = A2_DCH * A2_PLANE;
constexpr int A2_V_BYTES = (A2_DCH + 1) * A2_PLANE;   // + ones plane; the 6th plane aliases P (unused cols)
constexpr int A2_PN = 80;                   // keys per step
constexpr int A2_NP = POS_ROWS / A2_PN;     // 5 parts
constexpr int A2_P_BYTES = (A2_PN / 8) * TILE_M * 16; // 20 480 per buffer
constexpr int A2_Q_BYTES = A2_DCH * TILE_M * 16;
constexpr int A2_SMEM = A2_K_BYTES + A2_V_BYTES + 2 * A2_P_BYTES + 2 * A2_Q_BYTES + 416 + 32 + 112;
constexpr int A2_TMEM_COLS = 256;
constexpr int A2_O_COL = 2 * A2_PN;         // O buffers at columns 160 and 208
constexpr int A2_NV = 48;                   // V columns (32) + ones + padding to a multiple of 16

// exp2 on the FMA / ALU pipes for x <= 0: round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax
// 2^f (max relative error 7.5e-5, far below the bf16 rounding of P), exponent added as an integer.
// The MUFU pipe (16 ex2 per clock per SM) is the busiest pipe of this kernel (ncu: 60 %); moving a
// share of the exponentials off it lets both pipes work in parallel.
template <bool CLAMP = true>
__device__ __forceinline__ float exp2_fma(float x) {
  if (CLAMP) x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: n lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517162f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// 16 S columns -> 16 probabilities -> two 16-byte P chunks of this row. POLY: the second chunk's
// exponentials run on the FMA pipe instead of the MUFU.
// which of the 5 exp_store16 calls of a step use the FMA-pipe exponential for their second chunk
// (measured per layer at batch 256, share of the exps on the FMA pipe: 0 % 65.7 us, 10 % 65.0, 20 % 63.0,
// 30 % 61.7, 40 % 65.1, 50 % 67.0)
#ifndef MARU_ATT_POLY_MASK
#define MARU_ATT_POLY_MASK 0x15
#endif
constexpr int A2_POLY_MASK = MARU_ATT_POLY_MASK;
#ifndef MARU_ATT_POLY_NS_MASK
#define MARU_ATT_POLY_NS_MASK 0x09
#endif
// the unshifted path has fewer FMA-pipe instructions per exponential to begin with: its optimum is 20 % (forward at batch 256,
// both layers: 0 % 460.0 us, 10 % 455.0, 20 % 452.9-454.2 whichever calls, 30 % 455.8, 40 % 455.5, 50 % 461.6; shifted 463.7)
constexpr int A2_POLY_NS_MASK = MARU_ATT_POLY_NS_MASK;
// |score| bound (in the log2 units of the folded weights) below which a warp skips the shift: 400 keys x 2^64 and the
// products with V stay far inside fp32 / bf16 range on both sides (2^-64 is a normal number in both)
#ifndef MARU_ATT_NOSHIFT
#define MARU_ATT_NOSHIFT 64.0f
#endif
constexpr float A2_NOSHIFT = MARU_ATT_NOSHIFT;

// SHIFT = false: the scores of this warp's rows are bounded by A2_NOSHIFT (|s_ij| <= |q_i| max_j |k_j|), exp2(s) can
// neither overflow nor lose the row sum, and the subtraction of c_i — one FADD per exponential in an issue-bound loop —
// is dropped (softmax is shift invariant: c_i = 0 is as good as any other shift).
template <bool POLY, bool SHIFT>
__device__ __forceinline__ void exp_store16(const uint32_t* v, float c, uint8_t* s_p, int chunk, int trow) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float x0 = __uint_as_float(v[h * 8 + 2 * e]), x1 = __uint_as_float(v[h * 8 + 2 * e + 1]);
      if (SHIFT) {
        x0 -= c;
        x1 -= c;
      }
      o[e] = (POLY && h == 1) ? pack_bf16x2(exp2_fma<SHIFT>(x0), exp2_fma<SHIFT>(x1)) : pack_bf16x2(ex2_approx(x0), ex2_approx(x1));
    }
    *reinterpret_cast<uint4*>(s_p + ((chunk + h) * TILE_M + trow) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Key parts (80 padded rows each) that hold at least one on-board cell: a 9x9 board centred in the 19x19 frame lives in
// parts 1-3, a 13x13 board in parts 1-4. A part without on-board keys contributes exactly nothing (its V rows and its
// ones column are zero), so skipping it leaves the result bit-identical and saves its QK, its 128 x 80 exponentials and
// its PV. `m16` is this lane's 16 mask bytes (lanes 0..24 cover the 400 rows), the result is the same in every lane.
__device__ __forceinline__ void a2_active_parts(const uint4 m16, const int lane, int& np, uint32_t& plist) {
  const bool any = lane < POS_ROWS / 16 && (m16.x | m16.y | m16.z | m16.w) != 0u;
  const uint32_t b = __ballot_sync(0xffffffffu, any);           // bit i: rows 16i .. 16i+15
  np = 0;
  plist = 0u;
#pragma unroll
  for (int part = 0; part < A2_NP; part++) {
    if ((b >> (5 * part)) & 31u) {                               // 80 rows = 5 groups of 16
      plist |= (uint32_t)part << (3 * np);
      np++;
    }
  }
  if (np == 0) {                                                 // no on-board cell at all: keep the full sequence
    np = A2_NP;
    plist = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);
  }
}

__global__ void __launch_bounds__(A2_THREADS, 2) attention_tc_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_k = smem;
  uint8_t* s_v = s_k + A2_K_BYTES;
  uint8_t* s_p = s_v + A2_V_BYTES;                       // two P buffers
  uint8_t* s_q = s_p + 2 * A2_P_BYTES;                   // two Q tile buffers
  uint8_t* s_mask = s_q + 2 * A2_Q_BYTES;                // [400] (+pad to 416)
  float* s_red = reinterpret_cast<float*>(s_mask + 416); // [8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + 416 + 32);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;      // [2]
  uint64_t* s_full = bars + 3;      // [2]
  uint64_t* p_full = bars + 5;      // [2]
  uint64_t* o_full = bars + 7;      // [2]
  uint64_t* o_empty = bars + 9;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  pdl_launch_dependents();
  const int heads = p.channels / A2_D;
  const int pos = blockIdx.x / heads;
  const int head = blockIdx.x - pos * heads;
  if (pos >= ((p.n_dev != nullptr) ? *p.n_dev : p.n)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MARU_EXPERIMENTS
  long long* const tr = (p.trace != nullptr && (int)blockIdx.x == p.trace_cta) ? p.trace : nullptr;
#endif
  // frame geometry (common.cuh): 20 / 400 for the 19x19 frame; a compact frame of a small board has fewer rows, hence
  // fewer key parts and query tiles (9x9: 100 rows = 2 parts x 1 tile instead of 5 x 3)
  const int pos_rows = p.pos_rows, q0 = p.pitch;
  const int kp = (pos_rows + A2_PN - 1) / A2_PN;          // key parts of 80 rows
  const int plane_rows = kp * A2_PN;                       // rows of one 8-channel plane image in shared memory
  const int plane_b = plane_rows * 16;                     // (rows >= pos_rows are zero: they are keys like any other)
  const int nt = (pos_rows - q0 + TILE_M - 1) / TILE_M;   // query tiles, the first starts after the top halo row
  const bool full = pos_rows == POS_ROWS;
  const size_t prow0 = (size_t)GUARD_ROWS + (size_t)pos * pos_rows;
  const int cch = p.channels / 8;
  const int q_chunk0 = head * A2_DCH, k_chunk0 = cch + head * A2_DCH, v_chunk0 = 2 * cch + head * A2_DCH;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], A2_SM_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == A2_SM_WARPS) {
    tmem_alloc(tmem_slot, A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // qkv and mask are written by the previous kernels

  if (warp == A2_SM_WARPS) {
    // =========================== control warp: bulk loads + MMA issue ===========================
    auto load_q = [&](int tile) {
      const int buf = tile & 1;
      const int r0 = q0 + tile * TILE_M;
      mbar_expect_tx(&q_full[buf], A2_Q_BYTES);
      for (int c = 0; c < A2_DCH; c++)
        bulk_load_1d(s_q + buf * A2_Q_BYTES + c * TILE_M * 16,
                     p.qkv + ((size_t)(q_chunk0 + c) * p.rows_alloc + prow0 + r0) * 8, TILE_M * 16, &q_full[buf]);
    };
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * A2_DCH * pos_rows * 16);
      for (int c = 0; c < A2_DCH; c++) {
        bulk_load_1d(s_k + c * plane_b, p.qkv + ((size_t)(k_chunk0 + c) * p.rows_alloc + prow0) * 8, pos_rows * 16, kv_full);
        bulk_load_1d(s_v + c * plane_b, p.qkv + ((size_t)(v_chunk0 + c) * p.rows_alloc + prow0) * 8, pos_rows * 16, kv_full);
      }
      load_q(0);
      if (nt > 1) load_q(1);
    }
    __syncwarp();
    constexpr uint32_t idesc_qk = umma_idesc_bf16(TILE_M, A2_PN, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(TILE_M, A2_NV, 0, 1);   // B = V is MN-major
    const uint64_t q_d = umma_desc(0, TILE_M * 16, 128);      // Q, P tiles: K-major, LBO = 128 rows
    const uint64_t k_d = umma_desc(0, plane_b, 128);          // K image:    K-major, LBO = plane
    // V image read transposed (MN-major): SBO = distance between 8-channel planes (N direction),
    // LBO = distance between 8-key groups (K direction)
    const uint64_t v_d = umma_desc(0, 128, plane_b);
    const uint32_t q_hi = (uint32_t)(q_d >> 32), k_hi = (uint32_t)(k_d >> 32), v_hi = (uint32_t)(v_d >> 32);
    const uint32_t k_lo = (uint32_t)k_d + (smem_u32(s_k) >> 4);
    const uint32_t v_lo = (uint32_t)v_d + (smem_u32(s_v) >> 4);
    const uint32_t p_lo = (uint32_t)q_d + (smem_u32(s_p) >> 4);
    const uint32_t q_lo = (uint32_t)q_d + (smem_u32(s_q) >> 4);
    int np;
    uint32_t plist;
    if (full) {
      uint4 m16 = make_uint4(0u, 0u, 0u, 0u);
      if (lane < POS_ROWS / 16) m16 = reinterpret_cast<const uint4*>(p.mask + (size_t)pos * POS_ROWS)[lane];
      a2_active_parts(m16, lane, np, plist);
    } else {                      // a compact frame has no empty key part
      np = kp;
      plist = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);
    }
    const int n_steps = nt * np;
    // rows past the position in the last key part (compact frames only: 400 = 5 x 80) must be ZERO keys before the
    // first QK reads them — whatever the previous kernel left there could be Inf / NaN bit patterns, and Inf * 0 (their
    // V rows are zero) would poison O.  Written here, by the warp that issues the MMAs, and fenced for the async proxy.
    if (plane_rows > pos_rows) {
      const int pad = plane_rows - pos_rows;
      for (int i = lane; i < pad * A2_DCH; i += 32)
        *reinterpret_cast<uint4*>(s_k + (i / pad) * plane_b + (pos_rows + i % pad) * 16) = make_uint4(0u, 0u, 0u, 0u);
      fence_async_smem();
      __syncwarp();
    }
    mbar_wait(kv_full, 0);

    // S[step & 1] = Q(tile) K(part)^T ; the first QK of a tile waits for its Q buffer
    // (tile, part index) are advanced incrementally: np is a per-CTA runtime value and a division by it in these loops
    // cost 10 % of the kernel
    int q_tile = 0, q_pi = 0;
    auto issue_qk = [&](int step) {
      const int tile = q_tile, pi = q_pi;
      if (++q_pi == np) {
        q_pi = 0;
        q_tile++;
      }
      const int part = (int)((plist >> (3 * pi)) & 7u);
      if (pi == 0) mbar_wait(&q_full[tile & 1], (tile >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t t_s = tmem_base + (step & 1) * A2_PN;
        const uint32_t qa = q_lo + (tile & 1) * (A2_Q_BYTES >> 4);
        const uint32_t kb = k_lo + part * A2_PN;
        umma_bf16_lohi<0>(t_s, qa, q_hi, kb, k_hi, idesc_qk);
        umma_bf16_lohi<1>(t_s, qa + 2 * TILE_M, q_hi, kb + 2 * plane_rows, k_hi, idesc_qk);
        umma_commit(&s_full[step & 1]);
      }
      __syncwarp();
    };
    issue_qk(0);
    if (n_steps > 1) issue_qk(1);
    int m_tile = 0, m_pi = 0;
#pragma unroll 1
    for (int step = 0; step < n_steps; step++) {
      const int tile = m_tile, pi = m_pi;
      if (++m_pi == np) {
        m_pi = 0;
        m_tile++;
      }
      const int part = (int)((plist >> (3 * pi)) & 7u);
      const int sb = step & 1, ob = tile & 1;
      mbar_wait(&p_full[sb], (step >> 1) & 1);     // P[sb] written (and fenced), S[sb] consumed
      ATT_STAMP(lane == 0, step, 4);
      if (pi == 0 && tile >= 2) mbar_wait(&o_empty[ob], 0);     // epilogue of tile-2 has read O[ob]
      tc_fence_after();
      if (elect_one()) {
        const uint32_t t_o = tmem_base + A2_O_COL + ob * A2_NV;
        const uint32_t pa = p_lo + sb * (A2_P_BYTES >> 4);
        const uint32_t vb = v_lo + part * A2_PN;
        if (pi == 0) umma_bf16_lohi<0>(t_o, pa, q_hi, vb, v_hi, idesc_pv);
        else umma_bf16_lohi<1>(t_o, pa, q_hi, vb, v_hi, idesc_pv);
#pragma unroll
        for (int j = 1; j < A2_PN / 16; j++)
          umma_bf16_lohi<1>(t_o, pa + 2 * j * TILE_M, q_hi, vb + 16 * j, v_hi, idesc_pv);
        if (pi == np - 1) umma_commit(&o_full[ob]);
      }
      __syncwarp();
      ATT_STAMP(lane == 0, step, 5);
      // the last QK of this tile completed before p_full(step): its Q buffer can take tile+2
      if (pi == np - 1 && tile + 2 < nt && elect_one()) load_q(tile + 2);
      __syncwarp();
      // QK(step+2) is issued AFTER PV(step): its commit (s_full) then also covers PV(step), which is
      // what allows the softmax group to overwrite P[sb] at step+2. (Issuing it earlier, as soon as S[sb]
      // had been read out, was tried: no gain, and it needs an extra P-free barrier.)
      if (step + 2 < n_steps) issue_qk(step + 2);
      ATT_STAMP(lane == 0, step, 6);
    }
  } else {
    // =========================== softmax + epilogue warps ===========================
    const int lg = warp & 3, grp = warp >> 2;               // grp g handles steps s with (s & 1) == g
    const int trow = lg * 32 + lane;                        // TMEM lane = row within the tile
    const int st = threadIdx.x;                             // 0..255
    if (full) {
      if (st < POS_ROWS / 16)                               // 400 mask bytes = 25 16-byte vectors
        reinterpret_cast<uint4*>(s_mask)[st] = reinterpret_cast<const uint4*>(p.mask + (size_t)pos * POS_ROWS)[st];
    } else {                                                // compact frames: pos * pos_rows is not 16-byte aligned
      for (int r = st; r < pos_rows; r += 256) s_mask[r] = p.mask[(size_t)pos * pos_rows + r];
    }
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    int np;
    uint32_t plist;
    if (full) {
      uint4 m16 = make_uint4(0u, 0u, 0u, 0u);
      if (lane < POS_ROWS / 16) m16 = reinterpret_cast<const uint4*>(s_mask)[lane];
      a2_active_parts(m16, lane, np, plist);       // the same list the control warp derives from global memory
    } else {
      np = kp;
      plist = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);
    }
    const int n_steps = nt * np;
    // ones plane of V: 1.0 for on-board keys, 0 otherwise (row sums + key mask in one column)
    for (int r = st; r < plane_rows; r += 256)
      *reinterpret_cast<uint4*>(s_v + A2_DCH * plane_b + r * 16) =
          make_uint4((r < pos_rows && s_mask[r]) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    // rows past the position in the last key part (compact frames only): zero values (the control warp zeroes the keys);
    // ordered before the first PV by this group's p_full arrival, like the ones plane
    for (int i = st; i < (plane_rows - pos_rows) * A2_DCH; i += 256) {
      const int c = i / (plane_rows - pos_rows), r = pos_rows + i % (plane_rows - pos_rows);
      *reinterpret_cast<uint4*>(s_v + c * plane_b + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    mbar_wait(kv_full, 0);
    // max_j |k_j|^2
    float mk = 0.0f;
    for (int r = st; r < pos_rows; r += 256) {
      float ss = 0.0f;
#pragma unroll
      for (int c = 0; c < A2_DCH; c++) {
        const uint4 u = *reinterpret_cast<const uint4*>(s_k + c * plane_b + r * 16);
        const float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                            bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int e = 0; e < 8; e++) ss = fmaf(f[e], f[e], ss);
      }
      mk = fmaxf(mk, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mk = fmaxf(mk, __shfl_xor_sync(0xffffffffu, mk, o));
    if (lane == 0) s_red[warp] = mk;
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    float maxk2 = s_red[0];
#pragma unroll
    for (int w = 1; w < A2_SM_WARPS; w++) maxk2 = fmaxf(maxk2, s_red[w]);
    const uint32_t t_lane = tmem_base + ((uint32_t)(lg * 32) << 16);

    // O / l, query mask, store; both groups take part, each owns 16 of the head's 32 channels
    auto epilogue = [&](int tile) {
      const int ob = tile & 1;
      const int row = q0 + tile * TILE_M + trow;
      ATT_STAMP(warp == 0 && lane == 0, 32 + tile, 0);
      mbar_wait(&o_full[ob], (tile >> 1) & 1);
      tc_fence_after();
      ATT_STAMP(warp == 0 && lane == 0, 32 + tile, 1);
      uint32_t o[16];
      tmem_ld16(t_lane + A2_O_COL + ob * A2_NV + grp * 16, o);
      const uint32_t lbits = tmem_ld1(t_lane + A2_O_COL + ob * A2_NV + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
      ATT_STAMP(warp == 0 && lane == 0, 32 + tile, 2);
      if (row < pos_rows) {
        const bool on = s_mask[row] != 0;
        const float inv = on ? 1.0f / __uint_as_float(lbits) : 0.0f;
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          uint4 w;
          w.x = on ? pack_bf16x2(__uint_as_float(o[ch * 8 + 0]) * inv, __uint_as_float(o[ch * 8 + 1]) * inv) : 0u;
          w.y = on ? pack_bf16x2(__uint_as_float(o[ch * 8 + 2]) * inv, __uint_as_float(o[ch * 8 + 3]) * inv) : 0u;
          w.z = on ? pack_bf16x2(__uint_as_float(o[ch * 8 + 4]) * inv, __uint_as_float(o[ch * 8 + 5]) * inv) : 0u;
          w.w = on ? pack_bf16x2(__uint_as_float(o[ch * 8 + 6]) * inv, __uint_as_float(o[ch * 8 + 7]) * inv) : 0u;
          const int plane = q_chunk0 + grp * 2 + ch;
          const size_t off = p.out_blocked ? blocked_offset(cch, plane, pos * pos_rows + row)
                                           : ((size_t)plane * p.rows_alloc + prow0 + row) * 8;
          *reinterpret_cast<uint4*>(p.out + off) = w;
        }
      }
      ATT_STAMP(warp == 0 && lane == 0, 32 + tile, 3);
    };

    float c = 0.0f;
    bool shift = true;                                      // warp-uniform, per tile
    const bool force_shift = (p.dbg & 4) != 0;              // debug bit 23: A/B and parity of the shifted path
    int c_tile = -1, epi_done = 0;
    int s_tile = 0, s_pi = grp;                             // (tile, part index) of `step`, advanced by two per iteration
    while (s_pi >= np) {
      s_pi -= np;
      s_tile++;
    }
#pragma unroll 1
    for (int step = grp; step < n_steps; step += 2) {
      const int tile = s_tile;
      const int tile_first_step = step - s_pi;              // == tile * np
      s_pi += 2;
      while (s_pi >= np) {
        s_pi -= np;
        s_tile++;
      }
      const int buf = tile & 1;
      if (tile != c_tile) {                                 // new tile: shift c_i of this row
        c_tile = tile;
        {
          mbar_wait(&q_full[buf], (tile >> 1) & 1);
          float ss = 0.0f;
#pragma unroll
          for (int ch = 0; ch < A2_DCH; ch++) {
            const uint4 u = *reinterpret_cast<const uint4*>(s_q + buf * A2_Q_BYTES + (ch * TILE_M + trow) * 16);
            const float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int e = 0; e < 8; e++) ss = fmaf(f[e], f[e], ss);
          }
          const float bound = sqrtf(ss * maxk2) * 1.0001f + 1e-3f;
          c = fminf(bound, 96.0f);
          shift = force_shift || __any_sync(0xffffffffu, !(bound < A2_NOSHIFT));     // (a NaN bound keeps the shifted path)
        }
      }
      ATT_STAMP(lg == 0 && lane == 0, step, 0);
      mbar_wait(&s_full[grp], (step >> 1) & 1);
      tc_fence_after();
      ATT_STAMP(lg == 0 && lane == 0, step, 1);
      // s_full(step) was committed after PV(step-2) and, for step >= 5*tile+1, after the last PV of
      // the previous tile: its O is complete, the epilogue does not stall
      {
        const uint32_t t_s = t_lane + grp * A2_PN;
        uint8_t* pb = s_p + grp * A2_P_BYTES;
        // software pipelined: the next tcgen05.ld is in flight while this block's exps run
        // (a tcgen05.ld + wait round trip is ~170 cycles, profiles/r1_tmem_microbench.txt)
        uint32_t va[32], vb[32];
        tmem_ld32(t_s, va);
        tmem_ld_wait();
        tmem_ld32(t_s + 32, vb);
        if (shift) {
          exp_store16<(A2_POLY_MASK >> 0) & 1, true>(va, c, pb, 0, trow);
          exp_store16<(A2_POLY_MASK >> 1) & 1, true>(va + 16, c, pb, 2, trow);
          tmem_ld_wait();
          tmem_ld16(t_s + 64, va);
          exp_store16<(A2_POLY_MASK >> 2) & 1, true>(vb, c, pb, 4, trow);
          exp_store16<(A2_POLY_MASK >> 3) & 1, true>(vb + 16, c, pb, 6, trow);
          tmem_ld_wait();
          ATT_STAMP(lg == 0 && lane == 0, step, 2);
          exp_store16<(A2_POLY_MASK >> 4) & 1, true>(va, c, pb, 8, trow);
        } else {
          exp_store16<(A2_POLY_NS_MASK >> 0) & 1, false>(va, c, pb, 0, trow);
          exp_store16<(A2_POLY_NS_MASK >> 1) & 1, false>(va + 16, c, pb, 2, trow);
          tmem_ld_wait();
          tmem_ld16(t_s + 64, va);
          exp_store16<(A2_POLY_NS_MASK >> 2) & 1, false>(vb, c, pb, 4, trow);
          exp_store16<(A2_POLY_NS_MASK >> 3) & 1, false>(vb + 16, c, pb, 6, trow);
          tmem_ld_wait();
          ATT_STAMP(lg == 0 && lane == 0, step, 2);
          exp_store16<(A2_POLY_NS_MASK >> 4) & 1, false>(va, c, pb, 8, trow);
        }
      }
      fence_async_smem();     // P visible to the async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      ATT_STAMP(lg == 0 && lane == 0, step, 3);
      // (with fewer than three parts per tile a group may not see a later step of the tile: it takes the epilogue at once)
      if (epi_done < tile && (step >= tile_first_step + 1 || np < 3)) {
        epilogue(epi_done);
        epi_done++;
      }
    }
    while (epi_done < nt) {
      epilogue(epi_done);
      epi_done++;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == A2_SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A2_TMEM_COLS);
  }
}

#ifdef MARU_EXPERIMENTS
template <int D>
static constexpr int attn_smem_bytes() { return 2 * (D / 8) * POS_ROWS * 16 + 512 + 16; }
#endif

cudaError_t configure_attention() {
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#ifdef MARU_EXPERIMENTS
  cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<32>());
  cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<64>());
#endif
  return e;
}

// head_dim 32 only (every BASELINE.json config): the engine refuses other packs at load time.
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  const int heads = p.channels / p.head_dim;
#ifdef MARU_EXPERIMENTS
  if (p.head_dim == 32 && (p.dbg & 2))     // dbg bit1: v1 mma.sync kernel (A/B)
    return launch_kernel(attention_kernel<32>, dim3(p.n * heads), dim3(ATT_THREADS), attn_smem_bytes<32>(), stream,
                         p.pdl != 0, p);
  if (p.head_dim == 64)
    return launch_kernel(attention_kernel<64>, dim3(p.n * heads), dim3(ATT_THREADS), attn_smem_bytes<64>(), stream,
                         p.pdl != 0, p);
#endif
  if (p.head_dim != 32) return cudaErrorInvalidValue;
  return launch_kernel(attention_tc_kernel, dim3(p.n * heads), dim3(A2_THREADS), A2_SMEM, stream, p.pdl != 0, p);
}

}  // namespace maru
