// heads.cu — K5: policy / territory / value heads, fused (sm_100a, HBM-bound).
//
// Consumes g = relu(bn(conv1x1(h, C -> Ch))) * M  (written by conv_gemm) and produces the
// reference output row float32[2169] = planes [6][19][19] + 3 scalars
// (reference src/deepgo/config.py:43-50; consumers Evaluator.cpp:66-80, player.py:415-461):
//   planes 0,1 : softmax over ON-BOARD cells of the two policy logits (off-board exactly 0)
//   planes 2..4: 3-way territory softmax per cell, times M
//   plane 5    : sigmoid, times M
//   scalars    : fc2(relu(fc1([mean_onboard(g), max(g)])));  [0] through sigmoid
// and / or the compact result maru_leaf (masked move priors + the 3 scalars, 1456 B) that
// Evaluator::evaluate actually keeps (Evaluator.cpp:55-80).
// One CTA per position.  Algorithmic bytes per position: 361*Ch*2 read + 2169*4 written.
#include "kernels.h"

namespace maru {

constexpr int HEAD_THREADS = 416;      // one thread per padded row (400) in the per-cell phase
constexpr int HEAD_WARPS = HEAD_THREADS / 32;
constexpr int HEAD_MAX_CH = 32;
constexpr int HEAD_MAX_CV = 128;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// single-instruction warp reductions (REDUX): non-negative floats order like their bit patterns, and
// the channel sums are taken in 20.12 fixed point (g is bf16: 2^-12 is below its rounding noise)
__device__ __forceinline__ unsigned redux_max_u32(unsigned v) {
  unsigned r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;\n" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ unsigned redux_add_u32(unsigned v) {
  unsigned r;
  asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;\n" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per position, one thread per padded row: the row's 32 head features stay in registers
// from the single global read to the last use (logits, territory / value planes, pooling), the
// policy softmax is a two-level shuffle reduction, and only the tiny value MLP goes through smem.
__global__ void __launch_bounds__(HEAD_THREADS) heads_kernel(const HeadParams p) {
  __shared__ float s_wp[6][HEAD_MAX_CH];
  __shared__ float s_bp[6];
  __shared__ float s_red[4][HEAD_WARPS];                 // max0, max1 then sum0, sum1 per warp
  __shared__ float s_psum[HEAD_WARPS][HEAD_MAX_CH];      // per-warp channel sums
  __shared__ float s_pmax[HEAD_WARPS][HEAD_MAX_CH];
  __shared__ float s_cnt[HEAD_WARPS];
  __shared__ float s_pool[2 * HEAD_MAX_CH];
  __shared__ float s_hid[HEAD_MAX_CV];

  pdl_launch_dependents();
  const int pos = blockIdx.x;
  if (pos >= ((p.n_dev != nullptr) ? *p.n_dev : p.n)) return;
  const int tid = threadIdx.x;
  // 416 threads (one per padded row) on the 19x19 frame; a compact frame is launched with just enough warps for its rows
  // (128 threads for 9x9), so that several positions share an SM instead of leaving three quarters of a CTA idle
  const int nthreads = (int)blockDim.x, nwarps = nthreads >> 5;
  const int ch = p.ch;
  const int warp = tid >> 5, lane = tid & 31;
  const int n_chunks = ch / 8;

  for (int i = tid; i < 6 * ch; i += nthreads) s_wp[i / ch][i % ch] = p.w_planes[i];
  if (tid < 6) s_bp[tid] = p.b_planes[tid];
  pdl_wait();   // g and mask come from the previous kernels

  // frame geometry: the 19x19 frame, or the compact frame of a board_w x board_h board whose outputs are written at the
  // cells the reference uses for it (centred in 19x19, Evaluator.cpp:57-69); everything around the board is zero
  const int pitch = p.pitch, pos_rows = p.pos_rows;
  const bool compact = p.board_w > 0;
  const int fw = compact ? p.board_w : BOARD;            // cells per frame row
  const int ox = compact ? (BOARD - p.board_w) / 2 : 0, oy = compact ? (BOARD - p.board_h) / 2 : 0;
  const int r = tid;                                     // padded row of this thread
  const bool in_pos = r < pos_rows;
  const int gy = r / pitch, x = r - gy * pitch;
  const bool cell_ok = in_pos && gy >= 1 && x < fw;      // a cell of the frame (maybe off-board)
  const int cell = (gy - 1 + oy) * BOARD + x + ox;       // its index in the 19x19 model frame
  const bool on = in_pos && (p.mask[(size_t)pos * pos_rows + (in_pos ? r : 0)] != 0);
  float gv[HEAD_MAX_CH];
#pragma unroll
  for (int c8 = 0; c8 < HEAD_MAX_CH / 8; c8++) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (in_pos && c8 < n_chunks) {
      const size_t grow = (size_t)GUARD_ROWS + (size_t)pos * pos_rows + r;
      v = *reinterpret_cast<const uint4*>(p.g + ((size_t)c8 * p.rows_alloc + grow) * 8);
    }
    gv[c8 * 8 + 0] = bf16_lo(v.x); gv[c8 * 8 + 1] = bf16_hi(v.x);
    gv[c8 * 8 + 2] = bf16_lo(v.y); gv[c8 * 8 + 3] = bf16_hi(v.y);
    gv[c8 * 8 + 4] = bf16_lo(v.z); gv[c8 * 8 + 5] = bf16_hi(v.z);
    gv[c8 * 8 + 6] = bf16_lo(v.w); gv[c8 * 8 + 7] = bf16_hi(v.w);
  }
  __syncthreads();                                       // head-plane weights in smem

  // ---- per-cell logits
  float z[6];
#pragma unroll
  for (int k = 0; k < 6; k++) z[k] = s_bp[k];
#pragma unroll
  for (int c = 0; c < HEAD_MAX_CH; c++) {
    if (c < ch) {
#pragma unroll
      for (int k = 0; k < 6; k++) z[k] = fmaf(s_wp[k][c], gv[c], z[k]);
    }
  }
  float* out = (p.out != nullptr) ? p.out + (size_t)pos * OUT_ROW : nullptr;
  if (compact) {                                         // model cells outside the board are not visited below
    if (out != nullptr)
      for (int i = tid; i < 6 * CELLS; i += nthreads) out[i] = 0.0f;
    if (p.logits != nullptr)
      for (int i = tid; i < 2 * CELLS; i += nthreads) p.logits[(size_t)pos * 2 * CELLS + i] = 0.0f;
    __syncthreads();
  }
  if (cell_ok && p.logits != nullptr) {
    p.logits[((size_t)pos * 2 + 0) * CELLS + cell] = z[0];
    p.logits[((size_t)pos * 2 + 1) * CELLS + cell] = z[1];
  }
  if (cell_ok && out != nullptr) {
    if (on) {
      const float tm = fmaxf(z[2], fmaxf(z[3], z[4]));
      const float e2 = __expf(z[2] - tm), e3 = __expf(z[3] - tm), e4 = __expf(z[4] - tm);
      const float inv = 1.0f / (e2 + e3 + e4);
      out[2 * CELLS + cell] = e2 * inv;
      out[3 * CELLS + cell] = e3 * inv;
      out[4 * CELLS + cell] = e4 * inv;
      out[5 * CELLS + cell] = 1.0f / (1.0f + __expf(-z[5]));
    } else {
      out[2 * CELLS + cell] = 0.0f;
      out[3 * CELLS + cell] = 0.0f;
      out[4 * CELLS + cell] = 0.0f;
      out[5 * CELLS + cell] = 0.0f;
    }
  }
  // ---- block maxima of the two policy planes + per-warp pooling partials
  {
    const float m0 = warp_max(on ? z[0] : -INFINITY), m1 = warp_max(on ? z[1] : -INFINITY);
    const float cnt = warp_sum(on ? 1.0f : 0.0f);
    if (lane == 0) {
      s_red[0][warp] = m0;
      s_red[1][warp] = m1;
      s_cnt[warp] = cnt;
    }
    // channel sums / maxima over the warp's 32 rows (g >= 0 and exactly 0 off-board)
#pragma unroll
    for (int c = 0; c < HEAD_MAX_CH; c++) {
      const unsigned sv = redux_add_u32(__float2uint_rn(gv[c] * 4096.0f));
      const unsigned mv = redux_max_u32(__float_as_uint(gv[c]));
      if (lane == c) {
        s_psum[warp][c] = (float)sv * (1.0f / 4096.0f);
        s_pmax[warp][c] = __uint_as_float(mv);
      }
    }
  }
  __syncthreads();
  float bm0 = -INFINITY, bm1 = -INFINITY;
  for (int w = 0; w < nwarps; w++) {
    bm0 = fmaxf(bm0, s_red[0][w]);
    bm1 = fmaxf(bm1, s_red[1][w]);
  }
  const float e0 = on ? __expf(z[0] - bm0) : 0.0f, e1 = on ? __expf(z[1] - bm1) : 0.0f;
  {
    const float s0 = warp_sum(e0), s1 = warp_sum(e1);
    if (lane == 0) {
      s_red[2][warp] = s0;
      s_red[3][warp] = s1;
    }
  }
  if (tid < ch) {                                        // finish the pooling while the sums settle
    float sv = 0.0f, mv = 0.0f, cnt = 0.0f;
    for (int w = 0; w < nwarps; w++) {
      sv += s_psum[w][tid];
      mv = fmaxf(mv, s_pmax[w][tid]);
      cnt += s_cnt[w];
    }
    s_pool[tid] = sv / cnt;
    s_pool[ch + tid] = mv;
  }
  __syncthreads();
  float t0 = 0.0f, t1 = 0.0f;
  for (int w = 0; w < nwarps; w++) {
    t0 += s_red[2][w];
    t1 += s_red[3][w];
  }
  if (cell_ok && out != nullptr) {
    out[0 * CELLS + cell] = e0 / t0;
    out[1 * CELLS + cell] = e1 / t1;
  }
  if (p.leaf != nullptr) {
    // Compact result (Evaluator.cpp:55-72): the prior of board cell (bx,by) is output plane 0 at the
    // centred model cell; cells the host marked illegal / fixed territory get exactly 0. The same
    // division as the full row, so the two transports agree bit for bit.
    maru_leaf* lf = p.leaf + pos;
    const maru_state* st = p.states + pos;
    const int w = st->width, h = st->height;
    if (cell_ok && on) {
      const int bx = compact ? x : x - (BOARD - w) / 2, by = compact ? gy - 1 : (gy - 1) - (BOARD - h) / 2;
      const int idx = by * w + bx;
      const bool allowed = !(st->flags & MARU_STATE_HAS_MASK) || ((st->move_mask[idx >> 3] >> (idx & 7)) & 1u);
      lf->prior[idx] = allowed ? e0 / t0 : 0.0f;
    }
    for (int i = w * h + tid; i < CELLS; i += nthreads) lf->prior[i] = 0.0f;   // tail of the fixed-size record
  }
  // ---- value / score MLP: fc1 rows are read coalesced-by-row by one warp each
  for (int j = warp; j < p.cv; j += nwarps) {
    const float* w = p.w_fc1 + (size_t)j * 2 * ch;
    float a = 0.0f;
    for (int k = lane; k < 2 * ch; k += 32) a = fmaf(w[k], s_pool[k], a);
    a = warp_sum(a);
    if (lane == 0) s_hid[j] = fmaxf(a + p.b_fc1[j], 0.0f);
  }
  __syncthreads();
  if (warp < 3) {
    const float* w = p.w_fc2 + (size_t)warp * p.cv;
    float a = 0.0f;
    for (int k = lane; k < p.cv; k += 32) a = fmaf(w[k], s_hid[k], a);
    a = warp_sum(a);
    if (lane == 0) {
      a += p.b_fc2[warp];
      const float v = (warp == 0) ? 1.0f / (1.0f + __expf(-a)) : a;
      if (out != nullptr) out[6 * CELLS + warp] = v;
      if (p.leaf != nullptr) {
        if (warp == 0) p.leaf[pos].value = v;
        else p.leaf[pos].score[warp - 1] = v;
      }
    }
  }
}

cudaError_t configure_heads() {
  return cudaFuncSetAttribute(heads_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_heads(const HeadParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  if (p.ch > HEAD_MAX_CH || p.ch % 8 != 0 || p.cv > HEAD_MAX_CV) return cudaErrorInvalidValue;
  // Compact frames leave most of a 416-thread CTA idle.  With many positions in flight the machine is better used by
  // small CTAs (128 threads for 9x9: 7 instead of 2 CTAs per SM; measured 42 -> 20 us at batch 1024); a lone wave of
  // CTAs is faster with all 13 warps sharing the value MLP (11 us vs 17 at batch 256), so small batches keep them.
  int threads = HEAD_THREADS;
  if (p.board_w > 0 && p.n > 592) {                     // more than two full-size CTAs per SM on 148 SMs
    threads = (p.pos_rows + 31) / 32 * 32;
    if (threads < 128) threads = 128;                   // the value MLP's last layer runs on warps 0..2
  }
  return launch_kernel(heads_kernel, dim3(p.n), dim3(threads), 0, stream, p.pdl != 0, p);
}

}  // namespace maru
