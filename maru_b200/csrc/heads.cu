// heads.cu — K5: policy / territory / value heads, fused (sm_100a, HBM-bound).
//
// Consumes g = relu(bn(conv1x1(h, C -> Ch))) * M  (written by conv_gemm) and produces the
// reference output row float32[2169] = planes [6][19][19] + 3 scalars
// (reference src/deepgo/config.py:43-50; consumers Evaluator.cpp:66-80, player.py:415-461):
//   planes 0,1 : softmax over ON-BOARD cells of the two policy logits (off-board exactly 0)
//   planes 2..4: 3-way territory softmax per cell, times M
//   plane 5    : sigmoid, times M
//   scalars    : fc2(relu(fc1([mean_onboard(g), max(g)])));  [0] through sigmoid
// One CTA per position.  Algorithmic bytes per position: 361*Ch*2 read + 2169*4 written.
#include "kernels.h"

namespace maru {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_MAX_CH = 32;
constexpr int HEAD_MAX_CV = 128;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(HEAD_THREADS) heads_kernel(const HeadParams p) {
  __shared__ __align__(16) __nv_bfloat16 s_g[HEAD_MAX_CH / 8][POS_ROWS + 1][8];  // chunk-planar copy
  __shared__ float s_logit[2][POS_ROWS];
  __shared__ uint8_t s_mask[POS_ROWS];
  __shared__ float s_wp[6][HEAD_MAX_CH];
  __shared__ float s_bp[6];
  __shared__ float s_red[2][HEAD_THREADS / 32];
  __shared__ float s_part[8][2 * HEAD_MAX_CH];
  __shared__ float s_pool[2 * HEAD_MAX_CH];
  __shared__ float s_hid[HEAD_MAX_CV];
  __shared__ float s_stat[4];

  pdl_launch_dependents();
  const int pos = blockIdx.x;
  if (pos >= ((p.n_dev != nullptr) ? *p.n_dev : p.n)) return;
  const int tid = threadIdx.x;
  const int ch = p.ch;
  const int warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 6 * ch; i += HEAD_THREADS) s_wp[i / ch][i % ch] = p.w_planes[i];
  if (tid < 6) s_bp[tid] = p.b_planes[tid];
  pdl_wait();   // g and mask come from the previous kernels
  for (int r = tid; r < POS_ROWS; r += HEAD_THREADS) s_mask[r] = p.mask[(size_t)pos * POS_ROWS + r];
  // g tile: [ch/8 planes][400 rows][8] -> smem fp32 [row][ch]; 16-byte row-consecutive loads
  const int n_chunks = ch / 8;
  for (int i = tid; i < n_chunks * POS_ROWS; i += HEAD_THREADS) {
    const int c = i / POS_ROWS, r = i - c * POS_ROWS;
    const size_t grow = (size_t)GUARD_ROWS + (size_t)pos * POS_ROWS + r;
    *reinterpret_cast<uint4*>(&s_g[c][r][0]) =
        *reinterpret_cast<const uint4*>(p.g + ((size_t)c * p.rows_alloc + grow) * 8);
  }
  __syncthreads();

  float* out = p.out + (size_t)pos * OUT_ROW;

  // ---- per-cell logits; territory / point-value planes written directly
  float lmax0 = -INFINITY, lmax1 = -INFINITY;
  for (int r = tid; r < POS_ROWS; r += HEAD_THREADS) {
    const int gy = r / PITCH, x = r - gy * PITCH;
    if (gy < 1 || x >= BOARD) continue;
    const int cell = (gy - 1) * BOARD + x;
    const bool on = s_mask[r] != 0;
    float z[6];
#pragma unroll
    for (int k = 0; k < 6; k++) z[k] = s_bp[k];
    for (int c8 = 0; c8 < n_chunks; c8++) {
      const uint4 v = *reinterpret_cast<const uint4*>(&s_g[c8][r][0]);
      const float gv[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                           bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
      for (int e = 0; e < 8; e++) {
#pragma unroll
        for (int k = 0; k < 6; k++) z[k] = fmaf(s_wp[k][c8 * 8 + e], gv[e], z[k]);
      }
    }
    s_logit[0][r] = z[0];
    s_logit[1][r] = z[1];
    if (p.logits != nullptr) {
      p.logits[((size_t)pos * 2 + 0) * CELLS + cell] = z[0];
      p.logits[((size_t)pos * 2 + 1) * CELLS + cell] = z[1];
    }
    if (on) {
      lmax0 = fmaxf(lmax0, z[0]);
      lmax1 = fmaxf(lmax1, z[1]);
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
  // ---- policy softmax over on-board cells (two planes)
  lmax0 = warp_max(lmax0);
  lmax1 = warp_max(lmax1);
  if (lane == 0) {
    s_red[0][warp] = lmax0;
    s_red[1][warp] = lmax1;
  }
  __syncthreads();
  if (tid < 2) {
    float m = -INFINITY;
    for (int w = 0; w < HEAD_THREADS / 32; w++) m = fmaxf(m, s_red[tid][w]);
    s_stat[tid] = m;
  }
  __syncthreads();
  const float m0 = s_stat[0], m1 = s_stat[1];
  float sum0 = 0.0f, sum1 = 0.0f;
  for (int r = tid; r < POS_ROWS; r += HEAD_THREADS) {
    if (s_mask[r]) {
      const float e0 = __expf(s_logit[0][r] - m0), e1 = __expf(s_logit[1][r] - m1);
      s_logit[0][r] = e0;
      s_logit[1][r] = e1;
      sum0 += e0;
      sum1 += e1;
    }
  }
  sum0 = warp_sum(sum0);
  sum1 = warp_sum(sum1);
  __syncthreads();
  if (lane == 0) {
    s_red[0][warp] = sum0;
    s_red[1][warp] = sum1;
  }
  __syncthreads();
  if (tid < 2) {
    float s = 0.0f;
    for (int w = 0; w < HEAD_THREADS / 32; w++) s += s_red[tid][w];
    s_stat[2 + tid] = s;
  }
  __syncthreads();
  const float inv0 = 1.0f / s_stat[2], inv1 = 1.0f / s_stat[3];
  for (int r = tid; r < POS_ROWS; r += HEAD_THREADS) {
    const int gy = r / PITCH, x = r - gy * PITCH;
    if (gy < 1 || x >= BOARD) continue;
    const int cell = (gy - 1) * BOARD + x;
    const bool on = s_mask[r] != 0;
    out[0 * CELLS + cell] = on ? s_logit[0][r] * inv0 : 0.0f;
    out[1 * CELLS + cell] = on ? s_logit[1][r] * inv1 : 0.0f;
  }

  // ---- pooled features: mean over on-board cells and max (g >= 0, off-board rows are 0)
  {
    const int c = tid % HEAD_MAX_CH, grp = tid / HEAD_MAX_CH;  // 32 channels x 8 row groups
    float s = 0.0f, mx = 0.0f, cnt = 0.0f;
    if (c < ch) {
      for (int r = grp; r < POS_ROWS; r += 8) {
        if (s_mask[r]) {
          const float gv = __bfloat162float(s_g[c >> 3][r][c & 7]);
          s += gv;
          mx = fmaxf(mx, gv);
          cnt += 1.0f;
        }
      }
    }
    s_part[grp][c] = s;
    s_part[grp][HEAD_MAX_CH + c] = mx;
    if (c == 0) s_red[0][grp] = cnt;
  }
  __syncthreads();
  if (tid < ch) {
    float s = 0.0f, mx = 0.0f, cnt = 0.0f;
    for (int g8 = 0; g8 < 8; g8++) {
      s += s_part[g8][tid];
      mx = fmaxf(mx, s_part[g8][HEAD_MAX_CH + tid]);
      cnt += s_red[0][g8];
    }
    s_pool[tid] = s / cnt;
    s_pool[ch + tid] = mx;
  }
  __syncthreads();
  if (tid < p.cv) {
    float a = p.b_fc1[tid];
    const float* w = p.w_fc1 + (size_t)tid * 2 * ch;
    for (int k = 0; k < 2 * ch; k++) a = fmaf(w[k], s_pool[k], a);
    s_hid[tid] = fmaxf(a, 0.0f);
  }
  __syncthreads();
  if (tid < 3) {
    float a = p.b_fc2[tid];
    const float* w = p.w_fc2 + (size_t)tid * p.cv;
    for (int k = 0; k < p.cv; k++) a = fmaf(w[k], s_hid[k], a);
    out[6 * CELLS + tid] = (tid == 0) ? 1.0f / (1.0f + __expf(-a)) : a;
  }
}

cudaError_t configure_heads() {
  return cudaFuncSetAttribute(heads_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_heads(const HeadParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  if (p.ch > HEAD_MAX_CH || p.ch % 8 != 0 || p.cv > HEAD_MAX_CV) return cudaErrorInvalidValue;
  return launch_kernel(heads_kernel, dim3(p.n), dim3(HEAD_THREADS), 0, stream, p.pdl != 0, p);
}

}  // namespace maru
