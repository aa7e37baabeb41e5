// attention.cu — K4: multi-head self-attention core over the board tokens of one position.
//
//   O = softmax(Q K^T  with off-board KEYS masked to -inf) V        per (position, head)
//
// Replaces the ATen scaled-dot-product attention inside the TorchScript network (reference
// src/deepgo/native/cpp/Model.cpp:96).  Q is pre-scaled by log2(e)/sqrt(d) in the folded QKV
// weights (maru_b200/convert.py) so the softmax is a bare exp2.
//
// v1 implementation: warp-level online softmax with QK^T / PV on the tensor cores through
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  One CTA = (position, head): the head's K and V
// planes (400 padded rows x d) are bulk-copied to shared memory once — in the chunk-planar layout a
// plane slice of one position is a single contiguous 6 400-byte run — and 5 warps sweep the 25
// query tiles of 16 rows.  K fragments come from ldmatrix, V fragments from ldmatrix.trans, P never
// leaves registers (S accumulator fragment == A fragment of the PV MMA).
// Tokens are the 400 padded rows; halo / off-board rows are masked as keys and written as zeros as
// queries, which makes the result identical to attention over the w x h on-board cells.
#include "kernels.h"

namespace maru {

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
      __nv_bfloat16* dst = p.out + ((size_t)(q_chunk0 + i) * p.rows_alloc + prow0 + r0) * 8;
      *reinterpret_cast<uint32_t*>(dst + g * 8 + tq * 2) =
          on0 ? pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0) : 0u;
      *reinterpret_cast<uint32_t*>(dst + (g + 8) * 8 + tq * 2) =
          on1 ? pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1) : 0u;
    }
  }
}

template <int D>
static constexpr int attn_smem_bytes() { return 2 * (D / 8) * POS_ROWS * 16 + 512 + 16; }

cudaError_t configure_attention() {
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       attn_smem_bytes<32>());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              attn_smem_bytes<64>());
}

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  const int heads = p.channels / p.head_dim;
  if (p.head_dim == 32) {
    attention_kernel<32><<<p.n * heads, ATT_THREADS, attn_smem_bytes<32>(), stream>>>(p);
  } else if (p.head_dim == 64) {
    attention_kernel<64><<<p.n * heads, ATT_THREADS, attn_smem_bytes<64>(), stream>>>(p);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace maru
