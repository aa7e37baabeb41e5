// encode.cu — K1: board-to-feature-plane encoding on the GPU (bit-exact, integer logic).
//
// Restates deepgo::Board::getInputs (reference src/deepgo/native/cpp/Board.cpp:494-623) over the
// compact maru_state record (include/maru_b200.h) instead of the Board object:
//   plane 0        empty on-board cell                                   Board.cpp:517-520
//   planes 1..10   own stone / own ladder flag / own liberties one-hot   Board.cpp:523-531
//   planes 11..13  own last three non-pass moves                         Board.cpp:542-556
//   planes 14..23  opponent stone / ladder / liberties                   Board.cpp:533-537
//   planes 24..26  opponent last three moves                             Board.cpp:558-566
//   planes 27..30  1st..4th line rings (loop bounds replicated)          Board.cpp:569-584
//   plane 31       ko point WITHOUT the centring offset (reference quirk) Board.cpp:587-593
//   plane 32       on-board mask                                         Board.cpp:514
//   scalars        turn, komi*color/13 (double division), superko, ko, rule  Board.cpp:596-622
//
// Three kernels share one device function `plane_value`:
//   encode_rows_kernel   : records -> reference fp32 rows [n][11920], float4 coalesced stores
//   encode_trunk_kernel  : records -> trunk input planes (chunk-planar bf16, 16-byte stores) + mask
//   ingest_rows_kernel   : reference fp32 rows -> the same trunk input planes (+ mask)
#include "kernels.h"

namespace maru {

struct StateView {
  const uint8_t* cells;
  int w, h, ox, oy;
  int color;  // +1 / -1
  int own_code, opp_code;
  int ko_x, ko_y;  // -1 if none
  int hist[2][3];  // [0]=own,[1]=opp: model cell index of i-th newest move, -1 if none
  int rule, superko;
  float komi;
};

__device__ __forceinline__ void load_state_view(const maru_state* s, StateView& v) {
  v.cells = s->cells;
  v.w = s->width;
  v.h = s->height;
  v.ox = (BOARD - v.w) / 2;  // Board.cpp:496 with _width = w + 2
  v.oy = (BOARD - v.h) / 2;
  v.color = s->color;
  v.own_code = (v.color == MARU_BLACK) ? MARU_CELL_BLACK : MARU_CELL_WHITE;
  v.opp_code = (v.color == MARU_BLACK) ? MARU_CELL_WHITE : MARU_CELL_BLACK;
  v.ko_x = (s->ko_x == MARU_NONE) ? -1 : s->ko_x;
  v.ko_y = (s->ko_y == MARU_NONE) ? -1 : s->ko_y;
  const int own_side = (v.color == MARU_BLACK) ? 0 : 1;  // _histories[(1 - color) / 2]
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int side = k == 0 ? own_side : 1 - own_side;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int x = s->hist[side][i][0], y = s->hist[side][i][1];
      v.hist[k][i] = (x == MARU_NONE || y == MARU_NONE) ? -1 : (v.oy + y) * BOARD + (v.ox + x);
    }
  }
  v.rule = s->rule;
  v.superko = s->superko;
  v.komi = s->komi;
}

// value of input plane p (0..32) at model cell (my, mx); exactly 0 or 1
__device__ __forceinline__ int plane_value(const StateView& v, int p, int my, int mx) {
  const int bx = mx - v.ox, by = my - v.oy;
  const bool on = (bx >= 0) && (bx < v.w) && (by >= 0) && (by < v.h);
  if (p == 32) return on;
  if (p == 31) return (v.ko_x >= 0) && (my == v.ko_y) && (mx == v.ko_x);  // no offset: reference quirk
  if (p >= 27) {
    const int i = p - 27;
    const int bx0 = v.ox + i, ex = v.ox + v.w - i, by0 = v.oy + i, ey = v.oy + v.h - i;
    const bool in_y = (my >= by0) && (my < ey);
    const bool in_x = (mx >= bx0) && (mx < ex);
    return (in_y && (mx == bx0 || mx == ex - 1)) || (in_x && (my == by0 || my == ey - 1));
  }
  if (p >= 24) return v.hist[1][p - 24] == my * BOARD + mx;
  if (p >= 11 && p <= 13) return v.hist[0][p - 11] == my * BOARD + mx;
  if (!on) return 0;
  const int cell = v.cells[by * v.w + bx];
  const int code = cell & MARU_CELL_COLOR_MASK;
  if (p == 0) return code == MARU_CELL_EMPTY;
  const int side_code = (p <= 10) ? v.own_code : v.opp_code;
  if (code != side_code) return 0;
  const int q = (p <= 10) ? p : p - 13;  // 1: stone, 2: ladder, 3..10: liberties 1..8
  if (q == 1) return 1;
  if (q == 2) return (cell & MARU_CELL_LADDER) != 0;
  return ((cell >> MARU_CELL_LIB_SHIFT) & 15) == q - 2;
}

__device__ __forceinline__ float info_value(const StateView& v, int i) {
  switch (i) {
    case 0: return v.color == MARU_BLACK ? 1.0f : 0.0f;
    case 1: return v.color == MARU_BLACK ? 0.0f : 1.0f;
    // Board.cpp:605  inputs[..] = (komi * color) / 13.0;  float*int -> float, then double division
    case 2: return (float)((double)(v.komi * (float)v.color) / 13.0);
    case 3: return v.superko ? 1.0f : 0.0f;
    case 4: return v.ko_x >= 0 ? 1.0f : 0.0f;
    case 5: return v.rule != MARU_RULE_JP ? 1.0f : 0.0f;
    default: return v.rule == MARU_RULE_JP ? 1.0f : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------
// records -> reference rows.  One CTA per position; 2980 float4 stores per row, fully coalesced.
// Algorithmic bytes per position: 448 read + 47 680 written.
__global__ void __launch_bounds__(256) encode_rows_kernel(const maru_state* __restrict__ states,
                                                          float* __restrict__ rows, int n,
                                                          const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = *n_dev;
  __shared__ __align__(16) maru_state s_state;
  __shared__ StateView s_view;
  for (int pos = blockIdx.x; pos < n; pos += gridDim.x) {
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(states + pos);
    if (threadIdx.x < sizeof(maru_state) / 16)
      reinterpret_cast<uint4*>(&s_state)[threadIdx.x] = src[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) load_state_view(&s_state, s_view);
    __syncthreads();
    const StateView v = s_view;
    float4* dst = reinterpret_cast<float4*>(rows + (size_t)pos * IN_ROW);
    for (int i4 = threadIdx.x; i4 < IN_ROW / 4; i4 += blockDim.x) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int e = i4 * 4 + k;
        if (e < IN_PLANES * CELLS) {
          const int p = e / CELLS, c = e - p * CELLS;
          const int my = c / BOARD, mx = c - my * BOARD;
          o[k] = plane_value(v, p, my, mx) ? 1.0f : 0.0f;
        } else {
          o[k] = info_value(v, e - IN_PLANES * CELLS);
        }
      }
      dst[i4] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// bf16 split of the komi scalar so that the stem GEMM sees it with ~16 mantissa bits
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Packs the 64 trunk-input channels of one on-board/off-board row into 8 16-byte chunks.
// channel c<33: plane c; 33..39: info scalars (on-board cells only); 40: komi low part; else 0.
__device__ __forceinline__ void store_trunk_row(__nv_bfloat16* x0, int rows_alloc, size_t grow,
                                                unsigned long long bits, bool on,
                                                const __nv_bfloat16* info9) {
  // bits: bit c = plane c value (c < 33)
#pragma unroll
  for (int ch = 0; ch < STEM_CIN / 8; ch++) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      uint32_t lohi = 0;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int c = ch * 8 + e * 2 + half;
        uint32_t b16 = 0;
        if (c < IN_PLANES) {
          b16 = ((bits >> c) & 1ull) ? 0x3F80u : 0u;  // bf16(1.0)
        } else if (c <= STEM_KOMI_LO) {
          b16 = on ? (uint32_t)__bfloat16_as_ushort(info9[c - STEM_INFO0]) : 0u;
        }
        lohi |= b16 << (16 * half);
      }
      w[e] = lohi;
    }
    *reinterpret_cast<uint4*>(x0 + ((size_t)ch * rows_alloc + grow) * 8) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// records -> trunk input planes.  One CTA per position, one thread per padded row (400 rows).
// Algorithmic bytes per position: 448 read + 361*33*2 = 23 826 written; the padded
// 64-channel, 400-row image actually written is 51 200 B.
__global__ void __launch_bounds__(416) encode_trunk_kernel(const maru_state* __restrict__ states,
                                                           __nv_bfloat16* __restrict__ x0,
                                                           uint8_t* __restrict__ mask,
                                                           int rows_alloc, int n,
                                                           const int* __restrict__ n_dev) {
  __shared__ __align__(16) maru_state s_state;
  __shared__ StateView s_view;
  __shared__ __nv_bfloat16 s_info[9];
  pdl_launch_dependents();
  if (n_dev != nullptr) n = *n_dev;
  const int pos = blockIdx.x;
  if (pos >= n) {
    // the CTA after the last position writes the zero tail rows
    if (pos == n) {
      for (int r = threadIdx.x; r < TAIL_ROWS; r += blockDim.x) {
        const size_t grow = (size_t)GUARD_ROWS + (size_t)n * POS_ROWS + r;
        for (int ch = 0; ch < STEM_CIN / 8; ch++)
          *reinterpret_cast<uint4*>(x0 + ((size_t)ch * rows_alloc + grow) * 8) = make_uint4(0, 0, 0, 0);
        mask[(size_t)n * POS_ROWS + r] = 0;
      }
    }
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(states + pos);
  if (threadIdx.x < sizeof(maru_state) / 16)
    reinterpret_cast<uint4*>(&s_state)[threadIdx.x] = src[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    load_state_view(&s_state, s_view);
    for (int i = 0; i < 7; i++) s_info[i] = __float2bfloat16_rn(info_value(s_view, i));
    __nv_bfloat16 hi, lo;
    split_bf16(info_value(s_view, 2), hi, lo);
    s_info[2] = hi;
    s_info[7] = lo;  // channel 40
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= POS_ROWS) return;
  const StateView v = s_view;
  const int gy = r / PITCH, x = r - gy * PITCH;
  unsigned long long bits = 0;
  bool on = false;
  if (gy >= 1 && x < BOARD) {
    const int my = gy - 1, mx = x;
#pragma unroll
    for (int p = 0; p < IN_PLANES; p++)
      bits |= (unsigned long long)(plane_value(v, p, my, mx) & 1) << p;
    on = (bits >> 32) & 1ull;
  }
  // channels 33..39 <- infos 0..6, channel 40 <- komi low part
  __nv_bfloat16 info9[8];
#pragma unroll
  for (int i = 0; i < 8; i++) info9[i] = s_info[i];
  const size_t grow = (size_t)GUARD_ROWS + (size_t)pos * POS_ROWS + r;
  store_trunk_row(x0, rows_alloc, grow, bits, on, info9);
  mask[(size_t)pos * POS_ROWS + r] = on ? 1 : 0;
}

// reference rows -> trunk input planes.  Same output as encode_trunk_kernel for the same position.
// Reads 47 680 B per position (plane-major, coalesced along cells).
__global__ void __launch_bounds__(416) ingest_rows_kernel(const float* __restrict__ rows,
                                                          __nv_bfloat16* __restrict__ x0,
                                                          uint8_t* __restrict__ mask, int rows_alloc,
                                                          int n, const int* __restrict__ n_dev) {
  __shared__ __nv_bfloat16 s_info[9];
  pdl_launch_dependents();
  if (n_dev != nullptr) n = *n_dev;
  const int pos = blockIdx.x;
  if (pos >= n) {
    if (pos == n) {
      for (int r = threadIdx.x; r < TAIL_ROWS; r += blockDim.x) {
        const size_t grow = (size_t)GUARD_ROWS + (size_t)n * POS_ROWS + r;
        for (int ch = 0; ch < STEM_CIN / 8; ch++)
          *reinterpret_cast<uint4*>(x0 + ((size_t)ch * rows_alloc + grow) * 8) = make_uint4(0, 0, 0, 0);
        mask[(size_t)n * POS_ROWS + r] = 0;
      }
    }
    return;
  }
  const float* row = rows + (size_t)pos * IN_ROW;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; i++) s_info[i] = __float2bfloat16_rn(row[IN_PLANES * CELLS + i]);
    __nv_bfloat16 hi, lo;
    split_bf16(row[IN_PLANES * CELLS + 2], hi, lo);
    s_info[2] = hi;
    s_info[7] = lo;
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= POS_ROWS) return;
  const int gy = r / PITCH, x = r - gy * PITCH;
  unsigned long long bits = 0;
  bool on = false;
  if (gy >= 1 && x < BOARD) {
    const int c = (gy - 1) * BOARD + x;
#pragma unroll
    for (int p = 0; p < IN_PLANES; p++)
      bits |= (unsigned long long)(row[p * CELLS + c] != 0.0f ? 1 : 0) << p;
    on = (bits >> 32) & 1ull;
  }
  __nv_bfloat16 info9[8];
#pragma unroll
  for (int i = 0; i < 8; i++) info9[i] = s_info[i];
  const size_t grow = (size_t)GUARD_ROWS + (size_t)pos * POS_ROWS + r;
  store_trunk_row(x0, rows_alloc, grow, bits, on, info9);
  mask[(size_t)pos * POS_ROWS + r] = on ? 1 : 0;
}

// planes -> fp32 [n][channels][361] (parity/debug only); (pitch, pos_rows, bw, bh) = frame geometry of the tensor
// (bw = 0: the 19x19 frame; else the compact frame of a bw x bh board, reported centred in 19x19, zeros around it)
__global__ void planes_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int rows_alloc,
                                      int channels, int total_channels, int blocked,
                                      float* __restrict__ out, int n, int pitch, int pos_rows, int bw, int bh) {
  const size_t total = (size_t)n * channels * CELLS;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int cell = i % CELLS;
    const int c = (i / CELLS) % channels;
    const int pos = i / ((size_t)CELLS * channels);
    int my = cell / BOARD, mx = cell - my * BOARD;
    if (bw > 0) {
      mx -= (BOARD - bw) / 2;
      my -= (BOARD - bh) / 2;
      if (mx < 0 || my < 0 || mx >= bw || my >= bh) {
        out[i] = 0.0f;
        continue;
      }
    }
    const int m = pos * pos_rows + (my + 1) * pitch + mx;
    const size_t off = blocked ? blocked_offset(total_channels / 8, c / 8, m)
                               : ((size_t)(c / 8) * rows_alloc + GUARD_ROWS + m) * 8;
    out[i] = __bfloat162float(x[off + (c % 8)]);
  }
}

// ------------------------------------------------------------------------------------------ crop to the compact frame
// The reference embeds a w x h board centred in its 19x19 model frame (Board.cpp:496-497) and computes all 361
// cells.  Everything outside the board is zero after every layer (board mask), so the trunk of a batch of small
// boards can run on a frame that holds just the board: pitch w + 1, (w + 1)(h + 1) rows per position — 100 rows for
// 9x9 instead of 400.  Only the STEM needs the reference's frame: its input carries the ko plane at UN-centred
// coordinates on small boards (Board.cpp:587-593), possibly in the ring around the board.  So the stem runs on the
// 19x19 frame and this kernel moves its (masked) output, and the board mask, to the compact frame.
// One CTA per position (+ one for the zero tail rows), one thread per compact row, 16-byte moves per plane.
constexpr int CROP_THREADS = 416;
__global__ void __launch_bounds__(CROP_THREADS) crop_kernel(const __nv_bfloat16* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst,
                                                            const uint8_t* __restrict__ mask_src,
                                                            uint8_t* __restrict__ mask_dst, int planes, int src_blocked,
                                                            int dst_blocked, int rows_alloc, int w, int h, int n,
                                                            const int* __restrict__ n_dev) {
  pdl_launch_dependents();
  const int live = (n_dev != nullptr) ? *n_dev : n;
  const int pos = blockIdx.x;
  if (pos > live) return;
  pdl_wait();
  const int pitch = w + 1, pos_rows = (w + 1) * (h + 1);
  const int r = threadIdx.x;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (pos == live) {                                   // zero rows after the last position
    if (r < TAIL_ROWS) {
      const int m = live * pos_rows + r;
      for (int c = 0; c < planes; c++) {
        const size_t off = dst_blocked ? blocked_offset(planes, c, m) : ((size_t)c * rows_alloc + GUARD_ROWS + m) * 8;
        *reinterpret_cast<uint4*>(dst + off) = zero;
      }
      mask_dst[m] = 0;
    }
    return;
  }
  if (r >= pos_rows) return;
  const int gy = r / pitch, x = r - gy * pitch;
  const bool cell = gy >= 1 && x < w;
  const int ms = pos * POS_ROWS + flat_row((BOARD - h) / 2 + gy - 1, (BOARD - w) / 2 + x);   // row in the 19x19 frame
  const int md = pos * pos_rows + r;
  const uint8_t on = cell ? mask_src[ms] : 0;
  mask_dst[md] = on;
  for (int c = 0; c < planes; c++) {
    uint4 v = zero;
    if (on) {
      const size_t so = src_blocked ? blocked_offset(planes, c, ms) : ((size_t)c * rows_alloc + GUARD_ROWS + ms) * 8;
      v = *reinterpret_cast<const uint4*>(src + so);
    }
    const size_t off = dst_blocked ? blocked_offset(planes, c, md) : ((size_t)c * rows_alloc + GUARD_ROWS + md) * 8;
    *reinterpret_cast<uint4*>(dst + off) = v;
  }
}

cudaError_t configure_heads();
cudaError_t configure_misc_kernels() {
  cudaError_t e = configure_heads();
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(encode_trunk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ingest_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_encode_rows(const maru_state* states, float* rows, int n, const int* n_dev,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  encode_rows_kernel<<<n, 256, 0, stream>>>(states, rows, n, n_dev);
  return cudaGetLastError();
}
cudaError_t launch_encode_trunk(const maru_state* states, __nv_bfloat16* x0, uint8_t* mask,
                                int rows_alloc, int n, const int* n_dev, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  encode_trunk_kernel<<<n + 1, 416, 0, stream>>>(states, x0, mask, rows_alloc, n, n_dev);
  return cudaGetLastError();
}
cudaError_t launch_ingest_rows(const float* rows, __nv_bfloat16* x0, uint8_t* mask, int rows_alloc,
                               int n, const int* n_dev, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  ingest_rows_kernel<<<n + 1, 416, 0, stream>>>(rows, x0, mask, rows_alloc, n, n_dev);
  return cudaGetLastError();
}
cudaError_t launch_planes_to_nchw(const __nv_bfloat16* x, int rows_alloc, int channels, int total_channels,
                                  int blocked, float* out, int n, int pitch, int pos_rows, int bw, int bh,
                                  cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  planes_to_nchw_kernel<<<256, 256, 0, stream>>>(x, rows_alloc, channels, total_channels, blocked, out, n, pitch,
                                                 pos_rows, bw, bh);
  return cudaGetLastError();
}

cudaError_t launch_crop(const __nv_bfloat16* src, __nv_bfloat16* dst, const uint8_t* mask_src, uint8_t* mask_dst,
                        int channels, int src_blocked, int dst_blocked, int rows_alloc, int w, int h, int n,
                        const int* n_dev, int pdl, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (w < 1 || h < 1 || w > BOARD || h > BOARD || (w + 1) * (h + 1) > CROP_THREADS) return cudaErrorInvalidValue;
  return launch_kernel(crop_kernel, dim3(n + 1), dim3(CROP_THREADS), 0, stream, pdl != 0, src, dst, mask_src, mask_dst,
                       channels / 8, src_blocked, dst_blocked, rows_alloc, w, h, n, n_dev);
}

}  // namespace maru
