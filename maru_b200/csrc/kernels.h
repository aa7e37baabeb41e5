// kernels.h — launcher interface of the sm_100a kernels (internal to libmaru_b200.so).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "maru_b200.h"

namespace maru {

// Chunk-planar activation tensors:  [C/8][rows_alloc][8] bf16.  Row r of the padded-flat geometry
// lives at plane row GUARD_ROWS + r; the GUARD_ROWS rows in front of every plane stay zero so that
// the 3x3 slab of the first tile can be fetched without bounds logic.
constexpr int GUARD_ROWS = 32;
inline int rows_alloc_for(int max_batch) {
  const int m = total_rows(max_batch);
  return GUARD_ROWS + ((m + TILE_M - 1) / TILE_M) * TILE_M + 32;
}

struct ConvParams {
  const __nv_bfloat16* in;   // input planes
  int in_chunk0;             // first input plane (channel/8)
  const __nv_bfloat16* w;    // [taps][cin/8][cout_total][8]  (UMMA K-major SWIZZLE_NONE image)
  const float* bias;         // [cout_total]
  const uint8_t* blob;       // optional [groups][W image of the group | bias operand image]: one bulk copy per CTA
  const __nv_bfloat16* res;  // optional residual planes (may alias out)
  int res_chunk0;
  __nv_bfloat16* out;
  int out_chunk0;
  const uint8_t* mask;       // [m_rows] 1 = on-board cell
  int rows_alloc;            // plane stride in rows (same for in/res/out)
  int m_rows;                // total_rows_g(n, pos_rows)
  int pitch, pos_rows;       // frame geometry (common.cuh): 20 / 400 for the 19x19 frame, w + 1 / (w + 1)(h + 1) compact
  int cout_total;
  int relu;
  int stages;                // pipeline depth (filled by the launcher from the smem budget)
  int pdl;                   // launch with programmatic dependent launch (see common.cuh)
  // Tensors that are only ever consumed tile-locally (1x1 convs, residuals: the C-channel trunk h
  // and the attention output o) are stored TILE-BLOCKED: [tile of 128 rows][plane][128 rows][8 ch],
  // so that a tile's input slab / residual tile is ONE contiguous bulk copy (the TMA unit accepts a
  // request only every ~50-70 cycles; 33 per-plane requests per tile made the 1x1 layers
  // request-bound). Tensors read with a 3x3 halo or per position stay chunk-planar.
  // Cross-layer pipelining: instead of waiting for the whole previous grid (griddepcontrol.wait), a
  // conv that consumes another conv's output waits only for the 3 (3x3) or 1 (1x1) producer tiles
  // it reads. Every epilogue warp bumps out_flags[tile] after its stores; the consumer's producer
  // warp polls in_flags[tile-1..tile+1] >= in_expect. nullptr => plain grid dependency.
  const unsigned* in_flags;
  unsigned in_expect;
  unsigned* out_flags;
  int in_blocked, res_blocked, out_blocked;
  int in_planes, res_planes, out_planes;   // total planes (channels / 8) of those tensors
  const int* n_dev;          // optional: number of positions read on the device (graph replay with
                             // a smaller batch than the one captured); overrides m_rows
  long long* trace;          // optional [gridDim.x][TRACE_TILES][TRACE_SLOTS] clock64 stamps (bring-up)
  long long* timeline;       // optional [launch seq][2 CTAs][8] %globaltimer stamps of every conv launch
  int seq;
};
constexpr int TRACE_TILES = 32, TRACE_SLOTS = 8;

// ---- persistent multi-layer conv kernel (conv_flow.cu) ----------------------------------------
// A CHAIN of consecutive conv layers runs in ONE kernel, one CTA per SM.  Every CTA owns a contiguous
// range of M tiles for all layers; dependencies between layers are tracked per tile inside the CTA
// (shared-memory counters) and, for the two range ends, by gpu-scope flags between neighbouring CTAs.
// No grid barrier, no kernel boundary: a launch boundary costs ~3.5 of the ~10 us of a batch-256 layer.
struct FlowLayer {
  const __nv_bfloat16* in;
  const __nv_bfloat16* res;    // optional residual, same channels as the output (may alias out)
  __nv_bfloat16* out;
  const uint8_t* blob;         // [W image of the nt output channels | bias image 2*nt*16 B]: one bulk copy
  int in_blocked, in_planes, res_blocked, res_planes, out_blocked, out_planes;
  int out_chunk0;              // first output plane (channel / 8) this layer writes
  int cin, nt, taps, relu;
  int dep_back;                // the input was written by layer l - dep_back of this launch (0: by an earlier kernel)
};
constexpr int FLOW_MAX = 24;           // layers per chain
constexpr int FLOW_MAX_TILES = 48;     // M tiles per CTA
struct FlowParams {
  FlowLayer layer[FLOW_MAX];
  int n_layers;
  const uint8_t* mask;
  int rows_alloc;
  int m_rows;
  int pitch, pos_rows;         // frame geometry, as in ConvParams
  const int* n_dev;
  int w_region;                // shared-memory bytes of ONE weight region (two regions, double-buffered)
  int area;                    // bytes of the slab ring
  unsigned* flags;             // [2 * grid] first / last tile of CTA b finished layers < value - flag_base
  unsigned flag_base;          // flag value before the first layer of this launch (flags only ever grow)
  long long* stamps;           // optional [n_layers + 1][4] %globaltimer of CTA 0: layer start, weights landed, done
  long long* trace;            // optional [256 items][8] %globaltimer stamps of CTA trace_cta (bring-up, conv_flow.cu)
  int trace_cta;
  int dbg;                     // timing experiments only: bit0 skip the epilogue's proxy fence, bit1 its gpu fence
  int pdl;
};
cudaError_t configure_conv_flow();
bool conv_flow_supports(int cin, int nt, int taps);
int conv_flow_blob_bytes(int cin, int nt, int taps);
int conv_flow_area(int w_region);
// does a layer fit a chain whose weight regions are w_region bytes each (blob fits, two slabs fit the ring)?
bool conv_flow_fits(int cin, int nt, int taps, int w_region);
// fills w_region / area; returns cudaErrorInvalidConfiguration when a layer does not fit
cudaError_t launch_conv_flow(FlowParams& fp, int sm_count, cudaStream_t stream);
int conv_flow_ctas_per_sm();

// ---- 3x3, 128 input channels, half-slab pipeline (conv_ksplit.cu) -----------------------------------
cudaError_t configure_conv_ksplit();
bool conv_ksplit_supports(int cin, int nt, int taps);
cudaError_t launch_conv_ksplit(ConvParams p, int sm_count, cudaStream_t stream);

// ---- 1x1, 256 input channels, 256 output channels per CTA (conv_wide.cu) ------------------------------
cudaError_t configure_conv_wide();
bool conv_wide_supports(int cin, int cout, int taps);
cudaError_t launch_conv_wide(ConvParams p, int sm_count, cudaStream_t stream);

// ---- 3x3 128 -> 128 on CTA pairs, cta_group::2 (conv_pair.cu) ------------------------------------------
cudaError_t configure_conv_pair();
bool conv_pair_supports(int cin, int cout, int taps);
int conv_pair_blob_nt(int cout);
cudaError_t launch_conv_pair(ConvParams p, int cin, int sm_count, cudaStream_t stream);

cudaError_t configure_conv_gemm();
cudaError_t launch_conv_gemm(const ConvParams& p, int cin, int nt, int taps, int sm_count,
                             cudaStream_t stream);

// ---- K1: feature encoding (encode.cu) ------------------------------------------------------
// compact records -> reference fp32 rows [n][11920]   (bit-exact Board::getInputs)
// All launchers take `n` (grid sizing) and an optional device pointer `n_dev` holding the live
// number of positions (<= n), so that one captured CUDA graph serves every batch size of a bucket.
cudaError_t launch_encode_rows(const maru_state* states, float* rows, int n, const int* n_dev,
                               cudaStream_t stream);
// compact records -> trunk input planes (STEM_CIN channels) + board mask
cudaError_t launch_encode_trunk(const maru_state* states, __nv_bfloat16* x0, uint8_t* mask,
                                int rows_alloc, int n, const int* n_dev, cudaStream_t stream);
// reference fp32 rows -> trunk input planes + board mask
cudaError_t launch_ingest_rows(const float* rows, __nv_bfloat16* x0, uint8_t* mask, int rows_alloc,
                               int n, const int* n_dev, cudaStream_t stream);
// trunk input planes -> fp32 [n][channels][361] (debug / parity)
cudaError_t launch_planes_to_nchw(const __nv_bfloat16* x, int rows_alloc, int channels, int total_channels,
                                  int blocked, float* out, int n, int pitch, int pos_rows, int bw, int bh,
                                  cudaStream_t stream);
// stem output + mask on the 19x19 frame -> compact frame of a w x h board (encode.cu: crop_kernel)
cudaError_t launch_crop(const __nv_bfloat16* src, __nv_bfloat16* dst, const uint8_t* mask_src, uint8_t* mask_dst,
                        int channels, int src_blocked, int dst_blocked, int rows_alloc, int w, int h, int n,
                        const int* n_dev, int pdl, cudaStream_t stream);

// ---- K4: attention core (attention.cu) -------------------------------------------------------
struct AttnParams {
  const __nv_bfloat16* qkv;  // planes: q [0,C), k [C,2C), v [2C,3C); q pre-scaled by log2(e)/sqrt(d)
  __nv_bfloat16* out;        // planes [0,C)
  int out_blocked;           // output tensor is tile-blocked (consumed by the 1x1 projection only)
  const uint8_t* mask;
  int rows_alloc;
  int n;                     // positions
  int channels;              // C
  int head_dim;              // 32 or 64
  int pitch, pos_rows;       // frame geometry (20 / 400, or the compact frame of a small board)
  const int* n_dev;
  int pdl;
  int dbg;                   // bit0: swap LBO/SBO of the MN-major V descriptor, bit1: legacy mma.sync kernel, bit2: always shift
  long long* trace;          // EXPERIMENTS builds: [64 steps][8] clock64 stamps of CTA trace_cta (tools/att_trace.py)
  int trace_cta;
};
cudaError_t configure_attention();
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream);

// ---- K5: heads (heads.cu) --------------------------------------------------------------------
struct HeadParams {
  const __nv_bfloat16* g;    // planes [0,Ch): relu(bn(conv1x1(h))) * mask
  const uint8_t* mask;
  const float* w_planes;     // [6][Ch]
  const float* b_planes;     // [6]
  const float* w_fc1;        // [Cv][2*Ch]
  const float* b_fc1;        // [Cv]
  const float* w_fc2;        // [3][Cv]
  const float* b_fc2;        // [3]
  float* out;                // [n][2169] reference rows, or nullptr
  maru_leaf* leaf;           // [n] compact results (masked move priors + value scalars), or nullptr
  const maru_state* states;  // records of the batch (board geometry + move_mask), needed with `leaf`
  float* logits;             // optional [n][2][361] pre-softmax policy logits
  int rows_alloc;
  int pitch, pos_rows;       // frame geometry of g / mask
  int board_w, board_h;      // compact frame: the board it holds (outputs are written centred in 19x19); 0 = full frame
  int n;
  int ch;                    // head channels (32)
  int cv;                    // value channels (<= 128)
  const int* n_dev;
  int pdl;
};
cudaError_t launch_heads(const HeadParams& p, cudaStream_t stream);
cudaError_t configure_misc_kernels();   // same shared-memory carveout for encode / heads kernels

}  // namespace maru
