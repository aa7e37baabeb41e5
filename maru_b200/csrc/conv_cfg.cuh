// conv_cfg.cuh — shared-memory geometry of one conv layer, shared by the one-layer kernel
// (conv_gemm.cu) and the persistent multi-layer kernel (conv_chain.cu).
#pragma once
#include "kernels.h"

namespace maru {

constexpr int CONV_THREADS = 384;  // warp 0 producer, warps 1,10 MMA issue, warps 2..9 epilogue, warp 11 tile flags

template <int CIN, int NT, int TAPS>
struct ConvCfg {
  static constexpr int LEAD = (TAPS == 9) ? 24 : 0;            // slab rows before the tile
  static constexpr int SLAB_ROWS = TILE_M + 2 * LEAD;          // 176 or 128 (multiple of 8)
  static constexpr int KCH = CIN / 8;                          // 16-byte channel chunks
  static constexpr int W_BYTES = TAPS * CIN * NT * 2;
  static constexpr int ONES_BYTES = 2 * TILE_M * 16;           // A operand of the bias MMA  (K = 16)
  static constexpr int BIASB_BYTES = 2 * NT * 16;              // B operand of the bias MMA
  static constexpr int IDENT_BYTES = NT * NT * 2;              // B operand of the residual MMAs
  static constexpr int SLAB_BYTES = SLAB_ROWS * CIN * 2;
  static constexpr int RES_BYTES = TILE_M * NT * 2;            // residual tile [NT/8][128][16 B]
  static constexpr int MASK_BYTES = TILE_M;                    // board mask of the tile's rows
  static constexpr int MAX_STAGES = 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_LIMIT = 232448;                    // 227 KB per CTA
  static int fixed_bytes(bool has_res) {
    return W_BYTES + ONES_BYTES + BIASB_BYTES + BAR_BYTES + (has_res ? IDENT_BYTES : 0);
  }
  // a stage = input slab + mask (+ residual tile when the layer has one); as many as fit, <= 4
  static int stage_bytes(bool has_res) { return SLAB_BYTES + MASK_BYTES + (has_res ? RES_BYTES : 0); }
  static int stages(bool has_res) {
    int n = (SMEM_LIMIT - fixed_bytes(has_res)) / stage_bytes(has_res);
    return n > MAX_STAGES ? MAX_STAGES : n;
  }
  static int smem_bytes(bool has_res) { return fixed_bytes(has_res) + stages(has_res) * stage_bytes(has_res); }
  static constexpr int TMEM_COLS = (2 * NT <= 32) ? 32 : (2 * NT <= 64) ? 64 : (2 * NT <= 128) ? 128
                                   : (2 * NT <= 256) ? 256 : 512;
  static_assert(CIN % 16 == 0 && NT % 32 == 0 && NT >= 32 && NT <= 256, "UMMA shape");
  static_assert(TAPS == 1 || LEAD >= HALO, "slab lead must cover the 3x3 halo");
  static_assert(W_BYTES + ONES_BYTES + BIASB_BYTES + BAR_BYTES + SLAB_BYTES + MASK_BYTES <= SMEM_LIMIT,
                "one stage must fit");
};


}  // namespace maru
