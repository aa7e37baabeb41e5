/*
 * oracle/features.c — TEST INFRASTRUCTURE ONLY (oracle "port"). Never linked into the product.
 *
 * Plain-C restatement of deepgo::Board::getInputs (reference
 * src/deepgo/native/cpp/Board.cpp:494-623) over the compact maru_state record. It follows the
 * reference statement by statement (zero-fill, stone loop, history, line rings, ko, scalars) so
 * that it can be diffed against the C++ by eye; the CUDA kernel (maru_b200/csrc/encode.cu) is
 * written the other way round (value-per-output-element), which makes the two independent.
 *
 * Pinned: tests/test_oracle_features.py checks it bit-for-bit against rows produced by the
 * reference's own Board::getInputs (oracle/_ref live, and tests/golden/positions_*.npz).
 */
#include <string.h>

#include "maru_b200.h"

#define MS MARU_MODEL_SIZE
#define LEN (MS * MS)

/* Board.cpp:494-623 */
void oracle_get_inputs(const maru_state* st, float* inputs) {
  const int w = st->width, h = st->height, color = st->color;
  const int offset_x = (MS - w) / 2; /* :496  (MODEL_SIZE - _width + 2) / 2 with _width = w + 2 */
  const int offset_y = (MS - h) / 2; /* :497 */
  int x, y, i;

  for (i = 0; i < MARU_INPUT_SIZE; i++) inputs[i] = 0.0f; /* :503-505 */

  for (y = 0; y < h; y++) { /* :508-539 */
    for (x = 0; x < w; x++) {
      const int cell = st->cells[y * w + x];
      const int code = cell & MARU_CELL_COLOR_MASK;
      const int index = (offset_y + y) * MS + (offset_x + x);
      inputs[LEN * MARU_MODEL_FEATURES + index] = 1.0f; /* :514 mask */
      if (code == MARU_CELL_EMPTY) {                    /* :517-520 */
        inputs[LEN * 0 + index] = 1.0f;
        continue;
      }
      {
        const int stone = (code == MARU_CELL_BLACK) ? MARU_BLACK : MARU_WHITE;
        const float shicho = (cell & MARU_CELL_LADDER) ? 1.0f : 0.0f; /* :523 */
        int size = (cell >> MARU_CELL_LIB_SHIFT) & 15;                /* :524 min(liberties, 8) */
        if (size > 8) size = 8;
        if (stone * color == MARU_BLACK) { /* :527-531 own */
          inputs[LEN * 1 + index] = 1.0f;
          inputs[LEN * 2 + index] = shicho;
          inputs[LEN * (2 + size) + index] = 1.0f;
        } else { /* :533-537 opponent */
          inputs[LEN * 14 + index] = 1.0f;
          inputs[LEN * 15 + index] = shicho;
          inputs[LEN * (15 + size) + index] = 1.0f;
        }
      }
    }
  }

  /* :542-566 history: own side = _histories[(1 - color) / 2], newest first */
  {
    const int own = (1 - color) / 2, opp = (1 + color) / 2;
    for (i = 0; i < 3; i++) {
      if (st->hist[own][i][0] != MARU_NONE && st->hist[own][i][1] != MARU_NONE) {
        const int index = (offset_y + st->hist[own][i][1]) * MS + (offset_x + st->hist[own][i][0]);
        inputs[LEN * (11 + i) + index] = 1.0f;
      }
      if (st->hist[opp][i][0] != MARU_NONE && st->hist[opp][i][1] != MARU_NONE) {
        const int index = (offset_y + st->hist[opp][i][1]) * MS + (offset_x + st->hist[opp][i][0]);
        inputs[LEN * (24 + i) + index] = 1.0f;
      }
    }
  }

  /* :569-584 lines 1-4 (loop bounds kept verbatim, degenerate cases included) */
  for (i = 0; i < 4; i++) {
    const int begin_x = offset_x + i, end_x = offset_x + w - i;
    const int begin_y = offset_y + i, end_y = offset_y + h - i;
    for (y = begin_y; y < end_y; y++) {
      inputs[LEN * (27 + i) + y * MS + begin_x] = 1.0f;
      inputs[LEN * (27 + i) + y * MS + end_x - 1] = 1.0f;
    }
    for (x = begin_x; x < end_x; x++) {
      inputs[LEN * (27 + i) + begin_y * MS + x] = 1.0f;
      inputs[LEN * (27 + i) + (end_y - 1) * MS + x] = 1.0f;
    }
  }

  /* :587-593 ko point: written WITHOUT offset_x / offset_y (reference behaviour) */
  if (st->ko_x != MARU_NONE && st->ko_y != MARU_NONE) {
    inputs[LEN * 31 + st->ko_y * MS + st->ko_x] = 1.0f;
  }

  /* :596-622 scalars */
  {
    const int info = (MARU_MODEL_FEATURES + 1) * LEN;
    if (color == MARU_BLACK) inputs[info + 0] = 1.0f;
    else inputs[info + 1] = 1.0f;
    inputs[info + 2] = (float)((double)(st->komi * (float)color) / 13.0); /* :605 */
    if (st->superko) inputs[info + 3] = 1.0f;
    if (st->ko_x != MARU_NONE && st->ko_y != MARU_NONE) inputs[info + 4] = 1.0f;
    if (st->rule != MARU_RULE_JP) inputs[info + 5] = 1.0f;
    else inputs[info + 6] = 1.0f;
  }
}

void oracle_get_inputs_batch(const maru_state* st, float* inputs, int n) {
  int i;
  for (i = 0; i < n; i++) oracle_get_inputs(st + i, inputs + (size_t)i * MARU_INPUT_SIZE);
}

/* Evaluator.cpp:55-80 post-processing of one output row given the record's move mask */
int oracle_policies(const maru_state* st, const float* outputs, int* xs, int* ys, float* ps,
                    float* value) {
  const int w = st->width, h = st->height;
  const int offset_x = (MS - w) / 2, offset_y = (MS - h) / 2;
  int n = 0, x, y;
  for (y = 0; y < h; y++) {
    for (x = 0; x < w; x++) {
      const int bi = y * w + x;
      if (st->move_mask[bi >> 3] & (1u << (bi & 7))) {
        xs[n] = x;
        ys[n] = y;
        ps[n] = outputs[(offset_y + y) * MS + (offset_x + x)];
        n++;
      }
    }
  }
  *value = outputs[MARU_MODEL_PREDICTIONS * LEN + 0] * 2 - 1; /* :75 */
  if (st->color == MARU_WHITE) *value = -*value;              /* :78-80 */
  return n;
}
