// state_from_board.h — fill a compact maru_state from a reference-style Board, using only the
// Board's PUBLIC getters, so the reference's Board.{h,cpp} stays byte-identical.
//
// Getters used (reference src/deepgo/native/cpp/Board.h):
//   getWidth/getHeight :47-53   getColor(x,y) :85   getRenSpace(x,y) :108   isShicho(x,y) :116
//   getHistories(color) :77  (oldest -> newest, empty slots dropped, Board.cpp:226-240)
//   getKo(color) :70          ((-1,-1) unless color == _koColor, Board.cpp:213-219)
//   getEnableds / getTerritories (optional move mask, Evaluator.cpp:60-68)
//
// Everything Board::getInputs (Board.cpp:494-623) reads is captured:
//   stone colour, min(liberties,8), ladder flag per cell; 3 newest non-pass moves per colour;
//   ko point for the side to move; komi, rule, superko.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

#include "lean_rules.h"
#include "maru_b200.h"

namespace maru_b200 {

// lean_ladder: take the ladder flags from the flat-board read-out of lean_ladder.h (bit-identical to
// Board::isShicho, 20-45x cheaper: the reference copies a whole std::set-based Board per search step)
// instead of calling Board::isShicho, which triggers Board::_updateShicho (Board.cpp:303-313, 1045-1067),
// and the move mask from lean_rules.h instead of Board::getEnableds(checkSeki) / getTerritories
// (identical on every cell, tests/test_lean_rules.py).
template <class BoardT>
inline void state_from_board(BoardT* board, int32_t color, float komi, int32_t rule, bool superko,
                             bool with_move_mask, maru_state* st, bool lean_ladder = false) {
  std::memset(st, 0, sizeof(maru_state));
  const int32_t w = board->getWidth();
  const int32_t h = board->getHeight();
  st->width = (uint8_t)w;
  st->height = (uint8_t)h;
  st->color = (int8_t)color;
  st->rule = (uint8_t)rule;
  st->superko = superko ? 1 : 0;
  st->komi = komi;

  uint8_t ladder[361];
  LeanBoard lb;
  if (lean_ladder) {
    lb.init(w, h);
    for (int32_t y = 0; y < h; y++)
      for (int32_t x = 0; x < w; x++) {
        const int32_t c = board->getColor(x, y);
        lb.c[lb.index(x, y)] = (c == MARU_BLACK) ? LeanBoard::kBlack : (c == MARU_WHITE) ? LeanBoard::kWhite
                                                                                      : LeanBoard::kEmpty;
      }
    for (int32_t kc : {MARU_BLACK, MARU_WHITE}) {     // Board::_koIndex / _koColor through the public getter
      auto k = board->getKo(kc);
      if (k.first >= 0 && k.second >= 0) {
        lb.ko_index = lb.index(k.first, k.second);
        lb.ko_color = kc;
      }
    }
    lb.ladder_flags(ladder);
  }

  for (int32_t y = 0; y < h; y++) {
    for (int32_t x = 0; x < w; x++) {
      const int32_t c = board->getColor(x, y);
      uint8_t cell = MARU_CELL_EMPTY;
      if (c != 0) {
        int32_t libs = board->getRenSpace(x, y);
        if (libs > 8) libs = 8;
        cell = (c == MARU_BLACK) ? MARU_CELL_BLACK : MARU_CELL_WHITE;
        cell |= (uint8_t)(libs << MARU_CELL_LIB_SHIFT);
        if (lean_ladder ? (ladder[y * w + x] != 0) : board->isShicho(x, y)) cell |= MARU_CELL_LADDER;
      }
      st->cells[y * w + x] = cell;
    }
  }

  // history: reference returns oldest -> newest with empty slots dropped; planes want newest first
  std::memset(st->hist, MARU_NONE, sizeof(st->hist));
  for (int32_t side = 0; side < 2; side++) {
    auto moves = board->getHistories(side == 0 ? MARU_BLACK : MARU_WHITE);
    int32_t k = 0;
    for (auto it = moves.rbegin(); it != moves.rend() && k < 3; ++it, ++k) {
      st->hist[side][k][0] = (uint8_t)it->first;
      st->hist[side][k][1] = (uint8_t)it->second;
    }
  }

  auto ko = board->getKo(color);
  if (ko.first >= 0 && ko.second >= 0) {
    st->ko_x = (uint8_t)ko.first;
    st->ko_y = (uint8_t)ko.second;
  } else {
    st->ko_x = MARU_NONE;
    st->ko_y = MARU_NONE;
  }

  if (with_move_mask && lean_ladder) {
    LeanRules rules(lb);
    rules.move_mask(color, st->move_mask);
    st->flags |= MARU_STATE_HAS_MASK;
  } else if (with_move_mask) {
    std::unique_ptr<int32_t[]> enableds(new int32_t[w * h]);
    std::unique_ptr<int32_t[]> territories(new int32_t[w * h]);
    board->getEnableds(enableds.get(), color, true);
    board->getTerritories(territories.get(), color);
    for (int32_t i = 0; i < w * h; i++) {
      if (enableds[i] == 1 && territories[i] == 0) st->move_mask[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    st->flags |= MARU_STATE_HAS_MASK;
  }
}

}  // namespace maru_b200
