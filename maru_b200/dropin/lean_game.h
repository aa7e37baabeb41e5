// lean_game.h — a complete position for the leaf path on the flat board: stones, ko, move history, and
// the compact record (maru_state) built WITHOUT any reference Board (end state).
//
// LeanGame::play follows Board::play (reference src/deepgo/native/cpp/Board.cpp:140-205): an off-board
// move is a pass (ko cleared, nothing enters the history, Board.cpp:142-146), an illegal move returns -1
// and changes nothing, a legal one is appended to the mover's 3-entry history ring (History.cpp:36-58).
// LeanGame::to_state produces byte for byte the record maru_b200::state_from_board() builds from a
// reference Board through its public getters (tests/test_lean_game.py replays random games move by move
// on both and compares all 448 bytes at every ply, move mask included).
// One copy is ~0.6 KB of plain data (the reference copies ~440 Ren objects with three std::set each).
#pragma once

#include "lean_rules.h"
#include "maru_b200.h"

namespace maru_b200 {

struct LeanGame {
  LeanBoard board;
  int16_t hist[2][3];          // [0] black, [1] white: ring of board indices, -1 = empty slot
  uint8_t hist_pos[2];         // next slot to overwrite (History::_index)

  void init(int width, int height) {
    board.init(width, height);
    for (int s = 0; s < 2; s++) {
      hist_pos[s] = 0;
      for (int i = 0; i < 3; i++) hist[s][i] = -1;
    }
  }

  // Board::play: captured stones, 0 for a pass, -1 if not allowed
  int play(int x, int y, int color) {
    if (x < 0 || y < 0 || x >= board.w || y >= board.h) {
      board.ko_index = -1;
      board.ko_color = 0;
      return 0;
    }
    const int idx = board.index(x, y);
    const int r = board.play(idx, color);
    if (r < 0) return -1;
    const int s = (color == LeanBoard::kBlack) ? 0 : 1;
    hist[s][hist_pos[s]] = (int16_t)idx;
    hist_pos[s] = (uint8_t)((hist_pos[s] + 1) % 3);
    return r;
  }

  // the record of Board::getInputs' inputs for `color` to move (include/maru_b200.h: maru_state)
  void to_state(int color, float komi, int rule, bool superko, bool with_move_mask, maru_state* st) const {
    std::memset(st, 0, sizeof(maru_state));
    const int w = board.w, h = board.h;
    st->width = (uint8_t)w;
    st->height = (uint8_t)h;
    st->color = (int8_t)color;
    st->rule = (uint8_t)rule;
    st->superko = superko ? 1 : 0;
    st->komi = komi;

    LeanRules rules(board);                       // group labelling: liberties per stone
    // Board::_updateShicho: only groups in atari are read out (Board.cpp:1045-1067); the labelling above
    // already knows them, so the board is not flood-filled a second time (LeanBoard::ladder_flags does)
    uint8_t ladder[361];
    std::memset(ladder, 0, (size_t)(w * h));
    for (int idx = 0; idx < board.N; idx++) {
      if (rules.gid[idx] != idx || rules.n_libs[idx] != 1) continue;
      if (!board.ladder(idx)) continue;
      for (int i = 0; i < rules.n_stones[idx]; i++) {
        const int p = rules.stone_pool[rules.stone_off[idx] + i];
        ladder[(p / board.W - 1) * w + (p % board.W - 1)] = 1;
      }
    }
    for (int y = 0; y < h; y++) {
      for (int x = 0; x < w; x++) {
        const int idx = board.index(x, y);
        const int c = board.c[idx];
        uint8_t cell = MARU_CELL_EMPTY;
        if (c == LeanBoard::kBlack || c == LeanBoard::kWhite) {
          int libs = rules.n_libs[rules.gid[idx]];
          if (libs > 8) libs = 8;
          cell = (c == LeanBoard::kBlack) ? MARU_CELL_BLACK : MARU_CELL_WHITE;
          cell |= (uint8_t)(libs << MARU_CELL_LIB_SHIFT);
          if (ladder[y * w + x]) cell |= MARU_CELL_LADDER;
        }
        st->cells[y * w + x] = cell;
      }
    }

    // Board::getHistories: ring read oldest -> newest, empty slots dropped; the record wants newest first
    std::memset(st->hist, MARU_NONE, sizeof(st->hist));
    for (int s = 0; s < 2; s++) {
      int k = 0;
      for (int i = 2; i >= 0 && k < 3; i--) {
        const int idx = hist[s][(hist_pos[s] + i) % 3];
        if (idx < 0) continue;
        st->hist[s][k][0] = (uint8_t)(idx % board.W - 1);
        st->hist[s][k][1] = (uint8_t)(idx / board.W - 1);
        k++;
      }
    }

    // Board::getKo(color): only the side that may not retake sees the point (Board.cpp:213-219)
    if (board.ko_index != -1 && color == board.ko_color) {
      st->ko_x = (uint8_t)(board.ko_index % board.W - 1);
      st->ko_y = (uint8_t)(board.ko_index / board.W - 1);
    } else {
      st->ko_x = MARU_NONE;
      st->ko_y = MARU_NONE;
    }

    if (with_move_mask) {
      rules.move_mask(color, st->move_mask);
      st->flags |= MARU_STATE_HAS_MASK;
    }
  }
};

}  // namespace maru_b200
