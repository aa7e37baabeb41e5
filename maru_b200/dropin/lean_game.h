// lean_game.h — a complete position for the leaf path on the flat board: stones, ko, move history, and
// the compact record (maru_state) built WITHOUT any reference Board (end state).
//
// LeanGame::play follows Board::play (reference src/deepgo/native/cpp/Board.cpp:140-205): an off-board
// move is a pass (ko cleared, nothing enters the history, Board.cpp:142-146), an illegal move returns -1
// and changes nothing, a legal one is appended to the mover's 3-entry history ring (History.cpp:36-58).
// LeanGame::to_state produces byte for byte the record maru_b200::state_from_board() builds from a
// reference Board through its public getters (tests/test_lean_game.py replays random games move by move
// on both and compares all 448 bytes at every ply, move mask included).
// One copy is ~0.9 KB of plain data (the reference copies ~440 Ren objects with three std::set each).
//
// Liberties are kept INCREMENTALLY: `info[y*w+x]` = colour | min(liberties, 8) << 3 of the stone's group (the record's
// cell byte without the ladder bit) is updated by play() for the groups the move can have changed — the played stone's
// group, the neighbouring opponent groups, the groups around captured stones — instead of re-labelling the whole board
// for every evaluated leaf (which was the largest part of the host time per visit once the move filter had become lazy).
// The ladder flags cannot be kept: a ladder depends on stones far away, so every group in atari is read out again.
#pragma once

#include "lean_rules.h"
#include "maru_b200.h"

namespace maru_b200 {

struct LeanGame {
  LeanBoard board;
  int16_t hist[2][3];          // [0] black, [1] white: ring of board indices, -1 = empty slot
  uint8_t hist_pos[2];         // next slot to overwrite (History::_index)
  uint8_t info_valid;          // info[] describes the board (0 after `board` was edited behind play()'s back)
  uint8_t info[361];           // per board cell: MARU_CELL_* colour | min(#liberties, 8) << MARU_CELL_LIB_SHIFT

  void init(int width, int height) {
    board.init(width, height);
    std::memset(info, 0, sizeof info);
    info_valid = 1;
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
    int16_t captured[361];
    const int r = board.play(idx, color, captured);
    if (r < 0) return -1;
    if (info_valid) refresh_info(idx, captured, r);
    const int s = (color == LeanBoard::kBlack) ? 0 : 1;
    hist[s][hist_pos[s]] = (int16_t)idx;
    hist_pos[s] = (uint8_t)((hist_pos[s] + 1) % 3);
    return r;
  }

  // After a legal move at `idx` that emptied `captured[0..n_cap)`: liberties can only have changed for the group of the
  // played stone (which absorbed its friendly neighbours), for the opponent groups next to it (one liberty less) and for
  // every group next to a captured stone (liberties gained).  Each such group is flooded once.
  void refresh_info(int idx, const int16_t* captured, int n_cap) {
    const int w = board.w, W = board.W;
    const int d[4] = {-1, -W, 1, W};
    uint8_t seen[LeanBoard::MAXN];
    std::memset(seen, 0, (size_t)board.N);
    LeanBoard::Group g;
    auto relabel = [&](int start) {
      const int8_t c = board.c[start];
      if ((c != LeanBoard::kBlack && c != LeanBoard::kWhite) || seen[start]) return;
      board.group(start, g);
      const int libs = g.n_libs > 8 ? 8 : g.n_libs;
      const uint8_t cell = (uint8_t)(((c == LeanBoard::kBlack) ? MARU_CELL_BLACK : MARU_CELL_WHITE) | (libs << MARU_CELL_LIB_SHIFT));
      for (int k = 0; k < g.n_stones; k++) {
        const int p = g.stones[k];
        seen[p] = 1;
        info[(p / W - 1) * w + (p % W - 1)] = cell;
      }
    };
    for (int i = 0; i < n_cap; i++) info[(captured[i] / W - 1) * w + (captured[i] % W - 1)] = MARU_CELL_EMPTY;
    relabel(idx);
    for (int k = 0; k < 4; k++) relabel(idx + d[k]);
    for (int i = 0; i < n_cap; i++)
      for (int k = 0; k < 4; k++) relabel(captured[i] + d[k]);
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

    if (!info_valid) {                                  // (only after direct edits of `board`): label everything once
      LeanRules rules(board);
      LeanGame* self = const_cast<LeanGame*>(this);
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          const int idx = board.index(x, y);
          const int c = board.c[idx];
          uint8_t cell = MARU_CELL_EMPTY;
          if (c == LeanBoard::kBlack || c == LeanBoard::kWhite) {
            int libs = rules.n_libs[rules.gid[idx]];
            if (libs > 8) libs = 8;
            cell = (uint8_t)(((c == LeanBoard::kBlack) ? MARU_CELL_BLACK : MARU_CELL_WHITE) | (libs << MARU_CELL_LIB_SHIFT));
          }
          self->info[y * w + x] = cell;
        }
      self->info_valid = 1;
    }
    std::memcpy(st->cells, info, (size_t)(w * h));
    // Board::_updateShicho: only groups in atari are read out (Board.cpp:1045-1067)
    {
      uint8_t done[361];
      bool any = false;
      for (int i = 0; i < w * h; i++)
        if ((info[i] >> MARU_CELL_LIB_SHIFT) == 1) {
          if (!any) {
            std::memset(done, 0, (size_t)(w * h));
            any = true;
          }
          if (done[i]) continue;
          const int idx = board.index(i % w, i / w);
          LeanBoard::Group g;
          board.group(idx, g);
          const bool caught = board.ladder(idx);
          for (int k = 0; k < g.n_stones; k++) {
            const int p = g.stones[k], cell = (p / board.W - 1) * w + (p % board.W - 1);
            done[cell] = 1;
            if (caught) st->cells[cell] |= MARU_CELL_LADDER;
          }
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
      LeanRules rules(board);
      rules.move_mask(color, st->move_mask);
      st->flags |= MARU_STATE_HAS_MASK;
    }
  }
};

}  // namespace maru_b200
