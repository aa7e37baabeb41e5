// lean_ladder.h — flat-array Go board with the reference's exact ladder ("shicho") read-out
// (first half: the part of the host leaf path that dominates its cost).
//
// Board::getInputs marks stones whose group can be captured in a ladder (planes 2 and 15, reference
// src/deepgo/native/cpp/Board.cpp:523-536).  The reference decides that per group with a depth-first
// search that COPIES a whole Board — two std::set per group, heap arrays — at every step
// (Board::_updateShicho / _isShichoRen, Board.cpp:1045-1153): 0.25 ms (9x9) to 0.9 ms (19x19) per
// leaf on mid-game positions, ~3/4 of the host time the search spends per evaluated node, and every
// expanded node is a new position, so the result is never cached.
//
// This header restates the same decision procedure on a plain-old-data board (one byte per cell,
// groups and liberties found by flood fill, ~0.5 KB per copy), so that the substitute Evaluator.cpp
// can fill maru_state::cells without calling Board::isShicho.  Bit-identical results are a hard
// requirement (the planes must stay bit-exact): every rule below follows the reference line by line,
// including its quirks —
//   * play() legality = Board::_isEnabled(index, color, checkSeki=false)      Board.cpp:1177-1220
//   * captures, ko bookkeeping exactly as Board::play                         Board.cpp:140-205
//   * the read-out order: liberties in ascending index order, LIFO stack, "return true" as soon as
//     ANY branch ends in a capture, an illegal escape or a 1-liberty escape   Board.cpp:1074-1153
// tests/test_lean_ladder.py checks it against the compiled reference on thousands of random positions.
#pragma once

#include <cstdint>
#include <cstring>
#include <vector>

namespace maru_b200 {

struct LeanBoard {
  static constexpr int MAXW = 21;                 // 19 + border
  static constexpr int MAXN = MAXW * MAXW;
  static constexpr int8_t kEmpty = 0, kBlack = 1, kWhite = -1, kEdge = 2;   // (the reference's Config.h #defines EMPTY/BLACK/WHITE/EDGE)

  int32_t w = 0, h = 0, W = 0, N = 0;             // W = w + 2 (row pitch with border), N = W * (h + 2)
  int32_t ko_index = -1;
  int32_t ko_color = 0;
  int8_t c[MAXN];

  void init(int width, int height) {
    w = width;
    h = height;
    W = width + 2;
    N = W * (height + 2);
    ko_index = -1;
    ko_color = 0;
    for (int i = 0; i < N; i++) c[i] = kEdge;
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) c[index(x, y)] = kEmpty;
  }
  int index(int x, int y) const { return (y + 1) * W + (x + 1); }   // Board.cpp:1649-1651

  // stones and liberties of the group containing `start` (liberties in ASCENDING index order, the
  // iteration order of the reference's std::set<int32_t> spaces)
  struct Group {
    int16_t stones[361];
    int16_t libs[361];
    int n_stones = 0, n_libs = 0;
  };
  void group(int start, Group& g) const {
    uint8_t seen[MAXN];
    std::memset(seen, 0, (size_t)N);
    const int8_t col = c[start];
    const int d[4] = {-1, -W, 1, W};
    g.n_stones = g.n_libs = 0;
    int16_t stack[361];
    int sp = 0;
    stack[sp++] = (int16_t)start;
    seen[start] = 1;
    while (sp > 0) {
      const int p = stack[--sp];
      g.stones[g.n_stones++] = (int16_t)p;
      for (int k = 0; k < 4; k++) {
        const int q = p + d[k];
        if (seen[q]) continue;
        if (c[q] == kEmpty) {
          seen[q] = 1;
          g.libs[g.n_libs++] = (int16_t)q;
        } else if (c[q] == col) {
          seen[q] = 1;
          stack[sp++] = (int16_t)q;
        }
      }
    }
    // insertion sort: groups in atari fights have a handful of liberties
    for (int i = 1; i < g.n_libs; i++) {
      const int16_t v = g.libs[i];
      int j = i - 1;
      while (j >= 0 && g.libs[j] > v) {
        g.libs[j + 1] = g.libs[j];
        j--;
      }
      g.libs[j + 1] = v;
    }
  }
  int liberties(int start) const {
    Group g;
    group(start, g);
    return g.n_libs;
  }
  // min(#liberties, limit) of the group containing `start`: the flood stops at the limit-th liberty (every caller
  // below only asks "exactly one?" or "more than one?", and the groups next to a ladder are often large)
  int liberties_upto(int start, int limit) const {
    uint8_t seen[MAXN];
    std::memset(seen, 0, (size_t)N);
    const int8_t col = c[start];
    const int d[4] = {-1, -W, 1, W};
    int16_t stack[361];
    int sp = 0, libs = 0;
    stack[sp++] = (int16_t)start;
    seen[start] = 1;
    while (sp > 0) {
      const int p = stack[--sp];
      for (int k = 0; k < 4; k++) {
        const int q = p + d[k];
        if (seen[q]) continue;
        if (c[q] == kEmpty) {
          seen[q] = 1;
          if (++libs >= limit) return libs;
        } else if (c[q] == col) {
          seen[q] = 1;
          stack[sp++] = (int16_t)q;
        }
      }
    }
    return libs;
  }

  // Board::_isEnabled(index, color, false)
  bool enabled(int idx, int color) const {
    if (c[idx] != kEmpty) return false;
    if (idx == ko_index && color == ko_color) return false;
    const int d[4] = {-1, -W, 1, W};
    for (int k = 0; k < 4; k++) {
      const int q = idx + d[k];
      if (c[q] == kEmpty) return true;
      if (c[q] == kEdge) continue;
      const int nl = liberties_upto(q, 2);
      if (c[q] == color && nl > 1) return true;
      if (c[q] == -color && nl == 1) return true;
    }
    return false;
  }

  // Board::play for an on-board point: number of captured stones, -1 if not allowed (board unchanged).
  // `captured` (optional, room for 361 cells) receives the cells that were emptied.
  int play(int idx, int color, int16_t* captured = nullptr) {
    // Legality (== enabled(idx, color): an empty neighbour, an own neighbour group with more than one liberty, or an
    // opponent neighbour group with exactly one) and the capture test share their floods: an opponent group is captured
    // by this stone exactly when it has ONE liberty before the stone is placed (that liberty is idx).
    if (c[idx] != kEmpty) return -1;
    if (idx == ko_index && color == ko_color) return -1;
    const int d[4] = {-1, -W, 1, W};
    bool ok = false;
    bool atari[4] = {false, false, false, false};
    for (int k = 0; k < 4; k++) {
      const int q = idx + d[k];
      if (c[q] == kEmpty) ok = true;
      else if (c[q] == -color && liberties_upto(q, 2) == 1) ok = atari[k] = true;
    }
    for (int k = 0; k < 4 && !ok; k++) {
      const int q = idx + d[k];
      if (c[q] == color && liberties_upto(q, 2) > 1) ok = true;
    }
    if (!ok) return -1;
    c[idx] = (int8_t)color;
    int removed = 0;
    Group g;
    for (int k = 0; k < 4; k++) {
      const int q = idx + d[k];
      if (!atari[k] || c[q] != -color) continue;      // (already taken through another neighbour: same group)
      group(q, g);
      for (int i = 0; i < g.n_stones; i++) {
        c[g.stones[i]] = kEmpty;
        if (captured != nullptr) captured[removed + i] = g.stones[i];
      }
      removed += g.n_stones;
      ko_index = q;
    }
    // a ko needs exactly one captured stone AND a lone capturing stone with one liberty: the flood of the played stone's
    // group is only made when a single stone was taken (most moves, and most steps of a ladder read-out, capture nothing)
    bool ko = false;
    if (removed == 1) {
      group(idx, g);
      ko = g.n_stones == 1 && g.n_libs == 1;
    }
    if (!ko) {
      ko_index = -1;
      ko_color = 0;
    } else {
      ko_color = -color;
    }
    return removed;
  }

  // Board::_isShichoRen for the group containing `index`
  bool ladder(int index) const {
    Group g;
    group(index, g);
    if (g.n_libs > 1) return false;
    // (one scratch stack per thread: the read-out runs for every group in atari of every evaluated leaf, and a heap
    // allocation per call was a visible share of the host time per leaf)
    static thread_local std::vector<LeanBoard> stack;
    stack.clear();
    stack.push_back(*this);
    const int d[4] = {-1, -W, 1, W};
    while (!stack.empty()) {
      // (one copy per searched board: the board popped here is also the one the escaping side plays on)
      LeanBoard curr = stack.back();
      stack.pop_back();
      const LeanBoard& board = curr;
      const int color = board.c[index];
      const int op = -color;
      board.group(index, g);
      // an adjacent opponent group in atari can be captured: this branch escapes
      bool escaped = false;
      for (int i = 0; i < g.n_stones && !escaped; i++) {
        for (int k = 0; k < 4; k++) {
          const int q = g.stones[i] + d[k];
          if (board.c[q] == op && board.liberties_upto(q, 2) == 1) {
            escaped = true;
            break;
          }
        }
      }
      if (escaped) continue;
      // the escaping side extends at its (lowest-index) liberty
      const int curr_pos = g.libs[0];
      if (curr.play(curr_pos, color) < 0) return true;
      Group e;
      curr.group(index, e);
      if (e.n_libs == 1) return true;
      if (e.n_libs > 2) continue;
      // the chasing side tries both liberties (an illegal try leaves the board unchanged and is still searched)
      for (int i = 0; i < e.n_libs; i++) {
        stack.push_back(curr);
        stack.back().play(e.libs[i], op);
      }
    }
    return false;
  }

  // flags[y*w + x] = 1 for every stone whose group is ladder-capturable (Board::_updateShicho)
  void ladder_flags(uint8_t* flags) const {
    std::memset(flags, 0, (size_t)(w * h));
    uint8_t done[MAXN];
    std::memset(done, 0, (size_t)N);
    Group g;
    for (int y = 0; y < h; y++) {
      for (int x = 0; x < w; x++) {
        const int idx = index(x, y);
        if (done[idx] || (c[idx] != kBlack && c[idx] != kWhite)) continue;
        group(idx, g);
        for (int i = 0; i < g.n_stones; i++) done[g.stones[i]] = 1;
        if (g.n_libs != 1) continue;
        Group keep = g;                               // ladder() reuses its own scratch groups
        if (ladder(idx)) {
          for (int i = 0; i < keep.n_stones; i++) {
            const int p = keep.stones[i];
            flags[(p / W - 1) * w + (p % W - 1)] = 1;
          }
        }
      }
    }
  }
};

}  // namespace maru_b200
