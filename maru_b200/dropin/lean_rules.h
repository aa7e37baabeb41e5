// lean_rules.h — the reference's candidate-move filter on a flat board.
//
// Evaluator::evaluate keeps a move only if it is legal under the reference's seki-aware legality test
// AND does not lie in territory that is already unconditionally decided (reference
// src/deepgo/native/cpp/Evaluator.cpp:60-68):
//     Board::getEnableds(enableds, color, checkSeki=true)     Board.cpp:333-343, 1177-1220
//       -> Board::_isSeki / _isSekiRen / _isSekiArea           Board.cpp:1228-1465
//       -> Board::_isNakade / _isSingleArea                    Board.cpp:1480-1631
//     Board::getTerritories(territories, color)                Board.cpp:350-378
//       -> Board::_updateArea                                  Board.cpp:872-1040
// On the reference's Board (three std::set per group, Ren objects copied BY VALUE inside _isEnabled)
// these two calls cost 0.1 ms (9x9) to 0.4 ms (19x19) per leaf — what is left of the host time once the
// ladder read-out is lean.  This header restates them over LeanBoard (lean_ladder.h): groups are labelled
// once per position, sets are small sorted arrays.  The result must be IDENTICAL on every cell (it decides
// which moves the search may expand), so every quirk is kept: the 9-cell cut-offs, the exact nakade
// vital-point count with its corner rule (as a lookup over 4x4 window masks), "an area is decided only if EVERY cell of it touches exactly the
// own groups its first cell touches", and the iteration that un-decides groups with fewer than two
// decided areas.  tests/test_lean_rules.py compares against the compiled reference.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>

#include "lean_ladder.h"

namespace maru_b200 {

// sorted unique small set of cell / group indices (the reference uses std::set<int32_t>).  Every set these
// rules build is small by construction: the reference itself gives up at 9 cells / 7 stones, opponent groups
// that matter have exactly 2 liberties, and own groups with >= 9 liberties / >= 6 stones are cut off before
// their cells are inserted.  The largest reachable size is 4 + 4*8 = 36 (liberties of four groups).
struct IdxSet {
  static constexpr int CAP = 96;
  int n = 0;
  int16_t v[CAP];
  IdxSet() = default;
  IdxSet(const IdxSet& o) : n(o.n) { std::memcpy(v, o.v, (size_t)o.n * sizeof(int16_t)); }   // only the live part
  IdxSet& operator=(const IdxSet& o) {
    n = o.n;
    std::memcpy(v, o.v, (size_t)o.n * sizeof(int16_t));
    return *this;
  }
  void clear() { n = 0; }
  bool has(int x) const {
    for (int i = 0; i < n; i++)
      if (v[i] >= x) return v[i] == x;
    return false;
  }
  bool insert(int x) {
    int i = 0;
    while (i < n && v[i] < x) i++;
    if (i < n && v[i] == x) return false;
    if (n >= CAP) return false;                    // unreachable by the bounds above; never write past the array
    for (int j = n; j > i; j--) v[j] = v[j - 1];
    v[i] = (int16_t)x;
    n++;
    return true;
  }
  void erase(int x) {
    for (int i = 0; i < n; i++) {
      if (v[i] == x) {
        for (int j = i; j + 1 < n; j++) v[j] = v[j + 1];
        n--;
        return;
      }
    }
  }
  bool operator==(const IdxSet& o) const { return n == o.n && std::memcmp(v, o.v, (size_t)n * sizeof(int16_t)) == 0; }
  bool operator!=(const IdxSet& o) const { return !(*this == o); }
};

// the (at most four) own groups around one cell, sorted
struct Ids4 {
  int16_t v[4];
  int n = 0;
  void insert(int x) {
    int i = 0;
    while (i < n && v[i] < x) i++;
    if (i < n && v[i] == x) return;
    for (int j = n; j > i; j--) v[j] = v[j - 1];
    v[i] = (int16_t)x;
    n++;
  }
  bool operator!=(const Ids4& o) const {
    if (n != o.n) return true;
    for (int i = 0; i < n; i++)
      if (v[i] != o.v[i]) return true;
    return false;
  }
};

struct LeanRules {
  const LeanBoard& b;
  int d[4];
  // group labelling: gid[cell] = smallest cell index of the stone's group, -1 for empty, -2 for the border
  int16_t gid[LeanBoard::MAXN];
  int16_t n_libs[LeanBoard::MAXN];       // per group id
  int16_t n_stones[LeanBoard::MAXN];
  int32_t lib_off[LeanBoard::MAXN];      // per group id: offset of its liberty list in lib_pool (unordered)
  int32_t stone_off[LeanBoard::MAXN];
  int16_t lib_pool[4 * 361 + 8];
  int16_t stone_pool[361 + 8];
  // Connected components of the cells a colour does NOT hold (empty or opponent), per colour (0 black, 1 white): the
  // "areas" of Board::_updateArea and the regions Board::_isSekiArea floods.  Built on first use (one flood per colour),
  // shared by territories() and every is_seki_area() call on this position.
  mutable bool comp_built[2] = {false, false};
  mutable int16_t comp_id[2][LeanBoard::MAXN];     // id = the component's first cell in scan order; -1 for own stones / border
  mutable int16_t comp_size[2][LeanBoard::MAXN];   // per component id
  mutable int32_t comp_off[2][LeanBoard::MAXN];    // per component id: offset of its cell list in comp_pool
  mutable int16_t comp_pool[2][LeanBoard::MAXN];

  explicit LeanRules(const LeanBoard& board) : b(board) {
    d[0] = -1;
    d[1] = -b.W;
    d[2] = 1;
    d[3] = b.W;
    // One pass over the board: cells are scanned in ascending order, so the first unlabelled stone of a group
    // is its smallest cell = its id.  lib_mark[q] == id says "q is already counted as a liberty of this group".
    int16_t lib_mark[LeanBoard::MAXN];
    std::memset(lib_mark, 0xFF, (size_t)b.N * sizeof(int16_t));                    // -1
    std::memset(gid, 0xFF, (size_t)b.N * sizeof(int16_t));                         // -1: empty (stones are labelled below)
    {                                                                              // -2: the border ring
      const int W = b.W, H = b.h + 2;
      for (int x = 0; x < W; x++) gid[x] = gid[(H - 1) * W + x] = -2;
      for (int y = 1; y < H - 1; y++) gid[y * W] = gid[y * W + W - 1] = -2;
    }
    int lp = 0, sp = 0;
    int16_t stack[361];
    for (int i = 0; i < b.N; i++) {
      const int8_t col = b.c[i];
      if ((col != LeanBoard::kBlack && col != LeanBoard::kWhite) || gid[i] >= 0) continue;
      const int id = i;
      lib_off[id] = lp;
      stone_off[id] = sp;
      int top = 0;
      stack[top++] = (int16_t)i;
      gid[i] = (int16_t)id;
      while (top > 0) {
        const int p = stack[--top];
        stone_pool[sp++] = (int16_t)p;
        for (int k = 0; k < 4; k++) {
          const int q = p + d[k];
          const int8_t cq = b.c[q];
          if (cq == LeanBoard::kEmpty) {
            if (lib_mark[q] != id) {
              lib_mark[q] = (int16_t)id;
              lib_pool[lp++] = (int16_t)q;
            }
          } else if (cq == col && gid[q] < 0) {
            gid[q] = (int16_t)id;
            stack[top++] = (int16_t)q;
          }
        }
      }
      n_libs[id] = (int16_t)(lp - lib_off[id]);
      n_stones[id] = (int16_t)(sp - stone_off[id]);
    }
  }
  int color_of(int cell) const { return b.c[cell]; }     // kEmpty / kBlack / kWhite / kEdge

  // Two raster passes with a union-find whose roots are always the smallest cell of their set (= the reference's area
  // id, the first cell of the area in scan order), then a counting sort of the cells by component: predictable loads
  // instead of a stack flood (3x faster on mid-game boards; the flood was a third of move_mask()).
  void build_components(int c) const {
    if (comp_built[c]) return;
    comp_built[c] = true;
    const int op = (c == 0) ? LeanBoard::kWhite : LeanBoard::kBlack;
    const int8_t* const cc = b.c;
    const int N = b.N, W = b.W;
    int16_t* const id = comp_id[c];
    int16_t* const size = comp_size[c];
    const auto find = [&](int x) {
      while (id[x] != x) {
        id[x] = id[id[x]];                                      // path halving
        x = id[x];
      }
      return x;
    };
    // the border ring (rows 0 and h + 1, columns 0 and W - 1) is kEdge: never allowed, so i - 1 and i - W stay in range
    for (int i = 0; i < N; i++) {
      const int ic = cc[i];
      if (ic != LeanBoard::kEmpty && ic != op) {
        id[i] = -1;
        continue;
      }
      const int l = (id[i - 1] >= 0) ? find(i - 1) : -1;
      const int u = (id[i - W] >= 0) ? find(i - W) : -1;
      if (l < 0 && u < 0) {
        id[i] = (int16_t)i;
      } else if (l < 0 || u < 0 || l == u) {
        id[i] = (int16_t)((l < 0) ? u : l);
      } else {
        const int lo = (l < u) ? l : u, hi = (l < u) ? u : l;
        id[hi] = (int16_t)lo;
        id[i] = (int16_t)lo;
      }
    }
    for (int i = 0; i < N; i++) {
      if (id[i] < 0) continue;
      const int r = find(i);
      id[i] = (int16_t)r;
      if (r == i) size[r] = 0;                                  // (a root is the first cell of its set in this order)
      size[r]++;
    }
    int16_t cursor[LeanBoard::MAXN];
    int pp = 0;
    for (int i = 0; i < N; i++) {
      if (id[i] != i) continue;
      comp_off[c][i] = pp;
      cursor[i] = (int16_t)pp;
      pp += size[i];
    }
    int16_t* const pool = comp_pool[c];
    for (int i = 0; i < N; i++)
      if (id[i] >= 0) pool[cursor[id[i]]++] = (int16_t)i;
  }

  // ---------------------------------------------------------------- Board::_isNakade (Board.cpp:1480-1591)
  // "Is there a vital point": the shape (<= 6 cells inside a 4x4 window) is a 16-bit mask and the answer for
  // shapes that touch no board corner comes from a table over all 65 536 masks (NakadeTable below, filled
  // once by the same rule); shapes with a corner cell are evaluated directly.
  bool is_nakade(const IdxSet& positions) const {
    if (positions.n == 0 || positions.n >= 7) return false;
    int x0 = b.w, y0 = b.h, x1 = 0, y1 = 0;
    for (int k = 0; k < positions.n; k++) {
      const int x = positions.v[k] % b.W - 1, y = positions.v[k] / b.W - 1;
      x0 = std::min(x0, x);
      y0 = std::min(y0, y);
      x1 = std::max(x1, x);
      y1 = std::max(y1, y);
    }
    if (x1 - x0 > 3 || y1 - y0 > 3) return false;
    uint32_t shape = 0, corners = 0;
    for (int k = 0; k < positions.n; k++) {
      const int x = positions.v[k] % b.W - 1, y = positions.v[k] / b.W - 1;
      const uint32_t bit = 1u << ((y - y0) * 4 + (x - x0));
      shape |= bit;
      if ((x == 0 || x == b.w - 1) && (y == 0 || y == b.h - 1)) corners |= bit;
    }
    return corners == 0 ? NakadeTable::get().test(shape) : nakade_shape(shape, corners);
  }

  // The reference's vital-point rule on a window mask.  A cell of the window's upper-left 3x3 part is vital if
  // (orthogonal neighbours in the shape) + (one diagonal neighbour reached around a filled elbow, counted
  // once) + (one diagonal neighbour reached through a board-corner cell, counted once) >= size - 1.  The
  // diagonals are visited in the reference's order (down-right, down-left, up-right, up-left) because a
  // diagonal is booked as "corner" before it may be booked as "elbow" (Board.cpp:1557-1576).
  static bool nakade_shape(uint32_t shape, uint32_t corners) {
    const auto at = [](uint32_t m, int x, int y) -> bool {
      return x >= 0 && x < 4 && y >= 0 && y < 4 && ((m >> (y * 4 + x)) & 1u);
    };
    const int need = __builtin_popcount(shape) - 1;
    static const int diag[4][2] = {{1, 1}, {-1, 1}, {1, -1}, {-1, -1}};      // (dx, dy)
    for (int y = 0; y < 3; y++) {
      for (int x = 0; x < 3; x++) {
        if (!at(shape, x, y)) continue;
        int links = at(shape, x + 1, y) + at(shape, x - 1, y) + at(shape, x, y + 1) + at(shape, x, y - 1);
        bool elbow = false, via_corner = false;
        for (const auto& dg : diag) {
          const int dx = dg[0], dy = dg[1];
          if (!at(shape, x + dx, y + dy)) continue;
          if (!via_corner && at(corners, x, y + dy)) {
            via_corner = true;
          } else if (!via_corner && at(corners, x + dx, y)) {
            via_corner = true;
          } else if (!elbow && at(shape, x, y + dy) && at(shape, x + dx, y)) {
            elbow = true;
          }
        }
        if (links + (elbow ? 1 : 0) + (via_corner ? 1 : 0) >= need) return true;
      }
    }
    return false;
  }

  struct NakadeTable {
    uint64_t bits[1024];
    NakadeTable() {
      std::memset(bits, 0, sizeof bits);
      for (uint32_t m = 1; m < 65536u; m++)
        if (__builtin_popcount(m) < 7 && nakade_shape(m, 0)) bits[m >> 6] |= 1ull << (m & 63);
    }
    bool test(uint32_t m) const { return (bits[m >> 6] >> (m & 63)) & 1ull; }
    static const NakadeTable& get() {
      static const NakadeTable t;
      return t;
    }
  };

  // ---------------------------------------------------------------- Board::_isSingleArea (Board.cpp:1599-1631)
  bool is_single_area(const IdxSet& positions, int color, int excluded) const {
    const int op = -color;
    IdxSet areas;
    int16_t stack[LeanBoard::MAXN];
    int sp = 0;
    stack[sp++] = positions.v[0];
    areas.insert(positions.v[0]);
    while (sp > 0) {
      const int pos = stack[--sp];
      for (int a : d) {
        const int t = pos + a;
        if (t < 0 || t >= b.N) continue;                       // (the flood never leaves the bordered board)
        if ((gid[t] == -1 || (gid[t] >= 0 && b.c[t] == op)) && t != excluded && !areas.has(t)) {
          stack[sp++] = (int16_t)t;
          areas.insert(t);
        }
      }
    }
    for (int k = 0; k < positions.n; k++)
      if (!areas.has(positions.v[k])) return false;
    return true;
  }

  // ---------------------------------------------------------------- Board::_isSekiRen (Board.cpp:1301-1376)
  bool is_seki_ren(int index, int color, const IdxSet& ren_ids, int space_index) const {
    const int op = -color;
    IdxSet op_ids;
    for (int a : d) {
      const int targets[2] = {index + a, space_index + a};
      for (int t : targets) {
        const int id = gid[t];
        if (t != index && t != space_index && id == -1) return false;
        if (id >= 0 && b.c[t] == op) op_ids.insert(id);
      }
    }
    if (op_ids.n == 0) return false;
    for (int k = 0; k < op_ids.n; k++)
      if (n_libs[op_ids.v[k]] != 2) return false;
    // (a group of >= 6 stones plus the new stone is always "7 or more": same early exit as below)
    for (int k = 0; k < ren_ids.n; k++)
      if (n_stones[ren_ids.v[k]] >= 6) return true;
    IdxSet positions;
    positions.insert(index);
    for (int k = 0; k < ren_ids.n; k++) {
      const int id = ren_ids.v[k];
      for (int i = 0; i < n_stones[id]; i++) positions.insert(stone_pool[stone_off[id] + i]);
      if (positions.n >= 7) return true;
    }
    if (positions.n >= 4 && !is_nakade(positions)) return true;
    IdxSet op_spaces;
    for (int k = 0; k < op_ids.n; k++) {
      const int id = op_ids.v[k];
      for (int i = 0; i < n_libs[id]; i++) op_spaces.insert(lib_pool[lib_off[id] + i]);
    }
    op_spaces.erase(index);
    op_spaces.erase(space_index);
    return op_spaces.n != 0;
  }

  // ---------------------------------------------------------------- Board::_isSekiArea (Board.cpp:1386-1465)
  bool is_seki_area(int index, int color, const IdxSet& ren_ids_in, const IdxSet& spaces) const {
    const int op = -color;
    {
      // The flood below collects the component(s) of empty-or-opponent cells that hold `index` and `spaces` and gives up
      // (false) as soon as it has 9 cells: with the components known that is a sum over at most 9 seeds — the common exit
      // (every open area) without touching the board.
      const int c = (color == LeanBoard::kBlack) ? 0 : 1;
      build_components(c);
      int16_t seen[9];
      int n_seen = 0, total = 0;
      const auto add = [&](int cell) {
        const int16_t id = comp_id[c][cell];
        for (int k = 0; k < n_seen; k++)
          if (seen[k] == id) return;
        seen[n_seen++] = id;
        total += comp_size[c][id];
      };
      add(index);
      for (int k = 0; k < spaces.n && total < 9; k++) add(spaces.v[k]);
      if (total >= 9) return false;
    }
    IdxSet positions, ren_ids;
    int16_t stack[LeanBoard::MAXN];
    int sp = 0;
    positions.insert(index);
    stack[sp++] = (int16_t)index;
    for (int k = 0; k < spaces.n; k++) {
      positions.insert(spaces.v[k]);
      stack[sp++] = spaces.v[k];
    }
    while (sp > 0) {
      const int pos = stack[--sp];
      for (int a : d) {
        const int t = pos + a;
        const int id = gid[t];
        if ((id == -1 || (id >= 0 && b.c[t] == op)) && !positions.has(t)) {
          stack[sp++] = (int16_t)t;
          positions.insert(t);
        }
        if (id >= 0 && b.c[t] == color) ren_ids.insert(id);
      }
      if (positions.n >= 9) return false;
    }
    if (ren_ids != ren_ids_in) return false;
    if (is_single_area(positions, color, -1)) {
      for (int k = 0; k < positions.n; k++) {
        const int pos = positions.v[k];
        if (gid[pos] != -1) continue;
        IdxSet tmp = positions;
        tmp.erase(pos);
        if (is_nakade(tmp)) return false;
      }
    }
    positions.erase(index);
    if (!is_single_area(positions, color, index)) return false;
    for (int k = 0; k < positions.n; k++) {
      const int pos = positions.v[k];
      if (gid[pos] != -1) continue;
      IdxSet tmp = positions;
      tmp.erase(pos);
      if (is_nakade(tmp)) return true;
    }
    return false;
  }

  // ---------------------------------------------------------------- Board::_isSeki (Board.cpp:1228-1291)
  bool is_seki(int index, int color) const {
    const int op = -color;
    for (int a : d) {
      const int id = gid[index + a];
      if (id >= 0 && b.c[index + a] == op && n_libs[id] == 1) return false;
    }
    IdxSet ren_ids;
    for (int a : d) {
      const int id = gid[index + a];
      if (id >= 0 && b.c[index + a] == color) ren_ids.insert(id);
    }
    if (ren_ids.n == 0) return false;
    // the reference inserts every group's liberties and gives up once the union holds >= 9 cells (checked
    // after each group): a group with >= 9 liberties of its own always ends there — skip the insertion work
    for (int k = 0; k < ren_ids.n; k++)
      if (n_libs[ren_ids.v[k]] >= 9) return false;
    IdxSet spaces;
    for (int a : d)
      if (gid[index + a] == -1) spaces.insert(index + a);
    for (int k = 0; k < ren_ids.n; k++) {
      const int id = ren_ids.v[k];
      for (int i = 0; i < n_libs[id]; i++) spaces.insert(lib_pool[lib_off[id] + i]);
      if (spaces.n >= 9) return false;
    }
    spaces.erase(index);
    if (spaces.n == 0) return false;
    if (spaces.n == 1) return is_seki_ren(index, color, ren_ids, spaces.v[0]);
    return is_seki_area(index, color, ren_ids, spaces);
  }

  // ---------------------------------------------------------------- Board::_isEnabled (Board.cpp:1177-1220)
  bool enabled(int index, int color, bool check_seki) const {
    if (gid[index] != -1) return false;
    if (index == b.ko_index && color == b.ko_color) return false;
    if (check_seki && is_seki(index, color)) return false;
    const int op = -color;
    for (int a : d) {
      const int t = index + a;
      if (gid[t] == -1) return true;
      if (gid[t] < 0) continue;                                  // border
      if (b.c[t] == color && n_libs[gid[t]] > 1) return true;
      if (b.c[t] == op && n_libs[gid[t]] == 1) return true;
    }
    return false;
  }

  // ---------------------------------------------------------------- Board::_updateArea + getTerritories
  // territories[y*w + x] = owner * color in {-1, 0, +1} (Board.cpp:350-378)
  void territories(int color, int8_t* out) const {
    bool fixed[LeanBoard::MAXN];                                 // per group id
    std::memset(fixed, 0, sizeof fixed);
    bool area_flag[2][LeanBoard::MAXN];                          // per component id: the area is decided
    for (int c = 0; c < 2; c++) {
      const int col = (c == 0) ? LeanBoard::kBlack : LeanBoard::kWhite;
      build_components(c);
      int16_t ids_v[LeanBoard::MAXN];                            // groups of this colour (their smallest cells)
      int ids_n = 0;
      for (int i = 0; i < b.N; i++)
        if (gid[i] == i && b.c[i] == col) ids_v[ids_n++] = (int16_t)i;
      // (group, decided area) registrations of this colour: Ren::areas of the reference, as a flat pair list
      int16_t pair_group[4 * LeanBoard::MAXN], pair_area[4 * LeanBoard::MAXN];
      int n_pairs = 0;
      for (int k = 0; k < ids_n; k++) fixed[ids_v[k]] = true;
      // (local copies: through `this` / `b` every store to the scratch arrays forces the compiler to reload them)
      const int8_t* const cc = b.c;
      const int16_t* const gg = gid;
      const int16_t* const cid = comp_id[c];
      const int N = b.N, d0 = d[0], d1 = d[1], d2 = d[2], d3 = d[3];
      bool* const afl = area_flag[c];
      // An area is decided iff every cell of it touches exactly the own groups its first cell touches, and that set is
      // not empty (Board.cpp:905-960: the flag of the first cell falls at the first cell that differs, whatever the
      // order of the flood). An open area fails at one of its first cells.
      for (int index = 0; index < N; index++) {
        if (cid[index] != index) continue;
        Ids4 connected;
        {
          const int ts[4] = {index + d0, index + d1, index + d2, index + d3};
          for (int t : ts)
            if (gg[t] >= 0 && cc[t] == col) connected.insert(gg[t]);
        }
        bool live = connected.n != 0;
        const int16_t* const cells = comp_pool[c] + comp_off[c][index];
        const int n_cells = comp_size[c][index];
        for (int k = 1; live && k < n_cells; k++) {              // (cells[0] is `index` itself)
          const int pos = cells[k];
          const int ts[4] = {pos + d0, pos + d1, pos + d2, pos + d3};
          Ids4 around;
          for (int t : ts)
            if (gg[t] >= 0 && cc[t] == col) around.insert(gg[t]);
          if (around != connected) live = false;
        }
        afl[index] = live;
        if (live) {
          for (int k = 0; k < connected.n; k++) {
            pair_group[n_pairs] = connected.v[k];
            pair_area[n_pairs++] = (int16_t)index;
          }
        }
      }
      // Greatest fixed point of "a group stays decided while >= 2 of its areas are decided; an area of an
      // undecided group is undecided" (Board.cpp:988-1032). Flags only ever fall, so the order in which the
      // reference visits the groups does not change the result: count per sweep instead of per group.
      bool updated = true;
      int16_t cnt[LeanBoard::MAXN];
      while (updated) {
        updated = false;
        for (int k = 0; k < ids_n; k++) cnt[ids_v[k]] = 0;
        for (int i = 0; i < n_pairs; i++)
          if (afl[pair_area[i]]) cnt[pair_group[i]]++;
        for (int k = 0; k < ids_n; k++)
          if (fixed[ids_v[k]] && cnt[ids_v[k]] < 2) fixed[ids_v[k]] = false;
        for (int i = 0; i < n_pairs; i++) {
          if (!fixed[pair_group[i]] && afl[pair_area[i]]) {
            afl[pair_area[i]] = false;
            updated = true;
          }
        }
      }
    }
    for (int y = 0; y < b.h; y++) {
      for (int x = 0; x < b.w; x++) {
        const int index = b.index(x, y);
        const int id = gid[index];
        int8_t t = 0;
        if (id >= 0 && fixed[id]) {
          t = (int8_t)(b.c[index] * color);
        } else if (comp_id[0][index] != -1 && area_flag[0][comp_id[0][index]]) {
          t = (int8_t)(LeanBoard::kBlack * color);
        } else if (comp_id[1][index] != -1 && area_flag[1][comp_id[1][index]]) {
          t = (int8_t)(LeanBoard::kWhite * color);
        }
        out[y * b.w + x] = t;
      }
    }
  }

  // bit (y*w + x) of mask = legal (seki-checked) AND not decided territory (Evaluator.cpp:60-68)
  void move_mask(int color, uint8_t* mask48) const {
    int8_t terr[361];
    territories(color, terr);
    std::memset(mask48, 0, 48);
    for (int y = 0; y < b.h; y++)
      for (int x = 0; x < b.w; x++) {
        const int i = y * b.w + x;
        if (terr[i] == 0 && enabled(b.index(x, y), color, true)) mask48[i >> 3] |= (uint8_t)(1u << (i & 7));
      }
  }
};

}  // namespace maru_b200
