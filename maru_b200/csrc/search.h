// search.h — native tree search of ONE game over the flat host board.
//
// The reference runs one descent per thread-pool task: Player::_run hands every visit to a worker
// (reference src/deepgo/native/cpp/Player.cpp:282-313), the worker walks Node::evaluate from the root
// (Player.cpp:319-368, Node.cpp:71-225) and BLOCKS inside Evaluator::evaluate until the leaf's batch
// has come back; every expanded node deep-copies a std::set-based Board (Node.cpp:588-598).
// Here a game is a resumable state machine: `GameSearch::advance` walks descents until one of them
// needs the network, writes that leaf's 448-byte record and returns; `deliver` hands the result back
// and the same descent continues.  One host thread can therefore interleave hundreds of games and
// submit their leaves as ONE asynchronous batch (selfplay.cpp) — no thread hand-off per visit, no
// condition variable per leaf, ~0.6 KB copied per expansion.
//
// What is kept EXACTLY (tests/test_search.py plays the unmodified reference Player and this search on
// the same evaluator and compares every move and every root statistic at threads = 1):
//   * progressive widening: a candidate's priority is prior^t * e^Gumbel / (times picked + 1); the
//     best one joins a FIFO waiting list if it is not a child yet                 Node.cpp:92-154
//   * a waiting candidate becomes a child (new leaf) before any existing child is revisited, and
//     the first child of a node reports playouts = -1                             Node.cpp:159-181
//   * root-only switches: `equally` (even visits, 1 / (n + 1 - q/2)), `width` (only the best
//     `width` children by value LCB stay eligible), UCB1 instead of PUCB          Node.cpp:183-224
//   * PUCB with c = log((1 + N + 19652) / 19652) + 1.25 and exploration doubled  Node.cpp:476-486
//   * value LCB = q - 1.96 * 0.5 / sqrt(n + 1) in the mover's favour              Node.cpp:460-469
//   * backup: the leaf's value is added to every node of the path, playouts too   Player.cpp:332-346
//   * Player::play keeps the chosen child's subtree, drops the rest               Player.cpp:93-115
//   * candidates, policy fallback and pass                                         Player.cpp:229-261
//   * the candidate moves of a node = legal (seki-checked) and not decided territory, with the
//     network's prior of that cell, not renormalised                                Evaluator.cpp:55-72
// All arithmetic is done in the same types (float / double promotions) as the reference's
// expressions, so a deterministic evaluator gives bit-identical trees.
// Not kept: the per-node shared_mutexes (one descent in flight per game: the reference's threads = 1)
// and the process-wide random engine (one seeded engine per game: self-play is reproducible).
// LAZY LEGALITY: Evaluator::evaluate filters the candidate moves of every evaluated node (seki-aware
// legality + Benson territory, Evaluator.cpp:60-68; ~80 % of the host time per leaf even on the flat
// board), but a node's candidates are only ever read on its SECOND visit (Node.cpp:81-89) and most leaves
// of a 50-visit search are visited once.  A leaf is therefore sent WITHOUT a move mask, the network's
// whole prior plane is parked in the node, and the filter runs when the node is first expanded.  The
// priors are not renormalised by the reference, so the candidate list is bit-identical either way.
#pragma once

#include <cmath>
#include <cstdint>
#include <deque>
#include <random>
#include <utility>
#include <vector>

#include "../dropin/lean_game.h"
#include "maru_b200.h"

namespace maru {

// bring-up only: cycle counters around the phases of a visit (tools/search_prof.cpp defines the T_* globals)
#ifdef MARU_SEARCH_PROFILE
#define MARU_PROF(var, stmt) do { const unsigned long long _t = __rdtsc(); stmt; var += __rdtsc() - _t; } while (0)
#define MARU_PROF_SCOPE(var) struct _ProfScope { unsigned long long t0 = __rdtsc(); ~_ProfScope() { var += __rdtsc() - t0; } } _prof_scope
#else
#define MARU_PROF(var, stmt) do { stmt; } while (0)
#define MARU_PROF_SCOPE(var) do { } while (0)
#endif

struct SearchSetup {
  int width = 19, height = 19;
  float komi = 7.5f;
  int rule = MARU_RULE_CH;
  bool superko = false;
  bool eval_leaf_only = false;          // Player::_evalLeafOnly (Player.cpp:337-341)
};

// arguments of Node::evaluate that apply to the root only (Player.cpp:358-363 resets them below it)
struct RootMode {
  bool equally = false;
  int width = 0;
  bool use_ucb1 = false;
  float temperature = 1.0f;
  float noise = 0.0f;
};

struct Prior {                          // deepgo::Policy (Policy.h:10-39)
  int16_t x, y;
  float p;
  int32_t picked;                       // Policy::visits: how often widening chose this move
};

struct TreeNode {
  maru_b200::LeanGame pos;              // position AFTER the move (x, y) of `color`
  int32_t x = -1, y = -1;
  int32_t color = MARU_WHITE;           // colour of the stone that led here (root of a game: WHITE)
  int32_t captured = 0;
  float prior = 0.0f;
  bool evaluated = false;
  float net_value = 0.0f;               // Evaluator::getValue(): black's point of view, in [-1, 1]
  std::vector<Prior> moves;             // Node::_childPolicies (built on the second visit, see below)
  std::vector<float> plane;             // the network's prior of every cell, until `moves` is built
  bool moves_built = false;
  // widening bookkeeping (see GameSearch::step): the moves picked at least once, and the best never-picked moves
  std::vector<int16_t> picked_list;
  static constexpr int FREE_TOP = 8;
  int16_t free_top[FREE_TOP];           // the next never-picked moves in the order widening will want them
  int8_t free_n = 0, free_pos = 0;      // (highest prior first, lowest index among equals); refilled by one scan
  bool free_exhausted = false;          // a refill found nothing: every move has been picked
  struct Edge {
    int32_t cell, node;
  };
  std::vector<Edge> kids;               // Node::_children, insertion order
  std::vector<Prior> waiting;           // Node::_waitingQueue / _waitingSet (FIFO, at most a few)
  int32_t visits = 0, playouts = 0, count = 0;
  float value_sum = 0.0f;

  float mean() const { return count == 0 ? 0.0f : value_sum / count; }           // Node::getValue
  int kid_of(int cell) const {
    for (const Edge& e : kids)
      if (e.cell == cell) return e.node;
    return -1;
  }
  bool is_waiting(int cell, int board_w) const {
    for (const Prior& q : waiting)
      if (q.y * board_w + q.x == cell) return true;
    return false;
  }
  void reset_stats() {                                                            // Node::_reset
    evaluated = false;
    net_value = 0.0f;
    moves.clear();
    plane.clear();
    moves_built = false;
    picked_list.clear();
    free_n = free_pos = 0;
    free_exhausted = false;
    kids.clear();
    waiting.clear();
    visits = playouts = count = 0;
    value_sum = 0.0f;
  }
};

struct Candidate {                      // deepgo::Candidate as Player::getCandidates fills it
  int32_t x, y, color, visits, playouts;
  float prior, value;
};

class GameSearch {
 public:
  enum Status { WANT_LEAF, MOVE_READY };

  void init(const SearchSetup& s, uint64_t seed) {
    setup_ = s;
    rng_.seed((std::minstd_rand0::result_type)(seed % 2147483646u + 1u));
    pool_.clear();
    free_.clear();
    root_ = alloc();
    TreeNode& r = pool_[root_];
    r.pos.init(s.width, s.height);
    r.x = r.y = -1;
    r.color = MARU_WHITE;
    r.captured = 0;
    r.prior = 0.0f;
    r.reset_stats();
    path_.clear();
    pending_ = -1;
  }

  const TreeNode& root() const { return pool_[root_]; }
  int to_move() const { return -pool_[root_].color; }                             // Player::getColor
  size_t nodes_in_use() const { return pool_.size() - free_.size(); }

  // Arm a search: descents are made until the ROOT's visit count reaches `target`
  // (Player::startEvaluation / waitEvaluation, Player.cpp:180-223).
  void start(const RootMode& m, int target) {
    mode_ = m;
    target_ = target;
  }

  // Run descents.  WANT_LEAF: `*st` holds the record of the node that needs the network; call
  // deliver() and then advance() again.  MOVE_READY: the root has reached the target.
  Status advance(maru_state* st) {
    MARU_PROF_SCOPE(T_advance);
    for (;;) {
      if (path_.empty()) {
        if (pool_[root_].visits >= target_) return MOVE_READY;
        path_.push_back(root_);
        cur_ = mode_;
      }
      const int id = path_.back();
      TreeNode& n = pool_[id];
      if (!n.evaluated) {                                                         // Node::_evaluate
        request(id, st);
        return WANT_LEAF;
      }
      int next = -1, playouts = 0;
      MARU_PROF(T_step, step(n, cur_, &next, &playouts));
      const float v = n.net_value;
      if (playouts == 1) {
        for (int p : path_) {
          pool_[p].count += 1;
          pool_[p].value_sum += v;
        }
      } else if (playouts == -1 && setup_.eval_leaf_only) {
        for (int p : path_) {
          pool_[p].count -= 1;
          pool_[p].value_sum -= v;
        }
      }
      for (int p : path_) pool_[p].playouts += playouts;
      if (next < 0) {
        path_.clear();
        descents_++;
        continue;
      }
      path_.push_back(next);
      cur_ = RootMode();
    }
  }

  // The network's answer for the record advance() / need_root() produced last.
  void deliver(const maru_state& st, const maru_leaf& leaf) {
    MARU_PROF_SCOPE(T_deliver);
    TreeNode& n = pool_[pending_];
    const int cells = st.width * st.height;
    n.moves.clear();
    n.plane.assign(leaf.prior, leaf.prior + cells);
    n.moves_built = false;
    const float win = leaf.value;                                                 // Evaluator.cpp:75-80
    n.net_value = (st.color == MARU_WHITE) ? 1.0f - 2.0f * win : 2.0f * win - 1.0f;
    n.evaluated = true;
    pending_ = -1;
    evals_++;
  }

  // Player::getCandidates needs the root's priors when the root has no child yet (Player.cpp:246-254):
  // true -> `*st` holds the root's record, deliver() its result before calling candidates().
  bool need_root(maru_state* st) {
    TreeNode& r = pool_[root_];
    if (!r.kids.empty() || r.evaluated) return false;
    request(root_, st);
    return true;
  }

  void candidates(std::vector<Candidate>* out) {
    TreeNode& r = pool_[root_];
    out->clear();
    for (const TreeNode::Edge& e : r.kids) {
      const TreeNode& c = pool_[e.node];
      out->push_back(Candidate{c.x, c.y, c.color, c.visits, c.playouts, c.prior, c.mean()});
    }
    if (!out->empty()) return;
    int bx = -1, by = -1;                                                         // Node::getPolicyMove
    build_moves(r);
    if (!r.moves.empty()) {
      const Prior* best = &r.moves[0];
      for (const Prior& q : r.moves)
        if (best->p < q.p) best = &q;
      bx = best->x;
      by = best->y;
    }
    out->push_back(Candidate{bx, by, -r.color, 1, 1, 1.0f, r.mean()});
  }

  // Player::play: the child at (x, y) becomes the root with its subtree, or a fresh node does
  // (Node::getChild); everything else goes back to the pool.  Returns the stones captured.
  int play(int x, int y) {
    MARU_PROF_SCOPE(T_play);
    const int old = root_;
    const int cell = y * setup_.width + x;
    int keep = pool_[old].kid_of(cell);
    if (keep < 0) {
      keep = alloc();
      follow(pool_[keep], pool_[old], x, y, 1.0f);
    }
    root_ = keep;
    release_except(old, keep);
    path_.clear();
    pending_ = -1;
    return pool_[root_].captured;
  }

  uint64_t descents() const { return descents_; }
  uint64_t evals() const { return evals_; }
  uint64_t masks() const { return masks_; }     // nodes whose candidate filter actually ran

 private:
  SearchSetup setup_;
  RootMode mode_, cur_;
  int target_ = 0;
  std::deque<TreeNode> pool_;           // stable addresses; recycled through free_
  std::vector<int> free_;
  int root_ = -1;
  std::vector<int> path_;
  int pending_ = -1;
  uint64_t descents_ = 0, evals_ = 0, masks_ = 0;
  std::minstd_rand0 rng_;               // the reference's std::default_random_engine (Node.cpp:11-12)
  std::vector<std::pair<int, float>> scratch_;

  int alloc() {
    if (!free_.empty()) {
      const int id = free_.back();
      free_.pop_back();
      return id;
    }
    pool_.emplace_back();
    return (int)pool_.size() - 1;
  }

  void release_except(int from, int keep) {
    std::vector<int> stack{from};
    while (!stack.empty()) {
      const int id = stack.back();
      stack.pop_back();
      if (id == keep) continue;
      for (const TreeNode::Edge& e : pool_[id].kids) stack.push_back(e.node);
      pool_[id].kids.clear();
      free_.push_back(id);
    }
  }

  void follow(TreeNode& n, const TreeNode& prev, int x, int y, float prior) {     // Node::_setAsNextNode
    n.x = x;
    n.y = y;
    n.color = -prev.color;
    n.prior = prior;
    MARU_PROF(T_follow, n.pos = prev.pos; n.captured = n.pos.play(x, y, n.color));
    n.reset_stats();
  }

  void request(int id, maru_state* st) {
    const TreeNode& n = pool_[id];
    MARU_PROF(T_req, n.pos.to_state(-n.color, setup_.komi, setup_.rule, setup_.superko, /*with_move_mask=*/false, st));
    pending_ = id;
  }

  // Evaluator.cpp:55-72, deferred to the node's first expansion: candidate = legal with the seki veto and
  // outside decided territory; prior = the network's output at that cell, in cell order, not renormalised.
  void build_moves(TreeNode& n) {
    if (n.moves_built) return;
    n.moves_built = true;
    const int w = setup_.width, cells = setup_.width * setup_.height;
    uint8_t mask[48];
    MARU_PROF(T_mask, maru_b200::LeanRules rules(n.pos.board); rules.move_mask(-n.color, mask));
    n.moves.clear();
    n.picked_list.clear();
    n.free_n = n.free_pos = 0;
    n.free_exhausted = false;
    for (int cell = 0; cell < cells; cell++)
      if ((mask[cell >> 3] >> (cell & 7)) & 1u)
        n.moves.push_back(Prior{(int16_t)(cell % w), (int16_t)(cell / w), n.plane[(size_t)cell], 0});
    n.plane.clear();
    masks_++;
  }

  float lcb(const TreeNode& c) const {                                            // Node::getValueLCB
    if (c.count == 0) return 0.0f;
    const float q = c.value_sum / c.count;
    const float half = 1.96 * 0.5 / std::sqrt(c.visits + 1);
    return q - (half * c.color);
  }
  // The terms that depend on the parent alone are computed once per visited node, not once per child (same
  // expressions in the same types, so the same bits): c_puct as float, sqrt / log of the visit count as double.
  static float pucb_c(int total) { return std::log((1 + total + 19652.0) / 19652.0) + 1.25; }
  float pucb(const TreeNode& c, float c_puct, double sqrt_total) const {         // Node::getPriorityByPUCB
    if (c.count == 0) return -99.0f;
    const float q = (c.value_sum / c.count) * c.color;
    const float u = c_puct * c.prior * sqrt_total / (1 + c.visits);
    return q + 2 * u;
  }
  float ucb1(const TreeNode& c, double log_total) const {                         // Node::getPriorityByUCB1
    if (c.count == 0) return -99.0f;
    const float q = (c.value_sum / c.count) * c.color;
    const float u = 0.5 * std::sqrt(log_total / (c.visits + 1));
    return q + u;
  }

  // One Node::evaluate on an evaluated node: *next = node to descend into (-1: the descent ends here),
  // *playouts = NodeResult::getPlayouts.
  void step(TreeNode& n, const RootMode& m, int* next, int* playouts) {
    n.visits += 1;
    *next = -1;
    *playouts = 1;
    if (n.visits == 1) return;
    build_moves(n);
    if (n.moves.empty()) return;
    const int bw = setup_.width;

    // --- widening: which move deserves to be a child next
    const int have = (int)(n.kids.size() + n.waiting.size());
    if (have < (int)n.moves.size() && (m.width < 1 || have < m.width)) {
      int best = 0, best_class = 0;
      float best_pri = 0.0f;
      const float win = n.mean() * (-n.color) * 0.5 + 0.5;
      const float power = win + (1.0 / m.temperature) * (1 - win);
      const float scale = (have <= 4) ? 0.0f : m.noise;
      // pow(p, 1) == p and p * e^0 == p exactly: the two transcendental calls per candidate are only
      // made when they can change the result (below the root the reference always passes t = 1, noise = 0)
      const bool tempered = power != 1.0f, noisy = scale != 0.0f;
      if (!tempered && !noisy && !m.equally) {
        // The common case (every node below the root, and a PUCB root at temperature 1): priority = p / (picked + 1),
        // first index wins among equals (Node.cpp:109-140).  A never-picked move has priority p exactly, so the winner
        // is either the best never-picked move or one of the few moves picked before — O(children) instead of
        // O(legal moves) per visited node, same result bit for bit. With a flat policy the best never-picked move wins
        // EVERY time, so the never-picked moves are fetched eight at a time (one scan of the list per eight picks
        // instead of one per pick: the scan was a fifth of the host time per visit).
        if (n.free_pos >= n.free_n && !n.free_exhausted) {
          int cnt = 0;
          int16_t* const top = n.free_top;
          for (int i = 0; i < (int)n.moves.size(); i++) {
            if (n.moves[i].picked != 0) continue;
            const float p = n.moves[i].p;
            if (cnt == TreeNode::FREE_TOP && !(p > n.moves[top[cnt - 1]].p)) continue;     // (a later index never beats an equal prior)
            int j = (cnt < TreeNode::FREE_TOP) ? cnt++ : cnt - 1;
            for (; j > 0 && p > n.moves[top[j - 1]].p; j--) top[j] = top[j - 1];
            top[j] = (int16_t)i;
          }
          n.free_n = (int8_t)cnt;
          n.free_pos = 0;
          n.free_exhausted = cnt == 0;
        }
        const int best_free = (n.free_pos < n.free_n) ? n.free_top[n.free_pos] : -1;
        bool have_best = false;
        auto consider = [&](int i, float pri) {
          if (!have_best || pri > best_pri || (pri == best_pri && i < best)) {
            best = i;
            best_pri = pri;
            have_best = true;
          }
        };
        if (best_free >= 0) consider(best_free, n.moves[best_free].p / (0 + 1));
        for (int16_t i : n.picked_list) consider(i, n.moves[i].p / (n.moves[i].picked + 1));
        if (!have_best) best = 0;
      } else {
      std::extreme_value_distribution<float> gumbel(0.0f, noisy ? scale : 1.0f);
      for (int i = 0; i < (int)n.moves.size(); i++) {
        const Prior& q = n.moves[i];
        float pr = q.p;
        if (tempered) pr = std::pow(pr, power);
        if (noisy) pr *= std::exp(gumbel(rng_));
        int cls = 1;
        const float pri = pr / (q.picked + 1);
        if (m.equally) {
          const int cell = q.y * bw + q.x;
          if (n.kid_of(cell) >= 0 || n.is_waiting(cell, bw)) cls = 0;
        }
        if (cls > best_class || (cls == best_class && pri > best_pri)) {
          best = i;
          best_class = cls;
          best_pri = pri;
        }
      }
      }
      Prior& chosen = n.moves[best];
      const int cell = chosen.y * bw + chosen.x;
      if (n.kid_of(cell) < 0 && !n.is_waiting(cell, bw)) n.waiting.push_back(chosen);
      if (chosen.picked == 0) {
        n.picked_list.push_back((int16_t)best);
        // (a never-picked move is only ever chosen through the head of the list on the fast path; the general path
        // below may pick any of them, which invalidates the list)
        if (n.free_pos < n.free_n && best == n.free_top[n.free_pos]) n.free_pos++;
        else n.free_n = n.free_pos = 0;
      }
      chosen.picked += 1;
    }

    // --- a waiting move becomes a new leaf
    if (!n.waiting.empty() && (m.width <= 0 || (int)n.kids.size() < m.width)) {
      const Prior q = n.waiting.front();
      n.waiting.erase(n.waiting.begin());
      const int cell = q.y * bw + q.x;
      if (n.kid_of(cell) < 0) {
        const bool first = n.kids.empty();
        const int id = alloc();                       // (deque: `n` stays valid)
        follow(pool_[id], n, q.x, q.y, q.p);
        n.kids.push_back(TreeNode::Edge{cell, id});
        *next = id;
        *playouts = first ? -1 : 0;
        return;
      }
    }

    // --- otherwise revisit the most promising child
    // (the LCB ranking is only computed when the root width actually cuts the list; otherwise the children keep their
    // order, as in the reference, and the ranking would not be looked at)
    scratch_.clear();
    const bool cut = m.width > 0 && (int)n.kids.size() > m.width;
    for (const TreeNode::Edge& e : n.kids) {
      const TreeNode& c = pool_[e.node];
      scratch_.emplace_back(e.node, cut ? lcb(c) * c.color : 0.0f);
    }
    if (cut) {
      std::sort(scratch_.begin(), scratch_.end(),
                [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
      scratch_.resize(m.width);
    }
    int pick = scratch_[0].first;
    float top = -1.0;
    const float c_puct = pucb_c(n.visits);
    const double sqrt_total = std::sqrt(n.visits);
    const double log_total = m.use_ucb1 ? std::log(n.visits) : 0.0;
    for (const auto& kv : scratch_) {
      const TreeNode& c = pool_[kv.first];
      float pri;
      if (m.equally) {
        const float cv = c.visits;
        const float q = c.mean() * c.color;
        pri = 1.0 / (cv + 1 - q * 0.5);
      } else if (m.use_ucb1) {
        pri = ucb1(c, log_total);
      } else {
        pri = pucb(c, c_puct, sqrt_total);
      }
      if (top < pri) {
        pick = kv.first;
        top = pri;
      }
    }
    *next = pick;
    *playouts = 0;
  }
};

}  // namespace maru
