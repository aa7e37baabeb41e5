// selfplay.cpp — self-play game scheduler over the native search (SURVEY 8 rows f1 / e; north_star:
// "the self-play game scheduler ... games are independent, so self-play is sharded by game").
//
// The reference has no self-play driver; its pieces are Player::startEvaluation / waitEvaluation /
// getCandidates / play (reference src/deepgo/native/cpp/Player.cpp:93-115, 180-261) and the Python
// move decision around them (src/deepgo/player.py:295-366).  Driven as the reference is built — one
// Player (+ a search thread pool and a control thread) per game, every visit handed to a pool thread
// that blocks in Processor::execute — a GPU sees batches of a few leaves and the host spends its time
// in condition variables (profiles/r1_selfplay_visits.txt: 45 k visits/s flat, 0.33 efficiency at 4 GPUs).
//
// Here one scheduler owns G games and T worker threads.  A worker owns G/T games for their whole life
// (no locks on trees) split into `lanes`; per lane it (1) collects the previous batch's results and hands
// them to the waiting descents, (2) advances every game of the lane until it needs the network again or
// has nothing left to do, writing the 448-byte records side by side, (3) submits them as ONE asynchronous
// batch (maru_queue_submit_leaves) and turns to the next lane while the GPU works.  Nothing in a worker
// blocks except the wait for its oldest ticket.
#include <pthread.h>
#include <sched.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#ifdef MARU_TSAN_STUB_ENGINE
#include "engine_stub.h"   // `make tsan`: host-only stand-in, see tests/tsan_queue_selfplay.cpp
#else
#include "engine.h"
#endif
#include "search.h"

namespace maru {

// player.py:92-93: win chance of the mover and its lower confidence bound (Python floats = doubles)
static double win_chance_lcb(const Candidate& c) {
  const double win = (double)c.value * c.color * 0.5 + 0.5;
  return win - 1.96 * 0.25 / std::sqrt((double)c.visits + 1.0);
}

// Sequential-halving phases over `width` candidates (maru_b200/selfplay.py: halving_schedule):
// widths m, ceil(m/2), ..., 1; the budget is split evenly, earlier phases take the remainder.
static void halving_schedule(int visits, int width, std::vector<std::pair<int, int>>* out) {
  out->clear();
  if (visits <= 0) return;
  if (width < 1) width = 1;
  std::vector<int> widths{width};
  while (widths.back() > 1) widths.push_back((widths.back() + 1) / 2);
  const int n = (int)widths.size();
  if (visits < n) {
    for (int i = 0; i < visits; i++) out->emplace_back(widths[i], 1);
    return;
  }
  const int base = visits / n, rem = visits % n;
  for (int i = 0; i < n; i++) out->emplace_back(widths[i], base + (i < rem ? 1 : 0));
}

struct GameRunner {
  GameSearch search;
  const maru_selfplay_config* cfg = nullptr;
  int index = 0;
  uint64_t generation = 0;              // restarts so far (re-seeds the opening)
  int moves_played = 0, passes = 0, goal = 0;
  bool finished = false;
  enum Stage { PLAN, NEXT_PHASE, SEARCHING, SELECT } stage = PLAN;
  std::vector<std::pair<int, int>> phases;
  size_t phase = 0;
  int target = 0;
  std::vector<int16_t> record;          // (x, y) per move
  std::vector<Candidate> cand_log;      // cfg->record: candidates of every move, back to back
  std::vector<int32_t> cand_off;        // start of move i in cand_log (+ one end marker)
  std::vector<Candidate> cands;
  uint64_t visits_done = 0, moves_done = 0, games_done = 0;   // monotone counters of this game slot

  void setup_search() {
    SearchSetup s;
    s.width = cfg->width;
    s.height = cfg->height;
    s.komi = cfg->komi;
    s.rule = cfg->rule;
    s.superko = cfg->superko != 0;
    s.eval_leaf_only = cfg->eval_leaf_only != 0;
    search.init(s, cfg->seed * 1000003ull + (uint64_t)index * 7919ull + generation);
    moves_played = 0;
    passes = 0;
    finished = false;
    stage = PLAN;
    record.clear();
    cand_log.clear();
    cand_off.assign(1, 0);
    if (cfg->opening_moves > 0) play_opening();
  }

  // seeded random legal opening (mid-game positions for throughput runs); seki-checked legality
  void play_opening() {
    std::mt19937_64 rng(cfg->seed * 9176ull + (uint64_t)index * 31337ull + generation * 101ull + 17ull);
    for (int m = 0; m < cfg->opening_moves; m++) {
      const maru_b200::LeanGame& g = search.root().pos;
      const int color = search.to_move();
      maru_b200::LeanRules rules(g.board);
      int16_t legal[361];
      int n = 0;
      for (int y = 0; y < cfg->height; y++)
        for (int x = 0; x < cfg->width; x++)
          if (rules.enabled(g.board.index(x, y), color, true)) legal[n++] = (int16_t)(y * cfg->width + x);
      if (n == 0) break;
      const int cell = legal[rng() % (uint64_t)n];
      search.play(cell % cfg->width, cell / cfg->width);
    }
  }

  // true: *st holds a leaf for the network (deliver() its result, then call again); false: idle
  // (goal reached or game over).
  bool advance(maru_state* st) {
    for (;;) {
      if (moves_done_in_run >= goal) return false;
      if (finished) {
        if (!cfg->restart) return false;
        generation++;
        games_done++;
        setup_search();
      }
      switch (stage) {
        case PLAN: {
          if (cfg->root == MARU_ROOT_GUMBEL) {
            halving_schedule(cfg->visits, cfg->root_width, &phases);
          } else {
            phases.clear();
            if (cfg->visits > 0) phases.emplace_back(0, cfg->visits);
          }
          phase = 0;
          target = search.root().visits;
          visits_at_plan = search.descents();
          stage = NEXT_PHASE;
          break;
        }
        case NEXT_PHASE: {
          if (phase >= phases.size()) {
            stage = SELECT;
            break;
          }
          RootMode m;
          m.temperature = cfg->temperature;
          m.use_ucb1 = cfg->use_ucb1 != 0;
          if (cfg->root == MARU_ROOT_GUMBEL) {
            m.equally = true;
            m.width = phases[phase].first;
            m.noise = cfg->noise;
          }
          target += phases[phase].second;
          search.start(m, target);
          stage = SEARCHING;
          break;
        }
        case SEARCHING: {
          if (search.advance(st) == GameSearch::WANT_LEAF) return true;
          phase++;
          stage = NEXT_PHASE;
          break;
        }
        case SELECT: {
          if (search.need_root(st)) return true;
          search.candidates(&cands);
          choose_and_play();
          stage = PLAN;
          break;
        }
      }
    }
  }

  void choose_and_play() {
    // player.py:341-352 sorts by the criterion and the caller takes the first; max() over the
    // candidate order keeps the first of equals like Python's max / stable sort do
    size_t best = 0;
    for (size_t i = 1; i < cands.size(); i++) {
      const bool better = cfg->criterion == MARU_CRITERION_VISITS
                              ? cands[i].visits > cands[best].visits
                              : win_chance_lcb(cands[i]) > win_chance_lcb(cands[best]);
      if (better) best = i;
    }
    const Candidate mv = cands[best];
    if (cfg->record) {
      cand_log.insert(cand_log.end(), cands.begin(), cands.end());
      cand_off.push_back((int32_t)cand_log.size());
    }
    record.push_back((int16_t)mv.x);
    record.push_back((int16_t)mv.y);
    visits_done += search.descents() - visits_at_plan;
    search.play(mv.x, mv.y);
    moves_played++;
    moves_done++;
    moves_done_in_run++;
    const bool pass = mv.x < 0 || mv.y < 0;
    passes = pass ? passes + 1 : 0;
    if (passes >= 2 || moves_played >= cfg->max_moves) finished = true;
  }

  uint64_t visits_at_plan = 0;
  int moves_done_in_run = 0;
};

struct GameLane {
  std::vector<int> games;               // indices into maru_selfplay::games
  std::vector<maru_state> states;       // (queue / function back ends; an evaluator lane stages in pinned memory)
  std::vector<maru_leaf> leaves;
  std::vector<int> owner;               // game of row i of the batch in flight
  int n = 0;
  maru_ticket* ticket = nullptr;
  Lane* direct = nullptr;               // evaluator lane (maru_selfplay_create_eval)
  bool inflight = false;
  maru_state* in() { return direct != nullptr ? direct->h_states : states.data(); }
  const maru_leaf* out() const { return direct != nullptr ? direct->h_leaves : leaves.data(); }
};

struct WorkerStats {
  uint64_t submits = 0, max_submit = 0;
  double host_s = 0, stall_s = 0, device_s = 0;
};

}  // namespace maru

struct maru_selfplay {
  maru_selfplay_config cfg{};
  maru_queue* queue = nullptr;
  maru::Engine* engine = nullptr;        // evaluator lanes: the search threads submit their batches themselves
  maru_leaf_fn fn = nullptr;
  void* fn_user = nullptr;
  std::mutex fn_mu;                      // a caller-supplied evaluator is called by one worker at a time
  std::vector<std::unique_ptr<maru::GameRunner>> games;
  std::vector<maru::WorkerStats> wstats;
  std::atomic<int> failed{0};
  std::string error;
  std::mutex err_mu;
  double run_s = 0;
  bool running = false;
};

namespace maru {

using Clock = std::chrono::steady_clock;
static double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

static void fail(maru_selfplay* sp, const std::string& what) {
  std::lock_guard<std::mutex> l(sp->err_mu);
  if (!sp->failed.exchange(1)) sp->error = what;
}

static void worker_run(maru_selfplay* sp, int widx, int n_workers) {
  const maru_selfplay_config& cfg = sp->cfg;
  // Search workers are throughput threads; the leaf queue's worker thread (gather, launch, scatter: a few per cent of a
  // core) is latency-critical — the GPU idles until it has run.  When the workers occupy every core of the rank, a lower
  // priority lets the scheduler put the queue thread on a core the moment it wakes.  (Raising a priority needs
  // privileges, lowering one does not.)
  setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 5);
  if (cfg.cpu_count > 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(cfg.cpu_first + widx % cfg.cpu_count, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);   // best effort
  }
  const int n_lanes = cfg.lanes > 0 ? cfg.lanes : 2;
  std::vector<GameLane> lanes((size_t)n_lanes);
  int k = 0;
  for (int g = widx; g < (int)sp->games.size(); g += n_workers) lanes[(size_t)(k++ % n_lanes)].games.push_back(g);
  bool broken = false;
  for (GameLane& L : lanes) {
    L.owner.resize(L.games.size());
    if (L.games.empty()) continue;
    if (sp->engine != nullptr) {
      L.direct = sp->engine->lane_create((int)L.games.size());
      if (L.direct == nullptr) {
        fail(sp, std::string("evaluator lane: ") + g_last_error);
        broken = true;
      }
    } else {
      L.states.resize(L.games.size());
      L.leaves.resize(L.games.size());
    }
  }
  WorkerStats& ws = sp->wstats[(size_t)widx];
  bool any = true;
  while (any && !broken && !sp->failed.load(std::memory_order_relaxed)) {
    any = false;
    for (GameLane& L : lanes) {
      if (L.games.empty()) continue;
      auto t0 = Clock::now();
      if (L.inflight) {
        L.inflight = false;
        int rc = 0;
        if (L.direct != nullptr) {
          float ms = 0;
          rc = sp->engine->lane_wait(L.direct, &ms);
          ws.device_s += ms * 1e-3;
        } else if (sp->queue != nullptr) {
          rc = maru_queue_wait(sp->queue, L.ticket);
        }
        if (rc != 0) {
          fail(sp, std::string("leaf batch failed: ") + g_last_error);
          broken = true;
          break;
        }
        const auto t1 = Clock::now();
        ws.stall_s += secs(t0, t1);
        t0 = t1;
        const maru_state* in = L.in();
        const maru_leaf* out = L.out();
        for (int i = 0; i < L.n; i++) sp->games[(size_t)L.owner[(size_t)i]]->search.deliver(in[i], out[i]);
      }
      L.n = 0;
      maru_state* in = L.in();
      for (int g : L.games)
        if (sp->games[(size_t)g]->advance(&in[L.n])) L.owner[(size_t)L.n++] = g;
      ws.host_s += secs(t0, Clock::now());
      if (L.n == 0) continue;
      any = true;
      ws.submits++;
      if ((uint64_t)L.n > ws.max_submit) ws.max_submit = (uint64_t)L.n;
      if (L.direct != nullptr) {
        if (sp->engine->lane_submit(L.direct, L.n) != 0) {
          fail(sp, std::string("leaf submit failed: ") + g_last_error);
          broken = true;
          break;
        }
      } else if (sp->queue != nullptr) {
        if (maru_queue_submit_leaves(sp->queue, L.states.data(), L.leaves.data(), L.n, &L.ticket) != 0) {
          fail(sp, std::string("leaf submit failed: ") + g_last_error);
          broken = true;
          break;
        }
      } else {
        std::lock_guard<std::mutex> l(sp->fn_mu);
        if (sp->fn(L.states.data(), L.leaves.data(), L.n, sp->fn_user) != 0) {
          fail(sp, "leaf function failed");
          broken = true;
          break;
        }
      }
      L.inflight = true;
    }
  }
  // stopping early (a failure here or in another worker): still retire this worker's batches, the evaluator writes
  // into the lanes' buffers until then
  for (GameLane& L : lanes) {
    if (L.inflight && L.direct != nullptr) sp->engine->lane_wait(L.direct, nullptr);
    else if (L.inflight && sp->queue != nullptr) maru_queue_wait(sp->queue, L.ticket);
    if (L.direct != nullptr) sp->engine->lane_destroy(L.direct);
  }
}

static int check_cfg(const maru_selfplay_config* c) {
  if (c == nullptr) return 1;
  if (c->width < 1 || c->width > 19 || c->height < 1 || c->height > 19) return 1;
  if (c->games < 1 || c->threads < 1 || c->visits < 0 || c->max_moves < 1) return 1;
  if (c->root != MARU_ROOT_PUCB && c->root != MARU_ROOT_GUMBEL) return 1;
  if (c->criterion != MARU_CRITERION_LCB && c->criterion != MARU_CRITERION_VISITS) return 1;
  if (!(c->temperature > 0.0f) || c->noise < 0.0f || c->lanes < 0 || c->lanes > 16) return 1;
  return 0;
}

static int create(maru_queue* q, Engine* e, maru_leaf_fn fn, void* user, const maru_selfplay_config* cfg, maru_selfplay** out) {
  if (out == nullptr || (q == nullptr && fn == nullptr && e == nullptr) || check_cfg(cfg)) {
    set_error("maru_selfplay_create: bad argument");
    return 1;
  }
  *out = nullptr;
  try {
    auto sp = std::make_unique<maru_selfplay>();
    sp->cfg = *cfg;
    if (sp->cfg.threads > sp->cfg.games) sp->cfg.threads = sp->cfg.games;
    sp->queue = q;
    sp->engine = e;
    sp->fn = fn;
    sp->fn_user = user;
    for (int g = 0; g < cfg->games; g++) {
      auto r = std::make_unique<GameRunner>();
      r->cfg = &sp->cfg;
      r->index = g;
      sp->games.push_back(std::move(r));
    }
    for (auto& r : sp->games) r->setup_search();
    sp->wstats.resize((size_t)sp->cfg.threads);
    *out = sp.release();
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("maru_selfplay_create: ") + ex.what());
    return 1;
  }
}

}  // namespace maru

extern "C" {

int maru_selfplay_create(maru_queue* q, const maru_selfplay_config* cfg, maru_selfplay** out) {
  if (q == nullptr) {
    maru::set_error("maru_selfplay_create: no leaf-batch queue (there is no CPU evaluator)");
    return 1;
  }
  return maru::create(q, nullptr, nullptr, nullptr, cfg, out);
}

int maru_selfplay_create_fn(maru_leaf_fn fn, void* user, const maru_selfplay_config* cfg, maru_selfplay** out) {
  if (fn == nullptr) {
    maru::set_error("maru_selfplay_create_fn: no leaf function");
    return 1;
  }
  return maru::create(nullptr, nullptr, fn, user, cfg, out);
}

int maru_selfplay_create_eval(maru_eval* e, const maru_selfplay_config* cfg, maru_selfplay** out) {
  if (e == nullptr) {
    maru::set_error("maru_selfplay_create_eval: no evaluator (there is no CPU evaluator)");
    return 1;
  }
  return maru::create(nullptr, reinterpret_cast<maru::Engine*>(e), nullptr, nullptr, cfg, out);
}

void maru_selfplay_destroy(maru_selfplay* sp) { delete sp; }

int maru_selfplay_play(maru_selfplay* sp, int game, int x, int y) {
  if (sp == nullptr || game < 0 || game >= (int)sp->games.size() || sp->running) {
    maru::set_error("maru_selfplay_play: bad argument");
    return -1;
  }
  maru::GameRunner& r = *sp->games[(size_t)game];
  const bool pass = x < 0 || y < 0 || x >= sp->cfg.width || y >= sp->cfg.height;
  const int captured = r.search.play(pass ? -1 : x, pass ? -1 : y);
  if (captured < 0) {
    maru::set_error("maru_selfplay_play: illegal move");
    return -1;
  }
  r.stage = maru::GameRunner::PLAN;
  return captured;
}

int maru_selfplay_run(maru_selfplay* sp, int moves) {
  if (sp == nullptr || moves < 0 || sp->running) {
    maru::set_error("maru_selfplay_run: bad argument");
    return 1;
  }
  if (sp->failed.load()) {
    maru::set_error("maru_selfplay_run: scheduler failed earlier: " + sp->error);
    return 1;
  }
  sp->running = true;
  for (auto& r : sp->games) {
    r->goal = moves;
    r->moves_done_in_run = 0;
  }
  const auto t0 = maru::Clock::now();
  const int n = sp->cfg.threads;
  try {
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++) th.emplace_back(maru::worker_run, sp, i, n);
    for (auto& t : th) t.join();
  } catch (const std::exception& ex) {
    maru::fail(sp, std::string("worker thread: ") + ex.what());
  }
  sp->run_s += maru::secs(t0, maru::Clock::now());
  sp->running = false;
  if (sp->failed.load()) {
    maru::set_error("maru_selfplay_run: " + sp->error);
    return 1;
  }
  return 0;
}

int maru_selfplay_get_stats(maru_selfplay* sp, maru_selfplay_stats* st) {
  if (sp == nullptr || st == nullptr) return 1;
  std::memset(st, 0, sizeof(*st));
  for (auto& r : sp->games) {
    st->visits += r->visits_done;
    st->evals += r->search.evals();
    st->expansions += r->search.masks();
    st->moves += r->moves_done;
    st->games_finished += r->games_done + (r->finished ? 1 : 0);
  }
  for (const maru::WorkerStats& w : sp->wstats) {
    st->submits += w.submits;
    if (w.max_submit > st->max_submit) st->max_submit = w.max_submit;
    st->host_s += w.host_s;
    st->stall_s += w.stall_s;
    st->device_s += w.device_s;
  }
  st->run_s = sp->run_s;
  return 0;
}

int maru_selfplay_get_moves(maru_selfplay* sp, int game, int16_t* xy, int cap) {
  if (sp == nullptr || game < 0 || game >= (int)sp->games.size()) return -1;
  const std::vector<int16_t>& r = sp->games[(size_t)game]->record;
  const int n = (int)r.size() / 2;
  if (xy != nullptr)
    for (int i = 0; i < n && i < cap; i++) {
      xy[2 * i] = r[2 * (size_t)i];
      xy[2 * i + 1] = r[2 * (size_t)i + 1];
    }
  return n;
}

int maru_selfplay_get_candidates(maru_selfplay* sp, int game, int move, maru_candidate* out, int cap) {
  if (sp == nullptr || game < 0 || game >= (int)sp->games.size()) return -1;
  const maru::GameRunner& r = *sp->games[(size_t)game];
  if (move < 0 || move + 1 >= (int)r.cand_off.size()) return -1;
  const int a = r.cand_off[(size_t)move], b = r.cand_off[(size_t)move + 1];
  if (out != nullptr)
    for (int i = a; i < b && i - a < cap; i++) {
      const maru::Candidate& c = r.cand_log[(size_t)i];
      out[i - a] = maru_candidate{c.x, c.y, c.color, c.visits, c.playouts, c.prior, c.value};
    }
  return b - a;
}

}  // extern "C"
