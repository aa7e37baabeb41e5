// tsan_queue_selfplay.cpp — ThreadSanitizer harness for the host-side concurrency of the hot path
// (SURVEY 5 "race detection"; the reference has none).  Built by `make -C maru_b200/csrc tsan`:
// queue.cpp + selfplay.cpp compiled UNCHANGED with -fsanitize=thread over engine_stub.h (a host-only
// evaluator), no CUDA.  Exercises, concurrently:
//   * 12 threads of blocking maru_queue_execute_leaves with n = 1 (how Evaluator.cpp:52 submits),
//   * 4 threads of asynchronous tickets (submit / poll / wait, several in flight, retired out of order),
//   * coalescing on and off while jobs are in flight, two evaluators behind one queue (least-loaded routing),
//   * a self-play scheduler (4 workers x 3 lanes, 48 games, restarts) on the same queue at the same time,
//   * a second scheduler on evaluator lanes (maru_selfplay_create_eval: the workers submit their own batches).
// Exit code 0 and no "WARNING: ThreadSanitizer" on stderr = pass (tests/test_tsan.py runs it when built).
#include "../maru_b200/csrc/queue.cpp"
#include "../maru_b200/csrc/selfplay.cpp"

#include <atomic>
#include <cstdio>

int main() {
  maru::Engine e0(256), e1(256);
  maru_eval* evs[2] = {reinterpret_cast<maru_eval*>(&e0), reinterpret_cast<maru_eval*>(&e1)};
  maru_queue* q = nullptr;
  if (maru_queue_create(evs, 2, 256, &q)) return 2;
  std::atomic<int> bad{0};
  auto make_state = [](int seed) {
    maru_state s;
    std::memset(&s, 0, sizeof s);
    s.width = s.height = 9;
    s.color = 1;
    for (int i = 0; i < 81; i++) s.cells[i] = (uint8_t)((seed * 31 + i * 7) % 3);
    return s;
  };
  auto check = [&](const maru_state& s, const maru_leaf& got) {
    maru_leaf want;
    maru::Engine::eval(s, &want);
    if (std::memcmp(&want, &got, sizeof want) != 0) bad++;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < 12; t++)
    th.emplace_back([&, t] {
      for (int i = 0; i < 300; i++) {
        maru_state s = make_state(t * 1000 + i);
        maru_leaf out;
        if (maru_queue_execute_leaves(q, &s, &out, 1)) bad++;
        check(s, out);
      }
    });
  for (int t = 0; t < 4; t++)
    th.emplace_back([&, t] {
      for (int round = 0; round < 60; round++) {
        constexpr int K = 5, N = 7;
        maru_state s[K][N];
        maru_leaf out[K][N];
        maru_ticket* tk[K];
        for (int k = 0; k < K; k++) {
          for (int i = 0; i < N; i++) s[k][i] = make_state(t * 100000 + round * 100 + k * 10 + i);
          if (maru_queue_submit_leaves(q, s[k], out[k], N, &tk[k])) bad++;
        }
        int done = 0;
        maru_queue_poll(q, tk[K - 1], &done);
        for (int k = K - 1; k >= 0; k--) {               // retire newest first
          if (maru_queue_wait(q, tk[k])) bad++;
          for (int i = 0; i < N; i++) check(s[k][i], out[k][i]);
        }
      }
    });
  th.emplace_back([&] {
    for (int i = 0; i < 40; i++) {
      maru_queue_set_coalesce(q, (i & 1) ? 32 : 0, 200);
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    maru_queue_set_coalesce(q, 0, 0);
  });
  maru_selfplay_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.width = cfg.height = 7;
  cfg.komi = 7.5f;
  cfg.games = 48;
  cfg.threads = 4;
  cfg.visits = 12;
  cfg.max_moves = 6;
  cfg.root = MARU_ROOT_GUMBEL;
  cfg.root_width = 4;
  cfg.noise = 1.0f;
  cfg.temperature = 1.0f;
  cfg.criterion = MARU_CRITERION_VISITS;
  cfg.opening_moves = 2;
  cfg.restart = 1;
  cfg.seed = 3;
  cfg.lanes = 3;
  maru_selfplay* sp = nullptr;
  if (maru_selfplay_create(q, &cfg, &sp)) return 3;
  th.emplace_back([&] {
    if (maru_selfplay_run(sp, 10)) bad++;
  });
  // a second scheduler straight on an evaluator through lanes (no queue thread), at the same time
  maru_selfplay* sp2 = nullptr;
  maru_selfplay_config cfg2 = cfg;
  cfg2.games = 24;
  cfg2.threads = 3;
  cfg2.seed = 9;
  if (maru_selfplay_create_eval(evs[1], &cfg2, &sp2)) return 4;
  th.emplace_back([&] {
    if (maru_selfplay_run(sp2, 10)) bad++;
  });
  for (auto& t : th) t.join();
  maru_selfplay_stats st2;
  maru_selfplay_get_stats(sp2, &st2);
  if (st2.moves != 240 || st2.visits != 12ull * 240ull) bad++;
  maru_selfplay_destroy(sp2);
  maru_selfplay_stats st;
  maru_selfplay_get_stats(sp, &st);
  maru_queue_stats qs;
  maru_queue_get_stats(q, &qs);
  maru_selfplay_destroy(sp);
  maru_queue_destroy(q);
  std::printf("tsan harness: %d mismatches, %llu self-play visits, %llu moves, %llu queue batches, max batch %llu\n",
              bad.load(), (unsigned long long)st.visits, (unsigned long long)st.moves,
              (unsigned long long)qs.batches, (unsigned long long)qs.max_batch_seen);
  return (bad.load() == 0 && st.moves == 480 && st.visits == 12ull * 480ull) ? 0 : 1;
}
