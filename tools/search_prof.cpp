// search_prof.cpp — where do the host cycles of one visit go?  (bring-up tool, not part of the product)
// selfplay.cpp + search.h compiled with -DMARU_SEARCH_PROFILE (rdtsc around the phases of a visit) over a cheap
// in-process evaluator; one search thread, no GPU.
//   g++ -O2 -std=c++17 -DMARU_SEARCH_PROFILE -DMARU_TSAN_STUB_ENGINE -Iinclude -Imaru_b200/csrc -I/usr/local/cuda/include \
//       tools/search_prof.cpp -o tools/_search_prof -lpthread && tools/_search_prof 19 256 60 [moves] [sharpness 0..3]
#include <x86intrin.h>
unsigned long long T_follow, T_req, T_mask, T_step, T_deliver, T_advance, T_leaf, T_play;
#include "../maru_b200/csrc/queue.cpp"
#include "../maru_b200/csrc/selfplay.cpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>

static inline uint64_t mix(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  return x;
}
// peaked pseudo-policy from a hash of the record: weight = u^8, a handful of moves carry most of the mass
static int g_sharp = 3;   // weight = u^(2^g_sharp): 0 flat, 3 a handful of moves carry most of the mass
static int cheap_leaves(const maru_state* s, maru_leaf* out, int n, void*) {
  const unsigned long long t0 = __rdtsc();
  for (int k = 0; k < n; k++) {
    const maru_state& st = s[k];
    const int cells = st.width * st.height;
    uint64_t h = 0x9e3779b97f4a7c15ull;
    const uint64_t* w64 = reinterpret_cast<const uint64_t*>(st.cells);
    for (int i = 0; i < (cells + 7) / 8; i++) h = mix(h ^ w64[i]);
    h = mix(h ^ (uint64_t)(uint8_t)st.color);
    float sum = 0.f;
    for (int i = 0; i < cells; i++) {
      float u = (float)(mix(h + 0x1234567ull * (uint64_t)(i + 1)) >> 40) * (1.0f / 16777216.0f);
      for (int r = 0; r < g_sharp; r++) u *= u;
      out[k].prior[i] = u;
      sum += u;
    }
    const float inv = 1.0f / sum;
    for (int i = 0; i < cells; i++) out[k].prior[i] *= inv;
    out[k].value = 0.1f + 0.8f * (float)(mix(h ^ 0xabcdefull) >> 40) * (1.0f / 16777216.0f);
    out[k].score[0] = out[k].score[1] = 0.f;
  }
  T_leaf += __rdtsc() - t0;
  return 0;
}

int main(int argc, char** argv) {
  const int size = argc > 1 ? atoi(argv[1]) : 19, games = argc > 2 ? atoi(argv[2]) : 256, opening = argc > 3 ? atoi(argv[3]) : 60;
  const int moves = argc > 4 ? atoi(argv[4]) : 6;
  if (argc > 5) g_sharp = atoi(argv[5]);
  maru_selfplay_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.width = cfg.height = size;
  cfg.komi = 7.5f;
  cfg.rule = MARU_RULE_CH;
  cfg.games = games;
  cfg.threads = 1;
  cfg.visits = 50;
  cfg.max_moves = 400;
  cfg.root = MARU_ROOT_PUCB;
  cfg.temperature = 1.0f;
  cfg.criterion = MARU_CRITERION_LCB;
  cfg.opening_moves = opening;
  cfg.restart = 1;
  cfg.seed = 1;
  cfg.lanes = 1;
  maru_selfplay* sp = nullptr;
  if (maru_selfplay_create_fn(cheap_leaves, nullptr, &cfg, &sp)) { fprintf(stderr, "create failed: %s\n", maru_last_error()); return 1; }
  maru_selfplay_run(sp, 1);
  maru_selfplay_stats s0, s1;
  maru_selfplay_get_stats(sp, &s0);
  T_follow = T_req = T_mask = T_step = T_deliver = T_advance = T_leaf = T_play = 0;
  const unsigned long long c0 = __rdtsc();
  const auto w0 = std::chrono::steady_clock::now();
  maru_selfplay_run(sp, moves);
  const unsigned long long cyc = __rdtsc() - c0;
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
  maru_selfplay_get_stats(sp, &s1);
  const double v = (double)(s1.visits - s0.visits), e = (double)(s1.evals - s0.evals);
  const double ns_per_cyc = wall * 1e9 / (double)cyc;
  auto us = [&](unsigned long long t) { return (double)t * ns_per_cyc / v / 1e3; };
  printf("%dx%d, %d games, 1 thread: %.0f visits, %.0f evals, wall %.2f s = %.2f us per visit (evaluator %.2f us of it)\n", size, size,
         games, v, e, wall, wall * 1e6 / v, us(T_leaf));
  printf("per visit (us): advance %.2f [step %.2f incl. mask %.2f and follow %.2f; request/to_state %.2f]  deliver %.2f  play %.2f  other %.2f\n",
         us(T_advance), us(T_step), us(T_mask), us(T_follow), us(T_req), us(T_deliver), us(T_play),
         (wall * 1e6 / v) - us(T_advance) - us(T_deliver) - us(T_leaf) - us(T_play));
  printf("host_s per visit from the scheduler's own counters: %.2f us; candidate filters (move_mask) per visit %.3f = %.1f us each\n", (s1.host_s - s0.host_s) * 1e6 / v,
         (double)(s1.expansions - s0.expansions) / v, us(T_mask) * v / (double)(s1.expansions - s0.expansions + 1));
  maru_selfplay_destroy(sp);
  return 0;
}
