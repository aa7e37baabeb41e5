// engine_stub.h — TEST-ONLY stand-in for engine.h (selected by -DMARU_TSAN_STUB_ENGINE, `make tsan`).
//
// Lets queue.cpp and selfplay.cpp be compiled by plain g++ with -fsanitize=thread and no CUDA: the "evaluator"
// fills results on the host (a hash of the record), so ThreadSanitizer sees the real dispatcher / worker / ticket /
// scheduler synchronisation of the product code and none of the GPU runtime.  Never part of libmaru_b200.so.
#pragma once

#include <chrono>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "maru_b200.h"

inline int cudaSetDevice(int) { return 0; }

namespace maru {

inline thread_local std::string g_last_error;
inline void set_error(const std::string& s) { g_last_error = s; }

constexpr int IN_ROW = MARU_INPUT_SIZE, OUT_ROW = MARU_OUTPUT_SIZE;
enum { KIND_PLANES = 0, KIND_STATES = 1, KIND_LEAVES = 2 };
inline size_t in_row_bytes(int kind) { return kind == KIND_PLANES ? (size_t)IN_ROW * 4 : sizeof(maru_state); }
inline size_t out_row_bytes(int kind) { return kind == KIND_LEAVES ? sizeof(maru_leaf) : (size_t)OUT_ROW * 4; }

struct Slot {
  std::vector<float> rows, out;
  std::vector<maru_state> states;
  float* h_rows = nullptr;
  maru_state* h_states = nullptr;
  float* h_out = nullptr;
  int kind = 0, n = 0;
};

struct Engine;
struct Lane {
  Engine* e = nullptr;
  std::vector<maru_state> vs;
  std::vector<maru_leaf> vl;
  maru_state* h_states = nullptr;
  maru_leaf* h_leaves = nullptr;
  int rows = 0, n = 0;
  bool inflight = false;
};

struct Engine {
  int gpu_id = 0, max_batch = 0;
  double device_ms_last = 0;
  Slot slot[3];
  std::mutex mu;

  explicit Engine(int mb) : max_batch(mb) {
    for (Slot& s : slot) {
      s.rows.resize((size_t)mb * 8);           // plane jobs are not exercised by the harness
      s.states.resize((size_t)mb);
      s.out.resize((size_t)mb * OUT_ROW);
      s.h_rows = s.rows.data();
      s.h_states = s.states.data();
      s.h_out = s.out.data();
    }
  }
  static void eval(const maru_state& st, maru_leaf* leaf) {
    std::memset(leaf, 0, sizeof(*leaf));
    uint32_t h = 2166136261u;
    for (int i = 0; i < st.width * st.height; i++) h = (h ^ st.cells[i]) * 16777619u;
    for (int i = 0; i < st.width * st.height; i++) {
      h = h * 1664525u + 1013904223u;
      leaf->prior[i] = (float)((h >> 8) & 0xffff) / 65536.0f / (float)(st.width * st.height);
    }
    leaf->value = (float)((h >> 4) & 0xfff) / 4096.0f;
  }
  int enqueue_slot(int kind, int idx, int n) {
    std::lock_guard<std::mutex> l(mu);
    slot[idx].kind = kind;
    slot[idx].n = n;
    return 0;
  }
  int wait_slot(int idx, float* ms) {
    Slot& s = slot[idx];
    std::this_thread::sleep_for(std::chrono::microseconds(50));     // "GPU time": lets batches accumulate
    if (s.kind == KIND_LEAVES)
      for (int i = 0; i < s.n; i++) eval(s.h_states[i], reinterpret_cast<maru_leaf*>(s.h_out) + i);
    if (ms) *ms = 0.05f;
    return 0;
  }
  Lane* lane_create(int rows) {
    Lane* l = new Lane();
    l->e = this;
    l->rows = rows;
    l->vs.resize((size_t)rows);
    l->vl.resize((size_t)rows);
    l->h_states = l->vs.data();
    l->h_leaves = l->vl.data();
    return l;
  }
  void lane_destroy(Lane* l) { delete l; }
  int lane_submit(Lane* l, int n) {
    std::lock_guard<std::mutex> g(mu);
    l->n = n;
    l->inflight = true;
    return 0;
  }
  int lane_wait(Lane* l, float* ms) {
    if (l->inflight)
      for (int i = 0; i < l->n; i++) eval(l->h_states[i], &l->h_leaves[i]);
    l->inflight = false;
    if (ms) *ms = 0.05f;
    return 0;
  }
  int forward_host(int kind, const void* in, void* out, int n) {
    std::lock_guard<std::mutex> l(mu);
    if (kind == KIND_LEAVES)
      for (int i = 0; i < n; i++) eval(static_cast<const maru_state*>(in)[i], static_cast<maru_leaf*>(out) + i);
    return 0;
  }
};

}  // namespace maru

// the part of capi.cpp the two files call
extern "C" inline const char* maru_last_error(void) { return maru::g_last_error.c_str(); }
