// capi.cpp — extern "C" surface of libmaru_b200.so (declared in include/maru_b200.h).
#include "../dropin/lean_game.h"
#include <cstring>
#include <exception>
#include <new>

#include "engine.h"

using maru::Engine;

static Engine* E(maru_eval* e) { return reinterpret_cast<Engine*>(e); }

#define CKC(expr)                                                               \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      maru::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));      \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

extern "C" {

int maru_abi_version(void) { return 1; }

const char* maru_last_error(void) { return maru::g_last_error.c_str(); }

int maru_eval_create(const char* weights_path, int gpu_id, int max_batch, unsigned flags,
                     maru_eval** out) {
  if (weights_path == nullptr || out == nullptr) {
    maru::set_error("maru_eval_create: null argument");
    return 1;
  }
  *out = nullptr;
  Engine* e = new (std::nothrow) Engine();
  if (e == nullptr) {
    maru::set_error("out of memory");
    return 1;
  }
  int rc = 1;
  try {                                      // no C++ exception may cross the C ABI
    rc = e->init(weights_path, gpu_id, max_batch, flags);
  } catch (const std::exception& ex) {
    maru::set_error(std::string("maru_eval_create: ") + ex.what());
  } catch (...) {
    maru::set_error("maru_eval_create: unknown exception");
  }
  if (rc) {
    const std::string keep = maru::g_last_error;
    delete e;
    maru::set_error(keep);
    return 1;
  }
  *out = reinterpret_cast<maru_eval*>(e);
  return 0;
}

void maru_eval_destroy(maru_eval* e) { delete E(e); }

int maru_eval_get_info(maru_eval* ev, maru_eval_info* info) {
  if (ev == nullptr || info == nullptr) return 1;
  Engine* e = E(ev);
  std::memset(info, 0, sizeof(*info));
  info->blocks = e->blocks;
  info->channels = e->C;
  info->attn_every = e->attn_every;
  info->heads = e->C / e->head_dim;
  info->head_channels = e->Ch;
  info->value_channels = e->Cv;
  info->n_attn = e->n_attn;
  info->gpu_id = e->gpu_id;
  info->max_batch = e->max_batch;
  info->sm_count = e->sm_count;
  info->launches_per_forward = e->launches_per_forward;
  info->flops_per_eval = e->flops_alg;
  info->flops_per_eval_executed = e->flops_exec;
  info->n_forward = e->n_forward;
  info->n_positions = e->n_positions;
  info->device_ms_last = e->device_ms_last;
  return 0;
}

int maru_eval_forward_planes(maru_eval* e, const float* in, float* out, int n) {
  if (e == nullptr || in == nullptr || out == nullptr || n < 0) {
    maru::set_error("maru_eval_forward_planes: bad argument");
    return 1;
  }
  return E(e)->forward_host(maru::KIND_PLANES, in, out, n);
}

int maru_eval_forward_states(maru_eval* e, const maru_state* s, float* out, int n) {
  if (e == nullptr || s == nullptr || out == nullptr || n < 0) {
    maru::set_error("maru_eval_forward_states: bad argument");
    return 1;
  }
  return E(e)->forward_host(maru::KIND_STATES, s, out, n);
}

int maru_eval_forward_leaves(maru_eval* e, const maru_state* s, maru_leaf* out, int n) {
  if (e == nullptr || s == nullptr || out == nullptr || n < 0) {
    maru::set_error("maru_eval_forward_leaves: bad argument");
    return 1;
  }
  return E(e)->forward_host(maru::KIND_LEAVES, s, out, n);
}

int maru_eval_forward_leaves_device(maru_eval* ev, const maru_state* d_states, maru_leaf* d_out, int n) {
  if (ev == nullptr || d_states == nullptr || d_out == nullptr || n < 0) {
    maru::set_error("maru_eval_forward_leaves_device: bad argument");
    return 1;
  }
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  for (int off = 0; off < n; off += e->max_batch) {
    const int m = n - off < e->max_batch ? n - off : e->max_batch;
    if (e->forward_device(maru::KIND_LEAVES, d_states + off, d_out + off, m, e->dev_w, e->dev_h)) return 1;
  }
  return 0;
}

int maru_eval_forward_states_device(maru_eval* ev, const maru_state* d_states, float* d_out, int n) {
  if (ev == nullptr || d_states == nullptr || d_out == nullptr || n < 0) {
    maru::set_error("maru_eval_forward_states_device: bad argument");
    return 1;
  }
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  for (int off = 0; off < n; off += e->max_batch) {
    const int m = n - off < e->max_batch ? n - off : e->max_batch;
    if (e->forward_device(maru::KIND_STATES, d_states + off, d_out + (size_t)off * maru::OUT_ROW, m, e->dev_w, e->dev_h))
      return 1;
  }
  return 0;
}

int maru_eval_forward_planes_device(maru_eval* ev, const float* d_in, float* d_out, int n) {
  if (ev == nullptr || d_in == nullptr || d_out == nullptr || n < 0) {
    maru::set_error("maru_eval_forward_planes_device: bad argument");
    return 1;
  }
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  for (int off = 0; off < n; off += e->max_batch) {
    const int m = n - off < e->max_batch ? n - off : e->max_batch;
    if (e->forward_device(maru::KIND_PLANES, d_in + (size_t)off * maru::IN_ROW,
                          d_out + (size_t)off * maru::OUT_ROW, m, e->dev_w, e->dev_h))
      return 1;
  }
  return 0;
}

int maru_eval_sync(maru_eval* ev) {
  Engine* e = E(ev);
  CKC(cudaSetDevice(e->gpu_id));
  CKC(cudaStreamSynchronize(e->stream));
  return 0;
}

void* maru_eval_stream(maru_eval* ev) { return ev ? (void*)E(ev)->stream : nullptr; }

int maru_encode_planes(maru_eval* e, const maru_state* s, float* out_rows, int n) {
  if (e == nullptr || s == nullptr || out_rows == nullptr || n < 0) {
    maru::set_error("maru_encode_planes: bad argument");
    return 1;
  }
  return E(e)->encode_rows_host(s, out_rows, n);
}

int maru_encode_planes_device(maru_eval* ev, const maru_state* d_s, float* d_rows, int n) {
  if (ev == nullptr || d_s == nullptr || d_rows == nullptr || n < 0) {
    maru::set_error("maru_encode_planes_device: bad argument");
    return 1;
  }
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  CKC(maru::launch_encode_rows(d_s, d_rows, n, nullptr, e->stream));
  return 0;
}

int maru_eval_debug_planes(maru_eval* ev, int which, float* out, int n, int channels) {
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  const maru::Act* acts[7] = {&e->x0, &e->h, &e->r, &e->s, &e->qkv, &e->o, &e->g};
  if (which < 0 || which > 6 || acts[which]->p == nullptr || n > e->max_batch ||
      channels > acts[which]->channels) {
    maru::set_error("maru_eval_debug_planes: bad argument");
    return 1;
  }
  CKC(cudaSetDevice(e->gpu_id));
  float* d = nullptr;
  const size_t bytes = (size_t)n * channels * maru::CELLS * 4;
  CKC(cudaMalloc(&d, bytes));
  // x0 always lives on the 19x19 frame; the other tensors on the frame of the last forward (compact for small boards)
  const maru::Geom gm = (which == MARU_BUF_X0) ? maru::Geom() : e->last;
  cudaError_t le = maru::launch_planes_to_nchw(acts[which]->p, e->rows_alloc, channels, acts[which]->channels,
                                               acts[which]->blocked ? 1 : 0, d, n, gm.pitch, gm.pos_rows, gm.w, gm.h,
                                               e->stream);
  if (le == cudaSuccess) le = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, e->stream);
  if (le == cudaSuccess) le = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  CKC(le);
  return 0;
}

int maru_eval_profile_states(maru_eval* ev, const maru_state* d_states, float* d_out, int n,
                             maru_launch_time* out, int max_launches, int* n_launches) {
  if (ev == nullptr || out == nullptr || n_launches == nullptr) {
    maru::set_error("maru_eval_profile_states: null argument");
    return 1;
  }
  return E(ev)->profile_states(d_states, d_out, n, out, max_launches, n_launches);
}

int maru_eval_debug_trace(maru_eval* ev, long long* out, int n_ctas) {
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  if (n_ctas > 256) n_ctas = 256;
  CKC(cudaSetDevice(e->gpu_id));
  CKC(cudaStreamSynchronize(e->stream));
  CKC(cudaMemcpy(out, e->d_trace, (size_t)n_ctas * maru::TRACE_TILES * maru::TRACE_SLOTS * 8,
                 cudaMemcpyDeviceToHost));
  return 0;
}

int maru_eval_debug_timeline(maru_eval* ev, long long* out) {
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  CKC(cudaStreamSynchronize(e->stream));
  CKC(cudaMemcpy(out, e->d_timeline, 64 * 2 * 8 * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int maru_eval_debug_stamps(maru_eval* ev, long long* out) {
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  CKC(cudaStreamSynchronize(e->stream));
  CKC(cudaMemcpy(out, e->d_stamps, 512 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

struct maru_game {
  maru_b200::LeanGame g;
};

maru_game* maru_game_create(int width, int height) {
  if (width < 1 || height < 1 || width > 19 || height > 19) {
    maru::set_error("maru_game_create: bad board size");
    return nullptr;
  }
  maru_game* p = new (std::nothrow) maru_game();
  if (p != nullptr) p->g.init(width, height);
  return p;
}
maru_game* maru_game_clone(const maru_game* g) { return g ? new (std::nothrow) maru_game(*g) : nullptr; }
void maru_game_destroy(maru_game* g) { delete g; }
int maru_game_play(maru_game* g, int x, int y, int color) {
  if (g == nullptr || (color != 1 && color != -1)) return -1;
  return g->g.play(x, y, color);
}
int maru_game_state(const maru_game* g, int color, float komi, int rule, int superko, int with_move_mask,
                    maru_state* out) {
  if (g == nullptr || out == nullptr || (color != 1 && color != -1)) {
    maru::set_error("maru_game_state: bad argument");
    return 1;
  }
  g->g.to_state(color, komi, rule, superko != 0, with_move_mask != 0, out);
  return 0;
}

static void fill_lean_board(maru_b200::LeanBoard& b, const int8_t* colors, int width, int height, int ko_x,
                            int ko_y, int ko_color);

int maru_rules_eval(const int8_t* colors, int width, int height, int ko_x, int ko_y, int ko_color, int color,
                    uint8_t* enableds, int8_t* territories) {
  if (colors == nullptr || width < 1 || height < 1 || width > 19 || height > 19 || (color != 1 && color != -1)) {
    maru::set_error("maru_rules_eval: bad argument");
    return 1;
  }
  maru_b200::LeanBoard b;
  fill_lean_board(b, colors, width, height, ko_x, ko_y, ko_color);
  maru_b200::LeanRules r(b);
  if (enableds != nullptr)
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++) enableds[y * width + x] = r.enabled(b.index(x, y), color, true) ? 1 : 0;
  if (territories != nullptr) r.territories(color, territories);
  return 0;
}

static void fill_lean_board(maru_b200::LeanBoard& b, const int8_t* colors, int width, int height, int ko_x,
                            int ko_y, int ko_color) {
  b.init(width, height);
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const int8_t v = colors[y * width + x];
      b.c[b.index(x, y)] = v > 0 ? maru_b200::LeanBoard::kBlack : v < 0 ? maru_b200::LeanBoard::kWhite
                                                                    : maru_b200::LeanBoard::kEmpty;
    }
  if (ko_x >= 0 && ko_y >= 0 && ko_x < width && ko_y < height && ko_color != 0) {
    b.ko_index = b.index(ko_x, ko_y);
    b.ko_color = ko_color > 0 ? 1 : -1;
  }
}

int maru_ladder_flags(const int8_t* colors, int width, int height, int ko_x, int ko_y, int ko_color,
                      uint8_t* flags) {
  if (colors == nullptr || flags == nullptr || width < 1 || height < 1 || width > 19 || height > 19) {
    maru::set_error("maru_ladder_flags: bad argument");
    return 1;
  }
  maru_b200::LeanBoard b;
  fill_lean_board(b, colors, width, height, ko_x, ko_y, ko_color);
  b.ladder_flags(flags);
  return 0;
}

// ---- evaluator lanes (engine.cu: Engine::lane_*) ------------------------------------------------------------------
struct maru_lane {
  maru::Lane l;
};
static maru::Lane* LN(maru_lane* l) { return reinterpret_cast<maru::Lane*>(l); }

int maru_eval_lane_create(maru_eval* ev, int max_rows, maru_lane** out) {
  if (ev == nullptr || out == nullptr) {
    maru::set_error("maru_eval_lane_create: null argument");
    return 1;
  }
  *out = reinterpret_cast<maru_lane*>(E(ev)->lane_create(max_rows));
  return *out == nullptr ? 1 : 0;
}
void maru_lane_destroy(maru_lane* l) {
  if (l != nullptr) LN(l)->e->lane_destroy(LN(l));
}
maru_state* maru_lane_states(maru_lane* l) { return l ? LN(l)->h_states : nullptr; }
const maru_leaf* maru_lane_leaves(maru_lane* l) { return l ? LN(l)->h_leaves : nullptr; }
int maru_lane_submit(maru_lane* l, int n) {
  if (l == nullptr) {
    maru::set_error("maru_lane_submit: null lane");
    return 1;
  }
  return LN(l)->e->lane_submit(LN(l), n);
}
int maru_lane_poll(maru_lane* l, int* done) {
  if (l == nullptr || done == nullptr) {
    maru::set_error("maru_lane_poll: null argument");
    return 1;
  }
  return LN(l)->e->lane_poll(LN(l), done);
}
int maru_lane_wait(maru_lane* l) {
  if (l == nullptr) {
    maru::set_error("maru_lane_wait: null lane");
    return 1;
  }
  return LN(l)->e->lane_wait(LN(l), nullptr);
}

int maru_eval_set_device_board(maru_eval* ev, int width, int height) {
  if (ev == nullptr || width < 0 || height < 0 || width > 19 || height > 19 || ((width == 0) != (height == 0))) {
    maru::set_error("maru_eval_set_device_board: bad argument");
    return 1;
  }
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  e->dev_w = width;
  e->dev_h = height;
  return 0;
}

int maru_eval_debug_set(maru_eval* ev, int key, int value) {
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  CKC(cudaSetDevice(e->gpu_id));
  CKC(cudaStreamSynchronize(e->stream));
  e->clear_graphs();
  if (key == MARU_DBG_STOP_AFTER) e->dbg_stop_after = value;
  else if (key == MARU_DBG_DESC_BITS) e->dbg_desc = value;
  else if (key == MARU_DBG_TRACE_LAUNCH) e->dbg_trace_launch = value;
  else if (key == MARU_DBG_TIMELINE) e->dbg_timeline = value;
  else if (key == MARU_DBG_FLOW_TRACE_CTA) e->dbg_flow_trace_cta = value;
  else if (key == MARU_DBG_FLOW_TRACE_LAUNCH) e->dbg_flow_trace_chain = value;
  else if (key == MARU_DBG_ATT_TRACE_CTA) e->dbg_att_trace_cta = value;
  else {
    maru::set_error("maru_eval_debug_set: unknown key");
    return 1;
  }
  return 0;
}

int maru_eval_debug_policy_logits(maru_eval* ev, float* out, int n) {
  Engine* e = E(ev);
  std::lock_guard<std::mutex> lock(e->mu);
  if (n > e->max_batch) {
    maru::set_error("maru_eval_debug_policy_logits: n > max_batch");
    return 1;
  }
  CKC(cudaSetDevice(e->gpu_id));
  CKC(cudaMemcpyAsync(out, e->d_logits, (size_t)n * 2 * maru::CELLS * 4, cudaMemcpyDeviceToHost,
                      e->stream));
  CKC(cudaStreamSynchronize(e->stream));
  return 0;
}

}  // extern "C"
