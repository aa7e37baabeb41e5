// engine.h — internal C++ face of the evaluator (one GPU, one stream, one weight replica).
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.h"

namespace maru {

extern thread_local std::string g_last_error;
void set_error(const std::string& s);

struct Pack;

struct ConvLayer {
  std::string name;
  int cin = 0, cout = 0, taps = 0, nt = 0;
  __nv_bfloat16* w = nullptr;
  float* b = nullptr;
  // persistent flow kernel (conv_flow.cu): per channel group [W image | bias image], and the
  // channels-per-tile it was packed for (0: this layer cannot run inside a chain)
  uint8_t* blob = nullptr;
  int nt_flow = 0;
  uint8_t* blob_pair = nullptr;   // per group of cout/2 output channels (conv_pair.cu: one group per CTA of a pair)
  uint8_t* blob_wide = nullptr;   // the same per group of 256 output channels (conv_wide.cu), when the shape allows
  int nt_blob = 0;     // channels per group the blob was packed for (nt_flow, or nt for the K-split kernel)
};
struct Block {
  ConvLayer reduce, c[4], expand;
};
struct Attn {
  ConvLayer qkv, proj;
};

// transport of one forward: fp32 plane rows -> rows, compact records -> rows, compact records -> compact results
enum { KIND_PLANES = 0, KIND_STATES = 1, KIND_LEAVES = 2 };
inline size_t in_row_bytes(int kind) { return kind == KIND_PLANES ? (size_t)IN_ROW * 4 : sizeof(maru_state); }
inline size_t out_row_bytes(int kind) { return kind == KIND_LEAVES ? sizeof(maru_leaf) : (size_t)OUT_ROW * 4; }

// activation tensor: chunk-planar [C/8][rows_alloc][8] or tile-blocked [tile][C/8][128][8] (kernels.h)
struct Act {
  __nv_bfloat16* p = nullptr;
  int channels = 0;
  bool blocked = false;
  unsigned* flags = nullptr;   // per-tile completion counters (cross-layer pipelining)
  unsigned version = 0;        // host-side: counter value every tile reaches once the last writer is done
  bool flagged = false;        // the current contents were produced by a flag-bumping conv
};

// device + pinned staging for one in-flight batch (the queue double-buffers over two of these)
struct Slot {
  float* d_rows = nullptr;
  maru_state* d_states = nullptr;
  float* d_out = nullptr;
  float* h_rows = nullptr;
  maru_state* h_states = nullptr;
  float* h_out = nullptr;
  cudaEvent_t done = nullptr;              // D2H of the batch finished
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // kernel span of the batch (timing events)
};

// frame geometry of the layers being enqueued (common.cuh): the reference's 19x19 frame, or the compact frame of a
// batch of w x h boards (after the stem)
struct Geom {
  int pitch = PITCH, pos_rows = POS_ROWS, w = 0, h = 0;
  bool compact() const { return w != 0; }
  static Geom board(int bw, int bh) {
    Geom g;
    if (bw > 0 && bh > 0 && (bw < BOARD || bh < BOARD)) {
      g.pitch = bw + 1;
      g.pos_rows = (bw + 1) * (bh + 1);
      g.w = bw;
      g.h = bh;
    }
    return g;
  }
};

struct Engine;
// pinned staging of one in-flight batch (include/maru_b200.h: maru_lane)
struct Lane {
  Engine* e = nullptr;
  maru_state* h_states = nullptr;
  maru_leaf* h_leaves = nullptr;
  int rows = 0, n = 0;
  bool inflight = false;
  cudaEvent_t done = nullptr, t0 = nullptr, t1 = nullptr;
};

struct GraphKey {
  int kind, bucket;
  const void* in;
  const void* out;
  bool flow;
  int w = 0, h = 0;   // compact-frame board of the forward (0, 0: the 19x19 frame)   // captured with the persistent flow kernels (only while the engine has the device to itself)
  bool operator<(const GraphKey& o) const {
    return std::tie(kind, bucket, in, out, flow, w, h) < std::tie(o.kind, o.bucket, o.in, o.out, o.flow, o.w, o.h);
  }
};

struct Engine {
  bool dev_counted = false;
  int gpu_id = -1, max_batch = 0, sm_count = 0, rows_alloc = 0;
  unsigned flags = 0;
  int blocks = 0, C = 0, attn_every = 0, head_dim = 0, Ch = 0, Cv = 0, n_attn = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::mutex mu;  // serialises the blocking entry points (a Model is single-threaded, Executor.cpp:172)

  ConvLayer stem, headg;
  std::vector<Block> blk;
  std::vector<Attn> att;
  float *w_planes = nullptr, *b_planes = nullptr, *w_fc1 = nullptr, *b_fc1 = nullptr,
        *w_fc2 = nullptr, *b_fc2 = nullptr;

  Act x0, h, r, s, qkv, o, g;
  uint8_t* mask = nullptr;          // board mask on the 19x19 frame (written by encode / ingest)
  uint8_t* mask_c = nullptr;        // the same on the compact frame of the forward (written by the crop kernel)
  Geom geom;                        // geometry of the layers being enqueued right now
  Geom fwd;                         // trunk geometry of the forward being enqueued (compact for a batch of small boards)
  Geom last;                        // ... of the last forward enqueued (debug_planes)
  bool compact_on = true;           // MARU_B200_COMPACT=0 / debug bit 21: always the 19x19 frame
  int dev_w = 0, dev_h = 0;         // board size of the records of the *_device entry points (0: unknown -> 19x19 frame)
  const uint8_t* cur_mask() const { return geom.compact() ? mask_c : mask; }
  int* d_n = nullptr;
  int* h_n = nullptr;
  unsigned n_seq = 0;
  float* d_logits = nullptr;
  Slot slot[3];  // 0,1: leaf-queue double buffer; 2: direct blocking calls
  std::vector<void*> allocs;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::map<GraphKey, int> graph_launches;   // kernels per replay of each graph (info().launches_per_forward follows the path in use)

  int dbg_stop_after = -1;  // debug: stop the launch sequence after this many trunk launches
  int dbg_desc = 0;         // debug: bit0 = disable programmatic dependent launch
  int launches_done = 0;
  long long* d_trace = nullptr;   // conv pipeline trace (debug key 2 = index of the conv launch to trace)
  int dbg_trace_launch = -1;
  long long* d_timeline = nullptr;   // [64 launches][2][8] globaltimer stamps (debug key 3 = on/off)
  int dbg_timeline = 0;
  int dbg_flow_trace_cta = -1;     // debug key 4: CTA of the flow kernel to trace per item; key 5: which flow launch
  int dbg_flow_trace_chain = 0;
  int dbg_att_trace_cta = -1;      // debug key 6 (EXPERIMENTS builds)
  void clear_graphs();

  // per-launch profiling (maru_eval_profile_states): when prof != nullptr every launch is bracketed
  std::vector<maru_launch_time>* prof = nullptr;
  std::vector<cudaEvent_t> prof_ev;
  int prof_n = 0;
  int prof_reps = 1;        // each launch is repeated back to back this many times inside its event pair
  int prof_only = -1;       // >= 0: only that launch of the sequence is repeated (profile_states: one clean pass per launch)
  int reps_now() const;
  void prof_begin(const char* name, int kind, int cin, int cout, int taps, double flops, double bytes,
                  cudaStream_t st);
  void prof_end(cudaStream_t st);
  int profile_states(const maru_state* d_states, float* d_out, int n, maru_launch_time* out, int max,
                     int* count);

  double flops_alg = 0, flops_exec = 0;
  int launches_per_forward = 0;
  uint64_t n_forward = 0, n_positions = 0;
  double device_ms_last = 0;

  int init(const char* weights_path, int gpu, int max_batch, unsigned flags);
  ~Engine();

  int upload_conv(const Pack& pk, const std::string& name, int cin, int cout, int taps, ConvLayer& L);
  int upload_f32(const Pack& pk, const std::string& name, size_t count, float** dst);
  int alloc_act(int channels, bool blocked, Act* a);
  unsigned* d_flags = nullptr;   // all Act::flags live in this one allocation (one memset per forward)
  size_t flags_bytes = 0;
  int n_tiles_max = 0;

  int conv(const ConvLayer& L, Act& in, Act* res, Act& out, int relu, int n, const int* n_dev,
           cudaStream_t st);
  // consecutive conv layers are collected here and launched as ONE persistent kernel by flush_chain()
  FlowParams pending;
  ConvParams pending_single;        // the same layer as a one-layer launch (a chain of one falls back to it)
  const ConvLayer* pending_first = nullptr;
  int pending_w = 0;                // weight-blob region of the pending chain
  unsigned* d_sync = nullptr;       // neighbour flags of the flow kernels, [2][256] (zeroed per forward)
  unsigned flag_next = 0;           // next unused flag value of the forward being enqueued
  int n_flow_launches = 0;
  int pending_convs = 0;            // conv layers (not channel groups) in the pending chain
  double pending_flops = 0, pending_bytes = 0;
  bool flow_enabled();
  bool flow_device_ok();
  bool pair_pays(int cin, int n) const;
  bool pair64_on = false;           // measured slower than conv_gemm<64,64,9> at every batch size
  int pair64_min_batch = 1;
  bool flow_pays(int n) const;      // enough tiles per CTA for the persistent kernels to win
  bool flow_live = true;            // flow_choice(live batch size) of the forward being enqueued
  // Inside the window of flow_pays the two paths are within ~8 % of each other and WHICH one wins depends on the GPU
  // (same driver, same clocks: flow 448 vs per-layer 466 us on one B200, 484 vs 449 us on another - the SM -> die map is
  // per part). The first forward in the window on a board geometry times both graphs and the engine keeps the faster.
  std::map<int, int> flow_tuned;    // key w * 32 + h -> 1 flow, 0 one launch per layer
  bool flow_tune_on = true;         // MARU_B200_FLOW_TUNE=0: trust the window
  bool flow_choice(int n) const;    // flow_pays(n) and not measured slower
  int tune_flow(int kind, const void* d_in, void* d_out, int n);
  int enqueue_forward(int kind, const void* d_in, void* d_out, int n, bool* ok);
  bool flow_broken = false;         // a flow launch failed (cooperative grid does not fit: shared device): never again
  void device_ref(int delta);
  long long* d_stamps = nullptr;    // per-layer %globaltimer stamps of the chains (profiling)
  int stamps_used = 0;
  int n_launched = 0;               // kernels launched by the forward being enqueued
  int flush_chain(cudaStream_t st);
  int launch_trunk(int kind, const void* d_in, void* d_out, int n, const int* n_dev, cudaStream_t st);
  int launch_all(int kind, const void* d_in, void* d_out, int n, const int* n_dev, cudaStream_t st);
  // async on `stream`; (bw, bh) = board size of ALL n records when the caller knows it (0: 19x19 frame)
  int forward_device(int kind, const void* d_in, void* d_out, int n, int bw = 0, int bh = 0);
  // the board size shared by all n host records, or (0, 0)
  static void uniform_board(const maru_state* s, int n, int* bw, int* bh);
  // the same for fp32 rows (Board::getInputs layout): the board is read off plane 32, the on-board mask
  static void uniform_board_rows(const float* rows, int n, int* bw, int* bh);
  int forward_host(int kind, const void* in, void* out, int n);        // blocking, host buffers
  // leaf-queue pipeline: inputs already gathered in slot.h_*; enqueue H2D + forward + D2H (async)
  int enqueue_slot(int kind, int slot_idx, int n);
  int wait_slot(int slot_idx, float* device_ms);
  int encode_rows_host(const maru_state* s, float* out_rows, int n);
  // evaluator lanes (maru_lane_*): leaves staged by the caller in pinned memory, enqueued by the caller's thread
  Lane* lane_create(int rows);
  void lane_destroy(Lane* l);
  int lane_submit(Lane* l, int n);
  int lane_wait(Lane* l, float* device_ms);
  int lane_poll(Lane* l, int* done);
};

}  // namespace maru
