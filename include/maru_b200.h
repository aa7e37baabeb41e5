/*
 * maru_b200.h — C ABI of the B200-native leaf evaluator for takedarts/maru.
 *
 * This is the drop-in boundary for ONE path of the reference: batched neural-network
 * evaluation of MCTS leaf positions (feature planes -> trunk -> heads -> policy/value).
 * Everything here is plain C: pointers, sizes, int return codes (0 = ok, !=0 = error,
 * text via maru_last_error()). No torch / C++ types cross this boundary.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   S1  deepgo::Model::Model(std::string, int32_t gpu, bool fp16, bool deterministic)
 *       deepgo::Model::forward(float* in, float* out, uint32_t n)
 *       deepgo::Model::isCuda()                     src/deepgo/native/cpp/Model.h:22-41
 *                                                   src/deepgo/native/cpp/Model.cpp:62-100
 *   S2  deepgo::Processor::execute(float*, float*, int32_t)
 *                                                   src/deepgo/native/cpp/Processor.h:33
 *                                                   src/deepgo/native/cpp/Processor.cpp:36-60
 *       deepgo::Executor::{execute,_run,_forward}   src/deepgo/native/cpp/Executor.cpp:52-189
 *   S3  deepgo::Board::getInputs(float*, color, komi, rule, superko)
 *                                                   src/deepgo/native/cpp/Board.cpp:494-623
 *       deepgo::Evaluator::evaluate(Board*, color)  src/deepgo/native/cpp/Evaluator.cpp:37-84
 *   tensor contract                                 src/deepgo/config.py:37-50
 *
 * Row layouts (identical to the reference):
 *   input  row: float32[11920] = planes [33][19][19] (plane, y, x) + 7 scalars
 *   output row: float32[2169]  = planes [6][19][19] + 3 scalars
 */
#ifndef MARU_B200_H
#define MARU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MARU_MODEL_SIZE 19            /* config.py:37  MODEL_SIZE        */
#define MARU_MODEL_FEATURES 32        /* config.py:39  MODEL_FEATURES    */
#define MARU_MODEL_INFOS 7            /* config.py:41  MODEL_INFOS       */
#define MARU_MODEL_PREDICTIONS 6      /* config.py:43  MODEL_PREDICTIONS */
#define MARU_MODEL_VALUES 3           /* config.py:45  MODEL_VALUES      */
#define MARU_INPUT_SIZE 11920         /* config.py:48  (32+1)*361 + 7    */
#define MARU_OUTPUT_SIZE 2169         /* config.py:50  6*361 + 3         */

#define MARU_BLACK 1
#define MARU_WHITE (-1)
#define MARU_RULE_CH 0
#define MARU_RULE_JP 1
#define MARU_RULE_COM 2

/* cell byte of maru_state::cells */
#define MARU_CELL_EMPTY 0u
#define MARU_CELL_BLACK 1u
#define MARU_CELL_WHITE 2u
#define MARU_CELL_COLOR_MASK 3u
#define MARU_CELL_LADDER 4u           /* Ren::shicho of the stone's group (Board.cpp:523)        */
#define MARU_CELL_LIB_SHIFT 3         /* bits 3..6: min(#liberties, 8) of the group (Board.cpp:524) */
#define MARU_NONE 0xFFu               /* "no move / no ko" coordinate                              */

#define MARU_STATE_HAS_MASK 1u        /* maru_state::flags: move_mask is filled                    */

/*
 * Compact board record: everything Board::getInputs reads, 448 bytes instead of 47 680.
 * Filled from the reference Board's PUBLIC getters (getWidth/getHeight/getColor/getRenSpace/
 * isShicho/getHistories/getKo, Board.h:47-116) — see maru_b200/dropin/state_from_board.h.
 */
typedef struct maru_state {
  uint8_t cells[368];      /* y*width+x over the width x height board; [361..367] zero          */
  uint8_t width;           /* 1..19                                                              */
  uint8_t height;          /* 1..19                                                              */
  int8_t  color;           /* side to move: +1 black, -1 white                                   */
  uint8_t rule;            /* MARU_RULE_*                                                        */
  uint8_t superko;         /* 0/1                                                                */
  uint8_t ko_x, ko_y;      /* Board::getKo(color) (Board.cpp:213-219); MARU_NONE if none          */
  uint8_t flags;           /* MARU_STATE_*                                                       */
  uint8_t hist[2][3][2];   /* [0]=black,[1]=white; [i]: i-th newest non-pass move (x,y); MARU_NONE */
  float   komi;
  uint8_t reserved[8];
  uint8_t move_mask[48];   /* optional: bit (y*width+x) = legal && territory==EMPTY (Evaluator.cpp:60-68) */
} maru_state;              /* sizeof == 448 */

/*
 * Compact evaluation result: what Evaluator::evaluate keeps of the 2169-float output row
 * (Evaluator.cpp:55-80) — the move prior of every candidate cell and the value scalars; 1456 bytes
 * instead of 8676. prior[y*width + x] = output plane 0 at the centred model cell
 * ((19-h)/2 + y)*19 + (19-w)/2 + x (Evaluator.cpp:57-69) when the record's move_mask allows the
 * cell (or the record carries no mask), else exactly 0. No renormalisation (Evaluator.cpp:63-72).
 */
typedef struct maru_leaf {
  float prior[361];        /* board order y*width + x; entries >= width*height are 0               */
  float value;             /* output scalar 0: win probability of the side to move (row[2166])     */
  float score[2];          /* output scalars 1, 2 (row[2167], row[2168])                           */
} maru_leaf;               /* sizeof == 1456 */

typedef struct maru_eval maru_eval;     /* one evaluator = one GPU, one stream, one weight replica */
typedef struct maru_queue maru_queue;   /* leaf-batch queue over >=1 evaluators (Processor mirror)  */

/* flags for maru_eval_create */
#define MARU_FLAG_NO_GRAPH 1u           /* launch kernels directly instead of CUDA graphs           */

typedef struct maru_eval_info {
  int32_t blocks, channels, attn_every, heads, head_channels, value_channels;
  int32_t n_attn;                       /* number of attention layers                               */
  int32_t gpu_id, max_batch, sm_count;
  int32_t launches_per_forward;         /* kernels launched by one forward_states of <= max_batch   */
  int32_t reserved0;
  double  flops_per_eval;               /* algorithmic FLOPs per 19x19 position (SURVEY 8d formula)  */
  double  flops_per_eval_executed;      /* FLOPs the kernels execute per position (padding included) */
  uint64_t n_forward, n_positions;      /* counters since create                                    */
  double  device_ms_last;               /* CUDA-event time of the last forward (kernels only)       */
} maru_eval_info;

/* ---- lifetime (replaces Model::Model, Model.cpp:62-80) -------------------------------------- */
/* weights_path: a weight pack written by `python -m maru_b200.convert <file>.model`; if the path
 * does not end in ".b200" the sidecar "<path>.b200" is opened. gpu_id >= 0 (there is no CPU path:
 * gpu_id < 0 is an error). */
int  maru_eval_create(const char* weights_path, int gpu_id, int max_batch, unsigned flags,
                      maru_eval** out);
void maru_eval_destroy(maru_eval* e);
int  maru_eval_get_info(maru_eval* e, maru_eval_info* info);

/* ---- blocking forwards on HOST buffers (replace Model::forward, Model.cpp:88-100) ------------ */
/* in: [n][11920] fp32 rows exactly as Board::getInputs writes them; out: [n][2169] fp32. */
int  maru_eval_forward_planes(maru_eval* e, const float* in, float* out, int n);
/* s: [n] compact records; encoding happens on the GPU (K1). out: [n][2169] fp32. */
int  maru_eval_forward_states(maru_eval* e, const maru_state* s, float* out, int n);

/* s: [n] compact records in, [n] compact results out: the search's own transport (448 B up, 1456 B
 * down per leaf instead of 47 680 / 8 676). Policy gather and masking happen in the heads kernel. */
int  maru_eval_forward_leaves(maru_eval* e, const maru_state* s, maru_leaf* out, int n);

/* ---- device-resident forwards (inputs/outputs already in HBM; used by bench `value`) -------- */
int  maru_eval_forward_states_device(maru_eval* e, const maru_state* d_states, float* d_out, int n);
int  maru_eval_forward_planes_device(maru_eval* e, const float* d_in, float* d_out, int n);
int  maru_eval_forward_leaves_device(maru_eval* e, const maru_state* d_states, maru_leaf* d_out, int n);
/* Small boards: a batch whose records all hold the same w x h board (< 19x19) is evaluated on a COMPACT frame after
 * the stem — (w + 1)(h + 1) rows per position instead of the reference's 400 (it computes all 19x19 cells of the model
 * frame for every board size, Board.cpp:496-497) — with results equal to the 19x19-frame evaluation up to fp32
 * summation order.  The host entry points (forward_states / forward_leaves, the queue) read the sizes from the records;
 * the *_device entry points cannot, so the caller states them here (0, 0 = unknown / mixed: the 19x19 frame).
 * MARU_B200_COMPACT=0 or debug bit 21 disables the compact frame. */
int  maru_eval_set_device_board(maru_eval* e, int width, int height);
int  maru_eval_sync(maru_eval* e);
void* maru_eval_stream(maru_eval* e);   /* the evaluator's cudaStream_t (for CUDA-event timing)      */

/* ---- K1 alone: compact record -> reference fp32 rows (bit-exact Board::getInputs) ----------- */
int  maru_encode_planes(maru_eval* e, const maru_state* s, float* out_rows, int n);         /* host */
int  maru_encode_planes_device(maru_eval* e, const maru_state* d_s, float* d_rows, int n);  /* HBM  */

/* ---- debugging / parity: activations of the last forward ------------------------------------- */
/* Copies `channels` channels of an activation buffer of the last forward as fp32 [n][channels][361]
 * (19x19 cells, halo rows dropped). which: 0 trunk input x0 (64 ch), 1 h (C), 2 r (C/2), 3 s (C/2),
 * 4 qkv (3C), 5 attention output o (C), 6 head features g (Ch). */
#define MARU_BUF_X0 0
#define MARU_BUF_H 1
#define MARU_BUF_R 2
#define MARU_BUF_S 3
#define MARU_BUF_QKV 4
#define MARU_BUF_O 5
#define MARU_BUF_G 6
int  maru_eval_debug_planes(maru_eval* e, int which, float* out, int n, int channels);
/* Pre-softmax policy logits of the last forward: [n][2][361] fp32 (north_star tolerance is on logits). */
int  maru_eval_debug_policy_logits(maru_eval* e, float* out, int n);
/* Debug knobs. key 0: stop the launch sequence after `value` trunk launches (<0: run everything;
 * the heads then do not run). key 1 (bit mask): bit0 = disable programmatic dependent launch (A/B timing),
 * bit4 = conv_gemm instead of the K-split (3x3, 128 input channels) and wide (1x1, 256 -> k*256) kernels (A/B),
 * bit5 = DISABLE the persistent multi-layer flow kernels (conv_flow.cu): one PDL launch per conv layer
 * (A/B timing and the layer-by-layer parity tests), bit6 = record per-layer stamps of the flow kernels
 * (maru_eval_debug_stamps), bit7 = use the flow kernels at every batch size (default: inside 249..354 positions on
 * 148 SMs when the engine's own timing says they win), bit20 = conv_ksplit instead of the CTA-pair kernel conv_pair (A/B), bit21 = always the 19x19 frame (no
 * compact frame for small boards, A/B and parity), bit23 = attention always subtracts the row shift c_i before the
 * exponential (default: a warp skips it when its rows' scores are bounded by 64; A/B and parity of the shifted path).
 * Only in -DMARU_EXPERIMENTS builds of the library (not the product build): bit2 = the v1 mma.sync attention
 * kernel, bit3 = per-tile cross-layer flags in conv_gemm (measured slower), bits 8-17 = timing experiments inside
 * the flow kernels (tools/flow_ab.py; results are WRONG with any of them set). */
#define MARU_DBG_STOP_AFTER 0
#define MARU_DBG_DESC_BITS 1
#define MARU_DBG_TRACE_LAUNCH 2       /* index of the trunk launch whose pipeline is traced (-1: off) */
#define MARU_DBG_TIMELINE 3           /* 1: every conv launch stamps %globaltimer (see debug_timeline)  */
#define MARU_DBG_FLOW_TRACE_CTA 4     /* >= 0: that CTA of a flow kernel stamps every item (debug_trace, [256][8]) */
#define MARU_DBG_FLOW_TRACE_LAUNCH 5  /* which flow launch of the forward is traced (default 0)          */
#define MARU_DBG_ATT_TRACE_CTA 6      /* EXPERIMENTS builds: >= 0: that CTA of the first attention launch stamps clock64 per step
                                         (debug_trace, [64 steps][8]: softmax group {S seen, S read, P written}, control warp
                                         {P seen, PV issued, QK issued}) */
int  maru_eval_debug_set(maru_eval* e, int key, int value);
/* [64 launches][first CTA, last CTA][8] ns stamps: entry, prologue done, dependency resolved, first
 * operands landed, first accumulator ready, exit. */
int  maru_eval_debug_timeline(maru_eval* e, long long* out);
/* Persistent flow kernels (debug key 1 bit6 set): [128 layers][4] ns stamps of CTA 0 — producer enters
 * the layer, weights landed, epilogue done with the layer — chains back to back, each followed by one
 * [end, 0, 0, 0] row. */
int  maru_eval_debug_stamps(maru_eval* e, long long* out);
/* clock64 stamps of the traced conv launch: [n_ctas][32 tiles][8] = producer {wait,issue}, MMA warp
 * {wait start, operands ready, issued}, epilogue warp {wait start, accumulator ready, done}. */
int  maru_eval_debug_trace(maru_eval* e, long long* out, int n_ctas);

/* ---- per-launch timing (roofline evidence) --------------------------------------------------- */
/* Per-launch timing of the forward of n positions (device-resident compact records): after one
 * plain forward, every launch of the sequence is issued 10 times back to back inside one CUDA event
 * pair on the evaluator's stream; ms = that span / 10 (event overhead amortised, launch gaps kept).
 * flops / bytes are the ALGORITHMIC figures of that launch at T = 361 tokens per position
 * (DESIGN.md "Kernels"); kind: 0 encode, 1 conv3x3, 2 conv1x1, 3 attention core, 4 heads. */
typedef struct maru_launch_time {
  char   name[32];
  int32_t kind, cin, cout, taps;
  float  ms;
  double flops;
  double bytes;
} maru_launch_time;
int  maru_eval_profile_states(maru_eval* e, const maru_state* d_states, float* d_out, int n,
                              maru_launch_time* out, int max_launches, int* n_launches);

/* ---- leaf-batch queue (replaces Processor/Executor/ExecutorJob, Processor.cpp, Executor.cpp) -- */
/* Takes ownership of nothing: evaluators must outlive the queue. batch_size = drain cap
 * (Executor.cpp:116). One worker thread + pinned double buffers per evaluator. */
int  maru_queue_create(maru_eval** evals, int n_evals, int batch_size, maru_queue** out);
void maru_queue_destroy(maru_queue* q);
/* Thread-safe, blocking; same contract as Processor::execute(float*, float*, int32_t). */
int  maru_queue_execute(maru_queue* q, const float* in, float* out, int n);
/* Same, compact records in (the path the substitute Evaluator uses). */
int  maru_queue_execute_states(maru_queue* q, const maru_state* s, float* out, int n);
/* Compact records in, compact results out (what the substitute Evaluator.cpp calls). */
int  maru_queue_execute_leaves(maru_queue* q, const maru_state* s, maru_leaf* out, int n);

typedef struct maru_queue_stats {
  uint64_t jobs, positions, batches;
  uint64_t max_batch_seen;
  double   wait_ms_total;               /* sum over jobs of enqueue -> completion                   */
  double   device_ms_total;             /* sum over batches of CUDA-event forward time              */
} maru_queue_stats;
int  maru_queue_get_stats(maru_queue* q, maru_queue_stats* st);

/* Asynchronous variant (SURVEY 8b "async variant submit/poll for the new leaf queue"): the reference
 * parks every search thread on a per-job mutex + condition variable inside Processor::execute
 * (ExecutorJob.cpp:8-36); here a caller hands over a whole batch of leaves, keeps working, and
 * collects the result later.  `s` and `out` must stay valid until the ticket is retired.
 * submit: routed and batched exactly like maru_queue_execute_leaves (least-loaded evaluator, jobs never
 * split).  poll: *done = 1 once the results are in `out` (non-blocking; the ticket stays valid).
 * wait: blocks until done, returns the job's rc and RETIRES the ticket (every ticket must be waited on
 * exactly once). */
typedef struct maru_ticket maru_ticket;
int  maru_queue_submit_leaves(maru_queue* q, const maru_state* s, maru_leaf* out, int n, maru_ticket** t);
int  maru_queue_poll(maru_queue* q, maru_ticket* t, int* done);
int  maru_queue_wait(maru_queue* q, maru_ticket* t);
/* Batch forming when the evaluator is idle: the reference (and the default here) launches whatever is
 * queued the moment one job arrives (Executor.cpp:108-124).  With min_rows > 0 an idle worker waits up
 * to max_wait_us for at least min_rows queued rows first, so many small submitters still form batches
 * the tensor-core kernels are efficient at.  min_rows = 0 restores the default. */
int  maru_queue_set_coalesce(maru_queue* q, int min_rows, int max_wait_us);

/* ---- host-side helper: exact ladder flags without the reference's Board copies ------------------ */
/* colors: [height*width] raster, 0 empty / +1 black / -1 white. (ko_x, ko_y, ko_color): the point the
 * side ko_color may not retake (Board::_koIndex/_koColor, Board.cpp:140-205), or ko_x < 0 for none.
 * flags[y*width+x] = 1 for every stone Board::isShicho(x, y) reports (Board::_updateShicho /
 * _isShichoRen, Board.cpp:1045-1153) — same decision procedure on a flat board, ~20x cheaper
 * (maru_b200/dropin/lean_ladder.h). Pure host code, no GPU involved. */
int  maru_ladder_flags(const int8_t* colors, int width, int height, int ko_x, int ko_y, int ko_color,
                       uint8_t* flags);

/* Candidate-move filter of Evaluator::evaluate on the same flat board (Evaluator.cpp:60-68):
 * enableds[y*width+x] = Board::getEnableds(color, checkSeki=true) (Board.cpp:333-343, 1177-1631),
 * territories[y*width+x] = Board::getTerritories(color) in {-1, 0, +1} (Board.cpp:350-378, 872-1040);
 * a move is kept iff enableds != 0 && territories == 0 (maru_state::move_mask). Either output may be NULL. */
int  maru_rules_eval(const int8_t* colors, int width, int height, int ko_x, int ko_y, int ko_color,
                     int color, uint8_t* enableds, int8_t* territories);

/* A whole position on the flat board (maru_b200/dropin/lean_game.h): Board::play semantics (captures, ko,
 * pass = off-board coordinates, 3-move history per colour, Board.cpp:140-205) and the compact record built
 * without a reference Board — byte-identical to state_from_board() over the reference Board. Host code. */
typedef struct maru_game maru_game;
maru_game* maru_game_create(int width, int height);
maru_game* maru_game_clone(const maru_game* g);
void maru_game_destroy(maru_game* g);
int  maru_game_play(maru_game* g, int x, int y, int color);      /* captured stones, 0 = pass, -1 = illegal */
int  maru_game_state(const maru_game* g, int color, float komi, int rule, int superko, int with_move_mask,
                     maru_state* out);

/* ---- evaluator lanes: leaves staged by the caller, no queue thread ------------------------------ */
/* A lane is the pinned staging of ONE in-flight batch of an evaluator: the caller writes up to max_rows records
 * straight into maru_lane_states(), submits (H2D + forward + D2H are enqueued on the evaluator's stream, the call
 * returns), keeps working, and reads maru_lane_leaves() after maru_lane_wait().  Any number of lanes may be in flight
 * on one evaluator (they run in submission order); the call is thread-safe across lanes, a lane itself belongs to one
 * thread at a time.  Compared with the queue (whose worker thread gathers jobs into its own staging and scatters the
 * results back: two copies and two thread hand-offs per batch, Executor.cpp:144-189 in the reference) nothing is copied
 * on the host and no other thread has to run — what a search thread wants when the search threads own every core. */
typedef struct maru_lane maru_lane;
int  maru_eval_lane_create(maru_eval* e, int max_rows, maru_lane** out);
void maru_lane_destroy(maru_lane* l);
maru_state* maru_lane_states(maru_lane* l);         /* pinned host memory, [max_rows]                         */
const maru_leaf* maru_lane_leaves(maru_lane* l);    /* pinned host memory, [max_rows]; valid after wait        */
int  maru_lane_submit(maru_lane* l, int n);          /* asynchronous; one batch in flight per lane              */
int  maru_lane_poll(maru_lane* l, int* done);
int  maru_lane_wait(maru_lane* l);                   /* blocks (the thread sleeps) until the lane's batch is done */

/* ---- native search + self-play game scheduler (rows a10 / f1) -------------------------------- */
/* Many concurrent games per host thread over the flat board: the reference's tree policy
 * (Node::evaluate, Node.cpp:71-225; Player::_evaluate / play / getCandidates, Player.cpp:93-115,
 * 229-261, 319-368) as a resumable state machine per game (maru_b200/csrc/search.h) — one descent in
 * flight per game (the reference's threads = 1), leaves of all games of a worker submitted as ONE
 * asynchronous batch (maru_queue_submit_leaves), two batches per worker in flight so host work and GPU
 * overlap.  Move decision = the reference's Python driver (player.py:295-366: search, candidates,
 * LCB / visits criterion, pass when nothing is legal), optionally with a Gumbel sequential-halving
 * root over the reference's equally / width / noise hooks (Node.cpp:99-140, 191-210).
 * `visits` is the number of descents made for every move ON TOP of the subtree the new root
 * inherited (target = inherited root visits + visits; Player.cpp:189, 214-219). */
#define MARU_ROOT_PUCB 0                /* the reference's default search (run.py / gtp.py genmove)   */
#define MARU_ROOT_GUMBEL 1              /* sequential halving: widths m, m/2, .., 1; budget split evenly */
#define MARU_CRITERION_LCB 0            /* player.py:92-93, 349-352: win-chance lower confidence bound */
#define MARU_CRITERION_VISITS 1
typedef struct maru_selfplay_config {
  int32_t width, height;                /* board size, 1..19                                          */
  float   komi;
  int32_t rule, superko;                /* MARU_RULE_*; superko only travels in the record            */
  int32_t games;                        /* concurrent games of THIS scheduler (one GPU's shard)       */
  int32_t threads;                      /* host worker threads; each owns games/threads games         */
  int32_t visits;                       /* descents per move                                          */
  int32_t max_moves;                    /* a game stops after this many moves (or two passes)         */
  int32_t root;                         /* MARU_ROOT_*                                                */
  int32_t root_width;                   /* Gumbel: candidates of the first phase                      */
  float   noise;                        /* Gumbel(0, noise) on the root log-priors (Node.cpp:105-118) */
  float   temperature;                  /* root temperature (Node.cpp:100-114)                        */
  int32_t criterion;                    /* MARU_CRITERION_*                                           */
  int32_t eval_leaf_only;               /* Player::_evalLeafOnly                                      */
  int32_t opening_moves;                /* seeded random legal moves played before the first search   */
  int32_t restart;                      /* 1: a finished game starts over (steady load for benchmarks) */
  uint64_t seed;
  int32_t cpu_first, cpu_count;         /* cpu_count > 0: pin worker i to core cpu_first + i % cpu_count */
  int32_t lanes;                        /* batches in flight per worker (0 -> 2)                      */
  int32_t record;                       /* 1: keep every move's candidate list (maru_selfplay_get_candidates) */
  int32_t use_ucb1;                     /* PUCB root only: UCB1 instead of PUCB at the root (run.py --search ucb1, Node.cpp:211-212) */
} maru_selfplay_config;
typedef struct maru_selfplay_stats {
  uint64_t visits;                      /* root descents (Player.cpp:300)                              */
  uint64_t evals;                       /* leaves evaluated by the network                             */
  uint64_t expansions;                  /* nodes visited twice: their candidate filter (seki-aware legality +
                                           territory, Evaluator.cpp:60-68) ran; the other leaves never needed it */
  uint64_t moves, games_finished;
  uint64_t submits;                     /* asynchronous batches handed to the queue                    */
  uint64_t max_submit;                  /* largest of them                                             */
  double   run_s;                       /* wall time inside maru_selfplay_run                          */
  double   host_s;                      /* summed over workers: tree + record building                 */
  double   stall_s;                     /* summed over workers: blocked on a ticket                    */
  double   device_s;                    /* evaluator lanes only: CUDA-event time of the batches (GPU busy time) */
} maru_selfplay_stats;
typedef struct maru_candidate {         /* deepgo::Candidate (player.py:65-93)                         */
  int32_t x, y, color, visits, playouts;
  float   prior, value;
} maru_candidate;
typedef struct maru_selfplay maru_selfplay;
/* Any evaluator of leaves: [n] records in, [n] results out, rc 0.  The product binding is the leaf-batch
 * queue (maru_selfplay_create); a caller-supplied function lets the SAME search run over another
 * evaluator (tests drive it and the reference Player with one synthetic evaluator on the CPU). */
typedef int (*maru_leaf_fn)(const maru_state* s, maru_leaf* out, int n, void* user);
int  maru_selfplay_create(maru_queue* q, const maru_selfplay_config* cfg, maru_selfplay** out);
int  maru_selfplay_create_fn(maru_leaf_fn fn, void* user, const maru_selfplay_config* cfg, maru_selfplay** out);
/* The same scheduler straight on ONE evaluator through lanes (above): every search thread stages its leaves in pinned
 * memory and submits them itself; there is no queue thread to schedule, so all of a rank's cores can search. */
int  maru_selfplay_create_eval(maru_eval* e, const maru_selfplay_config* cfg, maru_selfplay** out);
void maru_selfplay_destroy(maru_selfplay* sp);
/* Force a move in one game (Player::play): openings, replays.  Not while maru_selfplay_run is active. */
int  maru_selfplay_play(maru_selfplay* sp, int game, int x, int y);
/* Every unfinished game plays `moves` more moves (search + play); blocks until all are done. */
int  maru_selfplay_run(maru_selfplay* sp, int moves);
int  maru_selfplay_get_stats(maru_selfplay* sp, maru_selfplay_stats* st);
/* Moves of a game so far as (x, y) int16 pairs ((-1, -1) = pass); returns the count, writes <= cap. */
int  maru_selfplay_get_moves(maru_selfplay* sp, int game, int16_t* xy, int cap);
/* Candidate list the move `move` of `game` was chosen from (cfg.record = 1); returns the count. */
int  maru_selfplay_get_candidates(maru_selfplay* sp, int game, int move, maru_candidate* out, int cap);

/* ---- errors ---------------------------------------------------------------------------------- */
const char* maru_last_error(void);      /* thread-local text of the last failing call              */
int  maru_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MARU_B200_H */
