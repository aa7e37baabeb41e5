// ref_shim.cpp — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// A thin extern "C" surface over the UNMODIFIED reference classes, compiled by oracle/Makefile
// straight from /root/reference/src/deepgo/native/cpp/*.cpp into oracle/_ref/*.so.
// It exists so that tests / bench `--impl reference` can drive the reference's own
// Board::getInputs, Processor::execute, Evaluator::evaluate and Player search from ctypes
// without Cython or cmake. No reference source is copied into this repository.
//
// Built twice:
//   libmaru_ref_board.so  (-DREF_BOARD_ONLY) : Board/Ren/History/Pattern only, no libtorch
//   libmaru_ref_nn.so                        : + Model/Executor/Processor/Evaluator/Node/Player (libtorch)
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "Board.h"
#include "Config.h"
#include "maru_b200.h"
#include "state_from_board.h"

#ifndef REF_BOARD_ONLY
#include "Evaluator.h"
#include "Player.h"
#include "Processor.h"
#endif

using deepgo::Board;

extern "C" {

// ---------------------------------------------------------------- Board (Board.h:17-175)
void* ref_board_new(int w, int h) { return new Board(w, h); }
void ref_board_free(void* b) { delete (Board*)b; }
void ref_board_clear(void* b) { ((Board*)b)->clear(); }
int ref_board_play(void* b, int x, int y, int color) { return ((Board*)b)->play(x, y, color); }
void ref_board_copy_from(void* dst, void* src) { ((Board*)dst)->copyFrom((Board*)src); }
int ref_board_get_color(void* b, int x, int y) { return ((Board*)b)->getColor(x, y); }
int ref_board_get_ren_space(void* b, int x, int y) { return ((Board*)b)->getRenSpace(x, y); }
int ref_board_is_shicho(void* b, int x, int y) { return ((Board*)b)->isShicho(x, y) ? 1 : 0; }
int ref_board_is_enabled(void* b, int x, int y, int color, int check_seki) {
  return ((Board*)b)->isEnabled(x, y, color, check_seki != 0) ? 1 : 0;
}
void ref_board_get_enableds(void* b, int32_t* out, int color, int check_seki) {
  ((Board*)b)->getEnableds(out, color, check_seki != 0);
}
void ref_board_get_territories(void* b, int32_t* out, int color) {
  ((Board*)b)->getTerritories(out, color);
}
void ref_board_get_ko(void* b, int color, int32_t* xy) {
  auto ko = ((Board*)b)->getKo(color);
  xy[0] = ko.first;
  xy[1] = ko.second;
}
// Board::getInputs, Board.cpp:494-623 — THE feature oracle.
void ref_board_get_inputs(void* b, float* out, int color, float komi, int rule, int superko) {
  ((Board*)b)->getInputs(out, color, komi, rule, superko != 0);
}
// Compact record through the public getters only (product header state_from_board.h).
void ref_board_get_state(void* b, maru_state* st, int color, float komi, int rule, int superko,
                         int with_mask) {
  maru_b200::state_from_board((Board*)b, color, komi, rule, superko != 0, with_mask != 0, st);
}
// Same record with the ladder flags taken from the flat-board read-out (lean_ladder.h) instead of
// Board::isShicho: the two must be byte-identical (tests/test_lean_ladder.py).
void ref_board_get_state_lean(void* b, maru_state* st, int color, float komi, int rule, int superko,
                              int with_mask) {
  maru_b200::state_from_board((Board*)b, color, komi, rule, superko != 0, with_mask != 0, st, true);
}
int ref_model_input_size() { return MODEL_INPUT_SIZE; }
int ref_model_output_size() { return MODEL_OUTPUT_SIZE; }

#ifndef REF_BOARD_ONLY
// ---------------------------------------------------------------- Processor (Processor.h:11-45)
void* ref_processor_new(const char* model, const int32_t* gpus, int n_gpus, int batch_size,
                        int fp16, int deterministic, int threads_par_gpu, char* err, int err_len) {
  try {
    std::vector<int32_t> g(gpus, gpus + n_gpus);
    return new deepgo::Processor(std::string(model), g, batch_size, fp16 != 0, deterministic != 0,
                                 threads_par_gpu);
  } catch (std::exception& e) {
    if (err && err_len > 0) {
      std::strncpy(err, e.what(), err_len - 1);
      err[err_len - 1] = 0;
    }
    return nullptr;
  }
}
void ref_processor_free(void* p) { delete (deepgo::Processor*)p; }
void ref_processor_execute(void* p, float* in, float* out, int n) {
  ((deepgo::Processor*)p)->execute(in, out, n);
}

// ---------------------------------------------------------------- Evaluator (Evaluator.cpp:37-84)
// Returns the number of policies; xs/ys/ps sized >= 361.
int ref_evaluator_evaluate(void* processor, void* board, int color, float komi, int rule,
                           int superko, int32_t* xs, int32_t* ys, float* ps, float* value) {
  deepgo::Evaluator ev((deepgo::Processor*)processor, komi, rule, superko != 0);
  ev.evaluate((Board*)board, color);
  auto pol = ev.getPolicies();
  for (size_t i = 0; i < pol.size(); i++) {
    xs[i] = pol[i].x;
    ys[i] = pol[i].y;
    ps[i] = pol[i].policy;
  }
  *value = ev.getValue();
  return (int)pol.size();
}

// ---------------------------------------------------------------- Player (Player.h:23-99)
void* ref_player_new(void* processor, int threads, int w, int h, float komi, int rule, int superko,
                     int eval_leaf_only) {
  return new deepgo::Player((deepgo::Processor*)processor, threads, w, h, komi, rule, superko != 0,
                            eval_leaf_only != 0);
}
void ref_player_free(void* p) { delete (deepgo::Player*)p; }
void ref_player_initialize(void* p) { ((deepgo::Player*)p)->initialize(); }
int ref_player_play(void* p, int x, int y) { return ((deepgo::Player*)p)->play(x, y); }
int ref_player_get_color(void* p) { return ((deepgo::Player*)p)->getColor(); }
// One search: startEvaluation + waitEvaluation(visits) + getCandidates (player.py:295-366).
// Writes up to max_c candidates as (x, y, visits, playouts, color) int32 and (policy, value) float.
int ref_player_evaluate(void* pp, int visits, int playouts, float timelimit, int equally,
                        int use_ucb1, int width, float temperature, float noise, int32_t* ints,
                        float* floats, int max_c) {
  auto* p = (deepgo::Player*)pp;
  p->startEvaluation(equally != 0, use_ucb1 != 0, width, temperature, noise);
  p->waitEvaluation(visits, playouts, timelimit, true);
  auto cands = p->getCandidates();
  int n = 0;
  for (auto& c : cands) {
    if (n >= max_c) break;
    ints[n * 5 + 0] = c.getX();
    ints[n * 5 + 1] = c.getY();
    ints[n * 5 + 2] = c.getVisits();
    ints[n * 5 + 3] = c.getPlayouts();
    ints[n * 5 + 4] = c.getColor();
    floats[n * 2 + 0] = c.getPolicy();
    floats[n * 2 + 1] = c.getValue();
    n++;
  }
  return n;
}
#endif  // REF_BOARD_ONLY

}  // extern "C"
