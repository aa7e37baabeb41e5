// Substitute for the reference's Evaluator.cpp (reference src/deepgo/native/cpp/Evaluator.cpp:37-84),
// compiled against the reference's UNCHANGED Evaluator.h.  Same observable behaviour:
//   * one network evaluation per node, cached by _evaluated                      (Evaluator.cpp:39, 83)
//   * candidate moves = legal (seki-checked) AND not fixed territory, with the prior read at the
//     centred model cell, no renormalisation, no pass entry                       (Evaluator.cpp:55-72)
//   * value = 2 p - 1 of output scalar 0, negated when White is to move           (Evaluator.cpp:75-80)
// What changes is the transport: instead of expanding 11 920 floats on the host (Board::getInputs)
// the position is captured as a 448-byte maru_state through the Board's public getters and the
// planes are built by the K1 kernel on the GPU; the result comes back as a 1456-byte maru_leaf
// (priors of the candidate cells + value) instead of the 2169-float row.
#include "Evaluator.h"

#include <cstdlib>

#include "state_from_board.h"

namespace deepgo {

// MARU_LEAN_LADDER=0 falls back to Board::isShicho (A/B measurements); the default is the flat-board
// read-out of lean_ladder.h, bit-identical and 20-45x cheaper.
static bool lean_ladder_enabled() {
  static const bool on = [] {
    const char* v = std::getenv("MARU_LEAN_LADDER");
    return v == nullptr || v[0] != '0';
  }();
  return on;
}

Evaluator::Evaluator(Processor* processor, float komi, int32_t rule, bool superko)
    : _processor(processor), _komi(komi), _rule(rule), _superko(superko), _policies(), _value(0.0f),
      _evaluated(false) {}

void Evaluator::clear() {
  _policies.clear();
  _value = 0.0f;
  _evaluated = false;
}

bool Evaluator::isEvaluated() { return _evaluated; }

std::vector<Policy> Evaluator::getPolicies() { return _policies; }

float Evaluator::getValue() { return _value; }

void Evaluator::evaluate(Board* board, int32_t color) {
  if (_evaluated) return;

  maru_state state;
  maru_b200::state_from_board(board, color, _komi, _rule, _superko, /*with_move_mask=*/true, &state,
                              lean_ladder_enabled());
  // The legal / non-territory filter (Evaluator.cpp:60-68) does not depend on the network, so it is
  // taken BEFORE the evaluation and travels in the record; the heads kernel gathers the priors of
  // exactly those cells (board order) next to the value scalars: 1456 B back instead of 8676 B.
  maru_leaf leaf;
  _processor->executeLeaves(&state, &leaf, 1);

  const int32_t w = state.width, h = state.height;
  for (int32_t cell = 0; cell < w * h; cell++) {
    if ((state.move_mask[cell >> 3] >> (cell & 7)) & 1u) {
      _policies.emplace_back(cell % w, cell / w, leaf.prior[cell], 0);
    }
  }
  const float win = leaf.value;
  _value = (color == WHITE) ? 1.0f - 2.0f * win : 2.0f * win - 1.0f;
  _evaluated = true;
}

}  // namespace deepgo
