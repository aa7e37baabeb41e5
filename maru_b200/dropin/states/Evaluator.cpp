// Substitute for the reference's Evaluator.cpp (reference src/deepgo/native/cpp/Evaluator.cpp:37-84),
// compiled against the reference's UNCHANGED Evaluator.h.  Same observable behaviour:
//   * one network evaluation per node, cached by _evaluated                      (Evaluator.cpp:39, 83)
//   * candidate moves = legal (seki-checked) AND not fixed territory, with the prior read at the
//     centred model cell, no renormalisation, no pass entry                       (Evaluator.cpp:55-72)
//   * value = 2 p - 1 of output scalar 0, negated when White is to move           (Evaluator.cpp:75-80)
// What changes is the transport: instead of expanding 11 920 floats on the host (Board::getInputs)
// the position is captured as a 448-byte maru_state through the Board's public getters and the
// planes are built by the K1 kernel on the GPU.
#include "Evaluator.h"

#include "state_from_board.h"

namespace deepgo {

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
  maru_b200::state_from_board(board, color, _komi, _rule, _superko, /*with_move_mask=*/true, &state);
  float result[MODEL_OUTPUT_SIZE];
  _processor->executeStates(&state, result, 1);

  const int32_t w = state.width, h = state.height;
  const int32_t left = (MODEL_SIZE - w) / 2, top = (MODEL_SIZE - h) / 2;
  for (int32_t cell = 0; cell < w * h; cell++) {
    if ((state.move_mask[cell >> 3] >> (cell & 7)) & 1u) {
      const int32_t x = cell % w, y = cell / w;
      _policies.emplace_back(x, y, result[(top + y) * MODEL_SIZE + (left + x)], 0);
    }
  }
  const float win = result[MODEL_PREDICTIONS * MODEL_SIZE * MODEL_SIZE];
  _value = (color == WHITE) ? 1.0f - 2.0f * win : 2.0f * win - 1.0f;
  _evaluated = true;
}

}  // namespace deepgo
