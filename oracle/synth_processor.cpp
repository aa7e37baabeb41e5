// synth_processor.cpp — TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
//
// deepgo::Processor as maru_b200/dropin/states/Processor.h declares it (the compact-record variant of
// the drop-in), backed by the synthetic evaluator of oracle/synth_eval.c instead of libmaru_b200.so:
// the reference's UNMODIFIED Node / Player / NodeManager / ThreadPool (+ the substitute Evaluator.cpp)
// then run on the CPU with a deterministic "network" (oracle/_ref/libmaru_ref_search_synth.so), which
// is what tests/test_search.py compares the native search against.
#include "Processor.h"

#include <chrono>
#include <stdexcept>
#include <thread>

extern "C" int oracle_synth_leaves(const maru_state* s, maru_leaf* out, int n, void* user);

namespace deepgo {

struct Processor::Api {};

Processor::Processor(std::string, std::vector<int32_t>, int32_t, bool, bool, int32_t)
    : _api(nullptr), _evaluators(), _queue(nullptr) {}
Processor::~Processor() {}
void Processor::execute(float*, float*, int32_t) {
  throw std::runtime_error("synthetic Processor: only executeLeaves is available");
}
void Processor::executeStates(const maru_state*, float*, int32_t) {
  throw std::runtime_error("synthetic Processor: only executeLeaves is available");
}
void Processor::executeLeaves(const maru_state* states, maru_leaf* leaves, int32_t size) {
  // A network answers in about a millisecond.  The reference relies on that: its waiter thread must see
  // "target reached" and stop the search before the running descent returns (Player.cpp:212-223, 282-313);
  // with an instantaneous evaluator its control thread often launches one descent too many.
  std::this_thread::sleep_for(std::chrono::microseconds(300));
  oracle_synth_leaves(states, leaves, size, nullptr);
}

}  // namespace deepgo
