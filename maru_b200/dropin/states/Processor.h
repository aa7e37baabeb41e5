// Substitute for the reference's deepgo::Processor (reference src/deepgo/native/cpp/Processor.h:11-45)
// for the COMPACT-RECORD variant of the drop-in: same constructor and execute() as the reference, plus
// executeStates() / executeLeaves(); the substitute Evaluator.cpp uses the latter so that leaves cross
// PCIe as 448-byte maru_state records up (instead of 47 680-byte fp32 rows) and 1456-byte maru_leaf
// results down (instead of 8 676-byte rows).  Backed by the leaf-batch queue of
// libmaru_b200.so (maru_queue_*), so deepgo::Executor / ExecutorJob / Model are not needed at all.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

// standard headers the rest of the reference tree used to receive through <torch/torch.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <unordered_map>
#include <unordered_set>

#include "maru_b200.h"

namespace deepgo {

class Processor {
 public:
  // gpus: CUDA ordinals (a negative id throws: no CPU path); one evaluator per gpu x threadsParGpu
  Processor(std::string model, std::vector<int32_t> gpus, int32_t batchSize, bool fp16,
            bool deterministic, int32_t threadsParGpu);
  ~Processor();
  Processor(const Processor&) = delete;
  Processor& operator=(const Processor&) = delete;

  // reference contract (Processor.cpp:36-60): thread-safe, blocking, size x 11920 -> size x 2169
  void execute(float* inputs, float* outputs, int32_t size);
  // added: compact records in, same outputs
  void executeStates(const maru_state* states, float* outputs, int32_t size);
  // added: compact records in, compact results out (masked move priors + value, 1456 B per leaf)
  void executeLeaves(const maru_state* states, maru_leaf* leaves, int32_t size);

 private:
  struct Api;
  Api* _api;
  std::vector<void*> _evaluators;
  void* _queue;
};

}  // namespace deepgo
