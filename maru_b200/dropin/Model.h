// Substitute for the reference's deepgo::Model (reference src/deepgo/native/cpp/Model.h:12-58).
//
// Drop this file and Model.cpp over the reference's own pair: Executor.cpp, Processor.cpp, the whole
// search (Node/Player) and the Cython layer then compile UNCHANGED and evaluate on the B200 kernels.
// The class keeps the exact public surface `Executor` uses (Executor.cpp:21 `new Model(model, gpu,
// fp16, deterministic)`, Executor.cpp:172 `_model->forward(in, out, n)`), but has no libtorch
// dependency: it dlopens libmaru_b200.so and speaks the C ABI of include/maru_b200.h.
#pragma once

#include <cstdint>
#include <string>

// The reference's Executor.cpp / Node.h lean on standard headers that used to arrive transitively
// through <torch/torch.h> (memcpy, std::unordered_map, ...). Keep them reachable from here so the
// rest of the tree compiles without edits.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace deepgo {

class Model {
 public:
  // filename: the TorchScript `.model` path the reference would load; the weight pack
  // "<filename>.b200" (python -m maru_b200.convert) is what is actually mapped to the device.
  // gpu < 0 (CPU) throws: this evaluator is B200-only. fp16 / deterministic are accepted for
  // signature compatibility (the kernels are bf16 tensor-core math, run-to-run deterministic).
  Model(std::string filename, int32_t gpu, bool fp16, bool deterministic);
  virtual ~Model();
  Model(const Model&) = delete;
  Model& operator=(const Model&) = delete;

  // inputs: size x 11920 fp32 host rows, outputs: size x 2169 fp32 host rows (Model.cpp:88-100)
  virtual void forward(float* inputs, float* outputs, uint32_t size);
  virtual int32_t isCuda();

 private:
  void* _library;
  void* _evaluator;
  int (*_forward)(void*, const float*, float*, int);
  void (*_destroy)(void*);
  const char* (*_error)(void);
};

}  // namespace deepgo
