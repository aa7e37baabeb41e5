// See Model.h. Errors keep the reference's convention (C++ exceptions out of the constructor and
// out of forward, Model.cpp:30): the C ABI's rc != 0 becomes std::runtime_error(maru_last_error()).
#include "Model.h"

#include <dlfcn.h>

#include <cstdlib>
#include <stdexcept>

namespace deepgo {

namespace {
template <class F>
F must_sym(void* lib, const char* name) {
  void* p = dlsym(lib, name);
  if (p == nullptr) throw std::runtime_error(std::string("libmaru_b200.so lacks ") + name);
  return reinterpret_cast<F>(p);
}
}  // namespace

Model::Model(std::string filename, int32_t gpu, bool /*fp16*/, bool /*deterministic*/)
    : _library(nullptr), _evaluator(nullptr), _forward(nullptr), _destroy(nullptr), _error(nullptr) {
  const char* path = std::getenv("MARU_B200_LIB");
  _library = dlopen(path != nullptr ? path : "libmaru_b200.so", RTLD_NOW | RTLD_GLOBAL);
  if (_library == nullptr) throw std::runtime_error(std::string("cannot load libmaru_b200.so: ") + dlerror());
  auto create = must_sym<int (*)(const char*, int, int, unsigned, void**)>(_library, "maru_eval_create");
  _forward = must_sym<int (*)(void*, const float*, float*, int)>(_library, "maru_eval_forward_planes");
  _destroy = must_sym<void (*)(void*)>(_library, "maru_eval_destroy");
  _error = must_sym<const char* (*)(void)>(_library, "maru_last_error");
  const char* mb = std::getenv("MARU_B200_MAX_BATCH");
  const int max_batch = mb != nullptr ? std::atoi(mb) : 2048;
  if (create(filename.c_str(), gpu, max_batch, 0u, &_evaluator) != 0) {
    throw std::runtime_error(_error());
  }
}

Model::~Model() {
  if (_evaluator != nullptr) _destroy(_evaluator);
}

void Model::forward(float* inputs, float* outputs, uint32_t size) {
  if (_forward(_evaluator, inputs, outputs, static_cast<int>(size)) != 0) {
    throw std::runtime_error(_error());
  }
}

int32_t Model::isCuda() { return 1; }

}  // namespace deepgo
