// See Processor.h.  rc != 0 from the C ABI becomes std::runtime_error (the reference's convention for
// constructor failures, Model.cpp:30; unlike the reference, runtime failures also reach the caller).
#include "Processor.h"

#include <dlfcn.h>

#include <cstdlib>
#include <stdexcept>

namespace deepgo {

struct Processor::Api {
  void* lib = nullptr;
  int (*eval_create)(const char*, int, int, unsigned, void**) = nullptr;
  void (*eval_destroy)(void*) = nullptr;
  int (*queue_create)(void**, int, int, void**) = nullptr;
  void (*queue_destroy)(void*) = nullptr;
  int (*queue_execute)(void*, const float*, float*, int) = nullptr;
  int (*queue_execute_states)(void*, const maru_state*, float*, int) = nullptr;
  int (*queue_execute_leaves)(void*, const maru_state*, maru_leaf*, int) = nullptr;
  const char* (*last_error)(void) = nullptr;

  template <class F>
  void bind(F& fn, const char* name) {
    fn = reinterpret_cast<F>(dlsym(lib, name));
    if (fn == nullptr) throw std::runtime_error(std::string("libmaru_b200.so lacks ") + name);
  }
  Api() {
    const char* path = std::getenv("MARU_B200_LIB");
    lib = dlopen(path != nullptr ? path : "libmaru_b200.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib == nullptr) throw std::runtime_error(std::string("cannot load libmaru_b200.so: ") + dlerror());
    bind(eval_create, "maru_eval_create");
    bind(eval_destroy, "maru_eval_destroy");
    bind(queue_create, "maru_queue_create");
    bind(queue_destroy, "maru_queue_destroy");
    bind(queue_execute, "maru_queue_execute");
    bind(queue_execute_states, "maru_queue_execute_states");
    bind(queue_execute_leaves, "maru_queue_execute_leaves");
    bind(last_error, "maru_last_error");
  }
};

Processor::Processor(std::string model, std::vector<int32_t> gpus, int32_t batchSize, bool /*fp16*/,
                     bool /*deterministic*/, int32_t threadsParGpu)
    : _api(new Api()), _evaluators(), _queue(nullptr) {
  const int max_batch = batchSize < 2048 ? batchSize : 2048;
  for (int32_t gpu : gpus) {
    for (int32_t i = 0; i < threadsParGpu; i++) {
      void* e = nullptr;
      if (_api->eval_create(model.c_str(), gpu, max_batch, 0u, &e) != 0) {
        const std::string msg = _api->last_error();
        for (void* p : _evaluators) _api->eval_destroy(p);
        delete _api;
        throw std::runtime_error(msg);
      }
      _evaluators.push_back(e);
    }
  }
  if (_evaluators.empty() ||
      _api->queue_create(_evaluators.data(), (int)_evaluators.size(), batchSize, &_queue) != 0) {
    const std::string msg = _evaluators.empty() ? "Processor: no GPU given" : _api->last_error();
    for (void* p : _evaluators) _api->eval_destroy(p);
    delete _api;
    throw std::runtime_error(msg);
  }
}

Processor::~Processor() {
  if (_queue != nullptr) _api->queue_destroy(_queue);
  for (void* p : _evaluators) _api->eval_destroy(p);
  delete _api;
}

void Processor::execute(float* inputs, float* outputs, int32_t size) {
  if (_api->queue_execute(_queue, inputs, outputs, size) != 0) throw std::runtime_error(_api->last_error());
}

void Processor::executeStates(const maru_state* states, float* outputs, int32_t size) {
  if (_api->queue_execute_states(_queue, states, outputs, size) != 0)
    throw std::runtime_error(_api->last_error());
}

void Processor::executeLeaves(const maru_state* states, maru_leaf* leaves, int32_t size) {
  if (_api->queue_execute_leaves(_queue, states, leaves, size) != 0)
    throw std::runtime_error(_api->last_error());
}

}  // namespace deepgo
