// queue.cpp — leaf-batch queue: the B200-side replacement of deepgo::Processor / Executor /
// ExecutorJob (reference src/deepgo/native/cpp/Processor.cpp:17-60, Executor.cpp:52-189,
// ExecutorJob.cpp:8-36).
//
// Same observable contract as Processor::execute(float*, float*, int32_t): thread-safe, blocking,
// any thread, any job size; jobs are routed to the least-loaded evaluator (waiting + reserved rows,
// Processor.cpp:42-55) and coalesced into batches of up to `batch_size` rows (Executor.cpp:116).
// What changed:
//   * no per-batch heap allocation: jobs are gathered straight into PINNED staging (Executor.cpp:153
//     did `new float[]` twice per batch and copied from pageable memory synchronously);
//   * two staging slots per evaluator: the worker gathers and enqueues batch i+1 while the GPU still
//     runs batch i (the reference had one batch in flight and a fully synchronous Model::forward);
//   * compact maru_state jobs (448 B/position) next to fp32-plane jobs (47 680 B/position);
//   * errors are returned to every waiter as rc != 0 (the reference printed to stderr and let the
//     worker thread die with its callers blocked forever, Executor.cpp:133-137).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#ifdef MARU_TSAN_STUB_ENGINE
#include "engine_stub.h"   // `make tsan`: host-only stand-in, see tests/tsan_queue_selfplay.cpp
#else
#include "engine.h"
#endif

namespace maru {

struct Job {
  int kind;
  const void* in;
  void* out;
  int n;
  int rc = 0;
  std::string err;
  bool done = false;  // guarded by mu; notify happens under the lock so the waiter's stack
  std::mutex mu;      // frame (which owns this Job) cannot disappear mid-notify
  std::condition_variable cv;
  std::chrono::steady_clock::time_point t_enq;
};

struct Worker {
  Engine* e = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job*> q;
  int waiting = 0;   // rows queued
  int reserved = 0;  // rows routed here but not yet queued (Processor.cpp:55)
  bool stop = false;
};

}  // namespace maru

struct maru_queue {
  std::vector<std::unique_ptr<maru::Worker>> workers;
  std::mutex mu;  // dispatcher (Processor::_mutex)
  int batch_size = 0;
  std::atomic<int> coalesce_rows{0};     // idle worker: wait for this many queued rows ...
  std::atomic<int> coalesce_us{0};       // ... but not longer than this (maru_queue_set_coalesce)
  std::mutex stat_mu;
  maru_queue_stats stats{};
};

namespace maru {

static void finish_job(maru_queue* q, Job* j, int rc, const std::string& err) {
  const double ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - j->t_enq).count();
  {
    std::lock_guard<std::mutex> l(q->stat_mu);
    q->stats.jobs++;
    q->stats.positions += j->n;
    q->stats.wait_ms_total += ms;
  }
  std::lock_guard<std::mutex> l(j->mu);
  j->rc = rc;
  if (rc) j->err = err;
  j->done = true;
  j->cv.notify_all();
}

struct InFlight {
  std::vector<Job*> jobs;
  int slot = -1;
  int rows = 0;
};

static void complete(maru_queue* q, Worker* w, InFlight& f) {
  if (f.slot < 0) return;
  float ms = 0;
  int rc = w->e->wait_slot(f.slot, &ms);
  const std::string err = rc ? g_last_error : std::string();
  if (!rc) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(w->e->slot[f.slot].h_out);
    const size_t row = out_row_bytes(f.jobs.empty() ? KIND_STATES : f.jobs[0]->kind);
    for (Job* j : f.jobs) {
      std::memcpy(j->out, src, (size_t)j->n * row);
      src += (size_t)j->n * row;
    }
  }
  {
    std::lock_guard<std::mutex> l(q->stat_mu);
    q->stats.batches++;
    q->stats.device_ms_total += ms;
    if ((uint64_t)f.rows > q->stats.max_batch_seen) q->stats.max_batch_seen = f.rows;
  }
  for (Job* j : f.jobs) finish_job(q, j, rc, err);
  f.jobs.clear();
  f.slot = -1;
  f.rows = 0;
}

static void worker_main(maru_queue* q, Worker* w) {
  cudaSetDevice(w->e->gpu_id);
  const int cap = q->batch_size < w->e->max_batch ? q->batch_size : w->e->max_batch;
  InFlight inflight;
  int next_slot = 0;
  for (;;) {
    std::vector<Job*> jobs;
    int rows = 0;
    Job* big = nullptr;
    {
      std::unique_lock<std::mutex> l(w->mu);
      if (inflight.slot < 0) {
        w->cv.wait(l, [w] { return !w->q.empty() || w->stop; });
        const int want = q->coalesce_rows.load(std::memory_order_relaxed);
        if (want > 0 && !w->stop && w->waiting < (want < cap ? want : cap)) {
          const int lim = want < cap ? want : cap;
          w->cv.wait_for(l, std::chrono::microseconds(q->coalesce_us.load(std::memory_order_relaxed)),
                         [w, lim] { return w->waiting >= lim || w->stop; });
        }
      }
      if (w->stop && w->q.empty()) {
        l.unlock();
        complete(q, w, inflight);
        return;
      }
      // drain: same kind, rows <= cap (a job never splits; one larger than cap runs alone)
      while (!w->q.empty()) {
        Job* j = w->q.front();
        if (j->n > cap) {
          if (jobs.empty()) {
            big = j;
            w->q.pop_front();
            w->waiting -= j->n;
          }
          break;
        }
        if (!jobs.empty() && (j->kind != jobs[0]->kind || rows + j->n > cap)) break;
        w->q.pop_front();
        w->waiting -= j->n;
        jobs.push_back(j);
        rows += j->n;
      }
    }
    if (big != nullptr) {
      complete(q, w, inflight);
      const int rc = w->e->forward_host(big->kind, big->in, big->out, big->n);
      {
        std::lock_guard<std::mutex> l(q->stat_mu);
        q->stats.batches++;
        q->stats.device_ms_total += w->e->device_ms_last;
        if ((uint64_t)big->n > q->stats.max_batch_seen) q->stats.max_batch_seen = big->n;
      }
      finish_job(q, big, rc, rc ? g_last_error : std::string());
      continue;
    }
    if (jobs.empty()) {  // nothing new: retire the batch in flight, then block for work
      complete(q, w, inflight);
      continue;
    }
    // gather into the free pinned slot while the previous batch (other slot) is still on the GPU
    Slot& sl = w->e->slot[next_slot];
    const int kind = jobs[0]->kind;
    if (kind != KIND_PLANES) {
      maru_state* dst = sl.h_states;
      for (Job* j : jobs) {
        std::memcpy(dst, j->in, (size_t)j->n * sizeof(maru_state));
        dst += j->n;
      }
    } else {
      float* dst = sl.h_rows;
      for (Job* j : jobs) {
        std::memcpy(dst, j->in, (size_t)j->n * IN_ROW * sizeof(float));
        dst += (size_t)j->n * IN_ROW;
      }
    }
    const int rc = w->e->enqueue_slot(kind, next_slot, rows);
    if (rc) {
      const std::string err = g_last_error;
      for (Job* j : jobs) finish_job(q, j, rc, err);
      complete(q, w, inflight);
      continue;
    }
    complete(q, w, inflight);  // retire the older batch (its slot becomes free)
    inflight.jobs = std::move(jobs);
    inflight.slot = next_slot;
    inflight.rows = rows;
    next_slot ^= 1;
  }
}

// Processor.cpp:42-55: least waiting + reserved rows, then reserve; Executor.cpp:52-67: enqueue + notify
static int enqueue_job(maru_queue* q, Job* job) {
  const int n = job->n;
  Worker* w = nullptr;
  {
    std::lock_guard<std::mutex> l(q->mu);
    int best = -1, best_cnt = 0;
    for (size_t i = 0; i < q->workers.size(); i++) {
      Worker* c = q->workers[i].get();
      std::lock_guard<std::mutex> lw(c->mu);
      const int cnt = c->waiting + c->reserved;
      if (best < 0 || cnt < best_cnt) {
        best = (int)i;
        best_cnt = cnt;
      }
    }
    w = q->workers[best].get();
    std::lock_guard<std::mutex> lw(w->mu);
    w->reserved += n;
  }
  job->t_enq = std::chrono::steady_clock::now();
  {
    std::lock_guard<std::mutex> l(w->mu);
    if (w->stop) {
      w->reserved = w->reserved - n > 0 ? w->reserved - n : 0;   // undo the reservation of this job
      set_error("maru_queue is shutting down");
      return 1;
    }
    w->q.push_back(job);
    w->waiting += n;
    w->reserved = w->reserved - n > 0 ? w->reserved - n : 0;  // Executor.cpp:60
  }
  w->cv.notify_all();
  return 0;
}

static int queue_execute(maru_queue* q, int kind, const void* in, void* out, int n) {
  if (q == nullptr || in == nullptr || out == nullptr || n < 0) {
    set_error("maru_queue_execute: bad argument");
    return 1;
  }
  if (n == 0) return 0;
  Job job;
  job.kind = kind;
  job.in = in;
  job.out = out;
  job.n = n;
  if (enqueue_job(q, &job)) return 1;
  {
    std::unique_lock<std::mutex> l(job.mu);
    job.cv.wait(l, [&job] { return job.done; });
  }
  if (job.rc) set_error(job.err);
  return job.rc;
}

}  // namespace maru

extern "C" {

int maru_queue_create(maru_eval** evals, int n_evals, int batch_size, maru_queue** out) {
  if (evals == nullptr || n_evals <= 0 || out == nullptr || batch_size <= 0) {
    maru::set_error("maru_queue_create: bad argument");
    return 1;
  }
  auto* q = new maru_queue();
  q->batch_size = batch_size;
  for (int i = 0; i < n_evals; i++) {
    auto w = std::make_unique<maru::Worker>();
    w->e = reinterpret_cast<maru::Engine*>(evals[i]);
    q->workers.push_back(std::move(w));
  }
  for (auto& w : q->workers) w->th = std::thread(maru::worker_main, q, w.get());
  *out = q;
  return 0;
}

void maru_queue_destroy(maru_queue* q) {
  if (q == nullptr) return;
  for (auto& w : q->workers) {
    {
      std::lock_guard<std::mutex> l(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
  }
  for (auto& w : q->workers)
    if (w->th.joinable()) w->th.join();
  delete q;
}

int maru_queue_execute(maru_queue* q, const float* in, float* out, int n) {
  return maru::queue_execute(q, maru::KIND_PLANES, in, out, n);
}

int maru_queue_execute_states(maru_queue* q, const maru_state* s, float* out, int n) {
  return maru::queue_execute(q, maru::KIND_STATES, s, out, n);
}

int maru_queue_execute_leaves(maru_queue* q, const maru_state* s, maru_leaf* out, int n) {
  return maru::queue_execute(q, maru::KIND_LEAVES, s, out, n);
}

// ---- asynchronous submit / poll / wait: a ticket is a heap Job ---------------------------------
int maru_queue_submit_leaves(maru_queue* q, const maru_state* s, maru_leaf* out, int n, maru_ticket** t) {
  if (q == nullptr || s == nullptr || out == nullptr || t == nullptr || n <= 0) {
    maru::set_error("maru_queue_submit_leaves: bad argument");
    return 1;
  }
  *t = nullptr;
  auto* job = new (std::nothrow) maru::Job();
  if (job == nullptr) {
    maru::set_error("out of memory");
    return 1;
  }
  job->kind = maru::KIND_LEAVES;
  job->in = s;
  job->out = out;
  job->n = n;
  if (maru::enqueue_job(q, job)) {
    delete job;
    return 1;
  }
  *t = reinterpret_cast<maru_ticket*>(job);
  return 0;
}

int maru_queue_poll(maru_queue* q, maru_ticket* t, int* done) {
  if (q == nullptr || t == nullptr || done == nullptr) {
    maru::set_error("maru_queue_poll: bad argument");
    return 1;
  }
  auto* job = reinterpret_cast<maru::Job*>(t);
  std::lock_guard<std::mutex> l(job->mu);
  *done = job->done ? 1 : 0;
  return 0;
}

int maru_queue_wait(maru_queue* q, maru_ticket* t) {
  if (q == nullptr || t == nullptr) {
    maru::set_error("maru_queue_wait: bad argument");
    return 1;
  }
  auto* job = reinterpret_cast<maru::Job*>(t);
  {
    std::unique_lock<std::mutex> l(job->mu);
    job->cv.wait(l, [job] { return job->done; });
  }
  const int rc = job->rc;
  if (rc) maru::set_error(job->err);
  delete job;
  return rc;
}

int maru_queue_set_coalesce(maru_queue* q, int min_rows, int max_wait_us) {
  if (q == nullptr || min_rows < 0 || max_wait_us < 0) {
    maru::set_error("maru_queue_set_coalesce: bad argument");
    return 1;
  }
  q->coalesce_rows.store(min_rows);
  q->coalesce_us.store(max_wait_us);
  return 0;
}

int maru_queue_get_stats(maru_queue* q, maru_queue_stats* st) {
  if (q == nullptr || st == nullptr) return 1;
  std::lock_guard<std::mutex> l(q->stat_mu);
  *st = q->stats;
  return 0;
}

}  // extern "C"
