// parallel_for.h -- split [0, n) over host threads (the per-SNP host arithmetic is independent per SNP).
//
// The threads belong to a persistent pool owned by the context (HostPool): spawning and joining std::threads per
// call cost ~0.3 ms with 32 cores, four times per 128 MB ingest chunk (whose transfer takes 2.4 ms), and with one
// process per GPU every rank spawned ALL cores' worth of threads (8 ranks x 32 threads on 32 cores).  The pool's size
// is the context's host-thread cap (gpca_set_host_threads / GPCA_HOST_THREADS; default: the CPUs the process may run
// on).  A pool is used by one caller at a time (calls on a context are serialised, gpca.h).
#pragma once
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

static inline unsigned host_cpu_count() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int n = CPU_COUNT(&set);
    if (n > 0) return (unsigned)n;
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return hw ? hw : 4;
}

class HostPool {
 public:
  explicit HostPool(unsigned nthreads) : n_(nthreads < 1 ? 1 : nthreads) {
    for (unsigned t = 1; t < n_; ++t) workers_.emplace_back([this] { worker(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;
  unsigned size() const { return n_; }

  // fn(lo, hi) over [0, n) in pieces of `piece` elements handed out through an atomic counter; the caller works too
  void run(uint64_t n, uint64_t piece, unsigned max_threads, const std::function<void(uint64_t, uint64_t)>& fn) {
    if (n == 0) return;
    if (piece == 0) piece = 1;
    const uint64_t n_pieces = (n + piece - 1) / piece;
    const unsigned helpers = (unsigned)std::min<uint64_t>(std::min<uint64_t>(n_ - 1, max_threads ? max_threads - 1 : 0),
                                                          n_pieces - 1);
    if (helpers == 0) {
      fn(0, n);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      total_ = n;
      piece_ = piece;
      next_.store(0, std::memory_order_relaxed);
      wanted_ = helpers;
      active_ = helpers;
      ++gen_;
    }
    cv_.notify_all();
    drain(fn, n, piece);
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [this] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  void drain(const std::function<void(uint64_t, uint64_t)>& fn, uint64_t n, uint64_t piece) {
    for (;;) {
      const uint64_t lo = next_.fetch_add(piece, std::memory_order_relaxed);
      if (lo >= n) break;
      fn(lo, std::min<uint64_t>(n, lo + piece));
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(uint64_t, uint64_t)>* fn = nullptr;
      uint64_t n = 0, piece = 1;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (wanted_ == 0) continue;      // enough helpers have taken this job already
        --wanted_;
        fn = fn_;
        n = total_;
        piece = piece_;
      }
      drain(*fn, n, piece);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--active_ == 0) done_cv_.notify_one();
      }
    }
  }

  unsigned n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(uint64_t, uint64_t)>* fn_ = nullptr;
  uint64_t total_ = 0, piece_ = 1, gen_ = 0;
  std::atomic<uint64_t> next_{0};
  unsigned wanted_ = 0, active_ = 0;
  bool stop_ = false;
};

// The pool parallel_for uses on this thread: the entry points of the library install their context's pool
// (HostPoolScope); code that runs without a context (tests of the host helpers) falls back to the calling thread.
inline HostPool*& current_host_pool() {
  static thread_local HostPool* p = nullptr;
  return p;
}
struct HostPoolScope {
  HostPool* prev;
  explicit HostPoolScope(HostPool* p) : prev(current_host_pool()) { current_host_pool() = p; }
  ~HostPoolScope() { current_host_pool() = prev; }
};

template <class F>
static inline void parallel_for(uint64_t n, F&& fn, uint64_t min_per_thread = 1u << 16) {
  HostPool* pool = current_host_pool();
  if (!pool || pool->size() <= 1 || n <= min_per_thread) {
    if (n) fn((uint64_t)0, n);
    return;
  }
  // pieces: at least min_per_thread elements, about four per thread so that a slow core does not hold the call up
  const uint64_t nt = std::min<uint64_t>(pool->size(), (n + min_per_thread - 1) / min_per_thread);
  uint64_t piece = std::max<uint64_t>(min_per_thread, (n + nt * 4 - 1) / (nt * 4));
  const std::function<void(uint64_t, uint64_t)> f = [&fn](uint64_t lo, uint64_t hi) { fn(lo, hi); };
  pool->run(n, piece, (unsigned)nt, f);
}
