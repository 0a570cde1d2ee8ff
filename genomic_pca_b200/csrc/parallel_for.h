// parallel_for.h -- split [0, n) over host threads (the per-SNP host arithmetic is independent per SNP).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

template <class F>
static inline void parallel_for(uint64_t n, F&& fn, uint64_t min_per_thread = 1u << 16) {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  uint64_t nt = std::min<uint64_t>(hw, (n + min_per_thread - 1) / min_per_thread);
  if (nt <= 1) {
    fn((uint64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  const uint64_t per = (n + nt - 1) / nt;
  for (uint64_t t = 0; t < nt; ++t) {
    const uint64_t lo = t * per, hi = std::min<uint64_t>(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([&fn, lo, hi] { fn(lo, hi); });
  }
  for (auto& x : th) x.join();
}
