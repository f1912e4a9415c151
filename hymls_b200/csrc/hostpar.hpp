// Host-side parallel loops for the symbolic phase (Initialize): plain std::thread, static chunks, results
// independent of the thread count (every iteration writes to its own slots).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace hymls {

int hostThreads();           // threads used by parallelFor (default: hardware concurrency, at most 16)
void setHostThreads(int n);  // n <= 0 restores the default; ranks sharing a node divide the cores

// f(begin, end, thread) over [0, n) split into one contiguous chunk per thread; f must not throw
template <class F>
void parallelFor(int64_t n, F&& f, int64_t minPerThread = 2048) {
  int T = (int)std::min<int64_t>(hostThreads(), std::max<int64_t>(1, n / std::max<int64_t>(1, minPerThread)));
  if (T <= 1) {
    f((int64_t)0, n, 0);
    return;
  }
  const int64_t chunk = (n + T - 1) / T;
  std::vector<std::thread> th;
  th.reserve(T - 1);
  for (int t = 1; t < T; ++t) {
    const int64_t b = std::min(n, t * chunk), e = std::min(n, b + chunk);
    th.emplace_back([&f, b, e, t] { f(b, e, t); });
  }
  f((int64_t)0, std::min(n, chunk), 0);
  for (auto& x : th) x.join();
}

}  // namespace hymls
