// Randomised check of planHostPipe / checkHostPipe (hymls_b200/csrc/symbolic.cpp) under ASan + UBSan, CPU only:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Ihymls_b200/csrc tools/fuzz_host_pipeline.cpp \
//       hymls_b200/csrc/{symbolic,partitioner,params}.cpp -lpthread -o /tmp/fuzz && /tmp/fuzz
// Every plan the planner returns must pass the row-by-row checker; plans with one row too few copied in time or a
// shifted work-list boundary must be rejected.  (Last run: 2636 plans, no finding.)
#include <cstdio>
#include <random>
#include <numeric>
#include "symbolic.hpp"
#include "hostpar.hpp"
using namespace hymls;
int main() {
  std::mt19937_64 rng(123);
  int ok = 0, none = 0;
  for (int it = 0; it < 3000; ++it) {
    const int M = 1 + rng() % 60;
    std::vector<int> n(M), nb(M);
    std::vector<int64_t> vecOff(M);
    int64_t nI = 0;
    for (int m = 0; m < M; ++m) { n[m] = rng() % 70; nb[m] = n[m] ? rng() % (n[m] + 1) : 0; vecOff[m] = nI; nI += n[m]; }
    const int64_t nS = rng() % 50;
    const int64_t nRows = nI + nS;
    std::vector<int> rows(nRows);
    std::iota(rows.begin(), rows.end(), 0);
    // locally shuffled: subdomains roughly along the row order, as in the real maps
    const int win = 1 + rng() % 40;
    for (int64_t i = 0; i + 1 < nRows; ++i) { int64_t j = std::min<int64_t>(nRows - 1, i + rng() % win); std::swap(rows[i], rows[j]); }
    std::vector<int> intRow(rows.begin(), rows.begin() + nI);
    const int K = rng() % 20;
    const bool taper = rng() & 1;
    const int rpi = 1 + rng() % 40;
    setHostThreads(1 + rng() % 8);
    HostPipePlan P = planHostPipe(n, nb, vecOff, intRow, nRows, rpi, K, taper);
    if (P.K == 0) { ++none; continue; }
    if (!checkHostPipe(P, n, nb, vecOff, intRow, nRows, rpi)) { printf("FAIL valid plan rejected it=%d\n", it); return 1; }
    ++ok;
    // mutations must be caught (when they change something that matters)
    HostPipePlan Q = P;
    int c = 1 + rng() % (P.K - 1 > 0 ? P.K - 1 : 1);
    if (c < P.K && Q.inRows[c] > 0) {
      Q.inRows[c] = Q.inRows[c] - 1;  // one row fewer copied before chunk c-1
      // valid only if no matrix of chunks < c touches that row: recheck by brute force
      bool needed = false;
      for (int m = 0; m < Q.matStart[c]; ++m) for (int q = 0; q < n[m]; ++q) if (intRow[vecOff[m] + q] == Q.inRows[c]) needed = true;
      bool mono = Q.inRows[c] >= Q.inRows[c - 1];
      if ((needed || !mono) && checkHostPipe(Q, n, nb, vecOff, intRow, nRows, rpi)) { printf("FAIL bad inRows accepted it=%d\n", it); return 1; }
    }
    Q = P;
    if (c < P.K) {
      Q.fullItem[c] += 1;
      if (checkHostPipe(Q, n, nb, vecOff, intRow, nRows, rpi)) { printf("FAIL bad fullItem accepted it=%d\n", it); return 1; }
    }
  }
  printf("fuzz ok: %d plans checked, %d without a plan\n", ok, none);
  return 0;
}
