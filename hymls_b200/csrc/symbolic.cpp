#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "symbolic.hpp"
#include "hostpar.hpp"

#include <algorithm>
#include <cmath>

#include "../../include/hymls_b200.h"

namespace hymls {

static int g_hostThreads = 0;
int hostThreads() {
  if (g_hostThreads > 0) return g_hostThreads;
  if (const char* e = getenv("HYMLS_B200_HOST_THREADS")) return std::max(1, atoi(e));
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(16u, hw ? hw : 1u));
}
void setHostThreads(int n) { g_hostThreads = n > 0 ? n : 0; }

static const double SMALL_ENTRY = 1e-14;  // HYMLS_SMALL_ENTRY, src/HYMLS_Macros.hpp:29

static inline int roundUp8(int n) { return (n + 7) & ~7; }
static inline double hsign(double x) { return (x < 0) ? -1.0 : (x > 0 ? 1.0 : 0.0); }  // Householder.cpp:15-18

struct SymTimer {
  bool on = getenv("HYMLS_B200_VERBOSE_SYM") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[hymls_b200 sym] %-40s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};
// Host version of the Tester::isDDcorrect count (src/HYMLS_Tester.cpp:253-455) for runs without a device (index maps
// only); with a device the count falls out of the A11 list construction (indexing.cu)
int64_t countInteriorCouplings(const LevelSym& L) {
  std::vector<int> sdOfInt(L.nI);
  for (int sd = 0; sd < L.nsd; ++sd)
    for (int64_t p = L.H.intPtr[sd]; p < L.H.intPtr[sd + 1]; ++p) sdOfInt[p] = sd;
  std::vector<int64_t> perThread(64, 0);
  parallelFor(L.nI, [&](int64_t p0, int64_t p1, int t) {
    int64_t ignored = 0;
    for (int64_t p = p0; p < p1; ++p) {
      const int r = L.intRow[p], sd = sdOfInt[p];
      for (int64_t e = L.rowptr[r]; e < L.rowptr[r + 1]; ++e) {
        const int cp = L.rowPos[L.colidx[e]];
        if (cp >= 0 && sdOfInt[cp] != sd) ++ignored;
      }
    }
    perThread[t & 63] += ignored;
  });
  int64_t total = 0;
  for (int64_t v : perThread) total += v;
  return total;
}

void buildLevelSym(LevelSym& L, const CartesianPartitioner& part, const std::vector<int>& gid2row) {
  SymTimer st;
  std::vector<char> present;
  if (L.level > 0) {
    present.assign(gid2row.size(), 0);
    for (gidx g : L.rowGid) present[g] = 1;
  }
  buildHierarchicalMap(part, present, L.H);
  const HierarchicalMap& H = L.H;
  L.nsd = H.nsd;
  L.nuniq = H.numUnique();
  L.nI = H.numInterior();
  L.nS = H.numSeparator();
  if (L.nI + L.nS != L.n)
    throw Error(HYMLS_B200_ERR_ARG, "partition does not cover the row map of level " + std::to_string(L.level) +
                                        " (interior " + std::to_string(L.nI) + " + separators " +
                                        std::to_string(L.nS) + " != " + std::to_string(L.n) + ")");
  st.lap("before: orderings");
  // ---- orderings ----
  L.intRow.resize(L.nI);
  L.sepRow.resize(L.nS);
  L.rowPos.assign(L.n, INT32_MIN);
  // Interior ordering inside a subdomain: the nodes that separator rows couple to (the columns of A21) come
  // first, in GID order, then the rest.  The first subdomain solve of ApplyInverse only feeds A21 x1, so it
  // needs just this leading block of rows of A11^-1 (engine.cu:applyLevel).
  {
    std::vector<char> isInt(L.n, 0), hit(L.n, 0);
    std::vector<int> bad(64, 0);
    parallelFor(L.nI, [&](int64_t p0, int64_t p1, int t) {
      for (int64_t p = p0; p < p1; ++p) {
        const int r = gid2row[H.intGid[p]];
        if (r < 0) bad[t & 63] = 1; else isInt[r] = 1;
      }
    });
    for (int b : bad)
      if (b) throw Error(HYMLS_B200_ERR_ARG, "interior node not in map / listed twice");
    // (several threads may store the same 1 into hit[]: relaxed atomic stores keep that well defined)
    parallelFor(L.nS, [&](int64_t p0, int64_t p1, int t) {
      for (int64_t p = p0; p < p1; ++p) {
        const int r = gid2row[H.sepGid[p]];
        if (r < 0) {
          bad[t & 63] = 1;
          continue;
        }
        for (int64_t e = L.rowptr[r]; e < L.rowptr[r + 1]; ++e)
          if (isInt[L.colidx[e]]) __atomic_store_n(&hit[L.colidx[e]], (char)1, __ATOMIC_RELAXED);
      }
    });
    for (int b : bad)
      if (b) throw Error(HYMLS_B200_ERR_ARG, "separator node not in map / listed twice");
    // ... and the interior rows that couple to separators (the rows of A12 with entries): A12 x2 is zero outside
    // them, so the correction A11^-1 (A12 x2) of the second solve needs the same leading COLUMNS only.  For a
    // structurally symmetric matrix both sets coincide.
    parallelFor(L.nI, [&](int64_t p0, int64_t p1, int) {
      for (int64_t p = p0; p < p1; ++p) {
        const int r = gid2row[H.intGid[p]];
        if (hit[r]) continue;
        for (int64_t e = L.rowptr[r]; e < L.rowptr[r + 1]; ++e)
          if (!isInt[L.colidx[e]]) {
            __atomic_store_n(&hit[r], (char)1, __ATOMIC_RELAXED);
            break;
          }
      }
    });
    L.sdNb.assign(H.nsd, 0);
    std::vector<double> sumPerThread(64, 0.0);
    parallelFor(H.nsd, [&](int64_t s0, int64_t s1, int t) {
      for (int64_t sd = s0; sd < s1; ++sd) {
        int64_t q = H.intPtr[sd];
        for (int pass = 0; pass < 2; ++pass)
          for (int64_t p = H.intPtr[sd]; p < H.intPtr[sd + 1]; ++p) {
            int r = gid2row[H.intGid[p]];
            if ((hit[r] != 0) == (pass == 0)) {
              L.intRow[q] = r;
              L.rowPos[r] = (int)q;
              ++q;
            }
          }
        for (int64_t p = H.intPtr[sd]; p < H.intPtr[sd + 1]; ++p) L.sdNb[sd] += hit[L.intRow[p]] ? 1 : 0;
      }
    }, 16);
    // (summed in subdomain order: the statistic must not depend on the thread count)
    L.sumNNb = 0;
    for (int sd = 0; sd < H.nsd; ++sd) L.sumNNb += (double)L.sdNb[sd] * (double)(H.intPtr[sd + 1] - H.intPtr[sd]);
  }
  {
    std::vector<int> bad(64, 0);
    parallelFor(L.nS, [&](int64_t p0, int64_t p1, int t) {
      for (int64_t p = p0; p < p1; ++p) {
        const int r = gid2row[H.sepGid[p]];
        if (r < 0 || L.rowPos[r] >= 0) {  // not in the map, or also an interior node
          bad[t & 63] = 1;
          continue;
        }
        L.sepRow[p] = r;
        L.rowPos[r] = -(int)p - 1;
      }
    });
    // interior + separator counts equal the map size (checked above), so a node listed twice leaves a row uncovered
    parallelFor(L.n, [&](int64_t r0, int64_t r1, int t) {
      for (int64_t r = r0; r < r1; ++r)
        if (L.rowPos[r] == INT32_MIN) bad[t & 63] = 1;
    });
    for (int b : bad)
      if (b) throw Error(HYMLS_B200_ERR_ARG, "separator node not in map / listed twice");
  }
  st.lap("before: A11 blocks");
  // ---- A11 blocks ----
  L.sdN.resize(L.nsd);
  L.sdNp.resize(L.nsd);
  L.a11Off.assign(L.nsd + 1, 0);
  L.sumNsq = 0;
  for (int sd = 0; sd < L.nsd; ++sd) {
    int n = (int)(H.intPtr[sd + 1] - H.intPtr[sd]);
    L.sdN[sd] = n;
    L.sdNp[sd] = roundUp8(n);
    L.a11Off[sd + 1] = L.a11Off[sd] + (int64_t)L.sdNp[sd] * L.sdNp[sd];
    L.sumNsq += (double)n * n;
  }
  // (the split of the matrix into A11 / A12 / A21 / A22 and the per-subdomain local pieces of the Schur assembly are
  // built on the device from the pattern: indexing.cu)
  st.lap("before: group instances, per-subdomain separator lists");
  // ---- group instances, per-subdomain separator lists ----
  L.sdM.resize(L.nsd);
  L.sdRowPtr.assign(L.nsd + 1, 0);
  L.sdInstPtr.assign(L.nsd + 1, 0);
  L.sdNumLink.resize(L.nsd);
  for (int sd = 0; sd < L.nsd; ++sd) {
    int m = 0;
    std::vector<int> types;
    for (int64_t g = H.sdGrpPtr[sd]; g < H.sdGrpPtr[sd + 1]; ++g) {
      int len = (int)(H.grpPtr[g + 1] - H.grpPtr[g]);
      int u = H.grpUnique[g];
      L.instLoc.push_back(m);
      L.instLen.push_back(len);
      L.instUniq.push_back(u);
      types.push_back(H.grpType[g]);
      for (int q = 0; q < len; ++q) L.sdSep.push_back((int)(H.uniqPtr[u] + q));
      m += len;
    }
    // linked sets as seen from this subdomain (hid_->GetLinkedSeparatorGroups(sd))
    std::vector<std::vector<int>> linked = linkGroups(types);
    std::vector<int> linkOf(types.size());
    for (size_t l = 0; l < linked.size(); ++l)
      for (int gi : linked[l]) linkOf[gi] = (int)l;
    for (int v : linkOf) L.instLink.push_back(v);
    L.sdNumLink[sd] = (int)linked.size();
    L.sdM[sd] = m;
    L.sdRowPtr[sd + 1] = L.sdRowPtr[sd] + m;
    L.sdInstPtr[sd + 1] = (int64_t)L.instLoc.size();
  }
  st.lap("before: blocks: per owner subdomain");
  // ---- blocks: per owner subdomain, linked sets among the groups it owns (InitializeBlocks :301-340) ----
  L.uniqBlk.assign(L.nuniq, -1);
  L.uniqBlkOff.assign(L.nuniq, 0);
  L.blkOff.assign(1, 0);
  L.blkRowPtr.assign(1, 0);
  L.sepBlk.assign(L.nS, -1);
  L.sepBlkIdx.assign(L.nS, -1);
  {
    int u = 0;
    while (u < L.nuniq) {
      int sd = H.uniqOwnerSd[u];
      int e = u;
      while (e < L.nuniq && H.uniqOwnerSd[e] == sd) ++e;  // unique groups of one owner are consecutive
      std::vector<int> types(H.uniqType.begin() + u, H.uniqType.begin() + e);
      for (auto& lg : linkGroups(types)) {
        int b = (int)L.blkN.size();
        int rows = 0;
        for (int gi : lg) {
          int uu = u + gi;
          L.uniqBlk[uu] = b;
          L.uniqBlkOff[uu] = rows;
          int64_t a = H.uniqPtr[uu], z = H.uniqPtr[uu + 1];
          for (int64_t p = a + 1; p < z; ++p) {
            L.sepBlk[p] = b;
            L.sepBlkIdx[p] = rows++;
            L.blkRows.push_back((int)p);
          }
        }
        L.blkN.push_back(rows);
        L.blkOwnerSd.push_back(sd);
        L.blkNp.push_back(roundUp8(rows));
        L.blkOff.push_back(L.blkOff.back() + (int64_t)roundUp8(rows) * roundUp8(rows));
        L.blkRowPtr.push_back((int64_t)L.blkRows.size());
      }
      u = e;
    }
  }
  L.nblk = (int)L.blkN.size();
  st.lap("before: reduced Schur pattern on the V-sums");
  // ---- reduced Schur pattern on the V-sums: union of per-subdomain cliques (:737-787) ----
  {
    // inverse index unique group -> the subdomains that see it, then one sorted union per group (parallel)
    std::vector<int64_t> occPtr(L.nuniq + 1, 0);
    for (int sd = 0; sd < L.nsd; ++sd)
      for (int64_t g = L.sdInstPtr[sd]; g < L.sdInstPtr[sd + 1]; ++g) occPtr[L.instUniq[g] + 1]++;
    for (int u = 0; u < L.nuniq; ++u) occPtr[u + 1] += occPtr[u];
    std::vector<int> occSd(occPtr[L.nuniq]);
    {
      std::vector<int64_t> f(occPtr.begin(), occPtr.end() - 1);
      for (int sd = 0; sd < L.nsd; ++sd)
        for (int64_t g = L.sdInstPtr[sd]; g < L.sdInstPtr[sd + 1]; ++g) occSd[f[L.instUniq[g]]++] = sd;
    }
    std::vector<std::vector<int>> rows(L.nuniq);
    parallelFor(L.nuniq, [&](int64_t u0, int64_t u1, int) {
      std::vector<int64_t> stamp(L.nuniq, -1);  // duplicates are dropped before sorting (a group is seen by up to
      std::vector<int> r;                        // 8 subdomains that share most of their groups)
      for (int64_t u = u0; u < u1; ++u) {
        r.clear();
        for (int64_t o = occPtr[u]; o < occPtr[u + 1]; ++o) {
          const int sd = occSd[o];
          for (int64_t h = L.sdInstPtr[sd]; h < L.sdInstPtr[sd + 1]; ++h) {
            const int v = L.instUniq[h];
            if (stamp[v] != u) {
              stamp[v] = u;
              r.push_back(v);
            }
          }
        }
        std::sort(r.begin(), r.end());
        rows[u] = r;
      }
    }, 256);
    L.redPtr.assign(L.nuniq + 1, 0);
    for (int u = 0; u < L.nuniq; ++u) L.redPtr[u + 1] = L.redPtr[u] + (int64_t)rows[u].size();
    L.redCol.resize(L.redPtr[L.nuniq]);
    parallelFor(L.nuniq, [&](int64_t u0, int64_t u1, int) {
      for (int64_t u = u0; u < u1; ++u) std::copy(rows[u].begin(), rows[u].end(), L.redCol.begin() + L.redPtr[u]);
    }, 256);
  }
  st.lap("before: Householder reflectors from the test vector");
  // ---- Householder reflectors from the test vector (InitializeOT :384-467, Householder::Construct) ----
  if ((int64_t)L.testVector.size() != L.n) L.testVector.assign(L.n, 1.0);
  L.what.assign(L.nS, 0.0);
  L.usign.assign(L.nuniq, 1.0);
  L.nextTestVector.assign(L.nuniq, 0.0);
  std::vector<double> v;
  for (int u = 0; u < L.nuniq; ++u) {
    int64_t a = H.uniqPtr[u], z = H.uniqPtr[u + 1];
    int len = (int)(z - a);
    v.resize(len);
    double nrm2 = 0;
    for (int q = 0; q < len; ++q) {
      v[q] = L.testVector[L.sepRow[a + q]];
      nrm2 += v[q] * v[q];
    }
    const double nrm = std::sqrt(nrm2);
    const double s = hsign(v[0]);
    for (int q = 0; q < len; ++q) v[q] *= s;
    // dense variant (Householder::Apply/ApplyR :38-126): identity when degenerate
    double nrmv2 = 0;
    for (int q = 0; q < len; ++q) nrmv2 += v[q] * v[q];
    const double nrmv = std::sqrt(nrmv2);
    const double v1 = v[0] + nrmv;
    const bool denseDegenerate = std::fabs(v1) < SMALL_ENTRY || nrmv < SMALL_ENTRY;
    L.usign[u] = denseDegenerate ? -1.0 : 1.0;
    // sparse variant (Householder::Construct :128-163): norm taken BEFORE the sign scaling
    v[0] += nrm;
    double n2 = 0;
    for (int q = 0; q < len; ++q) n2 += v[q] * v[q];
    n2 = std::sqrt(n2);
    double dotv = 0;
    if (n2 >= SMALL_ENTRY) {
      for (int q = 0; q < len; ++q) {
        L.what[a + q] = v[q] / n2;
        dotv += L.what[a + q] * L.testVector[L.sepRow[a + q]];
      }
    }
    // next level test vector = V-sum part of H*tv with H = 2ww'-I (ComputeNextLevel :569-573)
    L.nextTestVector[u] = 2.0 * L.what[a] * dotv - L.testVector[L.sepRow[a]];
  }
  st.lap("reflectors");
}

// ---------------------------------------------------------------------------------------------
// Copy / compute schedule of the host-buffer ApplyInverse (see HostPipePlan)
// ---------------------------------------------------------------------------------------------
HostPipePlan planHostPipe(const std::vector<int>& n, const std::vector<int>& nb, const std::vector<int64_t>& vecOff,
                          const std::vector<int>& intRow, int64_t nRows, int rowsPerItem, int K, bool taper) {
  HostPipePlan P;
  const int M = (int)n.size();
  if (M == 0 || K < 2 || rowsPerItem <= 0 || nRows <= 0) return P;
  // smallest / largest matrix row of every matrix, work = bytes of its inverse
  std::vector<int64_t> lo(M, nRows), hi(M, -1);
  std::vector<double> cum(M + 1, 0.0);
  parallelFor(M, [&](int64_t m0, int64_t m1, int) {
    for (int64_t m = m0; m < m1; ++m)
      for (int q = 0; q < n[m]; ++q) {
        const int64_t r = intRow[vecOff[m] + q];
        lo[m] = std::min(lo[m], r);
        hi[m] = std::max(hi[m], r);
      }
  }, 64);
  for (int m = 0; m < M; ++m) cum[m + 1] = cum[m] + (double)n[m] * (double)n[m];
  if (cum[M] <= 0.0) return P;
  // share of the work per chunk: equal, or tapered towards both ends (1 2 4 8 8 .. 8 4 2 1): the copy of b that the
  // FIRST chunk waits for and the copy of x that follows the LAST chunk are the parts that cannot be hidden
  std::vector<double> share(K, 1.0);
  if (taper)
    for (int c = 0; c < K; ++c) share[c] = (double)(1 << std::min(std::min(c, K - 1 - c), 3));
  double total = 0.0, acc = 0.0;
  for (double v : share) total += v;
  P.matStart.push_back(0);
  for (int c = 1; c < K; ++c) {
    acc += share[c - 1];
    const double target = cum[M] * acc / total;
    int m = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    m = std::min(m, M);
    if (m > P.matStart.back() && m < M) P.matStart.push_back(m);
  }
  P.matStart.push_back(M);
  P.K = (int)P.matStart.size() - 1;
  if (P.K < 2) return HostPipePlan();
  std::vector<int64_t> sufLo(M + 1, nRows);
  for (int m = M - 1; m >= 0; --m) sufLo[m] = std::min(sufLo[m + 1], lo[m]);
  P.leadItem.assign(P.K + 1, 0);
  P.fullItem.assign(P.K + 1, 0);
  P.inRows.assign(P.K + 1, 0);
  P.outRows.assign(P.K + 1, 0);
  for (int c = 0; c < P.K; ++c) {
    int lead = 0, full = 0;
    int64_t need = P.inRows[c];
    for (int m = P.matStart[c]; m < P.matStart[c + 1]; ++m) {
      lead += (nb[m] + rowsPerItem - 1) / rowsPerItem;
      full += (n[m] + rowsPerItem - 1) / rowsPerItem;
      need = std::max(need, hi[m] + 1);
    }
    P.leadItem[c + 1] = P.leadItem[c] + lead;
    P.fullItem[c + 1] = P.fullItem[c] + full;
    P.inRows[c + 1] = c + 1 == P.K ? nRows : need;
    P.outRows[c + 1] = c + 1 == P.K ? nRows : std::max(P.outRows[c], sufLo[P.matStart[c + 1]]);
  }
  return P;
}

bool checkHostPipe(const HostPipePlan& P, const std::vector<int>& n, const std::vector<int>& nb,
                   const std::vector<int64_t>& vecOff, const std::vector<int>& intRow, int64_t nRows,
                   int rowsPerItem) {
  const int M = (int)n.size(), K = P.K;
  if (K < 2 || (int)P.matStart.size() != K + 1 || (int)P.leadItem.size() != K + 1 ||
      (int)P.fullItem.size() != K + 1 || (int)P.inRows.size() != K + 1 || (int)P.outRows.size() != K + 1)
    return false;
  if (P.matStart[0] != 0 || P.matStart[K] != M || P.leadItem[0] != 0 || P.fullItem[0] != 0 || P.inRows[0] != 0 ||
      P.outRows[0] != 0 || P.inRows[K] != nRows || P.outRows[K] != nRows)
    return false;
  int64_t lead = 0, full = 0;
  int c = 0;
  std::vector<int> chunkOf(M, -1);
  for (int m = 0; m < M; ++m) {
    while (c < K && m >= P.matStart[c + 1]) {  // chunk boundary: the work lists must be cut exactly here
      if (P.matStart[c + 1] <= P.matStart[c] || P.leadItem[c + 1] != lead || P.fullItem[c + 1] != full ||
          P.inRows[c + 1] < P.inRows[c] || P.outRows[c + 1] < P.outRows[c])
        return false;
      ++c;
    }
    if (c >= K) return false;
    chunkOf[m] = c;
    for (int r0 = 0; r0 < nb[m]; r0 += rowsPerItem) ++lead;   // the loops of BatchedInverse::setup
    for (int r0 = 0; r0 < n[m]; r0 += rowsPerItem) ++full;
  }
  if (c != K - 1 || P.leadItem[K] != lead || P.fullItem[K] != full) return false;
  std::vector<char> bad(64, 0);
  parallelFor(M, [&](int64_t m0, int64_t m1, int t) {
    for (int64_t m = m0; m < m1; ++m) {
      const int cm = chunkOf[m];
      for (int q = 0; q < n[m]; ++q) {
        const int64_t r = intRow[vecOff[m] + q];
        // b[r] on the device before chunk cm of the first pass starts; x[r] not copied out before chunk cm of the
        // last pass has written it
        if (r < 0 || r >= nRows || r >= P.inRows[cm + 1] || r < P.outRows[cm]) bad[t & 63] = 1;
      }
    }
  }, 64);
  for (char f : bad)
    if (f) return false;
  return true;
}


}  // namespace hymls
