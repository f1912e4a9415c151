// Owner-computes multi-GPU path of level 0 (one process per GPU, NCCL over NVLink / NVSwitch).
//
// The reference distributes subdomains over MPI ranks (BasePartitioner::CreatePIDMap) and moves only what
// Epetra's Import/Export move: the off-rank columns of A21 x1 / A12 x2 (MatrixBlock::Apply,
// src/HYMLS_MatrixBlock.cpp:294-308), the V-sums to the next level (src/HYMLS_SchurPreconditioner.cpp:1076-1078)
// and the Krylov dot products (SumAll).  Here:
//   * a rank owns the interiors of its subdomains and the separator groups whose owner subdomain is its own
//     (HierarchicalMap: the first subdomain listing a group owns it, src/HYMLS_HierarchicalMap.cpp:261-271);
//   * rhs2 = b2 - A21 x1: every rank multiplies with the columns of its interiors; partial sums at separators
//     owned by a neighbour go to the owner with one grouped ncclSend/ncclRecv ("rev" halo) and are added there
//     in rank order (deterministic);
//   * Householder transform and separator-block solves run on the owned groups only; the V-sums (a vector of
//     nuniq doubles, 0.76 MB at 128^3) are summed to every rank for the next level, which stays replicated in /
//     replicated out because its vectors are 100x smaller;
//   * x2 of the owned separators goes to the ranks whose subdomains touch them ("fwd" halo) for A12 x2;
//   * the result is distributed by row owner.  The Krylov basis lives in this distribution, the operator apply
//     exchanges the matrix halo ("mat"), so the GMRES loop contains no full-vector collective at all.
// Vectors keep GLOBAL length and indexing (only owned + halo entries are touched), so every index array of the
// single-GPU path stays valid; kernels get row / group lists.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <random>

#include "engine.hpp"
#include "hostpar.hpp"

namespace hymls {

bool Engine::useDist() const {
  static const bool off = getenv("HYMLS_B200_DIST") && atoi(getenv("HYMLS_B200_DIST")) == 0;
  return !off && comm_.active() && comm_.size() > 1 && borderM_ == 0 && !levels_.empty() && levels_[0]->sharded &&
         !levels_[0]->exact;
}

static void uploadHalo(Halo& h, const std::vector<std::vector<int>>& send, const std::vector<std::vector<int>>& recv,
                       cudaStream_t s, bool device = true) {
  const int P = (int)send.size();
  h.peers.clear();
  h.sendPtr.assign(1, 0);
  h.recvPtr.assign(1, 0);
  std::vector<int> si, ri;
  for (int q = 0; q < P; ++q) {
    if (send[q].empty() && recv[q].empty()) continue;
    h.peers.push_back(q);
    si.insert(si.end(), send[q].begin(), send[q].end());
    ri.insert(ri.end(), recv[q].begin(), recv[q].end());
    h.sendPtr.push_back((int64_t)si.size());
    h.recvPtr.push_back((int64_t)ri.size());
  }
  if (!device) return;
  h.sendIdx.upload(si, s);
  h.recvIdx.upload(ri, s);
  h.sendBuf.alloc(std::max<size_t>(si.size(), 1));
  h.recvBuf.alloc(std::max<size_t>(ri.size(), 1));
}

static void sortUnique(std::vector<int>& v) {
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
}

void Engine::buildDistPlan(Level& L) {
  const LevelSym& S = L.sym;
  DistPlan& D = L.dist;
  cudaStream_t s = stream_;
  const int P = comm_.size(), me = comm_.rank();
  // owner of every separator position
  std::vector<int> sepOwner(S.nS, 0), uniqRank(S.nuniq, 0);
  std::vector<int> ownUniq, ownSepPos;
  for (int u = 0; u < S.nuniq; ++u) {
    uniqRank[u] = L.sdRank[S.H.uniqOwnerSd[u]];
    for (int64_t p = S.H.uniqPtr[u]; p < S.H.uniqPtr[u + 1]; ++p) sepOwner[p] = uniqRank[u];
    if (uniqRank[u] == me) {
      ownUniq.push_back(u);
      for (int64_t p = S.H.uniqPtr[u]; p < S.H.uniqPtr[u + 1]; ++p) ownSepPos.push_back((int)p);
    }
  }
  // separators touched by the subdomains of every rank: mine give the row list of the A21 product and the ghosts,
  // the others tell which of my separators they will send partial sums for
  std::vector<std::vector<int>> sendRev(P), recvRev(P);
  std::vector<int> rows21;
  for (int sd = 0; sd < S.nsd; ++sd) {
    const int q = L.sdRank[sd];
    for (int64_t R = S.sdRowPtr[sd]; R < S.sdRowPtr[sd + 1]; ++R) {
      const int p = S.sdSep[R];
      if (q == me) {
        rows21.push_back(p);
        if (sepOwner[p] != me) sendRev[sepOwner[p]].push_back(p);
      } else if (sepOwner[p] == me) {
        recvRev[q].push_back(p);
      }
    }
  }
  sortUnique(rows21);
  for (int q = 0; q < P; ++q) {
    sortUnique(sendRev[q]);
    sortUnique(recvRev[q]);
  }
  const bool dev = deviceOk_;  // without a device only the host side of the plan exists (owned_rows, tests)
  uploadHalo(D.rev, sendRev, recvRev, s, dev);
  uploadHalo(D.fwd, recvRev, sendRev, s, dev);  // the same lists, opposite direction
  // deterministic accumulation of the received partial sums: per owned node its sources in rank order
  {
    std::vector<std::pair<int, int64_t>> ent;  // (node, index in rev.recvBuf)
    int64_t off = 0;
    for (int q = 0; q < P; ++q) {
      for (size_t i = 0; i < recvRev[q].size(); ++i) ent.emplace_back(recvRev[q][i], off + (int64_t)i);
      off += (int64_t)recvRev[q].size();
    }
    std::stable_sort(ent.begin(), ent.end(),
                     [](const std::pair<int, int64_t>& a, const std::pair<int, int64_t>& b) { return a.first < b.first; });
    std::vector<int> node;
    std::vector<int64_t> ptr(1, 0), src;
    for (size_t i = 0; i < ent.size(); ++i) {
      if (node.empty() || node.back() != ent[i].first) {
        if (!node.empty()) ptr.push_back((int64_t)src.size());
        node.push_back(ent[i].first);
      }
      src.push_back(ent[i].second);
    }
    if (!node.empty()) ptr.push_back((int64_t)src.size());
    D.nAdd = (int64_t)node.size();
    if (dev) {
      D.addNode.upload(node, s);
      D.addPtr.upload(ptr, s);
      D.addSrc.upload(src, s);
    }
  }
  std::vector<int> rows12;
  for (int sd : L.ownSd)
    for (int64_t p = S.H.intPtr[sd]; p < S.H.intPtr[sd + 1]; ++p) rows12.push_back((int)p);
  // row ownership in matrix numbering
  D.rowOwner.assign(S.n, 0);
  for (int sd = 0; sd < S.nsd; ++sd)
    for (int64_t p = S.H.intPtr[sd]; p < S.H.intPtr[sd + 1]; ++p) D.rowOwner[S.intRow[p]] = L.sdRank[sd];
  for (int64_t p = 0; p < S.nS; ++p) D.rowOwner[S.sepRow[p]] = sepOwner[p];
  std::vector<int64_t> cnt(P, 0);
  for (int64_t r = 0; r < S.n; ++r) cnt[D.rowOwner[r]]++;
  D.maxOwn = *std::max_element(cnt.begin(), cnt.end());
  D.nOwn = cnt[me];
  std::vector<int> allRows((size_t)P * D.maxOwn, -1);
  std::vector<int64_t> fill(P, 0);
  D.hOwnRows.clear();
  for (int64_t r = 0; r < S.n; ++r) {
    const int q = D.rowOwner[r];
    allRows[(size_t)q * D.maxOwn + fill[q]++] = (int)r;
    if (q == me) D.hOwnRows.push_back((int)r);
  }
  D.nRows21 = (int64_t)rows21.size();
  D.nRows12 = (int64_t)rows12.size();
  D.nOwnUniq = (int64_t)ownUniq.size();
  D.nOwnSep = (int64_t)ownSepPos.size();
  if (dev) {
    D.rows21.upload(rows21, s);
    D.rows12.upload(rows12, s);
    D.ownUniq.upload(ownUniq, s);
    D.ownSepPos.upload(ownSepPos, s);
    D.ownRows.upload(D.hOwnRows, s);
    D.allRows.upload(allRows, s);
    D.gath.alloc((size_t)P * D.maxOwn);
    HY_CUDA(cudaStreamSynchronize(s));
  }
  D.ready = true;
  D.matReady = false;
  buildMatrixHalo();
}

// columns of my rows owned elsewhere (receive) and my rows that appear as columns of other ranks' rows (send).
// Part of Initialize (threaded over the rows): a rank that built this inside the first solve would make the others
// wait in their first collective.
void Engine::buildMatrixHalo() {
  Level& L = *levels_[0];
  DistPlan& D = L.dist;
  if (D.matReady) return;
  const int P = comm_.size(), me = comm_.rank();
  const int T = 64;
  std::vector<std::vector<std::vector<int>>> sendT(T, std::vector<std::vector<int>>(P)),
      recvT(T, std::vector<std::vector<int>>(P));
  parallelFor(n_, [&](int64_t r0, int64_t r1, int t) {
    auto& snd = sendT[t & (T - 1)];
    auto& rcv = recvT[t & (T - 1)];
    for (int64_t r = r0; r < r1; ++r) {
      const int q = D.rowOwner[r];
      for (int64_t e = hRowptr_[r]; e < hRowptr_[r + 1]; ++e) {
        const int c = hColidx_[e];
        const int qc = D.rowOwner[c];
        if (q == qc) continue;
        if (q == me) rcv[qc].push_back(c);
        else if (qc == me) snd[q].push_back(c);
      }
    }
  });
  std::vector<std::vector<int>> send(P), recv(P);
  for (int t = 0; t < T; ++t)
    for (int q = 0; q < P; ++q) {
      send[q].insert(send[q].end(), sendT[t][q].begin(), sendT[t][q].end());
      recv[q].insert(recv[q].end(), recvT[t][q].begin(), recvT[t][q].end());
    }
  for (int q = 0; q < P; ++q) {
    sortUnique(send[q]);
    sortUnique(recv[q]);
  }
  uploadHalo(D.mat, send, recv, stream_, deviceOk_);
  if (deviceOk_) HY_CUDA(cudaStreamSynchronize(stream_));
  D.matReady = true;
}

// src[sendIdx] -> peers; received values land in dst[recvIdx] (scatter) or stay in h.recvBuf (for haloAdd)
void Engine::haloExchange(Halo& h, const double* src, double* dst, bool scatter) {
  cudaStream_t s = stream_;
  packIdx(src, h.sendIdx.p, h.sendBuf.p, h.sendPtr.back(), s, &launches_);
  comm_.neighbourExchange(h.peers, h.sendBuf.p, h.sendPtr, h.recvBuf.p, h.recvPtr, s);
  if (scatter) scatterVec(h.recvBuf.p, h.recvIdx.p, dst, h.recvPtr.back(), s, &launches_);
}

// Level-0 ApplyInverse on distributed vectors: B and X have global length; B is read and X written at the rows this
// rank owns (Preconditioner::ApplyInverse, src/HYMLS_Preconditioner.cpp:930-1070, no border)
void Engine::applyLevel0Dist(const double* B, double* X) {
  Level& L = *levels_[0];
  const LevelSym& S = L.sym;
  DistPlan& D = L.dist;
  cudaStream_t s = stream_;
  static const bool verboseApply = getenv("HYMLS_B200_VERBOSE_APPLY") != nullptr;
  const bool lap = verboseApply && comm_.rank() == 0 && (stats_.num_apply_inverse % 16) == 5;
  auto t0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!lap) return;
    cudaStreamSynchronize(s);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[hymls_b200 apply dist] %-36s %8.3f ms\n", what,
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  };
  if (lap) cudaStreamSynchronize(s);
  t0 = std::chrono::steady_clock::now();
  // x1 = leading rows of A11 \ b1 on the owned subdomains
  GemvArgs g = L.a11.args();
  g.xin = B;
  g.gather = L.intRow.p;
  g.out = L.x1.p;
  g.mode = 0;
  g.itemMat = L.a11.itemMatLead.p;
  g.itemRow0 = L.a11.itemRow0Lead.p;
  g.nrows = L.a11.rowLimit.p;
  const bool timeIt = timeA11_;
  if (timeIt) HY_CUDA(cudaEventRecord(evA_, s));
  batchedGemv(g, L.a11.numItemsLead, L.a11.npMax, s, &launches_);
  if (timeIt) {
    HY_CUDA(cudaEventRecord(evB_, s));
    HY_CUDA(cudaEventSynchronize(evB_));
    float ms = 0;
    HY_CUDA(cudaEventElapsedTime(&ms, evA_, evB_));
    a11LeadMs_ += ms;
  }
  const bool split = splitActive(L, 0);  // see Engine::applyLevel
  if (split) {
    HY_CUDA(cudaEventRecord(evFork_, s));
    HY_CUDA(cudaStreamWaitEvent(side_, evFork_, 0));
    GemvArgs t = g;
    t.itemMat = L.a11.itemMatTrail.p;
    t.itemRow0 = L.a11.itemRow0Trail.p;
    t.nrows = nullptr;
    batchedGemv(t, L.a11.numItemsTrail, L.a11.npMax, side_, &launches_);
    HY_CUDA(cudaEventRecord(evJoin_, side_));
  }
  mark("A11 gemv 1 (leading rows)");
  // Z = -A21[:, owned interiors] x1 at the separators my subdomains touch; ghost parts go to their owners
  spmvRows(L.p21.p, L.c21.p, L.v21.p, L.x1.p, L.Z.p, D.rows21.p, D.nRows21, 0.0, nullptr, nullptr, -1.0, 0, s,
           &launches_);
  haloExchange(D.rev, L.Z.p, nullptr, false);
  haloAdd(L.Z.p, D.addNode.p, D.addPtr.p, D.addSrc.p, D.rev.recvBuf.p, D.nAdd, s, &launches_);
  gatherAddList(B, L.sepRow.p, L.Z.p, L.rhsS.p, D.ownSepPos.p, D.nOwnSep, s, &launches_);
  mark("A21 spmv + halo (partial sums)");
  // Householder on the owned groups; V-sum right-hand side summed to every rank (zeros elsewhere: exact)
  HY_CUDA(cudaMemsetAsync(L.vsRhs.p, 0, (size_t)S.nuniq * sizeof(double), s));
  householderList(L.uniqStart.p, D.ownUniq.p, (int)D.nOwnUniq, L.what.p, L.rhsS.p, L.Z.p, L.vsRhs.p, nullptr, nullptr,
                  nullptr, s, &launches_);
  comm_.allReduceSum(L.vsRhs.p, (size_t)S.nuniq, s);
  blockSolves(L, L.Z.p, L.Y.p);
  mark("householder + blocks + vsum allreduce");
  if (levels_.size() > 1) {
    applyLevel(1, L.vsRhs.p, L.vsSol.p, nullptr);
  } else {
    coarseSolve(L.vsRhs.p, L.vsSol.p, S.nuniq);
    comm_.broadcast(L.vsSol.p, (size_t)S.nuniq, 0, s);
  }
  mark("next level / coarse");
  householderList(L.uniqStart.p, D.ownUniq.p, (int)D.nOwnUniq, L.what.p, L.Y.p, L.Y.p, nullptr, L.vsSol.p, X,
                  L.sepRow.p, s, &launches_);
  haloExchange(D.fwd, L.Y.p, L.Y.p, true);
  mark("householder back + halo (x2)");
  spmvRows(L.p12.p, L.c12.p, L.v12.p, L.Y.p, L.y1.p, D.rows12.p, D.nRows12, 0.0, nullptr, nullptr, 1.0, 0, s,
           &launches_);
  g = L.a11.args();
  g.xin = B;
  g.gather = L.intRow.p;
  g.xsub = L.y1.p;
  g.out = X;
  g.scatter = L.intRow.p;
  g.mode = 0;
  if (split) {  // X[interior] = x1 - A11^-1[:, :nb] y1
    HY_CUDA(cudaStreamWaitEvent(s, evJoin_, 0));
    g.xin = L.y1.p;
    g.gather = nullptr;
    g.xsub = nullptr;
    g.xprev = L.x1.p;
    g.ncols = L.a11.rowLimit.p;
    g.mode = 1;
  }
  if (timeIt) HY_CUDA(cudaEventRecord(evA_, s));
  batchedGemv(g, L.a11.numItems, L.a11.npMax, s, &launches_);
  if (timeIt) {
    HY_CUDA(cudaEventRecord(evB_, s));
    HY_CUDA(cudaEventSynchronize(evB_));
    float ms = 0;
    HY_CUDA(cudaEventElapsedTime(&ms, evA_, evB_));
    a11Ms_ += ms;
    a11Launches_++;
  }
  mark("A12 spmv + A11 gemv 2");
}

// owned entries of a global-length vector -> the full vector on every rank (replicated output of the plain API)
void Engine::gatherOwned(const double* Xowned, double* Xfull) {
  DistPlan& D = levels_[0]->dist;
  cudaStream_t s = stream_;
  const int P = comm_.size();
  double* mine = D.gath.p + (int64_t)comm_.rank() * D.maxOwn;
  packIdx(Xowned, D.ownRows.p, mine, D.nOwn, s, &launches_);
  comm_.allGather(mine, D.gath.p, (size_t)D.maxOwn, s);
  scatterVecMasked(D.gath.p, D.allRows.p, Xfull, (int64_t)P * D.maxOwn, s, &launches_);
}

int64_t Engine::ownedRows(int64_t* rows, int64_t cap) {
  if (!initialized_) throw Error(HYMLS_B200_ERR_STATE, "not initialized");
  if (comm_.size() <= 1) {
    if (rows && cap >= n_)
      for (int64_t r = 0; r < n_; ++r) rows[r] = r;
    return n_;
  }
  DistPlan& D = levels_[0]->dist;
  if (!D.ready) throw Error(HYMLS_B200_ERR_STATE, "owned_rows: no distributed plan (Number of Levels = 0 is single-GPU only)");
  if (rows && cap >= D.nOwn)
    for (int64_t i = 0; i < D.nOwn; ++i) rows[i] = D.hOwnRows[i];
  return D.nOwn;
}

// ---------------------------------------------------------------------------------------------
// Krylov driver on row-owner-distributed vectors (same algorithm and parameters as Engine::solve)
// ---------------------------------------------------------------------------------------------
void Engine::solveDist(const double* b, double* x, int where, uint64_t seed, hymls_b200_solve_info* info, double* hist,
                       int histCap) {
  ParameterList& sol = params_.sublist("Solver");
  ParameterList& it = sol.sublist("Iterative Solver");
  const std::string method = sol.get("Krylov Method", "GMRES");
  const std::string side = sol.get("Left or Right Preconditioning", "Right");
  const std::string startVec = sol.get("Initial Vector", "Random");
  const int maxIters = it.get("Maximum Iterations", 1000);
  const double tol = it.get("Convergence Tolerance", 1e-8);
  int numBlocks = it.get("Num Blocks", 300);
  const int maxRestarts = it.get("Maximum Restarts", 20);
  const bool explicitTest = it.get("Explicit Residual Test", false);
  const std::string impScaling = it.get("Implicit Residual Scaling", "Norm of Preconditioned Initial Residual");
  const std::string expScaling = it.get("Explicit Residual Scaling", "Norm of Initial Residual");
  Level& L0 = *levels_[0];
  DistPlan& D = L0.dist;
  buildMatrixHalo();
  cudaStream_t s = stream_;
  const int64_t n = n_, nOwn = D.nOwn, ld = (D.maxOwn + 7) & ~(int64_t)7;
  numBlocks = std::max(1, std::min(numBlocks, maxIters));
  const int m = numBlocks;
  // global-length staging vectors (owned + halo entries in use) and compact vectors
  kX_.alloc(n);   // global: argument / result of the preconditioner and the operator
  kB_.alloc(n);   // global: second staging vector
  kR_.alloc(ld);  // compact r
  kW_.alloc(ld);  // compact scratch
  kZ_.alloc(ld);  // compact scratch
  DevBuf<double> xc, bc;  // compact solution and right-hand side
  xc.alloc(ld);
  bc.alloc(ld);
  kH_.alloc(2 * m + 8);
  kPartial_.alloc((size_t)(m + 2) * multiDotBlocks());
  const auto kind = where == HYMLS_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  HY_CUDA(cudaMemcpyAsync(kX_.p, b, n * sizeof(double), kind, s));
  packIdx(kX_.p, D.ownRows.p, bc.p, nOwn, s, &launches_);
  if (startVec == "Random") {
    std::vector<double> h(n), hc(nOwn);
    std::mt19937_64 rng(seed);
    for (auto& v : h) v = (double)(rng() >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    for (int64_t i = 0; i < nOwn; ++i) hc[i] = h[D.hOwnRows[i]];
    HY_CUDA(cudaMemcpyAsync(xc.p, hc.data(), nOwn * sizeof(double), cudaMemcpyHostToDevice, s));
    HY_CUDA(cudaStreamSynchronize(s));
  } else if (startVec == "Zero") {
    HY_CUDA(cudaMemsetAsync(xc.p, 0, ld * sizeof(double), s));
  } else {
    HY_CUDA(cudaMemcpyAsync(kX_.p, x, n * sizeof(double), kind, s));
    packIdx(kX_.p, D.ownRows.p, xc.p, nOwn, s, &launches_);
  }
  // the ranks start the timed region together (a late rank would otherwise bill its delay to the others' first
  // collective)
  if (method == "GMRES") kV_.alloc((size_t)(m + 1) * ld);  // workspace (grow-only): allocated outside the timed region
  comm_.allReduceSum(kH_.p, 1, s);
  HY_CUDA(cudaStreamSynchronize(s));
  HY_CUDA(cudaEventRecord(ev0_, s));
  // operators: "global" vectors have length n with the owned entries valid
  auto Aglobal = [&](double* full, double* outC) {  // halo of `full` is filled in place
    haloExchange(D.mat, full, full, true);
    spmvRows(L0.rowptr.p, L0.colidx.p, L0.val.p, full, outC, D.ownRows.p, nOwn, 0.0, nullptr, nullptr, 1.0, 1, s,
             &launches_);
  };
  auto toGlobal = [&](const double* c, double* full) { scatterVec(c, D.ownRows.p, full, nOwn, s, &launches_); };
  auto A = [&](const double* inC, double* outC) {
    toGlobal(inC, kB_.p);
    Aglobal(kB_.p, outC);
  };
  auto Mglobal = [&](const double* inC, double* outFull) {
    toGlobal(inC, kB_.p);
    applyLevel0Dist(kB_.p, outFull);
    stats_.num_apply_inverse++;
  };
  auto M = [&](const double* inC, double* outC) {
    Mglobal(inC, kX_.p);
    packIdx(kX_.p, D.ownRows.p, outC, nOwn, s, &launches_);
  };
  auto dot = [&](const double* u, const double* v) {
    double* d = kH_.p + 2 * m + 4;
    multiDot(u, ld, 1, v, nOwn, kPartial_.p, d, 0, s, &launches_);
    comm_.allReduceSum(d, 1, s);
    double h;
    HY_CUDA(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, s));
    HY_CUDA(cudaStreamSynchronize(s));
    return h;
  };
  std::vector<double> history;
  int iters = 0;
  bool converged = false;
  double lastRel = 0;
  const bool left = side == "Left", right = side == "Right";
  const double bnorm = std::sqrt(dot(bc.p, bc.p));

  if (method == "CG") {
    DevBuf<double> P, Ap;
    P.alloc(ld);
    Ap.alloc(ld);
    A(xc.p, kR_.p);
    axpby(1.0, bc.p, -1.0, kR_.p, nOwn, s, &launches_);
    const double r0 = std::sqrt(dot(kR_.p, kR_.p));
    history.push_back(1.0);
    if (r0 == 0) {
      converged = true;
    } else {
      M(kR_.p, kZ_.p);
      axpby(1.0, kZ_.p, 0.0, P.p, nOwn, s, &launches_);
      double rz = dot(kR_.p, kZ_.p);
      while (iters < maxIters) {
        A(P.p, Ap.p);
        const double alpha = rz / dot(P.p, Ap.p);
        axpby(alpha, P.p, 1.0, xc.p, nOwn, s, &launches_);
        axpby(-alpha, Ap.p, 1.0, kR_.p, nOwn, s, &launches_);
        ++iters;
        lastRel = std::sqrt(dot(kR_.p, kR_.p)) / r0;
        history.push_back(lastRel);
        if (lastRel <= tol) {
          converged = true;
          break;
        }
        M(kR_.p, kZ_.p);
        const double rzNew = dot(kR_.p, kZ_.p);
        axpby(1.0, kZ_.p, rzNew / rz, P.p, nOwn, s, &launches_);
        rz = rzNew;
      }
    }
  } else if (method == "GMRES") {
    kV_.alloc((size_t)(m + 1) * ld);
    HY_CUDA(cudaMemsetAsync(kV_.p, 0, (size_t)(m + 1) * ld * sizeof(double), s));
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), hbuf(2 * m + 8);
    double* dH1 = kH_.p;
    double* dH2 = kH_.p + (m + 1);
    double* dNrm = kH_.p + 2 * m + 2;
    A(xc.p, kR_.p);
    axpby(1.0, bc.p, -1.0, kR_.p, nOwn, s, &launches_);
    const double r0norm = std::sqrt(dot(kR_.p, kR_.p));
    double* r = kR_.p;
    if (left) {
      M(kR_.p, kZ_.p);
      r = kZ_.p;
    }
    const double pr0norm = left ? std::sqrt(dot(r, r)) : r0norm;
    auto scaleOf = [&](const std::string& k) {
      double v = 1.0;
      if (k == "Norm of RHS") v = bnorm;
      else if (k == "Norm of Initial Residual") v = r0norm;
      else if (k == "Norm of Preconditioned Initial Residual") v = pr0norm;
      else if (k == "None") v = 1.0;
      else throw Error(HYMLS_B200_ERR_ARG, "unknown residual scaling '" + k + "'");
      return v == 0.0 ? 1.0 : v;
    };
    const double impScale = scaleOf(impScaling), expScale = scaleOf(expScaling);
    double beta = pr0norm;
    double trueRes = r0norm;
    for (int restart = 0; restart <= maxRestarts; ++restart) {
      if (restart == 0) history.push_back(beta / impScale);
      lastRel = beta / impScale;
      if (beta == 0.0 || (lastRel <= tol && !explicitTest)) {
        converged = true;
        break;
      }
      axpby(1.0 / beta, r, 0.0, kV_.p, nOwn, s, &launches_);
      std::fill(g.begin(), g.end(), 0.0);
      g[0] = beta;
      int kDone = 0;
      for (int k = 0; k < m && iters < maxIters; ++k) {
        double* vk = kV_.p + (size_t)k * ld;
        double* w = kV_.p + (size_t)(k + 1) * ld;
        static const bool verboseSolve = getenv("HYMLS_B200_VERBOSE_SOLVE") != nullptr;
        const bool lapIt = verboseSolve && comm_.rank() == 0 && (k % 50) == 20;
        auto tl = std::chrono::steady_clock::now();
        auto lapS = [&](const char* what) {
          if (!lapIt) return;
          cudaStreamSynchronize(s);
          auto t1 = std::chrono::steady_clock::now();
          fprintf(stderr, "[hymls_b200 solve dist] it %d %-28s %8.3f ms\n", k, what,
                  std::chrono::duration<double, std::milli>(t1 - tl).count());
          tl = t1;
        };
        if (lapIt) { cudaStreamSynchronize(s); tl = std::chrono::steady_clock::now(); }
        if (right) {
          Mglobal(vk, kX_.p);
          lapS("ApplyInverse");
          Aglobal(kX_.p, w);
          lapS("matrix halo + spmv");
        } else if (left) {
          A(vk, kW_.p);
          M(kW_.p, w);
        } else {
          A(vk, w);
        }
        multiDot(kV_.p, ld, k + 1, w, nOwn, kPartial_.p, dH1, 0, s, &launches_);
        comm_.allReduceSum(dH1, (size_t)(k + 1), s);
        multiAxpy(kV_.p, ld, k + 1, dH1, w, nOwn, -1.0, s, &launches_);
        multiDot(kV_.p, ld, k + 1, w, nOwn, kPartial_.p, dH2, 0, s, &launches_);
        comm_.allReduceSum(dH2, (size_t)(k + 1), s);
        multiAxpy(kV_.p, ld, k + 1, dH2, w, nOwn, -1.0, s, &launches_);
        multiDot(w, ld, 1, w, nOwn, kPartial_.p, dNrm, 0, s, &launches_);
        comm_.allReduceSum(dNrm, 1, s);
        scaleByInvNorm(w, dNrm, w, nOwn, s, &launches_);
        lapS("CGS2 (4 sweeps, 3 allreduces)");
        HY_CUDA(cudaMemcpyAsync(hbuf.data(), kH_.p, (2 * m + 3) * sizeof(double), cudaMemcpyDeviceToHost, s));
        HY_CUDA(cudaStreamSynchronize(s));
        lapS("Hessenberg column to host");
        for (int i = 0; i <= k; ++i) H[(size_t)i * m + k] = hbuf[i] + hbuf[m + 1 + i];
        const double hn = std::sqrt(hbuf[2 * m + 2]);
        double colNorm = 0.0;
        for (int i = 0; i <= k; ++i) colNorm = std::hypot(colNorm, hbuf[i] + hbuf[m + 1 + i]);
        const bool breakdown = !(hn > 1e-300) || hn <= 1e-15 * colNorm;
        H[(size_t)(k + 1) * m + k] = breakdown ? 0.0 : hn;
        for (int i = 0; i < k; ++i) {
          const double t = cs[i] * H[(size_t)i * m + k] + sn[i] * H[(size_t)(i + 1) * m + k];
          H[(size_t)(i + 1) * m + k] = -sn[i] * H[(size_t)i * m + k] + cs[i] * H[(size_t)(i + 1) * m + k];
          H[(size_t)i * m + k] = t;
        }
        const double d = std::hypot(H[(size_t)k * m + k], H[(size_t)(k + 1) * m + k]);
        cs[k] = H[(size_t)k * m + k] / d;
        sn[k] = H[(size_t)(k + 1) * m + k] / d;
        H[(size_t)k * m + k] = d;
        H[(size_t)(k + 1) * m + k] = 0.0;
        g[k + 1] = -sn[k] * g[k];
        g[k] = cs[k] * g[k];
        ++iters;
        kDone = k + 1;
        lastRel = std::fabs(g[k + 1]) / impScale;
        history.push_back(lastRel);
        if (lastRel <= tol || breakdown) break;
      }
      if (kDone > 0) {
        std::vector<double> y(kDone);
        for (int i = kDone - 1; i >= 0; --i) {
          double t = g[i];
          for (int j = i + 1; j < kDone; ++j) t -= H[(size_t)i * m + j] * y[j];
          y[i] = t / H[(size_t)i * m + i];
        }
        HY_CUDA(cudaMemcpyAsync(dH1, y.data(), kDone * sizeof(double), cudaMemcpyHostToDevice, s));
        HY_CUDA(cudaMemsetAsync(kW_.p, 0, ld * sizeof(double), s));
        multiAxpy(kV_.p, ld, kDone, dH1, kW_.p, nOwn, 1.0, s, &launches_);
        HY_CUDA(cudaStreamSynchronize(s));
        if (right) {
          M(kW_.p, kZ_.p);
          axpby(1.0, kZ_.p, 1.0, xc.p, nOwn, s, &launches_);
        } else {
          axpby(1.0, kW_.p, 1.0, xc.p, nOwn, s, &launches_);
        }
      }
      A(xc.p, kR_.p);
      axpby(1.0, bc.p, -1.0, kR_.p, nOwn, s, &launches_);
      trueRes = std::sqrt(dot(kR_.p, kR_.p));
      r = kR_.p;
      beta = trueRes;
      if (left) {
        M(kR_.p, kZ_.p);
        r = kZ_.p;
        beta = std::sqrt(dot(r, r));
      }
      if (history.back() <= tol) {
        if (!explicitTest || trueRes / expScale <= tol) {
          converged = true;
          break;
        }
      }
      if (iters >= maxIters) break;
    }
  } else {
    throw Error(HYMLS_B200_ERR_ARG, "Krylov Method '" + method + "' not supported (GMRES, CG)");
  }
  HY_CUDA(cudaEventRecord(ev1_, s));
  HY_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  HY_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
  A(xc.p, kR_.p);
  axpby(1.0, bc.p, -1.0, kR_.p, nOwn, s, &launches_);
  const double res = std::sqrt(dot(kR_.p, kR_.p));
  // replicated result, like the single-GPU entry point
  toGlobal(xc.p, kB_.p);
  gatherOwned(kB_.p, kX_.p);
  const auto back = where == HYMLS_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  HY_CUDA(cudaMemcpyAsync(x, kX_.p, n * sizeof(double), back, s));
  HY_CUDA(cudaStreamSynchronize(s));
  if (info) {
    info->iterations = iters;
    info->converged = converged ? 1 : 0;
    info->rel_residual = lastRel;
    info->explicit_rel_residual = bnorm > 0 ? res / bnorm : res;
    info->solve_seconds = ms * 1e-3;
    info->history_len = (int)history.size();
  }
  if (hist)
    for (int i = 0; i < (int)history.size() && i < histCap; ++i) hist[i] = history[i];
}

}  // namespace hymls
