// Drop-in boundary for a DISTRIBUTED caller (one MPI rank per GPU): the reference takes a row-distributed
// Epetra_CrsMatrix on any map and imports it to its partitioner's map (src/HYMLS_Preconditioner.cpp:420-431), and
// vectors on the matrix' row map (:978-979, 1050-1052).  Here
//   * setMatrixDist: every rank passes its rows (GIDs, any distribution); rows are gathered over NCCL so that every
//     rank holds the pattern (the symbolic phase is replicated) and the values;
//   * setRowMap / applyInverseMap: vectors in the CALLER's distribution; they are moved to the owner distribution
//     of the library (hymls_b200_owned_rows) and back with one grouped ncclSend/ncclRecv each way.  When the
//     caller's map IS the owner map the exchange degenerates to a local permutation.
#include <algorithm>
#include <numeric>

#include "engine.hpp"

namespace hymls {

// ranks contribute `local` (any length); returns the concatenation in rank order and the per-rank counts
static std::vector<double> allGatherHost(const Comm& comm, const std::vector<double>& local, cudaStream_t s,
                                         std::vector<int64_t>& counts) {
  const int P = comm.size();
  counts.assign(P, (int64_t)local.size());
  if (P == 1 || !comm.active()) return local;
  DevBuf<double> cnt;
  cnt.alloc(P);
  const double mine = (double)local.size();
  HY_CUDA(cudaMemcpyAsync(cnt.p + comm.rank(), &mine, sizeof(double), cudaMemcpyHostToDevice, s));
  comm.allGather(cnt.p + comm.rank(), cnt.p, 1, s);
  std::vector<double> hc(P);
  HY_CUDA(cudaMemcpyAsync(hc.data(), cnt.p, P * sizeof(double), cudaMemcpyDeviceToHost, s));
  HY_CUDA(cudaStreamSynchronize(s));
  int64_t mx = 0;
  for (int q = 0; q < P; ++q) {
    counts[q] = (int64_t)hc[q];
    mx = std::max(mx, counts[q]);
  }
  if (mx == 0) return std::vector<double>();
  DevBuf<double> buf;
  buf.alloc((size_t)P * mx);
  if (!local.empty())
    HY_CUDA(cudaMemcpyAsync(buf.p + (size_t)comm.rank() * mx, local.data(), local.size() * sizeof(double),
                            cudaMemcpyHostToDevice, s));
  comm.allGather(buf.p + (size_t)comm.rank() * mx, buf.p, (size_t)mx, s);
  std::vector<double> all((size_t)P * mx), out;
  HY_CUDA(cudaMemcpyAsync(all.data(), buf.p, all.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  HY_CUDA(cudaStreamSynchronize(s));
  for (int q = 0; q < P; ++q) out.insert(out.end(), all.begin() + (size_t)q * mx, all.begin() + (size_t)q * mx + counts[q]);
  return out;
}

void Engine::setMatrixDist(int64_t nGlobal, int64_t nLocal, const int64_t* rowGids, const int64_t* rowptr,
                           const int64_t* colGids, const double* values) {
  if (nGlobal <= 0 || nLocal < 0 || (nLocal > 0 && (!rowGids || !rowptr || !colGids)))
    throw Error(HYMLS_B200_ERR_ARG, "set_matrix_csr_dist: bad arguments");
  if (comm_.size() > 1) needComm();
  if (values) needDevice();
  cudaStream_t s = stream_;
  const int64_t nnzLocal = nLocal ? rowptr[nLocal] : 0;
  // the local pattern decides whether the cached assembly map can be reused (NOX: same pattern every Newton step)
  std::vector<int64_t> sig;
  sig.reserve((size_t)(2 * nLocal + nnzLocal + 2));
  sig.push_back(nGlobal);
  sig.insert(sig.end(), rowGids, rowGids + nLocal);
  sig.insert(sig.end(), rowptr, rowptr + nLocal + (nLocal ? 1 : 0));
  sig.insert(sig.end(), colGids, colGids + nnzLocal);
  double same = (distMat_.ready && sig == distMat_.signature) ? 1.0 : 0.0;
  if (comm_.active()) {  // every rank must take the same branch
    DevBuf<double> f;
    f.alloc(1);
    same = 1.0 - same;
    HY_CUDA(cudaMemcpyAsync(f.p, &same, sizeof(double), cudaMemcpyHostToDevice, s));
    comm_.allReduceSum(f.p, 1, s);
    HY_CUDA(cudaMemcpyAsync(&same, f.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    HY_CUDA(cudaStreamSynchronize(s));
    same = same == 0.0 ? 1.0 : 0.0;
  }
  std::vector<int64_t> counts;
  if (same == 0.0) {
    // rows: (gid, length) pairs; columns as doubles (GIDs < 2^53 are exact)
    std::vector<double> rows((size_t)2 * nLocal), cols((size_t)nnzLocal);
    for (int64_t i = 0; i < nLocal; ++i) {
      rows[2 * i] = (double)rowGids[i];
      rows[2 * i + 1] = (double)(rowptr[i + 1] - rowptr[i]);
    }
    for (int64_t e = 0; e < nnzLocal; ++e) cols[e] = (double)colGids[e];
    const std::vector<double> allRows = allGatherHost(comm_, rows, s, counts);
    const std::vector<double> allCols = allGatherHost(comm_, cols, s, counts);
    const int64_t nrows = (int64_t)allRows.size() / 2;
    if (nrows != nGlobal)
      throw Error(HYMLS_B200_ERR_ARG, "set_matrix_csr_dist: the ranks passed " + std::to_string(nrows) +
                                          " rows in total, n_global is " + std::to_string(nGlobal));
    std::vector<int64_t> ptr(nGlobal + 1, 0);
    std::vector<char> seen(nGlobal, 0);
    for (int64_t i = 0; i < nrows; ++i) {
      const int64_t g = (int64_t)allRows[2 * i];
      if (g < 0 || g >= nGlobal || seen[g]) throw Error(HYMLS_B200_ERR_ARG, "set_matrix_csr_dist: row GIDs must tile [0, n)");
      seen[g] = 1;
      ptr[g + 1] = (int64_t)allRows[2 * i + 1];
    }
    for (int64_t g = 0; g < nGlobal; ++g) ptr[g + 1] += ptr[g];
    const int64_t nnz = ptr[nGlobal];
    std::vector<int> col(nnz);
    distMat_.slot.assign(nnz, 0);  // gathered entry k -> position in the assembled (column-sorted) CSR
    int64_t k = 0;
    std::vector<std::pair<int, int64_t>> rowEnt;
    for (int64_t i = 0; i < nrows; ++i) {
      const int64_t g = (int64_t)allRows[2 * i], len = (int64_t)allRows[2 * i + 1];
      rowEnt.resize(len);
      for (int64_t q = 0; q < len; ++q) rowEnt[q] = std::make_pair((int)allCols[k + q], k + q);
      std::sort(rowEnt.begin(), rowEnt.end());
      for (int64_t q = 0; q < len; ++q) {
        col[ptr[g] + q] = rowEnt[q].first;
        distMat_.slot[rowEnt[q].second] = ptr[g] + q;
      }
      k += len;
    }
    distMat_.ptr.swap(ptr);
    distMat_.col.swap(col);
    distMat_.signature.swap(sig);
    distMat_.ready = true;
  }
  std::vector<double> vals;
  if (values) {
    std::vector<double> loc(values, values + nnzLocal);
    const std::vector<double> all = allGatherHost(comm_, loc, s, counts);
    if (all.size() != distMat_.slot.size()) throw Error(HYMLS_B200_ERR_ARG, "set_matrix_csr_dist: value count changed");
    vals.resize(all.size());
    for (size_t e = 0; e < all.size(); ++e) vals[distMat_.slot[e]] = all[e];
  }
  setMatrix(nGlobal, distMat_.ptr.data(), distMat_.col.data(), values ? vals.data() : nullptr, HYMLS_B200_HOST);
}

// The caller's distribution of vectors: row_gids of this rank, in the caller's local order.  Collective.
void Engine::setRowMap(int64_t nLocal, const int64_t* rowGids) {
  if (!initialized_) throw Error(HYMLS_B200_ERR_STATE, "set_row_map: call Initialize first (the owner map is needed)");
  if (nLocal < 0 || (nLocal > 0 && !rowGids)) throw Error(HYMLS_B200_ERR_ARG, "set_row_map: bad arguments");
  if (comm_.size() > 1) needComm();
  needDevice();
  cudaStream_t s = stream_;
  const int P = comm_.size(), me = comm_.rank();
  RowMapPlan& R = rowMap_;
  R.ready = false;
  R.nLocal = nLocal;
  std::vector<double> mine(rowGids, rowGids + nLocal);
  std::vector<int64_t> counts;
  const std::vector<double> all = allGatherHost(comm_, mine, s, counts);
  if ((int64_t)all.size() != n_) throw Error(HYMLS_B200_ERR_ARG, "set_row_map: the ranks' rows must tile [0, n)");
  // holder (rank in the caller's map) of every row
  std::vector<int> holder(n_, -1);
  {
    int64_t k = 0;
    for (int q = 0; q < P; ++q)
      for (int64_t i = 0; i < counts[q]; ++i, ++k) {
        const int64_t g = (int64_t)all[k];
        if (g < 0 || g >= n_ || holder[g] >= 0) throw Error(HYMLS_B200_ERR_ARG, "set_row_map: rows must tile [0, n)");
        holder[g] = q;
      }
  }
  const std::vector<int>* owner = nullptr;
  std::vector<int> zero;
  if (P > 1) {
    if (!levels_[0]->dist.ready) throw Error(HYMLS_B200_ERR_STATE, "set_row_map: no distributed plan");
    owner = &levels_[0]->dist.rowOwner;
  } else {
    zero.assign(n_, 0);
    owner = &zero;
  }
  // to the owners: I send my local entries (ascending GID per destination); I receive the rows I own from their
  // holders (ascending GID per source) into the global-length work vector
  std::vector<std::vector<int>> sendTo(P), recvFrom(P);
  {
    std::vector<std::pair<int64_t, int>> byGid(nLocal);
    for (int64_t i = 0; i < nLocal; ++i) byGid[i] = std::make_pair(rowGids[i], (int)i);
    std::sort(byGid.begin(), byGid.end());
    for (auto& pr : byGid) sendTo[(*owner)[pr.first]].push_back(pr.second);
    for (int64_t g = 0; g < n_; ++g)
      if ((*owner)[g] == me) recvFrom[holder[g]].push_back((int)g);
  }
  auto build = [&](Halo& h, const std::vector<std::vector<int>>& snd, const std::vector<std::vector<int>>& rcv) {
    h.peers.clear();
    h.sendPtr.assign(1, 0);
    h.recvPtr.assign(1, 0);
    std::vector<int> si, ri;
    for (int q = 0; q < P; ++q) {
      if (snd[q].empty() && rcv[q].empty()) continue;
      h.peers.push_back(q);
      si.insert(si.end(), snd[q].begin(), snd[q].end());
      ri.insert(ri.end(), rcv[q].begin(), rcv[q].end());
      h.sendPtr.push_back((int64_t)si.size());
      h.recvPtr.push_back((int64_t)ri.size());
    }
    h.sendIdx.upload(si, s);
    h.recvIdx.upload(ri, s);
    h.sendBuf.alloc(std::max<size_t>(si.size(), 1));
    h.recvBuf.alloc(std::max<size_t>(ri.size(), 1));
  };
  build(R.toOwner, sendTo, recvFrom);
  build(R.fromOwner, recvFrom, sendTo);
  R.stage.alloc((size_t)std::max<int64_t>(2 * nLocal, 1));
  HY_CUDA(cudaStreamSynchronize(s));
  R.ready = true;
}

// ApplyInverse on vectors in the caller's distribution (Epetra_Operator::ApplyInverse of the adapter); with a border
// set and T, S given it is BorderedOperator::ApplyInverse(X, T, Y, S) (T, S: m x nvec, column major, replicated)
void Engine::applyInverseMap(const double* B, int64_t ldb, double* X, int64_t ldx, int nvec, int where, const double* T,
                             double* S) {
  needDevice();
  needComm();
  if (!computed_) throw Error(HYMLS_B200_ERR_STATE, "The preconditioner has not yet been computed.");
  RowMapPlan& R = rowMap_;
  if (!R.ready) throw Error(HYMLS_B200_ERR_STATE, "apply_inverse_map: call hymls_b200_set_row_map first");
  if (!B || !X || nvec < 0 || ldb < R.nLocal || ldx < R.nLocal) throw Error(HYMLS_B200_ERR_ARG, "apply_inverse_map: bad arguments");
  const int bm = borderM_;
  if (bm > 0 && (T != nullptr) != (S != nullptr)) throw Error(HYMLS_B200_ERR_ARG, "apply_inverse_map: T and S go together");
  cudaStream_t s = stream_;
  bufB_.alloc(n_);
  bufX_.alloc(n_);
  const auto in = where == HYMLS_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const auto out = where == HYMLS_B200_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  double* cB = R.stage.p;
  double* cX = R.stage.p + std::max<int64_t>(R.nLocal, 1);
  for (int k = 0; k < nvec; ++k) {
    if (R.nLocal) HY_CUDA(cudaMemcpyAsync(cB, B + (int64_t)k * ldb, R.nLocal * sizeof(double), in, s));
    haloExchange(R.toOwner, cB, bufB_.p, true);
    const double* dT = nullptr;
    if (bm > 0) {  // plain ApplyInverse with a border set: T = 0, S discarded (src/HYMLS_Preconditioner.cpp:594-605)
      bTin_.alloc(bm);
      if (T) HY_CUDA(cudaMemcpyAsync(bTin_.p, T + (size_t)k * bm, bm * sizeof(double), in, s));
      else HY_CUDA(cudaMemsetAsync(bTin_.p, 0, bm * sizeof(double), s));
      dT = bTin_.p;
    }
    applyOwned(bufB_.p, bufX_.p, dT);
    if (bm > 0 && S) HY_CUDA(cudaMemcpyAsync(S + (size_t)k * bm, bS_.p, bm * sizeof(double), out, s));
    haloExchange(R.fromOwner, bufX_.p, cX, true);
    if (R.nLocal) HY_CUDA(cudaMemcpyAsync(X + (int64_t)k * ldx, cX, R.nLocal * sizeof(double), out, s));
  }
  if (where == HYMLS_B200_HOST) HY_CUDA(cudaStreamSynchronize(s));
}

// full-length vector(s) in GID order from distributed pieces (host): test vector, border columns
static std::vector<double> gatherByGid(const Comm& comm, int64_t n, int64_t nLocal, const int64_t* gids,
                                       const double* v, int64_t ld, int ncol, cudaStream_t s) {
  std::vector<double> loc((size_t)nLocal * (1 + ncol));
  for (int64_t i = 0; i < nLocal; ++i) {
    loc[(size_t)i * (1 + ncol)] = (double)gids[i];
    for (int j = 0; j < ncol; ++j) loc[(size_t)i * (1 + ncol) + 1 + j] = v[i + (int64_t)j * ld];
  }
  std::vector<int64_t> counts;
  const std::vector<double> all = allGatherHost(comm, loc, s, counts);
  if ((int64_t)all.size() != n * (1 + ncol)) throw Error(HYMLS_B200_ERR_ARG, "distributed vector: the ranks' rows must tile [0, n)");
  std::vector<double> out((size_t)n * ncol);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t g = (int64_t)all[(size_t)i * (1 + ncol)];
    if (g < 0 || g >= n) throw Error(HYMLS_B200_ERR_ARG, "distributed vector: row GID out of range");
    for (int j = 0; j < ncol; ++j) out[g + (int64_t)j * n] = all[(size_t)i * (1 + ncol) + 1 + j];
  }
  return out;
}

void Engine::setTestVectorDist(int64_t nLocal, const int64_t* gids, const double* tv) {
  if (!haveMatrix_) throw Error(HYMLS_B200_ERR_STATE, "set the matrix before the test vector");
  if (!tv) { setTestVector(nullptr); return; }
  if (comm_.size() > 1) { needDevice(); needComm(); }
  const std::vector<double> full = gatherByGid(comm_, n_, nLocal, gids, tv, nLocal, 1, stream_);
  setTestVector(full.data());
}

void Engine::setBorderDist(int64_t nLocal, const int64_t* gids, const double* V, const double* W, const double* C, int m) {
  if (!haveMatrix_) throw Error(HYMLS_B200_ERR_STATE, "set the matrix before the border");
  if (!V || m <= 0) { setBorder(nullptr, nullptr, nullptr, 0); return; }
  if (comm_.size() > 1) { needDevice(); needComm(); }
  const std::vector<double> fv = gatherByGid(comm_, n_, nLocal, gids, V, nLocal, m, stream_);
  std::vector<double> fw;
  if (W) fw = gatherByGid(comm_, n_, nLocal, gids, W, nLocal, m, stream_);
  setBorder(fv.data(), W ? fw.data() : nullptr, C, m);
}

}  // namespace hymls
