// Coarse (last-level) solver beyond the size a single dense inverse can serve.
//
// Reference: CoarseSolver::Compute / ApplyInverse (src/HYMLS_CoarseSolver.cpp:131-323): drop (RelFullDiag), Dirichlet
// rows for the "Fix GID" entries, a sparse direct factorization (Amesos KLU on the root rank), solves with the fixed
// rows' right-hand side zeroed.  Up to COARSE_DENSE_MAX rows this library inverts the matrix densely (engine.cu);
// beyond, a dense inverse is neither affordable (95 356 V-sums at 128^3 / sx = 8 with one level: 73 GB) nor applicable
// (the GEMV keeps the right-hand side in shared memory).  Here: BLOCK-TRIDIAGONAL factorization on breadth-first level
// sets.  A BFS of the matrix graph from a pseudo-peripheral node puts every neighbour of a level-i node into levels
// i-1, i, i+1, so in the level ordering the matrix is block tridiagonal with sparse off-diagonal blocks:
//       D_0' = D_0,   D_i' = D_i - L_i inv(D_{i-1}') U_{i-1},          (L_i = A[i, i-1], U_{i-1} = A[i-1, i])
// and with F_i = inv(D_i') explicit (batched Gauss-Jordan + Newton-Schulz like every other block of the library)
//       forward   y_i = b_i - L_i t_{i-1},  t_i = F_i y_i      backward   x_i = t_i - F_i (U_i x_{i+1}).
// Only the F_i are dense: sum b_i^2 doubles (about 5 GB for the 95 356 V-sums), streamed twice per solve; the
// products with L_i, U_i are CSR SpMVs on row ranges.  Small consecutive levels are merged into blocks of at least
// 1024 rows (merging neighbours keeps the structure).  The exact block LU of the SAME matrix the reference
// factors: a direct solve, no approximation.  Replicated on every rank like the dense variant.
#include <algorithm>
#include <queue>

#include "engine.hpp"

namespace hymls {

static int coarseMinBlock() {  // rows per merged block (HYMLS_B200_COARSE_MIN_BLOCK: tests use small blocks)
  if (const char* e = getenv("HYMLS_B200_COARSE_MIN_BLOCK")) return std::max(1, atoi(e));
  return 1024;
}

// row -> identity row, entries of its column -> 0 (MatrixUtils::PutDirichlet on the CSR values; the column entries
// are found through the row's own pattern like the reference does: structurally symmetric matrix)
__global__ void k_put_dirichlet_csr(double* __restrict__ val, const int64_t* __restrict__ ptr, const int* __restrict__ col,
                                    int row) {
  const int64_t a = ptr[row], z = ptr[row + 1];
  for (int64_t e = a + threadIdx.x; e < z; e += blockDim.x) {
    const int c = col[e];
    val[e] = (c == row) ? 1.0 : 0.0;
    if (c != row)
      for (int64_t f = ptr[c]; f < ptr[c + 1]; ++f)
        if (col[f] == row) val[f] = 0.0;
  }
}

// dense block (rows [r0, r0+nr) x columns [c0, c0+nc) in the PERMUTED numbering) of a CSR matrix given in that
// numbering with values referenced through src: D[(r-r0)*ld + (col-c0)] = val[src[e]]
__global__ void k_csr_block_dense(const int64_t* __restrict__ ptr, const int* __restrict__ col,
                                  const int64_t* __restrict__ src, const double* __restrict__ val,
                                  double* __restrict__ D, int r0, int nr, int c0, int nc, int ld) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr) return;
  for (int64_t e = ptr[r0 + r]; e < ptr[r0 + r + 1]; ++e) {
    const int c = col[e] - c0;
    if (c >= 0 && c < nc) D[(int64_t)r * ld + c] = val[src[e]];
  }
}

void putDirichletCsr(double* val, const int64_t* ptr, const int* col, int row, cudaStream_t s, int64_t* launches) {
  k_put_dirichlet_csr<<<1, 128, 0, s>>>(val, ptr, col, row);
  ++*launches;
}

// level structure + index arrays from the pattern (host; once per pattern)
void Engine::planCoarseBT(const std::vector<int64_t>& ptr, const std::vector<int>& col, int n) {
  CoarseBT& C = coarseBT_;
  cudaStream_t s = stream_;
  C.n = n;
  // breadth-first level sets, every connected component from a pseudo-peripheral node (two sweeps)
  std::vector<int> level(n, -1), order;
  order.reserve(n);
  auto bfs = [&](int start, std::vector<int>& lv, std::vector<int>& visited) {
    visited.clear();
    std::queue<int> q;
    q.push(start);
    lv[start] = 0;
    int last = start;
    while (!q.empty()) {
      const int v = q.front();
      q.pop();
      visited.push_back(v);
      last = v;
      for (int64_t e = ptr[v]; e < ptr[v + 1]; ++e) {
        const int w = col[e];
        if (lv[w] < 0) {
          lv[w] = lv[v] + 1;
          q.push(w);
        }
      }
    }
    return last;
  };
  std::vector<int> visited;
  int numLevels = 0;
  for (int v0 = 0; v0 < n; ++v0) {
    if (level[v0] >= 0) continue;
    const int far = bfs(v0, level, visited);
    for (int v : visited) level[v] = -1;
    bfs(far, level, visited);
    for (int v : visited) numLevels = std::max(numLevels, level[v] + 1);
  }
  // merge consecutive levels into blocks of at least COARSE_MIN_BLOCK rows
  std::vector<int> cnt(numLevels, 0);
  for (int v = 0; v < n; ++v) cnt[level[v]]++;
  std::vector<int> blockOf(numLevels, 0);
  const int minBlock = coarseMinBlock();
  int nb = 0, acc = 0;
  for (int l = 0; l < numLevels; ++l) {
    blockOf[l] = nb;
    acc += cnt[l];
    if (acc >= minBlock && l + 1 < numLevels) { ++nb; acc = 0; }
  }
  const int m = nb + 1;
  C.m = m;
  C.lvlPtr.assign(m + 1, 0);
  for (int v = 0; v < n; ++v) C.lvlPtr[blockOf[level[v]] + 1]++;
  for (int b = 0; b < m; ++b) C.lvlPtr[b + 1] += C.lvlPtr[b];
  std::vector<int> fill(C.lvlPtr.begin(), C.lvlPtr.end() - 1), pos(n);
  C.perm.assign(n, 0);
  for (int v = 0; v < n; ++v) {  // rows keep their relative order inside a block
    const int p = fill[blockOf[level[v]]]++;
    C.perm[p] = v;
    pos[v] = p;
  }
  // dense inverses of the diagonal blocks
  std::vector<int> bn(m), bnp(m);
  std::vector<int64_t> moff(m + 1, 0), voff(m);
  C.itemPtr.assign(m + 1, 0);
  const int rows = gemvRowsPerItem();
  for (int b = 0; b < m; ++b) {
    bn[b] = C.lvlPtr[b + 1] - C.lvlPtr[b];
    bnp[b] = (bn[b] + 7) & ~7;
    moff[b + 1] = moff[b] + (int64_t)bnp[b] * bnp[b];
    voff[b] = C.lvlPtr[b];
    C.itemPtr[b + 1] = C.itemPtr[b] + (bn[b] + rows - 1) / rows;
  }
  C.inv.setup(bn, bnp, moff, voff, s);
  // diagonal scatter lists and the strictly lower / upper block parts as CSR in the permuted numbering
  std::vector<int64_t> dSrc, dDst, loPtr(n + 1, 0), upPtr(n + 1, 0), loSrc, upSrc;
  std::vector<int> loCol, upCol;
  C.dListPtr.assign(m + 1, 0);
  std::vector<std::pair<int, int64_t>> rowEnt;
  for (int b = 0; b < m; ++b) {
    for (int p = C.lvlPtr[b]; p < C.lvlPtr[b + 1]; ++p) {
      const int v = C.perm[p];
      rowEnt.clear();
      for (int64_t e = ptr[v]; e < ptr[v + 1]; ++e) rowEnt.emplace_back(pos[col[e]], e);
      std::sort(rowEnt.begin(), rowEnt.end());
      for (auto& pr : rowEnt) {
        const int q = pr.first;
        if (q < C.lvlPtr[b]) {
          if (b == 0 || q < C.lvlPtr[b - 1]) throw Error(HYMLS_B200_ERR_NUMERIC, "coarse solver: level structure broken");
          loCol.push_back(q);
          loSrc.push_back(pr.second);
        } else if (q >= C.lvlPtr[b + 1]) {
          if (b + 1 >= m || q >= C.lvlPtr[b + 2]) throw Error(HYMLS_B200_ERR_NUMERIC, "coarse solver: level structure broken");
          upCol.push_back(q);
          upSrc.push_back(pr.second);
        } else {
          dSrc.push_back(pr.second);
          dDst.push_back(moff[b] + (int64_t)(p - C.lvlPtr[b]) * bnp[b] + (q - C.lvlPtr[b]));
        }
      }
      loPtr[p + 1] = (int64_t)loCol.size();
      upPtr[p + 1] = (int64_t)upCol.size();
    }
    C.dListPtr[b + 1] = (int64_t)dSrc.size();
  }
  C.dSrc.upload(dSrc, s);
  C.dDst.upload(dDst, s);
  C.loPtr.upload(loPtr, s);
  C.upPtr.upload(upPtr, s);
  C.loCol.upload(loCol, s);
  C.upCol.upload(upCol, s);
  C.loSrc.upload(loSrc, s);
  C.upSrc.upload(upSrc, s);
  C.loVal.alloc(loCol.size());
  C.upVal.alloc(upCol.size());
  C.dPerm.upload(C.perm, s);
  C.y.alloc(n);
  C.t.alloc(n);
  C.r.alloc(n);
  C.xp.alloc(n);
  HY_CUDA(cudaStreamSynchronize(s));
  C.planned = true;
}

// numeric factorization; val holds the (already dropped, Dirichlet-adjusted) CSR values on the device
void Engine::factorCoarseBT(const int64_t* ptr, const int* col, const double* val) {
  CoarseBT& C = coarseBT_;
  cudaStream_t s = stream_;
  gatherValues(val, C.loSrc.p, C.loVal.p, (int64_t)C.loVal.n, s, &launches_);
  gatherValues(val, C.upSrc.p, C.upVal.p, (int64_t)C.upVal.n, s, &launches_);
  const int m = C.m;
  int npMax = 0;
  for (int b = 0; b < m; ++b) npMax = std::max(npMax, C.inv.hNp[b]);
  const size_t blk = (size_t)npMax * npMax;
  work_.alloc(blk);
  work2_.alloc(blk);
  blkA_.alloc(blk);
  a21d_.alloc(blk);  // dense L_i
  a12d_.alloc(blk);  // dense U_{i-1}
  dmat_.alloc(blk);  // T = F_{i-1} U_{i-1}
  for (int b = 0; b < m; ++b) {
    const int nb = C.inv.hN[b], np = C.inv.hNp[b];
    const size_t len = (size_t)np * np;
    HY_CUDA(cudaMemsetAsync(work_.p, 0, len * sizeof(double), s));
    const int64_t e0 = C.dListPtr[b], e1 = C.dListPtr[b + 1];
    scatterValues(val, C.dSrc.p + e0, C.dDst.p + e0, C.inv.hMatOff[b], work_.p, e1 - e0, s, &launches_);
    if (b > 0) {
      const int pn = C.inv.hN[b - 1], pnp = C.inv.hNp[b - 1];
      // dense L_b (nb x pnp) and U_{b-1} (pnp x np); the padding stays zero
      HY_CUDA(cudaMemsetAsync(a21d_.p, 0, (size_t)np * pnp * sizeof(double), s));
      HY_CUDA(cudaMemsetAsync(a12d_.p, 0, (size_t)pnp * np * sizeof(double), s));
      k_csr_block_dense<<<(nb + 127) / 128, 128, 0, s>>>(C.loPtr.p, C.loCol.p, C.loSrc.p, val, a21d_.p, C.lvlPtr[b], nb,
                                                          C.lvlPtr[b - 1], pn, pnp);
      k_csr_block_dense<<<(pn + 127) / 128, 128, 0, s>>>(C.upPtr.p, C.upCol.p, C.upSrc.p, val, a12d_.p, C.lvlPtr[b - 1],
                                                          pn, C.lvlPtr[b], nb, np);
      launches_ += 2;
      const double* Fprev = C.inv.F.p + C.inv.hMatOff[b - 1];
      denseGemm(Fprev, pnp, a12d_.p, np, dmat_.p, np, pnp, np, pnp, 1.0, 0.0, s, &launches_);     // T = F_{b-1} U_{b-1}
      denseGemm(a21d_.p, pnp, dmat_.p, np, work_.p, np, np, np, pnp, -1.0, 1.0, s, &launches_);  // D_b -= L_b T
      stats_.flops_compute += 2.0 * (double)pnp * pnp * np + 2.0 * (double)np * np * pnp;
    }
    if (refine_) HY_CUDA(cudaMemcpyAsync(blkA_.p, work_.p, len * sizeof(double), cudaMemcpyDeviceToDevice, s));
    auto refill = [&]() {
      HY_CUDA(cudaMemcpyAsync(work_.p, blkA_.p, len * sizeof(double), cudaMemcpyDeviceToDevice, s));
    };
    // (the padding identity of work_ is written by the inversion itself)
    std::vector<int64_t> rel(1, 0);
    relOff_.upload(rel, s);
    piv_.alloc((size_t)np + 128);
    perm_.alloc((size_t)np);
    invertBatched(work_.p, C.inv.F.p + C.inv.hMatOff[b], relOff_.p, C.inv.n.p + b, C.inv.np.p + b, 1, np, piv_.p, perm_.p,
                  piv_.p + np, info_.p, s, &launches_);
    if (refine_) {
      refill();
      refineInverseBatched(work_.p, C.inv.F.p + C.inv.hMatOff[b], work2_.p, relOff_.p, C.inv.n.p + b, C.inv.np.p + b, 1, np,
                           s, &launches_);
    }
    HY_CUDA(cudaStreamSynchronize(s));
    stats_.flops_compute += 2.0 * (double)nb * nb * nb;
  }
  checkInfo("coarse solver (block-tridiagonal factorization)");
  C.active = true;
}

// sol = S^-1 rhs with the block-tridiagonal factors (rhs, sol in the coarse numbering; may alias)
void Engine::solveCoarseBT(const double* rhs, double* sol) {
  CoarseBT& C = coarseBT_;
  cudaStream_t s = stream_;
  const int m = C.m;
  packIdx(rhs, C.dPerm.p, C.y.p, C.n, s, &launches_);
  auto gemvBlock = [&](int b, const double* xin, const double* xprev, double* out, int mode) {
    GemvArgs a = C.inv.args();
    a.itemMat = C.inv.itemMat.p + C.itemPtr[b];
    a.itemRow0 = C.inv.itemRow0.p + C.itemPtr[b];
    a.xin = xin;
    a.xprev = xprev;
    a.out = out;
    a.mode = mode;
    batchedGemv(a, C.itemPtr[b + 1] - C.itemPtr[b], C.inv.hNp[b], s, &launches_);
  };
  for (int b = 0; b < m; ++b) {
    const int a0 = C.lvlPtr[b], nb = C.lvlPtr[b + 1] - a0;
    if (b > 0)  // y_b -= L_b t_{b-1}
      spmv(C.loPtr.p + a0, C.loCol.p, C.loVal.p, C.t.p, C.y.p + a0, nb, 1.0, C.y.p + a0, nullptr, -1.0, s, &launches_);
    gemvBlock(b, C.y.p, nullptr, C.t.p, 0);  // t_b = F_b y_b
  }
  const int aL = C.lvlPtr[m - 1];
  HY_CUDA(cudaMemcpyAsync(C.xp.p + aL, C.t.p + aL, (size_t)(C.n - aL) * sizeof(double), cudaMemcpyDeviceToDevice, s));
  for (int b = m - 2; b >= 0; --b) {
    const int a0 = C.lvlPtr[b], nb = C.lvlPtr[b + 1] - a0;
    spmv(C.upPtr.p + a0, C.upCol.p, C.upVal.p, C.xp.p, C.r.p + a0, nb, 0.0, nullptr, nullptr, 1.0, s, &launches_);
    gemvBlock(b, C.r.p, C.t.p, C.xp.p, 1);   // x_b = t_b - F_b (U_b x_{b+1})
  }
  scatterVec(C.xp.p, C.dPerm.p, sol, C.n, s, &launches_);
}

}  // namespace hymls
