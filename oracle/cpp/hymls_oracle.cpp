// ORACLE -- test infrastructure and CPU baseline only.  Nothing under hymls_b200/ may include, link or load this.
//
// C++/OpenMP restatement of the reference's CPU path (one process x T threads standing in for `mpirun -np T`;
// the reference cannot be built here: it needs Trilinos + MPI):
//   Preconditioner::Compute / ApplyInverse            src/HYMLS_Preconditioner.cpp:400-517, 930-1070
//   MatrixBlock (A11/A12/A21/A22, subdomain solvers)   src/HYMLS_MatrixBlock.cpp:74-385
//   SparseDirectSolver (KLU: left-looking sparse LU)   src/HYMLS_SparseDirectSolver.cpp:735-856
//   SchurComplement::Construct11 / Construct22         src/HYMLS_SchurComplement.cpp:131-306
//   SchurPreconditioner::{InitializeOT, InitializeBlocks, AssembleTransformAndDrop, ConstructSCPart,
//        ComputeNextLevel, ApplyInverse, ApplyOT, ApplyBlockDiagonal}
//                                                      src/HYMLS_SchurPreconditioner.cpp:234-1093
//   Householder::{Apply, ApplyR, Construct}            src/HYMLS_Householder.cpp:38-163
//   CoarseSolver::{Compute, ApplyInverse}              src/HYMLS_CoarseSolver.cpp:131-323
//   MatrixUtils::{DropByValue, PutDirichlet}           src/HYMLS_MatrixUtils.cpp:1010-1309
//   BaseSolver::ApplyInverse -> Belos GMRES / CG       src/HYMLS_BaseSolver.cpp:309-359
// The same algorithm as oracle/hymls.py + oracle/krylov.py (which tests/ pin to the reference's fixtures and
// integration targets); tests/test_oracle_cpp.py checks this file against them.  The index maps (interior
// nodes and ordered separator groups per subdomain and level) are INPUT: from oracle/partitioner.py in the
// tests, from the library's host partitioner (bit-exact with the former, tests/test_host_maps.py) at sizes the
// Python partitioner cannot reach.  The per-subdomain solver is a left-looking sparse LU (Gilbert-Peierls, the
// algorithm inside KLU) in the reference's F-matrix ordering with its scaling and static pivots
//   MatrixUtils::FillReducingOrdering                  src/HYMLS_MatrixUtils.cpp:1311-1740
//   SparseDirectSolver::ComputeScaling, KLU settings   src/HYMLS_SparseDirectSolver.cpp:238-254, 632-664
// pinned to the reference's known-answer fill counts (unit_tests/HYMLS_SparseDirectSolver.cpp:117-152).
// Not covered: bordering, deflation, non-default variants (oracle/hymls.py covers bordering at small sizes).
#include <omp.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <queue>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace ho {

static const double SMALL = 1e-14;  // HYMLS_SMALL_ENTRY, src/HYMLS_Macros.hpp:29

struct Csr {
  int64_t n = 0, m = 0;  // rows, cols
  std::vector<int64_t> ptr;
  std::vector<int> col;
  std::vector<double> val;
  void matvec(const double* x, double* y) const {  // y = A x
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      double s = 0;
      for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) s += val[e] * x[col[e]];
      y[r] = s;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Sparse LU, left looking (Gilbert-Peierls).  General matrices (coarse solver): column ordering = minimum degree on
// A + A', partial pivoting with a preference for the diagonal (threshold 0.001 like KLU's default).  Subdomain
// matrices: the reference's F-matrix ordering, scaling and static pivots (fmatrixOrdering below).
// ---------------------------------------------------------------------------------------------
static std::vector<int> minimumDegreeOrderAdj(std::vector<std::vector<int>>& adj);
static std::vector<int> minimumDegreeOrder(int n, const std::vector<int>& Ap, const std::vector<int>& Ai) {
  std::vector<std::vector<int>> adj(n);
  for (int j = 0; j < n; ++j)
    for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
      const int i = Ai[p];
      if (i != j) {
        adj[i].push_back(j);
        adj[j].push_back(i);
      }
    }
  return minimumDegreeOrderAdj(adj);
}
// adj: undirected graph as (possibly unsorted, repeated) neighbour lists without self loops; consumed
static std::vector<int> minimumDegreeOrderAdj(std::vector<std::vector<int>>& adj) {
  const int n = (int)adj.size();
  for (auto& a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  typedef std::pair<int, int> DI;  // (degree, node), lazy deletion
  std::priority_queue<DI, std::vector<DI>, std::greater<DI>> heap;
  for (int v = 0; v < n; ++v) heap.push(DI((int)adj[v].size(), v));
  std::vector<char> done(n, 0);
  std::vector<int> order, merged;
  order.reserve(n);
  while (!heap.empty()) {
    const DI top = heap.top();
    heap.pop();
    const int v = top.second;
    if (done[v] || top.first != (int)adj[v].size()) continue;
    done[v] = 1;
    order.push_back(v);
    const std::vector<int> nb = adj[v];
    for (int u : nb) {  // neighbours of v become a clique, v disappears
      merged.clear();
      std::set_union(adj[u].begin(), adj[u].end(), nb.begin(), nb.end(), std::back_inserter(merged));
      merged.erase(std::remove_if(merged.begin(), merged.end(), [&](int w) { return w == u || w == v; }), merged.end());
      adj[u].swap(merged);
      heap.push(DI((int)adj[u].size(), u));
    }
    std::vector<int>().swap(adj[v]);
  }
  return order;
}

// ---------------------------------------------------------------------------------------------
// Ordering and scaling of the reference's subdomain solver ("Custom Ordering" / "Custom Scaling", both default true,
// src/HYMLS_SparseDirectSolver.cpp:238-239): MatrixUtils::FillReducingOrdering (src/HYMLS_MatrixUtils.cpp:1311-1740)
// and SparseDirectSolver::ComputeScaling (src/HYMLS_SparseDirectSolver.cpp:632-664).
//   * nodes with a zero diagonal are P-nodes, the others V-nodes (:1338-1363); the matrix must be an F-matrix: every
//     V-row has at most two P-entries (:1414);
//   * a fill-reducing ordering q of the V-nodes on the graph of A_VV + B B' (:1428-1432, AMD in the reference, exact
//     minimum degree here);
//   * walking through q, a V-node that still couples to two different (merged) P-nodes takes one of them along right
//     behind it -- the one with fewer couplings; merged P-nodes are tracked with a union-find (:1644-1694) -- and the two
//     ROWS are interchanged so that the pivots are  b 0 / a b  (:1690-1692); P-nodes never picked come last (:1706-1717);
//   * rows and columns of P-nodes are scaled by the largest diagonal entry.
// KLU then factors in exactly this order (ordering = 2: given, pivot tolerance 0: the diagonal whenever it is non-zero,
// :244-254).  Returns false when the matrix is not of that form (the caller orders it like a general matrix).
// ---------------------------------------------------------------------------------------------
static bool fmatrixOrdering(int N, const std::vector<int>& Ap, const std::vector<int>& Ai, const std::vector<double>& Ax,
                            std::vector<int>& rowperm, std::vector<int>& colperm, std::vector<double>& scale) {
  std::vector<double> diag(N, 0.0);
  // row-wise pattern (columns ascending)
  std::vector<int> Rp(N + 1, 0), Rc(Ai.size());
  for (int i : Ai) Rp[i + 1]++;
  for (int i = 0; i < N; ++i) Rp[i + 1] += Rp[i];
  {
    std::vector<int> fill(Rp.begin(), Rp.end() - 1);
    for (int j = 0; j < N; ++j)
      for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
        Rc[fill[Ai[p]]++] = j;
        if (Ai[p] == j) diag[j] = Ax[p];
      }
  }
  std::vector<int> vIdx(N, -1), pIdx(N, -1), V, P;
  double dmax = 0.0;
  for (int i = 0; i < N; ++i) {
    dmax = std::max(dmax, std::fabs(diag[i]));
    if (diag[i] == 0.0) { pIdx[i] = (int)P.size(); P.push_back(i); }
    else { vIdx[i] = (int)V.size(); V.push_back(i); }
  }
  const int n = (int)V.size(), m = (int)P.size();
  scale.assign(N, 1.0);
  for (int i = 0; i < N; ++i)
    if (std::fabs(diag[i]) <= SMALL * dmax) scale[i] = dmax;
  rowperm.assign(N, 0);
  colperm.assign(N, 0);
  if (m == 0) {
    std::vector<int> q = minimumDegreeOrder(N, Ap, Ai);
    rowperm = q;
    colperm = q;
    return true;
  }
  // Gr: the (at most two) P-nodes of every V-row; cont: V-couplings of every P-row
  std::vector<std::array<int, 2>> Gr(n, std::array<int, 2>{m, m});
  std::vector<int> cont(m, 0);
  int maxB = 0;
  for (int i = 0; i < n; ++i) {
    int cnt = 0;
    for (int e = Rp[V[i]]; e < Rp[V[i] + 1]; ++e) {
      const int j = pIdx[Rc[e]];
      if (j < 0) continue;
      if (Gr[i][0] == m) Gr[i][0] = j; else Gr[i][1] = j;
      ++cnt;
    }
    maxB = std::max(maxB, cnt);
  }
  if (maxB != 1 && maxB != 2) return false;
  for (int j = 0; j < m; ++j)
    for (int e = Rp[P[j]]; e < Rp[P[j] + 1]; ++e)
      if (vIdx[Rc[e]] >= 0) cont[j]++;
  // graph of A_VV + B Bt, Bt = A_PV
  std::vector<std::vector<int>> adj(n);
  for (int i = 0; i < n; ++i)
    for (int e = Rp[V[i]]; e < Rp[V[i] + 1]; ++e) {
      const int c = Rc[e];
      if (vIdx[c] >= 0) {
        if (vIdx[c] != i) { adj[i].push_back(vIdx[c]); adj[vIdx[c]].push_back(i); }
      } else {
        for (int f = Rp[c]; f < Rp[c + 1]; ++f) {
          const int k = vIdx[Rc[f]];
          if (k >= 0 && k != i) { adj[i].push_back(k); adj[k].push_back(i); }
        }
      }
    }
  const std::vector<int> q = minimumDegreeOrderAdj(adj);
  std::vector<int> pid(m + 1), symperm(N, -1), perm(N);
  for (int i = 0; i <= m; ++i) pid[i] = i;
  for (int i = 0; i < N; ++i) perm[i] = i;
  int jj = 0;
  for (int i = 0; i < n; ++i) {
    const int qi = q[i];
    symperm[jj] = V[qi];
    int g1 = Gr[qi][0], g2 = Gr[qi][1];
    while (pid[g1] != g1) g1 = pid[g1];
    while (pid[g2] != g2) g2 = pid[g2];
    if (g1 == g2) { jj += 1; continue; }  // no P-coupling (left)
    int take;
    if (g1 == m) { pid[g2] = pid[g1]; take = g2; }
    else if (g2 == m) { pid[g1] = pid[g2]; take = g1; }
    else if (cont[g2] > cont[g1]) { pid[g1] = pid[g2]; take = g1; cont[g2] = cont[g1] + cont[g2] - 2; }
    else { pid[g2] = pid[g1]; take = g2; cont[g1] = cont[g1] + cont[g2] - 2; }
    symperm[jj + 1] = P[take];
    perm[jj] = jj + 1;
    perm[jj + 1] = jj;
    jj += 2;
  }
  std::vector<char> placed(N, 0);
  for (int i = 0; i < jj; ++i) placed[symperm[i]] = 1;
  for (int i = 0; i < N; ++i)
    if (!placed[i]) symperm[jj++] = i;
  if (jj != N) return false;
  for (int i = 0; i < N; ++i) {
    colperm[i] = symperm[i];
    rowperm[i] = symperm[perm[i]];
  }
  return true;
}

struct SparseLU {
  int n = 0;
  std::vector<int> Lp, Li, Up, Ui, pinv, q;  // L unit lower (diagonal first in every column), U (diagonal last)
  std::vector<double> Lx, Ux;
  std::vector<double> scale;  // F-matrix path: rows and columns of the P-nodes scaled (empty: none)
  bool fOrdered = false;      // factored in the reference's F-matrix order with static pivots
  // A in CSC (Ap, Ai, Ax).  Returns false when a pivot column is exactly zero.
  // fmatrix: order, scale and pivot like the reference's subdomain solver when the matrix has that form
  bool factor(int n_, const std::vector<int>& Ap, const std::vector<int>& Ai, const std::vector<double>& AxIn,
              bool fmatrix = false) {
    n = n_;
    std::vector<int> prefRow;  // preferred pivot row of elimination step k (the diagonal of the permuted matrix)
    std::vector<double> AxScaled;
    scale.clear();
    fOrdered = false;
    if (fmatrix) {
      std::vector<int> rp, cp;
      if (fmatrixOrdering(n, Ap, Ai, AxIn, rp, cp, scale)) {
        q = cp;
        prefRow = rp;
        fOrdered = true;
        AxScaled = AxIn;
        for (int j = 0; j < n; ++j)
          for (int p = Ap[j]; p < Ap[j + 1]; ++p) AxScaled[p] *= scale[Ai[p]] * scale[j];
      } else {
        scale.clear();
      }
    }
    const std::vector<double>& Ax = fOrdered ? AxScaled : AxIn;
    if (!fOrdered) q = minimumDegreeOrder(n, Ap, Ai);
    pinv.assign(n, -1);
    Lp.assign(n + 1, 0);
    Up.assign(n + 1, 0);
    Li.clear(); Lx.clear(); Ui.clear(); Ux.clear();
    Li.reserve(4 * Ai.size()); Lx.reserve(4 * Ai.size()); Ui.reserve(4 * Ai.size()); Ux.reserve(4 * Ai.size());
    std::vector<double> x(n, 0.0);
    std::vector<int> xi(2 * n), stackPos(n), mark(n, -1);
    for (int k = 0; k < n; ++k) {
      Lp[k] = (int)Li.size();
      Up[k] = (int)Ui.size();
      const int col = q[k];
      // symbolic: reach of the column pattern in the graph of L (depth-first search, topological order)
      int top = n;
      for (int p = Ap[col]; p < Ap[col + 1]; ++p) {
        const int start = Ai[p];
        if (mark[start] == k) continue;
        int head = 0;
        xi[0] = start;
        while (head >= 0) {
          const int j = xi[head];
          const int jn = pinv[j];
          if (mark[j] != k) {
            mark[j] = k;
            stackPos[head] = jn < 0 ? 0 : Lp[jn] + 1;  // skip the unit diagonal
          }
          bool finished = true;
          const int pend = jn < 0 ? 0 : Lp[jn + 1];
          for (int p2 = stackPos[head]; p2 < pend; ++p2) {
            const int i = Li[p2];
            if (mark[i] == k) continue;
            stackPos[head] = p2 + 1;
            xi[++head] = i;
            finished = false;
            break;
          }
          if (finished) {
            --head;
            xi[--top] = j;  // xi[top..n-1] holds the output, the DFS stack lives in xi[0..head]
          }
        }
      }
      // the DFS stack and the output share xi: they cannot collide because head + (n - top) <= n
      for (int p = top; p < n; ++p) x[xi[p]] = 0.0;
      for (int p = Ap[col]; p < Ap[col + 1]; ++p) x[Ai[p]] = Ax[p];
      // numeric: x = L \ A(:, col) restricted to the reach
      for (int px = top; px < n; ++px) {
        const int j = xi[px];
        const int jn = pinv[j];
        if (jn < 0) continue;
        const double xj = x[j];
        for (int p = Lp[jn] + 1; p < Lp[jn + 1]; ++p) x[Li[p]] -= Lx[p] * xj;
      }
      // pivot: largest non-pivotal entry, the diagonal if it is within 0.001 of it
      int ipiv = -1;
      double a = -1.0;
      for (int p = top; p < n; ++p) {
        const int i = xi[p];
        if (pinv[i] < 0) {
          const double t = std::fabs(x[i]);
          if (t > a) { a = t; ipiv = i; }
        } else {
          Ui.push_back(pinv[i]);
          Ux.push_back(x[i]);
        }
      }
      if (ipiv < 0 || a <= 0.0) return false;
      if (fOrdered) {  // given order, pivot tolerance 0: the diagonal of the permuted matrix whenever it is non-zero
        const int pr = prefRow[k];
        if (pinv[pr] < 0 && mark[pr] == k && x[pr] != 0.0) ipiv = pr;
      } else if (pinv[col] < 0 && mark[col] == k && std::fabs(x[col]) >= 0.001 * a) {
        ipiv = col;
      }
      const double pivot = x[ipiv];
      Ui.push_back(k);
      Ux.push_back(pivot);
      pinv[ipiv] = k;
      Li.push_back(ipiv);
      Lx.push_back(1.0);
      for (int p = top; p < n; ++p) {
        const int i = xi[p];
        if (pinv[i] < 0) {
          Li.push_back(i);
          Lx.push_back(x[i] / pivot);
        }
        x[i] = 0.0;
      }
    }
    Lp[n] = (int)Li.size();
    Up[n] = (int)Ui.size();
    for (size_t p = 0; p < Li.size(); ++p) Li[p] = pinv[Li[p]];  // row indices of L in pivot order
    return true;
  }
  // b <- A^-1 b  (work: n doubles)
  void solve(double* b, double* work) const {
    if (!scale.empty())
      for (int i = 0; i < n; ++i) b[i] *= scale[i];
    for (int i = 0; i < n; ++i) work[pinv[i]] = b[i];
    for (int j = 0; j < n; ++j) {
      const double xj = work[j];
      if (xj != 0.0)
        for (int p = Lp[j] + 1; p < Lp[j + 1]; ++p) work[Li[p]] -= Lx[p] * xj;
    }
    for (int j = n - 1; j >= 0; --j) {
      const double xj = work[j] / Ux[Up[j + 1] - 1];
      work[j] = xj;
      if (xj != 0.0)
        for (int p = Up[j]; p < Up[j + 1] - 1; ++p) work[Ui[p]] -= Ux[p] * xj;
    }
    for (int k = 0; k < n; ++k) b[q[k]] = work[k];
    if (!scale.empty())
      for (int i = 0; i < n; ++i) b[i] *= scale[i];
  }
  int64_t nnzFactors() const { return (int64_t)Li.size() + (int64_t)Ui.size(); }
};

// Optional iterative refinement of the direct solves (ho_set_refinement): residuals in x87 extended precision
// (long double), corrections with the FP64 factors.  Used by tests that compare with the GPU path at sizes where
// the plain FP64 LU solves of this oracle are themselves 1e-10 away from the exact-arithmetic result
// (tests/test_gpu_baseline_sizes.py); never by the timed CPU baseline.
static void refinedSolve(const SparseLU& lu, const std::vector<int64_t>& ptr, const std::vector<int>& col,
                         const std::vector<double>& val, int steps, double* b, double* work, std::vector<double>& tmp) {
  const int n = lu.n;
  if (steps <= 0) { lu.solve(b, work); return; }
  std::vector<double> rhs(b, b + n);
  lu.solve(b, work);
  tmp.resize(n);
  for (int it = 0; it < steps; ++it) {
    for (int i = 0; i < n; ++i) {
      long double r = rhs[i];
      for (int64_t e = ptr[i]; e < ptr[i + 1]; ++e) r -= (long double)val[e] * (long double)b[col[e]];
      tmp[i] = (double)r;
    }
    lu.solve(tmp.data(), work);
    for (int i = 0; i < n; ++i) b[i] += tmp[i];
  }
}

// dense LU with partial pivoting (Ifpack_DenseContainer -> dgetrf / dgetrs), row major
struct DenseLU {
  int n = 0;
  std::vector<double> a;
  std::vector<int> piv;
  bool factor(int n_, const double* M) {
    n = n_;
    a.assign(M, M + (size_t)n * n);
    piv.resize(n);
    for (int k = 0; k < n; ++k) {
      int p = k;
      double best = std::fabs(a[(size_t)k * n + k]);
      for (int i = k + 1; i < n; ++i)
        if (std::fabs(a[(size_t)i * n + k]) > best) { best = std::fabs(a[(size_t)i * n + k]); p = i; }
      piv[k] = p;
      if (best == 0.0) return false;
      if (p != k)
        for (int j = 0; j < n; ++j) std::swap(a[(size_t)k * n + j], a[(size_t)p * n + j]);
      const double d = 1.0 / a[(size_t)k * n + k];
      for (int i = k + 1; i < n; ++i) {
        const double l = a[(size_t)i * n + k] * d;
        a[(size_t)i * n + k] = l;
        if (l != 0.0)
          for (int j = k + 1; j < n; ++j) a[(size_t)i * n + j] -= l * a[(size_t)k * n + j];
      }
    }
    return true;
  }
  std::vector<double> orig;  // kept when refinement is requested
  void solveRefined(double* b, int steps) const {
    if (steps <= 0 || orig.empty()) { solve(b); return; }
    std::vector<double> rhs(b, b + n), r(n);
    solve(b);
    for (int it = 0; it < steps; ++it) {
      for (int i = 0; i < n; ++i) {
        long double t = rhs[i];
        for (int j = 0; j < n; ++j) t -= (long double)orig[(size_t)i * n + j] * (long double)b[j];
        r[i] = (double)t;
      }
      solve(r.data());
      for (int i = 0; i < n; ++i) b[i] += r[i];
    }
  }
  void solve(double* b) const {
    for (int k = 0; k < n; ++k)
      if (piv[k] != k) std::swap(b[k], b[piv[k]]);
    for (int i = 1; i < n; ++i) {
      double s = b[i];
      for (int j = 0; j < i; ++j) s -= a[(size_t)i * n + j] * b[j];
      b[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = b[i];
      for (int j = i + 1; j < n; ++j) s -= a[(size_t)i * n + j] * b[j];
      b[i] = s / a[(size_t)i * n + i];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Householder (src/HYMLS_Householder.cpp)
// ---------------------------------------------------------------------------------------------
static double sgn(double x) { return x < 0 ? -1.0 : (x > 0 ? 1.0 : 0.0); }  // :15-18, sign(0) == 0

// Householder::Apply(SerialDenseMatrix, v) :38-80 on rows [a, a+len) of the ld x ncols row-major matrix X;
// colStride/rowStride let the same code do ApplyR (:83-126) on the transpose
static void householderDense(double* X, int64_t rowStride, int64_t colStride, int a, int len, int ncols,
                             const double* vin, std::vector<double>& v, std::vector<double>& fac) {
  v.assign(vin, vin + len);
  const double s = sgn(v[0]);
  double nrm2 = 0;
  for (int i = 0; i < len; ++i) { v[i] *= s; nrm2 += v[i] * v[i]; }
  const double nrmv = std::sqrt(nrm2);
  const double v1 = v[0] + nrmv;
  if (std::fabs(v1) < SMALL || nrmv < SMALL) return;
  const double fac1 = 1.0 / (nrmv * v1);
  fac.assign(ncols, 0.0);
  for (int i = 0; i < len; ++i) {
    const double* row = X + (int64_t)(a + i) * rowStride;
    const double vi = v[i];
    for (int j = 0; j < ncols; ++j) fac[j] += vi * row[j * colStride];
  }
  double* row0 = X + (int64_t)a * rowStride;
  for (int j = 0; j < ncols; ++j) fac[j] = fac1 * (nrmv * row0[j * colStride] + fac[j]);
  for (int i = 1; i < len; ++i) {
    double* row = X + (int64_t)(a + i) * rowStride;
    const double vi = v[i];
    for (int j = 0; j < ncols; ++j) row[j * colStride] = vi * fac[j] - row[j * colStride];
  }
  for (int j = 0; j < ncols; ++j) row0[j * colStride] = v1 * fac[j] - row0[j * colStride];
}

// Householder::Construct :128-163: normalised reflector (empty when degenerate)
static std::vector<double> householderW(const double* vin, int len) {
  std::vector<double> v(vin, vin + len);
  double nrm2 = 0;
  for (double t : v) nrm2 += t * t;
  const double nrm = std::sqrt(nrm2);
  const double s = sgn(v[0]);
  for (double& t : v) t *= s;
  v[0] += nrm;
  nrm2 = 0;
  for (double t : v) nrm2 += t * t;
  const double n2 = std::sqrt(nrm2);
  if (n2 < SMALL) return std::vector<double>();
  for (double& t : v) t /= n2;
  return v;
}

// ---------------------------------------------------------------------------------------------
// MatrixUtils::DropByValue :1010-1194 (in place on a CSR matrix; dropped entries are removed)
//   RelDropDiag: keep |a_ij| > tol max(|a_ii|,|a_jj|) and > tol; diagonal entries: keep when > tol
//   RelFullDiag: the same off the diagonal; every diagonal entry is kept (set to 0 when <= tol)
// ---------------------------------------------------------------------------------------------
static void dropByValue(Csr& A, double tol, bool fullDiag) {
  const int64_t n = A.n;
  std::vector<double> diag(n, 0.0);
  for (int64_t r = 0; r < n; ++r)
    for (int64_t e = A.ptr[r]; e < A.ptr[r + 1]; ++e)
      if (A.col[e] == r) diag[r] = A.val[e];
  std::vector<int64_t> ptr(n + 1, 0);
  std::vector<int> col;
  std::vector<double> val;
  col.reserve(A.col.size());
  val.reserve(A.val.size());
  for (int64_t r = 0; r < n; ++r) {
    bool haveDiag = false;
    for (int64_t e = A.ptr[r]; e < A.ptr[r + 1]; ++e) {
      const int c = A.col[e];
      const double v = A.val[e], av = std::fabs(v);
      if (c == r) {
        haveDiag = true;
        if (fullDiag) { col.push_back(c); val.push_back(av > tol ? v : 0.0); }
        else if (av > tol) { col.push_back(c); val.push_back(v); }
        continue;
      }
      const double scal = std::max(std::fabs(diag[r]), std::fabs(diag[c]));
      if (av > scal * tol && av > tol) { col.push_back(c); val.push_back(v); }
    }
    if (fullDiag && !haveDiag) {  // the full diagonal is inserted even where the pattern had none
      int64_t p = (int64_t)col.size();
      col.push_back((int)r);
      val.push_back(0.0);
      while (p > ptr[r] && col[p - 1] > col[p]) { std::swap(col[p - 1], col[p]); std::swap(val[p - 1], val[p]); --p; }
    }
    ptr[r + 1] = (int64_t)col.size();
  }
  A.ptr.swap(ptr);
  A.col.swap(col);
  A.val.swap(val);
}

// MatrixUtils::PutDirichlet :1229-1309: row -> unit row, its column entries (rows found through the row's own
// pattern, i.e. a structurally symmetric matrix is assumed like in the reference) -> 0
static void putDirichlet(Csr& A, int row) {
  std::vector<int> cols(A.col.begin() + A.ptr[row], A.col.begin() + A.ptr[row + 1]);
  for (int64_t e = A.ptr[row]; e < A.ptr[row + 1]; ++e) A.val[e] = (A.col[e] == row) ? 1.0 : 0.0;
  for (int r : cols)
    if (r != row)
      for (int64_t e = A.ptr[r]; e < A.ptr[r + 1]; ++e)
        if (A.col[e] == row) A.val[e] = 0.0;
}

static bool factorCsr(const Csr& A, SparseLU& lu, bool fmatrix = false) {  // CSR -> CSC -> LU
  const int n = (int)A.n;
  std::vector<int> Ap(n + 1, 0), Ai(A.col.size());
  std::vector<double> Ax(A.col.size());
  for (int c : A.col) Ap[c + 1]++;
  for (int j = 0; j < n; ++j) Ap[j + 1] += Ap[j];
  std::vector<int> fill(Ap.begin(), Ap.end() - 1);
  for (int r = 0; r < n; ++r)
    for (int64_t e = A.ptr[r]; e < A.ptr[r + 1]; ++e) {
      const int p = fill[A.col[e]]++;
      Ai[p] = r;
      Ax[p] = A.val[e];
    }
  return lu.factor(n, Ap, Ai, Ax, fmatrix);
}

// ---------------------------------------------------------------------------------------------
struct Partition {  // index maps of one level (input)
  int nsd = 0;
  std::vector<int64_t> intPtr, intGid;  // interior GIDs per subdomain (sorted)
  std::vector<int64_t> sdGrpPtr;        // nsd+1: groups of sd (ALL separator groups around it, reference order)
  std::vector<int64_t> grpPtr, grpGid;  // nodes of every group (sorted)
  std::vector<int> grpType;
};

struct Timings {
  double factorA11 = 0, schur = 0, blocks = 0, coarse = 0;
};

class Solver;

class Level {
 public:
  int level = 0, maxLevel = 1;
  Csr A;
  std::vector<int64_t> gids;
  std::vector<double> tv;
  Partition part;
  std::vector<int64_t> fixGids;
  // orderings
  int64_t nI = 0, nS = 0;
  std::vector<int> intRow, sepRow;
  int nuniq = 0;
  std::vector<int64_t> uniqPtr;          // nuniq+1, positions in the separator ordering
  std::vector<int> grpUniq;              // unique group of every (sd, group)
  std::vector<int> uniqOwner;
  std::vector<std::vector<double>> uniqW;  // reflector per unique group (empty: degenerate)
  std::vector<double> tvSep;
  std::vector<int64_t> sdSepPtr;         // nsd+1
  std::vector<int> sdSep;                // separator positions of all separator nodes around sd, group order
  // blocks of non-V-sum nodes: one per (owner subdomain, linked set of its own groups)
  std::vector<std::vector<int>> blocks;
  std::vector<int> sepBlk, sepBlkIdx;    // per separator position (-1 for V-sums)
  std::vector<DenseLU> blockLU;
  // matrix blocks
  Csr A12, A21, A22;
  std::vector<SparseLU> sdLU;
  std::vector<Csr> sdA;                  // the A11 blocks themselves (refinement only)
  Csr coarseA;
  int refineSteps = 0;
  // subdomain solvers ordered / scaled / pivoted like the reference's ("Custom Ordering", "Custom Scaling" = true);
  // false: general minimum-degree ordering with threshold partial pivoting (ho_set_subdomain_ordering)
  bool fmatrixSolver = true;
  // next level / coarse
  std::unique_ptr<Level> next;
  Csr reduced;                           // reduced Schur complement on the V-sums after dropping
  SparseLU coarseLU;
  std::vector<int> coarseFixRows;
  std::vector<int64_t> vsumGids;
  int64_t factorNnz = 0;
  Timings tm;

  void initialize();
  void compute(std::vector<Partition>& parts, const std::vector<int64_t>& fix);
  void applyInverse(const double* b, double* x) const;
  void schurApplyInverse(const double* rhs, double* sol) const;
  void applyOT(const double* v, double* out) const;

 private:
  void a11Solve(const double* b, double* x) const;
  void assemble(std::vector<double>& redVal, const std::vector<int64_t>& redPtr, const std::vector<int>& redCol);
};

static std::vector<std::vector<int>> linkByType(const std::vector<int>& types, const std::vector<int>& ids) {
  // HierarchicalMap::LinkSeparators :120-142: equal type >= 0 -> one set, first-seen order
  std::vector<std::vector<int>> out;
  for (int gi : ids) {
    bool found = false;
    if (types[gi] >= 0)
      for (auto& lg : out)
        if (types[lg[0]] == types[gi]) { lg.push_back(gi); found = true; break; }
    if (!found) out.push_back(std::vector<int>(1, gi));
  }
  return out;
}

void Level::initialize() {
  const int64_t n = A.n;
  const int nsd = part.nsd;
  std::unordered_map<int64_t, int> g2r;
  g2r.reserve((size_t)n * 2);
  for (int64_t r = 0; r < n; ++r) g2r[gids[r]] = (int)r;
  auto rowOf = [&](int64_t g) {
    auto it = g2r.find(g);
    if (it == g2r.end()) throw std::runtime_error("partition refers to a GID that is not a row of the matrix");
    return it->second;
  };
  if (tv.empty()) tv.assign(n, 1.0);
  // interior ordering = concatenation of the subdomain interiors (SpawnInterior :436-466)
  nI = part.intPtr[nsd];
  intRow.resize(nI);
  for (int64_t p = 0; p < nI; ++p) intRow[p] = rowOf(part.intGid[p]);
  // unique groups: the first subdomain listing a group (identified by its first GID) owns it (:248-275)
  const int64_t ngrp = (int64_t)part.grpType.size();
  std::unordered_map<int64_t, int> byFirst;
  grpUniq.assign(ngrp, -1);
  uniqPtr.assign(1, 0);
  uniqOwner.clear();
  sepRow.clear();
  for (int sd = 0; sd < nsd; ++sd)
    for (int64_t g = part.sdGrpPtr[sd]; g < part.sdGrpPtr[sd + 1]; ++g) {
      const int64_t first = part.grpGid[part.grpPtr[g]];
      auto it = byFirst.find(first);
      if (it == byFirst.end()) {
        const int u = (int)uniqOwner.size();
        byFirst.emplace(first, u);
        uniqOwner.push_back(sd);
        for (int64_t q = part.grpPtr[g]; q < part.grpPtr[g + 1]; ++q) sepRow.push_back(rowOf(part.grpGid[q]));
        uniqPtr.push_back((int64_t)sepRow.size());
        grpUniq[g] = u;
      } else {
        grpUniq[g] = it->second;
      }
    }
  nuniq = (int)uniqOwner.size();
  nS = (int64_t)sepRow.size();
  if (nI + nS != n) throw std::runtime_error("partition does not cover the map");
  std::vector<int> posSep(n, -1);
  for (int64_t p = 0; p < nS; ++p) posSep[sepRow[p]] = (int)p;
  tvSep.resize(nS);
  for (int64_t p = 0; p < nS; ++p) tvSep[p] = tv[sepRow[p]];
  // all separator nodes around every subdomain, in group order
  sdSepPtr.assign(nsd + 1, 0);
  sdSep.clear();
  for (int sd = 0; sd < nsd; ++sd) {
    for (int64_t g = part.sdGrpPtr[sd]; g < part.sdGrpPtr[sd + 1]; ++g)
      for (int64_t q = part.grpPtr[g]; q < part.grpPtr[g + 1]; ++q) sdSep.push_back(posSep[rowOf(part.grpGid[q])]);
    sdSepPtr[sd + 1] = (int64_t)sdSep.size();
  }
  // reflectors (InitializeOT :384-467) and blocks (InitializeBlocks :301-340) of the owned groups
  uniqW.assign(nuniq, std::vector<double>());
  for (int u = 0; u < nuniq; ++u)
    uniqW[u] = householderW(tvSep.data() + uniqPtr[u], (int)(uniqPtr[u + 1] - uniqPtr[u]));
  blocks.clear();
  sepBlk.assign(nS, -1);
  sepBlkIdx.assign(nS, -1);
  for (int sd = 0; sd < nsd; ++sd) {
    const int64_t g0 = part.sdGrpPtr[sd], g1 = part.sdGrpPtr[sd + 1];
    std::vector<int> types((size_t)(g1 - g0)), own;
    for (int64_t g = g0; g < g1; ++g) {
      types[g - g0] = part.grpType[g];
      if (uniqOwner[grpUniq[g]] == sd) {
        // a group is "local" to the FIRST subdomain that lists it; a later listing by the same subdomain
        // cannot occur (groups of one subdomain have distinct first nodes)
        own.push_back((int)(g - g0));
      }
    }
    for (const auto& linked : linkByType(types, own)) {
      std::vector<int> rows;
      for (int gi : linked) {
        const int u = grpUniq[g0 + gi];
        for (int64_t p = uniqPtr[u] + 1; p < uniqPtr[u + 1]; ++p) rows.push_back((int)p);
      }
      const int b = (int)blocks.size();
      for (size_t k = 0; k < rows.size(); ++k) { sepBlk[rows[k]] = b; sepBlkIdx[rows[k]] = (int)k; }
      blocks.push_back(rows);
    }
  }
  vsumGids.resize(nuniq);
  for (int u = 0; u < nuniq; ++u) vsumGids[u] = gids[sepRow[uniqPtr[u]]];
  // A12, A21, A22 in ordering positions (MatrixBlock::Compute :74-134)
  std::vector<int> posInt(n, -1);
  for (int64_t p = 0; p < nI; ++p) posInt[intRow[p]] = (int)p;
  auto extract = [&](const std::vector<int>& rows, const std::vector<int>& colPos, int64_t ncols, Csr& out) {
    out.n = (int64_t)rows.size();
    out.m = ncols;
    out.ptr.assign(rows.size() + 1, 0);
    out.col.clear();
    out.val.clear();
    for (size_t i = 0; i < rows.size(); ++i) {
      const int r = rows[i];
      std::vector<std::pair<int, double>> ent;
      for (int64_t e = A.ptr[r]; e < A.ptr[r + 1]; ++e)
        if (colPos[A.col[e]] >= 0) ent.emplace_back(colPos[A.col[e]], A.val[e]);
      std::sort(ent.begin(), ent.end());
      for (auto& pr : ent) { out.col.push_back(pr.first); out.val.push_back(pr.second); }
      out.ptr[i + 1] = (int64_t)out.col.size();
    }
  };
  extract(intRow, posSep, nS, A12);
  extract(sepRow, posInt, nI, A21);
  extract(sepRow, posSep, nS, A22);
}

void Level::a11Solve(const double* b, double* x) const {  // MatrixBlock::ApplyInverse :311-385
  const int nsd = part.nsd;
#pragma omp parallel
  {
    std::vector<double> work, tmp;
#pragma omp for schedule(dynamic, 4)
    for (int sd = 0; sd < nsd; ++sd) {
      const int64_t p0 = part.intPtr[sd];
      const int k = (int)(part.intPtr[sd + 1] - p0);
      if (k == 0) continue;
      work.resize(k);
      std::memcpy(x + p0, b + p0, k * sizeof(double));
      if (refineSteps > 0) refinedSolve(sdLU[sd], sdA[sd].ptr, sdA[sd].col, sdA[sd].val, refineSteps, x + p0, work.data(), tmp);
      else sdLU[sd].solve(x + p0, work.data());
    }
  }
}

void Level::applyOT(const double* v, double* out) const {  // ApplyOT :1236-1265: v <- 2 T'(T v) - v
#pragma omp parallel for schedule(static)
  for (int u = 0; u < nuniq; ++u) {
    const int64_t p0 = uniqPtr[u], len = uniqPtr[u + 1] - p0;
    const std::vector<double>& w = uniqW[u];
    double d = 0;
    if (!w.empty())
      for (int64_t q = 0; q < len; ++q) d += w[q] * v[p0 + q];
    for (int64_t q = 0; q < len; ++q) out[p0 + q] = (w.empty() ? 0.0 : 2.0 * w[q] * d) - v[p0 + q];
  }
}

// AssembleTransformAndDrop :698-875 / ConstructSCPart :877-986: per subdomain the dense Sk (A22 part, then
// -A21 A11^-1 A12), transformed group by group from the left and the right, of which the V-sum x V-sum entries
// and the non-V-sum entries inside each linked set are kept.  Pass 1 replaces (last writer wins), pass 2 sums.
void Level::assemble(std::vector<double>& redVal, const std::vector<int64_t>& redPtr, const std::vector<int>& redCol) {
  const int nsd = part.nsd;
  std::vector<std::vector<double>> blkVal(blocks.size());
  for (size_t b = 0; b < blocks.size(); ++b) blkVal[b].assign(blocks[b].size() * blocks[b].size(), 0.0);
  std::vector<std::vector<double>> blkVal2 = blkVal;
  std::vector<double> redVal2(redVal.size(), 0.0);
  struct Contribution {  // what one subdomain adds
    std::vector<double> vs[2];                 // G x G for the two passes
    std::vector<std::vector<double>> ll[2];    // per linked set: (non-V-sum)^2
    std::vector<std::vector<int>> llPos;       // separator positions of those nodes
  };
  const int batch = 4 * std::max(1, omp_get_max_threads());
  for (int sd0 = 0; sd0 < nsd; sd0 += batch) {
    const int sd1 = std::min(nsd, sd0 + batch);
    std::vector<Contribution> contrib(sd1 - sd0);
#pragma omp parallel
    {
      std::vector<double> Sk, col, work, vtmp, fac;
#pragma omp for schedule(dynamic, 1)
      for (int sd = sd0; sd < sd1; ++sd) {
        Contribution& C = contrib[sd - sd0];
        const int64_t g0 = part.sdGrpPtr[sd], g1 = part.sdGrpPtr[sd + 1];
        const int G = (int)(g1 - g0);
        const int* s = sdSep.data() + sdSepPtr[sd];
        const int m = (int)(sdSepPtr[sd + 1] - sdSepPtr[sd]);
        const int64_t i0 = part.intPtr[sd];
        const int nint = (int)(part.intPtr[sd + 1] - i0);
        std::vector<int> gstart(G + 1, 0), types(G), all(G);
        for (int g = 0; g < G; ++g) {
          gstart[g + 1] = gstart[g] + (int)(part.grpPtr[g0 + g + 1] - part.grpPtr[g0 + g]);
          types[g] = part.grpType[g0 + g];
          all[g] = g;
        }
        const std::vector<std::vector<int>> linked = linkByType(types, all);
        std::unordered_map<int, int> loc;  // separator position -> local index
        loc.reserve((size_t)m * 2);
        for (int i = 0; i < m; ++i) loc[s[i]] = i;
        C.llPos.resize(linked.size());
        for (size_t l = 0; l < linked.size(); ++l)
          for (int g : linked[l])
            for (int q = gstart[g] + 1; q < gstart[g + 1]; ++q) C.llPos[l].push_back(s[q]);
        for (int pass = 0; pass < 2; ++pass) {
          Sk.assign((size_t)m * m, 0.0);
          if (pass == 0) {  // Construct22 :258-306
            for (int i = 0; i < m; ++i)
              for (int64_t e = A22.ptr[s[i]]; e < A22.ptr[s[i] + 1]; ++e) {
                auto it = loc.find(A22.col[e]);
                if (it != loc.end()) Sk[(size_t)i * m + it->second] = A22.val[e];
              }
          } else if (nint > 0) {  // Construct11 :131-256: B = A11 \ A12(sd) column by column, Sk = -A21(sd) B
            // columns of A12(sd): gather the entries (interior row of sd, separator column in s)
            std::vector<std::vector<std::pair<int, double>>> cols(m);
            for (int r = 0; r < nint; ++r)
              for (int64_t e = A12.ptr[i0 + r]; e < A12.ptr[i0 + r + 1]; ++e) {
                auto it = loc.find(A12.col[e]);
                if (it != loc.end()) cols[it->second].emplace_back(r, A12.val[e]);
              }
            col.resize(nint);
            work.resize(nint);
            for (int j = 0; j < m; ++j) {
              if (cols[j].empty()) continue;
              std::fill(col.begin(), col.end(), 0.0);
              for (auto& pr : cols[j]) col[pr.first] = pr.second;
              if (refineSteps > 0)
                refinedSolve(sdLU[sd], sdA[sd].ptr, sdA[sd].col, sdA[sd].val, refineSteps, col.data(), work.data(), vtmp);
              else
                sdLU[sd].solve(col.data(), work.data());
              for (int i = 0; i < m; ++i) {
                double t = 0;
                for (int64_t e = A21.ptr[s[i]]; e < A21.ptr[s[i] + 1]; ++e) {
                  const int c = A21.col[e] - (int)i0;
                  if (c >= 0 && c < nint) t += A21.val[e] * col[c];
                }
                Sk[(size_t)i * m + j] = -t;
              }
            }
          }
          // RestrictedOT::Apply: rows, then columns, group by group (src/HYMLS_RestrictedOT.hpp:21-36)
          for (int g = 0; g < G; ++g) {
            const int a = gstart[g], len = gstart[g + 1] - a;
            const double* v = tvSep.data() + uniqPtr[grpUniq[g0 + g]];
            householderDense(Sk.data(), m, 1, a, len, m, v, vtmp, fac);
            householderDense(Sk.data(), 1, m, a, len, m, v, vtmp, fac);
          }
          C.vs[pass].resize((size_t)G * G);
          for (int g = 0; g < G; ++g)
            for (int h = 0; h < G; ++h) C.vs[pass][(size_t)g * G + h] = Sk[(size_t)gstart[g] * m + gstart[h]];
          C.ll[pass].resize(linked.size());
          for (size_t l = 0; l < linked.size(); ++l) {
            std::vector<int> li;
            for (int g : linked[l])
              for (int q = gstart[g] + 1; q < gstart[g + 1]; ++q) li.push_back(q);
            const size_t k = li.size();
            C.ll[pass][l].resize(k * k);
            for (size_t a = 0; a < k; ++a)
              for (size_t b = 0; b < k; ++b) C.ll[pass][l][a * k + b] = Sk[(size_t)li[a] * m + li[b]];
          }
        }
      }
    }
    // merge in subdomain order: deterministic, independent of the thread count
    for (int sd = sd0; sd < sd1; ++sd) {
      Contribution& C = contrib[sd - sd0];
      const int64_t g0 = part.sdGrpPtr[sd];
      const int G = (int)(part.sdGrpPtr[sd + 1] - g0);
      for (int g = 0; g < G; ++g) {
        const int ug = grpUniq[g0 + g];
        for (int h = 0; h < G; ++h) {
          const int uh = grpUniq[g0 + h];
          const int64_t pos = std::lower_bound(redCol.begin() + redPtr[ug], redCol.begin() + redPtr[ug + 1], uh) -
                              redCol.begin();
          redVal[pos] = C.vs[0][(size_t)g * G + h];
          redVal2[pos] += C.vs[1][(size_t)g * G + h];
        }
      }
      for (size_t l = 0; l < C.llPos.size(); ++l) {
        const std::vector<int>& pos = C.llPos[l];
        const size_t k = pos.size();
        for (size_t a = 0; a < k; ++a) {
          const int b = sepBlk[pos[a]];
          const size_t nb = blocks[b].size();
          for (size_t c = 0; c < k; ++c) {
            if (sepBlk[pos[c]] != b) continue;  // an entry no block solver ever reads
            const size_t idx = (size_t)sepBlkIdx[pos[a]] * nb + sepBlkIdx[pos[c]];
            blkVal[b][idx] = C.ll[0][l][a * k + c];
            blkVal2[b][idx] += C.ll[1][l][a * k + c];
          }
        }
      }
    }
  }
  for (size_t e = 0; e < redVal.size(); ++e) redVal[e] += redVal2[e];
  // separator-block LU (SchurPreconditioner::Compute :284-291)
  blockLU.assign(blocks.size(), DenseLU());
  bool ok = true;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t b = 0; b < (int64_t)blocks.size(); ++b) {
    if (blocks[b].empty()) continue;
    for (size_t e = 0; e < blkVal[b].size(); ++e) blkVal[b][e] += blkVal2[b][e];
    if (refineSteps > 0) blockLU[b].orig = blkVal[b];
    if (!blockLU[b].factor((int)blocks[b].size(), blkVal[b].data())) {
#pragma omp critical
      ok = false;
    }
  }
  if (!ok) throw std::runtime_error("singular separator block");
}

void Level::compute(std::vector<Partition>& parts, const std::vector<int64_t>& fix) {
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const int nsd = part.nsd;
  fixGids = fix;
  // subdomain solvers (ComputeSubdomainSolvers :210-292): sparse LU of every A11(sd)
  auto t0 = now();
  sdLU.assign(nsd, SparseLU());
  sdA.assign(refineSteps > 0 ? nsd : 0, Csr());
  std::vector<int> posInt(A.n, -1);
  for (int64_t p = 0; p < nI; ++p) posInt[intRow[p]] = (int)p;
  bool ok = true;
  int64_t nnzF = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : nnzF)
  for (int sd = 0; sd < nsd; ++sd) {
    const int64_t p0 = part.intPtr[sd];
    const int k = (int)(part.intPtr[sd + 1] - p0);
    if (k == 0) continue;
    Csr B;
    B.n = B.m = k;
    B.ptr.assign(k + 1, 0);
    for (int i = 0; i < k; ++i) {
      const int r = intRow[p0 + i];
      std::vector<std::pair<int, double>> ent;
      for (int64_t e = A.ptr[r]; e < A.ptr[r + 1]; ++e) {
        const int c = posInt[A.col[e]];
        if (c >= p0 && c < p0 + k) ent.emplace_back((int)(c - p0), A.val[e]);
      }
      std::sort(ent.begin(), ent.end());
      for (auto& pr : ent) { B.col.push_back(pr.first); B.val.push_back(pr.second); }
      B.ptr[i + 1] = (int64_t)B.col.size();
    }
    if (!factorCsr(B, sdLU[sd], fmatrixSolver)) {
#pragma omp critical
      ok = false;
    }
    nnzF += sdLU[sd].nnzFactors();
    if (refineSteps > 0) sdA[sd] = std::move(B);
  }
  if (!ok) throw std::runtime_error("singular subdomain matrix");
  factorNnz = nnzF;
  auto t1 = now();
  tm.factorA11 = secs(t0, t1);
  // reduced pattern: union of the V-sum cliques of the subdomains
  std::vector<std::vector<int>> rows(nuniq);
  for (int sd = 0; sd < nsd; ++sd) {
    const int64_t g0 = part.sdGrpPtr[sd], g1 = part.sdGrpPtr[sd + 1];
    for (int64_t g = g0; g < g1; ++g)
      for (int64_t h = g0; h < g1; ++h) rows[grpUniq[g]].push_back(grpUniq[h]);
  }
  std::vector<int64_t> redPtr(nuniq + 1, 0);
  std::vector<int> redCol;
  for (int u = 0; u < nuniq; ++u) {
    std::sort(rows[u].begin(), rows[u].end());
    rows[u].erase(std::unique(rows[u].begin(), rows[u].end()), rows[u].end());
    redCol.insert(redCol.end(), rows[u].begin(), rows[u].end());
    redPtr[u + 1] = (int64_t)redCol.size();
  }
  std::vector<double> redVal(redCol.size(), 0.0);
  assemble(redVal, redPtr, redCol);
  auto t2 = now();
  tm.schur = secs(t1, t2);
  // ComputeNextLevel :520-629
  reduced.n = reduced.m = nuniq;
  reduced.ptr = redPtr;
  reduced.col = redCol;
  reduced.val = redVal;
  dropByValue(reduced, SMALL, false);  // RelDropDiag
  if (level + 1 < maxLevel) {
    next.reset(new Level());
    next->level = level + 1;
    next->maxLevel = maxLevel;
    next->refineSteps = refineSteps;
    next->fmatrixSolver = fmatrixSolver;
    next->A = reduced;
    next->gids = vsumGids;
    std::vector<double> ttv(nS);
    applyOT(tvSep.data(), ttv.data());
    next->tv.resize(nuniq);
    for (int u = 0; u < nuniq; ++u) next->tv[u] = ttv[uniqPtr[u]];
    next->part = std::move(parts.at(level + 1));
    next->initialize();
    next->compute(parts, fix);
  } else {
    // CoarseSolver::Compute :131-248
    Csr S = reduced;
    dropByValue(S, SMALL, true);  // RelFullDiag
    coarseFixRows.clear();
    for (int64_t g : fix) {
      int row = -1;
      for (int u = 0; u < nuniq; ++u)
        if (vsumGids[u] == g) row = u;
      if (row < 0) throw std::runtime_error("fix GID not in matrix row map");
      coarseFixRows.push_back(row);
      putDirichlet(S, row);
    }
    if (nuniq > 0 && !factorCsr(S, coarseLU)) throw std::runtime_error("singular coarse matrix");
    if (refineSteps > 0) coarseA = S;
  }
  tm.coarse = secs(t2, now());
}

// SchurPreconditioner::ApplyInverse :1010-1093
void Level::schurApplyInverse(const double* rhs, double* sol) const {
  std::vector<double> B(nS), Y(nS, 0.0);
  applyOT(rhs, B.data());
  // ApplyBlockDiagonal :1311-1346
#pragma omp parallel
  {
    std::vector<double> tmp;
#pragma omp for schedule(dynamic, 16)
    for (int64_t b = 0; b < (int64_t)blocks.size(); ++b) {
      const std::vector<int>& rows = blocks[b];
      if (rows.empty()) continue;
      tmp.resize(rows.size());
      for (size_t k = 0; k < rows.size(); ++k) tmp[k] = B[rows[k]];
      blockLU[b].solveRefined(tmp.data(), refineSteps);
      for (size_t k = 0; k < rows.size(); ++k) Y[rows[k]] = tmp[k];
    }
  }
  // V-sums: next level or coarse solver (UpdateVsumRhs :1435-1459)
  std::vector<double> vr(nuniq), vs(nuniq);
  for (int u = 0; u < nuniq; ++u) vr[u] = B[uniqPtr[u]];
  if (next) {
    next->applyInverse(vr.data(), vs.data());
  } else if (nuniq > 0) {
    for (int row : coarseFixRows)
      if (row > 0) vr[row] = 0.0;  // sic: 'lid > 0', src/HYMLS_CoarseSolver.cpp:289
    std::vector<double> work(nuniq), tmp;
    vs = vr;
    if (refineSteps > 0) refinedSolve(coarseLU, coarseA.ptr, coarseA.col, coarseA.val, refineSteps, vs.data(), work.data(), tmp);
    else coarseLU.solve(vs.data(), work.data());
  }
  for (int u = 0; u < nuniq; ++u) Y[uniqPtr[u]] = vs[u];
  applyOT(Y.data(), sol);
}

// Preconditioner::ApplyInverse :930-1070 (b, x on this level's rows)
void Level::applyInverse(const double* b, double* x) const {
  std::vector<double> b1(nI), b2(nS), x1(nI), y(std::max(nI, nS)), x2(nS), z(nI);
  for (int64_t p = 0; p < nI; ++p) b1[p] = b[intRow[p]];
  for (int64_t p = 0; p < nS; ++p) b2[p] = b[sepRow[p]];
  a11Solve(b1.data(), x1.data());
  A21.matvec(x1.data(), y.data());
  for (int64_t p = 0; p < nS; ++p) b2[p] -= y[p];
  if (nS > 0) schurApplyInverse(b2.data(), x2.data());
  A12.matvec(x2.data(), y.data());
  a11Solve(y.data(), z.data());
  for (int64_t p = 0; p < nI; ++p) x[intRow[p]] = x1[p] - z[p];
  for (int64_t p = 0; p < nS; ++p) x[sepRow[p]] = x2[p];
}

// ---------------------------------------------------------------------------------------------
// Krylov (the same restatement of Belos BlockGmres / BlockCG, block size 1, as oracle/krylov.py)
// ---------------------------------------------------------------------------------------------
static double dotp(const double* a, const double* b, int64_t n) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

struct SolveResult {
  int iters = 0, converged = 0;
  double seconds = 0;
  std::vector<double> history;
};

static void gmres(const Csr& A, const Level& M, const double* b, double* x, double tol, int maxIters, int numBlocks,
                  int maxRestarts, SolveResult& res) {
  // right preconditioning, implicit residual scaled by the norm of the initial residual, zero or given x
  const int64_t n = A.n;
  const int m = std::max(1, std::min(numBlocks, maxIters));
  std::vector<double> r(n), w(n), z(n);
  std::vector<std::vector<double>> V;  // basis vectors, allocated as the iteration proceeds
  auto basis = [&](int i) -> std::vector<double>& {
    while ((int)V.size() <= i) V.emplace_back(n);
    return V[i];
  };
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), h(m + 1), h2(m + 1);
  auto t0 = std::chrono::steady_clock::now();
  A.matvec(x, r.data());
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
  const double r0 = std::sqrt(dotp(r.data(), r.data(), n));
  const double scale = r0 == 0.0 ? 1.0 : r0;
  bool first = true;
  for (int restart = 0; restart <= maxRestarts; ++restart) {
    const double beta = std::sqrt(dotp(r.data(), r.data(), n));
    if (first) res.history.push_back(beta / scale);
    first = false;
    if (beta / scale <= tol) { res.converged = 1; break; }
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    {
      std::vector<double>& v0 = basis(0);
      for (int64_t i = 0; i < n; ++i) v0[i] = r[i] / beta;
    }
    int kDone = 0;
    for (int k = 0; k < m && res.iters < maxIters; ++k) {
      double* vk = basis(k).data();
      basis(k + 1);
      M.applyInverse(vk, z.data());
      A.matvec(z.data(), w.data());
      for (int pass = 0; pass < 2; ++pass) {  // two passes of classical Gram-Schmidt (Belos ICGS)
        std::vector<double>& hh = pass ? h2 : h;
        // all k+1 dot products in one sweep over row chunks (w stays in cache); chunk partial sums are combined
        // in chunk order, so the result does not depend on the thread count
        const int64_t CH = 2048, nch = (n + CH - 1) / CH;
        std::vector<double> part((size_t)nch * (k + 1));
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nch; ++c) {
          const int64_t q0 = c * CH, q1 = std::min(n, q0 + CH);
          for (int i = 0; i <= k; ++i) {
            const double* vi = V[i].data();
            double t = 0;
            for (int64_t q = q0; q < q1; ++q) t += vi[q] * w[q];
            part[(size_t)c * (k + 1) + i] = t;
          }
        }
        for (int i = 0; i <= k; ++i) {
          double t = 0;
          for (int64_t c = 0; c < nch; ++c) t += part[(size_t)c * (k + 1) + i];
          hh[i] = t;
        }
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nch; ++c) {
          const int64_t q0 = c * CH, q1 = std::min(n, q0 + CH);
          for (int i = 0; i <= k; ++i) {
            const double* vi = V[i].data();
            const double hi = hh[i];
            for (int64_t q = q0; q < q1; ++q) w[q] -= hi * vi[q];
          }
        }
      }
      const double hn = std::sqrt(dotp(w.data(), w.data(), n));
      for (int i = 0; i <= k; ++i) H[(size_t)i * m + k] = h[i] + h2[i];
      H[(size_t)(k + 1) * m + k] = hn;
      if (hn > 0) {
        double* vn = V[k + 1].data();
        for (int64_t q = 0; q < n; ++q) vn[q] = w[q] / hn;
      }
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
      ++res.iters;
      kDone = k + 1;
      const double rel = std::fabs(g[k + 1]) / scale;
      res.history.push_back(rel);
      if (rel <= tol) break;
    }
    if (kDone > 0) {
      std::vector<double> y(kDone);
      for (int i = kDone - 1; i >= 0; --i) {
        double t = g[i];
        for (int j = i + 1; j < kDone; ++j) t -= H[(size_t)i * m + j] * y[j];
        y[i] = t / H[(size_t)i * m + i];
      }
      std::fill(w.begin(), w.end(), 0.0);
      for (int i = 0; i < kDone; ++i)
        for (int64_t q = 0; q < n; ++q) w[q] += y[i] * V[i][q];
      M.applyInverse(w.data(), z.data());
      for (int64_t q = 0; q < n; ++q) x[q] += z[q];
    }
    A.matvec(x, r.data());
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
    if (res.history.back() <= tol) { res.converged = 1; break; }
    if (res.iters >= maxIters) break;
  }
  res.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static void cg(const Csr& A, const Level& M, const double* b, double* x, double tol, int maxIters, SolveResult& res) {
  const int64_t n = A.n;
  std::vector<double> r(n), z(n), p(n), Ap(n);
  auto t0 = std::chrono::steady_clock::now();
  A.matvec(x, r.data());
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
  const double r0 = std::sqrt(dotp(r.data(), r.data(), n));
  res.history.push_back(r0 == 0 ? 0.0 : 1.0);
  if (r0 == 0) { res.converged = 1; return; }
  M.applyInverse(r.data(), z.data());
  p = z;
  double rz = dotp(r.data(), z.data(), n);
  while (res.iters < maxIters) {
    A.matvec(p.data(), Ap.data());
    const double alpha = rz / dotp(p.data(), Ap.data(), n);
    for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    ++res.iters;
    const double rel = std::sqrt(dotp(r.data(), r.data(), n)) / r0;
    res.history.push_back(rel);
    if (rel <= tol) { res.converged = 1; break; }
    M.applyInverse(r.data(), z.data());
    const double rzNew = dotp(r.data(), z.data(), n);
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + (rzNew / rz) * p[i];
    rz = rzNew;
  }
  res.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

struct Handle {
  Level L0;
  std::vector<Partition> parts;
  std::vector<int64_t> fix;
  std::string error;
  bool computed = false;
  double computeSeconds = 0;
  SolveResult last;
};

}  // namespace ho

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes: oracle/cpp_oracle.py)
// ---------------------------------------------------------------------------------------------
using namespace ho;
#define HO_TRY(h, body)                       \
  try {                                       \
    body;                                     \
    return 0;                                 \
  } catch (const std::exception& e) {         \
    (h)->error = e.what();                    \
    return -1;                                \
  }

extern "C" {

void* ho_create(int numLevels) {
  Handle* h = new Handle();
  h->L0.maxLevel = numLevels;
  h->parts.resize(std::max(numLevels, 1));
  return h;
}
void ho_destroy(void* hv) { delete (Handle*)hv; }
const char* ho_last_error(void* hv) { return ((Handle*)hv)->error.c_str(); }

int ho_set_matrix(void* hv, int64_t n, const int64_t* ptr, const int32_t* col, const double* val, const double* tv) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, {
    Csr& A = h->L0.A;
    A.n = A.m = n;
    A.ptr.assign(ptr, ptr + n + 1);
    A.col.assign(col, col + ptr[n]);
    A.val.assign(val, val + ptr[n]);
    h->L0.gids.resize(n);
    std::iota(h->L0.gids.begin(), h->L0.gids.end(), (int64_t)0);
    if (tv) h->L0.tv.assign(tv, tv + n); else h->L0.tv.clear();
    h->computed = false;
  })
}

// index maps of one level: interiors (intPtr[nsd+1], intGid) and groups (sdGrpPtr[nsd+1] -> group index,
// grpPtr[ngrp+1] -> grpGid, grpType[ngrp])
int ho_set_partition(void* hv, int level, int nsd, const int64_t* intPtr, const int64_t* intGid, const int64_t* sdGrpPtr,
                     const int64_t* grpPtr, const int64_t* grpGid, const int32_t* grpType) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, {
    Partition& P = h->parts.at(level);
    P.nsd = nsd;
    P.intPtr.assign(intPtr, intPtr + nsd + 1);
    P.intGid.assign(intGid, intGid + intPtr[nsd]);
    P.sdGrpPtr.assign(sdGrpPtr, sdGrpPtr + nsd + 1);
    const int64_t ngrp = sdGrpPtr[nsd];
    P.grpPtr.assign(grpPtr, grpPtr + ngrp + 1);
    P.grpGid.assign(grpGid, grpGid + grpPtr[ngrp]);
    P.grpType.assign(grpType, grpType + ngrp);
    h->computed = false;
  })
}

int ho_set_fix_gids(void* hv, int count, const int64_t* gids) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, h->fix.assign(gids, gids + count))
}

// steps > 0: every direct solve gets that many refinement steps with extended-precision residuals (checker use)
int ho_set_refinement(void* hv, int steps) {
  ((Handle*)hv)->L0.refineSteps = steps;
  ((Handle*)hv)->computed = false;
  return 0;
}

// 1 (default): the reference's F-matrix ordering + scaling + static pivots in the subdomain solvers; 0: general path
int ho_set_subdomain_ordering(void* hv, int fmatrix) {
  ((Handle*)hv)->L0.fmatrixSolver = fmatrix != 0;
  return 0;
}

int ho_compute(void* hv, int threads) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, {
    if (threads > 0) omp_set_num_threads(threads);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<Partition> parts = h->parts;  // levels >= 1 are moved into the level objects
    h->L0.level = 0;
    h->L0.part = parts[0];
    h->L0.initialize();
    h->L0.compute(parts, h->fix);
    h->computeSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    h->computed = true;
  })
}

int ho_apply_inverse(void* hv, const double* b, double* x) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, {
    if (!h->computed) throw std::runtime_error("The preconditioner has not yet been computed.");
    h->L0.applyInverse(b, x);
  })
}

// method 0 = GMRES (right preconditioned), 1 = CG; x holds the initial vector on entry
int ho_solve(void* hv, int method, const double* b, double* x, double tol, int maxIters, int numBlocks,
             int maxRestarts, int* iters, int* converged, double* seconds, double* hist, int histCap, int* histLen) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, {
    if (!h->computed) throw std::runtime_error("The preconditioner has not yet been computed.");
    SolveResult r;
    if (method == 0) gmres(h->L0.A, h->L0, b, x, tol, maxIters, numBlocks, maxRestarts, r);
    else cg(h->L0.A, h->L0, b, x, tol, maxIters, r);
    *iters = r.iters;
    *converged = r.converged;
    *seconds = r.seconds;
    *histLen = (int)r.history.size();
    for (int i = 0; i < (int)r.history.size() && i < histCap; ++i) hist[i] = r.history[i];
  })
}

// statistics: [0] compute seconds, [1] A11 factor s, [2] Schur assembly s, [3] coarse/next level s,
// [4] nnz(L+U) of all subdomain factors (level 0), [5] V-sums of level 0, [6] threads
int ho_stats(void* hv, double* out) {
  Handle* h = (Handle*)hv;
  out[0] = h->computeSeconds;
  out[1] = h->L0.tm.factorA11;
  out[2] = h->L0.tm.schur;
  out[3] = h->L0.tm.coarse;
  out[4] = (double)h->L0.factorNnz;
  out[5] = (double)h->L0.nuniq;
  out[6] = (double)omp_get_max_threads();
  return 0;
}

// reduced Schur complement of `level` after dropping (stage checks): sizes first, then the arrays
int ho_get_reduced(void* hv, int level, int64_t* n, int64_t* nnz, int64_t* ptr, int32_t* col, double* val) {
  Handle* h = (Handle*)hv;
  HO_TRY(h, {
    const Level* L = &h->L0;
    for (int l = 0; l < level; ++l) {
      if (!L->next) throw std::runtime_error("no such level");
      L = L->next.get();
    }
    *n = L->reduced.n;
    *nnz = (int64_t)L->reduced.col.size();
    if (ptr) std::copy(L->reduced.ptr.begin(), L->reduced.ptr.end(), ptr);
    if (col) std::copy(L->reduced.col.begin(), L->reduced.col.end(), col);
    if (val) std::copy(L->reduced.val.begin(), L->reduced.val.end(), val);
  })
}

// Fill of the sparse LU on its own (CSC input): nnz(L) incl. the unit diagonal and nnz(U) incl. the diagonal, and the
// solution of A x = b as a check.  Lets the tests compare the fill of the CPU baseline's factorization with the
// reference's known-answer counts for KLU (testSuite/unit_tests/HYMLS_SparseDirectSolver.cpp:62-152).
int ho_lu_fill(int n, const int32_t* Ap, const int32_t* Ai, const double* Ax, int fmatrix, const double* b, double* x,
               int64_t* nnzL, int64_t* nnzU) {
  try {
    std::vector<int> ap(Ap, Ap + n + 1), ai(Ai, Ai + Ap[n]);
    std::vector<double> ax(Ax, Ax + Ap[n]);
    SparseLU lu;
    if (!lu.factor(n, ap, ai, ax, fmatrix != 0)) return -2;
    if (fmatrix && !lu.fOrdered) return -3;  // not an F-matrix
    *nnzL = (int64_t)lu.Li.size();
    *nnzU = (int64_t)lu.Ui.size();
    if (b && x) {
      std::copy(b, b + n, x);
      std::vector<double> w(n);
      lu.solve(x, w.data());
    }
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

}  // extern "C"
