// Host-side symbolic phase of one level (Preconditioner::Initialize, src/HYMLS_Preconditioner.cpp:279-394;
// MatrixBlock ctor/Compute index work, src/HYMLS_MatrixBlock.cpp:30-134; SchurPreconditioner::Initialize /
// InitializeBlocks / CreateVSumMap, src/HYMLS_SchurPreconditioner.cpp:182-231,301-340,469-518).
// Everything here is integer work done once per sparsity pattern; the result is a set of flat index
// arrays that the CUDA kernels consume.
#pragma once
#include <cstdint>
#include <vector>

#include "partitioner.hpp"

namespace hymls {

struct CsrPattern {
  std::vector<int64_t> ptr;
  std::vector<int> col;
  std::vector<int64_t> src;  // index of the entry in the level matrix' value array
  int64_t nnz() const { return (int64_t)col.size(); }
};

struct LevelSym {
  int level = 0;
  int64_t n = 0;
  std::vector<gidx> rowGid;  // GID of every row of this level's matrix
  // pattern of this level's matrix: a VIEW of the caller's arrays (level 0: the engine's copy of the Jacobian pattern,
  // level l: redPtr / redCol of level l-1), valid as long as those live
  const int64_t* rowptr = nullptr;
  const int* colidx = nullptr;
  int64_t nnz = 0;

  HierarchicalMap H;
  int nsd = 0, nuniq = 0, nblk = 0;
  int64_t nI = 0, nS = 0;
  std::vector<int> intRow, sepRow;  // ordering -> row
  std::vector<int> rowPos;          // row -> interior position p (>=0) or -(separator position)-1

  // A11: one padded (np = roundup8(n)) row-major block per subdomain
  std::vector<int> sdN, sdNp;
  std::vector<int> sdNb;        // leading interior nodes of sd that separator rows couple to (columns of A21)
  double sumNNb = 0;            // sum n_sd * nb_sd
  std::vector<int64_t> a11Off;  // nsd+1, in doubles
  int64_t ignoredInteriorCouplings = 0;  // entries between interiors of different subdomains


  // ---- Schur-complement assembly (SchurComplement::Construct11/22, ConstructSCPart) ----
  std::vector<int> sdM;              // separator nodes around sd
  std::vector<int64_t> sdRowPtr;     // nsd+1: offset of sd's first local separator row in the s* arrays
  std::vector<int> sdSep;            // [sdRowPtr[sd] + i] -> separator position of local node i
  // group instances (a separator group as seen from one subdomain)
  std::vector<int64_t> sdInstPtr;    // nsd+1
  std::vector<int> instLoc, instLen, instUniq, instLink;  // local offset, length, unique id, linked-set id (per sd)
  std::vector<int> sdNumLink;        // linked sets per sd
  // unique groups (separator ordering): start/len via H.uniqPtr; block assignment
  std::vector<int> uniqBlk, uniqBlkOff;
  // blocks (one per owner subdomain and local linked set)
  std::vector<int> blkN, blkNp, blkOwnerSd;
  std::vector<int64_t> blkOff;       // nblk+1, in doubles
  std::vector<int64_t> blkRowPtr;    // nblk+1
  std::vector<int> blkRows;          // separator positions of the block rows
  std::vector<int> sepBlk, sepBlkIdx;  // per separator position: block and row in block (-1 for V-sums)

  // next level (reduced Schur complement on the V-sum nodes): pattern = union of per-sd cliques
  std::vector<int64_t> redPtr;
  std::vector<int> redCol;

  // test vector chain (numeric but matrix independent): Householder data
  std::vector<double> testVector;    // on this level's rows
  std::vector<double> what;          // per separator position: normalised reflector entry (0 if degenerate)
  std::vector<double> usign;         // per unique group: +1, or -1 when the dense transform is the identity
  std::vector<double> nextTestVector;

  double sumNsq = 0;  // sum n_sd^2
};

// Schedule that overlaps the host <-> device copies of a host-buffer ApplyInverse (one GPU) with the two passes over
// the level-0 subdomain inverses.  The subdomains (= matrices of the batched GEMV, in storage order) are cut into K
// consecutive chunks of about equal bytes.  Chunk c of the first pass needs b only on the rows its interiors occupy, so
// it may start once rows [0, inRows[c+1]) have arrived; after chunk c of the last pass every row below the smallest
// interior row of the later chunks is final, so rows [outRows[c], outRows[c+1]) can leave while chunk c+1 runs.
// (Subdomains are numbered along z, matrix rows likewise: a subdomain spans ~10 % of the row range at 128^3.)
struct HostPipePlan {
  int K = 0;                             // 0: no plan
  std::vector<int> matStart;             // K+1: chunk c = matrices [matStart[c], matStart[c+1])
  std::vector<int> leadItem, fullItem;   // K+1: the chunks' ranges in the leading-rows / full GEMV work lists
  std::vector<int64_t> inRows, outRows;  // K+1, nondecreasing, first 0, last nRows
};
// n / nb / vecOff: per matrix its order, its leading rows and the offset of its segment in the interior ordering;
// intRow: interior position -> matrix row; rowsPerItem: rows of one GEMV work item
HostPipePlan planHostPipe(const std::vector<int>& n, const std::vector<int>& nb, const std::vector<int64_t>& vecOff,
                          const std::vector<int>& intRow, int64_t nRows, int rowsPerItem, int K, bool taper = false);
// row-by-row verification of a plan against its definition (independent formulation; cheap: one pass over intRow)
bool checkHostPipe(const HostPipePlan& P, const std::vector<int>& n, const std::vector<int>& nb,
                   const std::vector<int64_t>& vecOff, const std::vector<int>& intRow, int64_t nRows, int rowsPerItem);

// Builds everything above from the matrix pattern and the partitioner of this level.
// `gid2row`: dense map GID -> row (or -1) over the fine-grid GID space.
int64_t countInteriorCouplings(const LevelSym& L);
void buildLevelSym(LevelSym& L, const CartesianPartitioner& part, const std::vector<int>& gid2row);

}  // namespace hymls
