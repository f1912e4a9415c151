// Index arrays of a level that are built ON the device from the sparsity pattern (count -> exclusive scan -> fill)
// instead of on the host + upload.  The dense-fill scatter list of the subdomain matrices A11 (one (source entry,
// dense destination) pair per nonzero: 42 M pairs = 670 MB at 128^3) took 0.9 s of host loops and pageable copies in
// Initialize; from the pattern already resident in HBM it is three small kernels.
// Order of the list: by interior position, the entries of a row in CSR order (every pair is written exactly once, so
// the order only matters for reproducibility).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "device.cuh"
#include "kernels.hpp"

namespace hymls {

// rowPos[row] = interior position p (>= 0) or -(separator position) - 1
__global__ void k_row_pos(const int* __restrict__ intRow, int64_t nI, const int* __restrict__ sepRow, int64_t nS,
                          int* __restrict__ rowPos) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nI) rowPos[intRow[i]] = (int)i;
  else if (i < nI + nS) rowPos[sepRow[i - nI]] = -(int)(i - nI) - 1;
}

// posMat[p] = index of the (owned) matrix whose rows contain interior position p; the caller presets -1
__global__ void k_pos_mat(const int* __restrict__ n, const int64_t* __restrict__ vecOff, int count, int* __restrict__ posMat) {
  const int m = blockIdx.x;
  if (m >= count) return;
  const int64_t v0 = vecOff[m];
  for (int i = threadIdx.x; i < n[m]; i += blockDim.x) posMat[v0 + i] = m;
}

// pass 0: cnt[p] = entries of row intRow[p] inside the matrix of p;  pass 1: write the pairs at ptr[p].
// posSd[p] = subdomain of interior position p (all subdomains): couplings between interiors of two different
// subdomains are counted in *ignored (the reference drops them too: A11 is block diagonal by construction)
template <int PASS>
__global__ void k_a11_list(const int64_t* __restrict__ rowptr, const int* __restrict__ colidx,
                           const int* __restrict__ intRow, const int* __restrict__ rowPos,
                           const int* __restrict__ posMat, const int* __restrict__ posSd, int64_t nI,
                           const int* __restrict__ np, const int64_t* __restrict__ matOff,
                           const int64_t* __restrict__ vecOff, int64_t* __restrict__ ptr, int64_t* __restrict__ src,
                           int64_t* __restrict__ dst, unsigned long long* __restrict__ ignored) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nI) return;
  const int m = posMat[p];
  const int r = intRow[p];
  if (PASS == 0) {
    const int sd = posSd[p];
    int64_t f = 0;
    unsigned long long ign = 0;
    for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const int cp = rowPos[colidx[e]];
      if (cp < 0) continue;
      if (posSd[cp] == sd) ++f; else ++ign;
    }
    ptr[p] = m >= 0 ? f : 0;
    if (ign) atomicAdd(ignored, ign);
    return;
  }
  if (m < 0) return;
  const int64_t v0 = vecOff[m];
  int64_t f = ptr[p];
  const int64_t base = matOff[m] + (p - v0) * np[m] - v0;
  for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
    const int cp = rowPos[colidx[e]];
    if (cp < 0 || posMat[cp] != m) continue;
    src[f] = e;
    dst[f] = base + cp;
    ++f;
  }
}

__global__ void k_list_ptr(const int64_t* __restrict__ ptr, const int64_t* __restrict__ vecOff, int count, int64_t nI,
                           int64_t* __restrict__ listPtr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) listPtr[k] = ptr[vecOff[k]];
  else if (k == count) listPtr[k] = ptr[nI];
}

void buildRowPos(const int* intRow, int64_t nI, const int* sepRow, int64_t nS, int* rowPos, cudaStream_t s,
                 int64_t* launches) {
  const int64_t tot = nI + nS;
  if (tot == 0) return;
  k_row_pos<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(intRow, nI, sepRow, nS, rowPos);
  ++*launches;
}

void buildPosMat(const int* n, const int64_t* vecOff, int count, int64_t nI, int* posMat, cudaStream_t s,
                 int64_t* launches) {
  if (nI == 0) return;
  HY_CUDA(cudaMemsetAsync(posMat, 0xff, nI * sizeof(int), s));
  if (count == 0) return;
  k_pos_mat<<<count, 128, 0, s>>>(n, vecOff, count, posMat);
  ++*launches;
}

// in-place exclusive sum of counts[0 .. n) (+ the total in counts[n]); returns the total
static int64_t scanCounts(int64_t* counts, int64_t n, DevBuf<char>& tmp, cudaStream_t s) {
  HY_CUDA(cudaMemsetAsync(counts + n, 0, sizeof(int64_t), s));
  size_t tmpBytes = 0;
  HY_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, counts, counts, n + 1, s));
  if (tmpBytes > tmp.cap) tmp.alloc(tmpBytes);
  HY_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, counts, counts, n + 1, s));
  int64_t total = 0;
  HY_CUDA(cudaMemcpyAsync(&total, counts + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  HY_CUDA(cudaStreamSynchronize(s));
  return total;
}

int64_t buildA11List(const int64_t* rowptr, const int* colidx, const int* intRow, const int* rowPos,
                     const int* posMat, const int* posSd, int64_t nI, const int* np, const int64_t* matOff,
                     const int64_t* vecOff, int count, DevBuf<int64_t>& src, DevBuf<int64_t>& dst,
                     DevBuf<int64_t>& listPtrDev, std::vector<int64_t>& listPtr, DeviceArena* scratch, cudaStream_t s,
                     int64_t* launches) {
  listPtr.assign(count + 1, 0);
  listPtrDev.alloc(count + 1);
  src.alloc(0);
  dst.alloc(0);
  HY_CUDA(cudaMemsetAsync(listPtrDev.p, 0, (count + 1) * sizeof(int64_t), s));
  if (nI == 0) return 0;
  DeviceArena* levelArena = g_arena;
  ArenaScope tmpScope(scratch);  // scratch of this call only, outside the level's arena
  DevBuf<int64_t> ptr;
  DevBuf<char> tmp;
  DevBuf<unsigned long long> ign;
  ptr.alloc(nI + 1);
  ign.alloc(1);
  HY_CUDA(cudaMemsetAsync(ign.p, 0, sizeof(unsigned long long), s));
  const unsigned grid = (unsigned)((nI + 255) / 256);
  k_a11_list<0><<<grid, 256, 0, s>>>(rowptr, colidx, intRow, rowPos, posMat, posSd, nI, np, matOff, vecOff, ptr.p,
                                     nullptr, nullptr, ign.p);
  const int64_t total = scanCounts(ptr.p, nI, tmp, s);
  unsigned long long ignored = 0;
  HY_CUDA(cudaMemcpyAsync(&ignored, ign.p, sizeof(ignored), cudaMemcpyDeviceToHost, s));
  if (count) {
    k_list_ptr<<<(count + 1 + 255) / 256, 256, 0, s>>>(ptr.p, vecOff, count, nI, listPtrDev.p);
    HY_CUDA(cudaMemcpyAsync(listPtr.data(), listPtrDev.p, (count + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  }
  HY_CUDA(cudaStreamSynchronize(s));
  {
    ArenaScope level(levelArena);  // the lists themselves belong to the level
    src.alloc(total);
    dst.alloc(total);
  }
  if (total)
    k_a11_list<1><<<grid, 256, 0, s>>>(rowptr, colidx, intRow, rowPos, posMat, posSd, nI, np, matOff, vecOff, ptr.p,
                                       src.p, dst.p, nullptr);
  *launches += 4;
  HY_CUDA(cudaStreamSynchronize(s));  // the scratch goes out of scope
  return (int64_t)ignored;
}

// ---------------------------------------------------------------------------------------------
// Split of the level matrix:  A12 (interior rows, separator columns), A21 / A22 (separator rows).  Columns are
// POSITIONS (interior / separator numbering), src = index of the entry in the level matrix (values are gathered
// with it at every Compute).  ownedOnly: A21 keeps the columns of the owned interiors only (sharded levels).
// ---------------------------------------------------------------------------------------------
template <int PASS>
__global__ void k_split_int(const int64_t* __restrict__ rowptr, const int* __restrict__ colidx,
                            const int* __restrict__ intRow, const int* __restrict__ rowPos, int64_t nI,
                            int64_t* __restrict__ p12, int* __restrict__ c12, int64_t* __restrict__ src12,
                            int* __restrict__ row12) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nI) return;
  const int r = intRow[p];
  int64_t f = PASS == 0 ? 0 : p12[p];
  for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
    const int cp = rowPos[colidx[e]];
    if (cp >= 0) continue;
    if (PASS == 1) {
      c12[f] = -cp - 1;
      src12[f] = e;
      row12[f] = (int)p;
    }
    ++f;
  }
  if (PASS == 0) p12[p] = f;
}

template <int PASS>
__global__ void k_split_sep(const int64_t* __restrict__ rowptr, const int* __restrict__ colidx,
                            const int* __restrict__ sepRow, const int* __restrict__ rowPos,
                            const int* __restrict__ posMat, int ownedOnly, int64_t nS, int64_t* __restrict__ p21,
                            int* __restrict__ c21, int64_t* __restrict__ src21, int64_t* __restrict__ p22,
                            int* __restrict__ c22, int64_t* __restrict__ src22) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nS) return;
  const int r = sepRow[p];
  int64_t f21 = PASS == 0 ? 0 : p21[p], f22 = PASS == 0 ? 0 : (p22 ? p22[p] : 0);
  for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
    const int cp = rowPos[colidx[e]];
    if (cp >= 0) {
      if (ownedOnly && posMat[cp] < 0) continue;
      if (PASS == 1) {
        c21[f21] = cp;
        src21[f21] = e;
      }
      ++f21;
    } else if (p22) {
      if (PASS == 1) {
        c22[f22] = -cp - 1;
        src22[f22] = e;
      }
      ++f22;
    }
  }
  if (PASS == 0) {
    p21[p] = f21;
    if (p22) p22[p] = f22;
  }
}

__global__ void k_lower_bounds(const int* __restrict__ sortedKeys, int64_t nKeys, int64_t nRows, int64_t* __restrict__ ptr) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c > nRows) return;
  int64_t lo = 0, hi = nKeys;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sortedKeys[mid] < (int)c) lo = mid + 1; else hi = mid;
  }
  ptr[c] = lo;
}

__global__ void k_gather_int(const int* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, int* __restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_iota64(int64_t* __restrict__ v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

void buildSplit(const int64_t* rowptr, const int* colidx, const int* intRow, const int* sepRow, const int* rowPos,
                const int* posMat, bool ownedOnly, bool want22, int64_t nI, int64_t nS, SplitOut& o,
                DeviceArena* scratch, cudaStream_t s, int64_t* launches) {
  DevBuf<char> tmp;
  DevBuf<int> row12;
  DevBuf<int> keys;
  DevBuf<int64_t> iota;
  o.p12->alloc(nI + 1);
  o.p21->alloc(nS + 1);
  if (want22) o.p22->alloc(nS + 1);
  o.t12Ptr->alloc(nS + 1);
  int64_t nnz12 = 0, nnz21 = 0, nnz22 = 0;
  const unsigned gI = (unsigned)((nI + 255) / 256), gS = (unsigned)((nS + 255) / 256);
  {
    ArenaScope tmpScope(scratch);
    if (nI) {
      k_split_int<0><<<gI, 256, 0, s>>>(rowptr, colidx, intRow, rowPos, nI, o.p12->p, nullptr, nullptr, nullptr);
      nnz12 = scanCounts(o.p12->p, nI, tmp, s);
    } else {
      HY_CUDA(cudaMemsetAsync(o.p12->p, 0, sizeof(int64_t), s));
    }
    if (nS) {
      k_split_sep<0><<<gS, 256, 0, s>>>(rowptr, colidx, sepRow, rowPos, posMat, ownedOnly ? 1 : 0, nS, o.p21->p, nullptr,
                                        nullptr, want22 ? o.p22->p : nullptr, nullptr, nullptr);
      nnz21 = scanCounts(o.p21->p, nS, tmp, s);
      if (want22) nnz22 = scanCounts(o.p22->p, nS, tmp, s);
    } else {
      HY_CUDA(cudaMemsetAsync(o.p21->p, 0, sizeof(int64_t), s));
      if (want22) HY_CUDA(cudaMemsetAsync(o.p22->p, 0, sizeof(int64_t), s));
    }
    row12.alloc(nnz12);
    keys.alloc(nnz12);
    iota.alloc(nnz12);
  }
  o.c12->alloc(nnz12);
  o.src12->alloc(nnz12);
  o.c21->alloc(nnz21);
  o.src21->alloc(nnz21);
  if (want22) {
    o.c22->alloc(nnz22);
    o.src22->alloc(nnz22);
  }
  o.t12Col->alloc(nnz12);
  o.t12Idx->alloc(nnz12);
  o.nnz12 = nnz12;
  o.nnz21 = nnz21;
  o.nnz22 = nnz22;
  if (nI && nnz12)
    k_split_int<1><<<gI, 256, 0, s>>>(rowptr, colidx, intRow, rowPos, nI, o.p12->p, o.c12->p, o.src12->p, row12.p);
  if (nS && (nnz21 || nnz22))
    k_split_sep<1><<<gS, 256, 0, s>>>(rowptr, colidx, sepRow, rowPos, posMat, ownedOnly ? 1 : 0, nS, o.p21->p, o.c21->p,
                                      o.src21->p, want22 ? o.p22->p : nullptr, want22 ? o.c22->p : nullptr,
                                      want22 ? o.src22->p : nullptr);
  // transposed index of A12: stable sort of the entries by column keeps (row, entry) order inside a column
  if (nnz12) {
    ArenaScope tmpScope(scratch);
    k_iota64<<<(unsigned)((nnz12 + 255) / 256), 256, 0, s>>>(iota.p, nnz12);
    int bits = 1;
    while (((int64_t)1 << bits) < nS + 1) ++bits;
    size_t tmpBytes = 0;
    HY_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, o.c12->p, keys.p, iota.p, o.t12Idx->p, nnz12, 0, bits, s));
    if (tmpBytes > tmp.cap) tmp.alloc(tmpBytes);
    HY_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, o.c12->p, keys.p, iota.p, o.t12Idx->p, nnz12, 0, bits, s));
    k_gather_int<<<(unsigned)((nnz12 + 255) / 256), 256, 0, s>>>(row12.p, o.t12Idx->p, nnz12, o.t12Col->p);
    k_lower_bounds<<<(unsigned)((nS + 1 + 255) / 256), 256, 0, s>>>(keys.p, nnz12, nS, o.t12Ptr->p);
  } else {
    HY_CUDA(cudaMemsetAsync(o.t12Ptr->p, 0, (nS + 1) * sizeof(int64_t), s));
  }
  *launches += 8;
  HY_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------
// Per-subdomain local pieces for the Schur assembly (SchurComplement::Construct / Construct11 / Construct22,
// src/HYMLS_SchurComplement.cpp:126-260: the reference extracts A12, A21, A22 of one subdomain with local indices).
// Local row R = (subdomain sd, i-th separator node of sd).
//   s21[R]: entries of the separator's matrix row in the interior of sd      -> (interior index in sd, entry)
//   s22[R]: entries in separator columns that also surround sd               -> (local separator index, entry)
//   s12[R]: entries of A12 in the COLUMN of this separator, rows in sd       -> (interior index in sd, entry)
// The local separator index of a separator position c in sd is found through its group u: occ lists all (sd, offset)
// pairs in which group u occurs (at most a handful).
// ---------------------------------------------------------------------------------------------
struct LocalArgs {
  const int64_t* rowptr;
  const int* colidx;
  const int* sepRow;
  const int* rowPos;
  const int* rowSd;        // per local row
  const int* sdSep;        // per local row: separator position
  const int64_t* intPtr;   // per subdomain: first interior position (nsd+1)
  const int* sepUniq;      // per separator position: unique group
  const int* uniqStart;    // per group: first separator position
  const int64_t* occPtr;
  const int* occSd;
  const int* occLoc;
  const int64_t* t12Ptr;
  const int* t12Col;
  const int64_t* t12Idx;
  const int64_t* src12;
  int64_t totalRows;
};

__device__ __forceinline__ int localSep(const LocalArgs& a, int sd, int c) {
  const int u = a.sepUniq[c];
  for (int64_t o = a.occPtr[u]; o < a.occPtr[u + 1]; ++o)
    if (a.occSd[o] == sd) return a.occLoc[o] + (c - a.uniqStart[u]);
  return -1;
}

template <int PASS>
__global__ void k_local_pieces(LocalArgs a, int64_t* __restrict__ s21Ptr, int* __restrict__ s21Col,
                               int64_t* __restrict__ s21Src, int64_t* __restrict__ s22Ptr, int* __restrict__ s22Col,
                               int64_t* __restrict__ s22Src, int64_t* __restrict__ s12Ptr, int* __restrict__ s12Row,
                               int64_t* __restrict__ s12Src) {
  const int64_t R = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (R >= a.totalRows) return;
  const int sd = a.rowSd[R], ps = a.sdSep[R];
  const int64_t i0 = a.intPtr[sd], i1 = a.intPtr[sd + 1];
  const int r = a.sepRow[ps];
  int64_t f21 = PASS == 0 ? 0 : s21Ptr[R], f22 = PASS == 0 ? 0 : s22Ptr[R], f12 = PASS == 0 ? 0 : s12Ptr[R];
  for (int64_t e = a.rowptr[r]; e < a.rowptr[r + 1]; ++e) {
    const int cp = a.rowPos[a.colidx[e]];
    if (cp >= 0) {
      if (cp >= i0 && cp < i1) {
        if (PASS == 1) {
          s21Col[f21] = (int)(cp - i0);
          s21Src[f21] = e;
        }
        ++f21;
      }
    } else {
      const int j = localSep(a, sd, -cp - 1);
      if (j >= 0) {
        if (PASS == 1) {
          s22Col[f22] = j;
          s22Src[f22] = e;
        }
        ++f22;
      }
    }
  }
  for (int64_t q = a.t12Ptr[ps]; q < a.t12Ptr[ps + 1]; ++q) {
    const int p = a.t12Col[q];
    if (p >= i0 && p < i1) {
      if (PASS == 1) {
        s12Row[f12] = (int)(p - i0);
        s12Src[f12] = a.src12[a.t12Idx[q]];
      }
      ++f12;
    }
  }
  if (PASS == 0) {
    s21Ptr[R] = f21;
    s22Ptr[R] = f22;
    s12Ptr[R] = f12;
  }
}

__global__ void k_sep_uniq(const int* __restrict__ uniqStart, int nuniq, int* __restrict__ sepUniq) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nuniq) return;
  for (int p = uniqStart[u]; p < uniqStart[u + 1]; ++p) sepUniq[p] = u;
}

void buildLocalPieces(const int64_t* rowptr, const int* colidx, const int* sepRow, const int* rowPos, const int* rowSd,
                      const int* sdSep, const int64_t* intPtr, const int* uniqStart, int nuniq, int64_t nS,
                      const int64_t* occPtr, const int* occSd, const int* occLoc, const int64_t* t12Ptr,
                      const int* t12Col, const int64_t* t12Idx, const int64_t* src12, int64_t totalRows, LocalOut& o,
                      DeviceArena* scratch, cudaStream_t s, int64_t* launches) {
  o.s21Ptr->alloc(totalRows + 1);
  o.s22Ptr->alloc(totalRows + 1);
  o.s12Ptr->alloc(totalRows + 1);
  o.nnz21 = o.nnz22 = o.nnz12 = 0;
  if (totalRows == 0) {
    HY_CUDA(cudaMemsetAsync(o.s21Ptr->p, 0, sizeof(int64_t), s));
    HY_CUDA(cudaMemsetAsync(o.s22Ptr->p, 0, sizeof(int64_t), s));
    HY_CUDA(cudaMemsetAsync(o.s12Ptr->p, 0, sizeof(int64_t), s));
    o.s21Col->alloc(0); o.s21Src->alloc(0); o.s22Col->alloc(0); o.s22Src->alloc(0); o.s12Row->alloc(0); o.s12Src->alloc(0);
    return;
  }
  DevBuf<char> tmp;
  DevBuf<int> sepUniq;
  {
    ArenaScope tmpScope(scratch);
    sepUniq.alloc(nS);
  }
  k_sep_uniq<<<(nuniq + 127) / 128, 128, 0, s>>>(uniqStart, nuniq, sepUniq.p);
  LocalArgs a{rowptr, colidx, sepRow, rowPos, rowSd, sdSep, intPtr, sepUniq.p, uniqStart, occPtr, occSd, occLoc,
              t12Ptr, t12Col, t12Idx, src12, totalRows};
  const unsigned grid = (unsigned)((totalRows + 127) / 128);
  k_local_pieces<0><<<grid, 128, 0, s>>>(a, o.s21Ptr->p, nullptr, nullptr, o.s22Ptr->p, nullptr, nullptr, o.s12Ptr->p,
                                         nullptr, nullptr);
  {
    ArenaScope tmpScope(scratch);
    o.nnz21 = scanCounts(o.s21Ptr->p, totalRows, tmp, s);
    o.nnz22 = scanCounts(o.s22Ptr->p, totalRows, tmp, s);
    o.nnz12 = scanCounts(o.s12Ptr->p, totalRows, tmp, s);
  }
  o.s21Col->alloc(o.nnz21);
  o.s21Src->alloc(o.nnz21);
  o.s22Col->alloc(o.nnz22);
  o.s22Src->alloc(o.nnz22);
  o.s12Row->alloc(o.nnz12);
  o.s12Src->alloc(o.nnz12);
  k_local_pieces<1><<<grid, 128, 0, s>>>(a, o.s21Ptr->p, o.s21Col->p, o.s21Src->p, o.s22Ptr->p, o.s22Col->p, o.s22Src->p,
                                         o.s12Ptr->p, o.s12Row->p, o.s12Src->p);
  *launches += 3;
  HY_CUDA(cudaStreamSynchronize(s));
}

}  // namespace hymls
