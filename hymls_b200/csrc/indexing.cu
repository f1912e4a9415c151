// Index arrays of a level that are built ON the device from the sparsity pattern (count -> exclusive scan -> fill)
// instead of on the host + upload.  The dense-fill scatter list of the subdomain matrices A11 (one (source entry,
// dense destination) pair per nonzero: 42 M pairs = 670 MB at 128^3) took 0.9 s of host loops and pageable copies in
// Initialize; from the pattern already resident in HBM it is three small kernels.
// Order of the list: by interior position, the entries of a row in CSR order (every pair is written exactly once, so
// the order only matters for reproducibility).
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

// pass 0: cnt[p] = entries of row intRow[p] inside the matrix of p;  pass 1: write the pairs at ptr[p]
template <int PASS>
__global__ void k_a11_list(const int64_t* __restrict__ rowptr, const int* __restrict__ colidx,
                           const int* __restrict__ intRow, const int* __restrict__ rowPos,
                           const int* __restrict__ posMat, int64_t nI, const int* __restrict__ np,
                           const int64_t* __restrict__ matOff, const int64_t* __restrict__ vecOff,
                           int64_t* __restrict__ ptr, int64_t* __restrict__ src, int64_t* __restrict__ dst) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nI) return;
  const int m = posMat[p];
  if (m < 0) {
    if (PASS == 0) ptr[p] = 0;
    return;
  }
  const int r = intRow[p];
  const int64_t v0 = vecOff[m];
  int64_t f = PASS == 0 ? 0 : ptr[p];
  const int64_t base = PASS == 0 ? 0 : matOff[m] + (p - v0) * np[m] - v0;
  for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
    const int cp = rowPos[colidx[e]];
    if (cp < 0 || posMat[cp] != m) continue;
    if (PASS == 1) {
      src[f] = e;
      dst[f] = base + cp;
    }
    ++f;
  }
  if (PASS == 0) ptr[p] = f;
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

void buildA11List(const int64_t* rowptr, const int* colidx, const int* intRow, const int* rowPos, int64_t nI,
                  const int* n, const int* np, const int64_t* matOff, const int64_t* vecOff, int count,
                  DevBuf<int64_t>& src, DevBuf<int64_t>& dst,
                  DevBuf<int64_t>& listPtrDev, std::vector<int64_t>& listPtr, cudaStream_t s, int64_t* launches) {
  listPtr.assign(count + 1, 0);
  listPtrDev.alloc(count + 1);
  if (count == 0 || nI == 0) {
    src.alloc(0);
    dst.alloc(0);
    HY_CUDA(cudaMemsetAsync(listPtrDev.p, 0, (count + 1) * sizeof(int64_t), s));
    return;
  }
  // scratch of this call only: plain allocations, outside the level's arena
  ArenaScope plain(nullptr);
  DevBuf<int> posMat;
  DevBuf<int64_t> ptr;
  DevBuf<char> tmp;
  posMat.alloc(nI);
  ptr.alloc(nI + 1);
  HY_CUDA(cudaMemsetAsync(posMat.p, 0xff, nI * sizeof(int), s));
  k_pos_mat<<<count, 128, 0, s>>>(n, vecOff, count, posMat.p);
  const unsigned grid = (unsigned)((nI + 255) / 256);
  k_a11_list<0><<<grid, 256, 0, s>>>(rowptr, colidx, intRow, rowPos, posMat.p, nI, np, matOff, vecOff, ptr.p, nullptr, nullptr);
  HY_CUDA(cudaMemsetAsync(ptr.p + nI, 0, sizeof(int64_t), s));
  size_t tmpBytes = 0;
  HY_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, ptr.p, ptr.p, nI + 1, s));
  tmp.alloc(tmpBytes);
  HY_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, ptr.p, ptr.p, nI + 1, s));
  k_list_ptr<<<(count + 1 + 255) / 256, 256, 0, s>>>(ptr.p, vecOff, count, nI, listPtrDev.p);
  HY_CUDA(cudaMemcpyAsync(listPtr.data(), listPtrDev.p, (count + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  HY_CUDA(cudaStreamSynchronize(s));
  const int64_t total = listPtr[count];
  {
    ArenaScope back(plain.prev);  // the lists themselves belong to the level
    src.alloc(total);
    dst.alloc(total);
  }
  if (total)
    k_a11_list<1><<<grid, 256, 0, s>>>(rowptr, colidx, intRow, rowPos, posMat.p, nI, np, matOff, vecOff, ptr.p, src.p, dst.p);
  *launches += 5;
  HY_CUDA(cudaStreamSynchronize(s));  // the scratch goes out of scope
}

}  // namespace hymls
