// Batched FP64 dense inversion for the per-subdomain interior blocks A11(sd) and the separator blocks.
//
// Replaces Ifpack_DenseContainer / Ifpack_SparseContainer<KLU>::Compute of the reference
// (src/HYMLS_MatrixBlock.cpp:210-292, src/HYMLS_SchurPreconditioner.cpp:284-291): the factor that
// ApplyInverse streams is the explicit inverse (8 n^2 bytes, the same as L+U), obtained by a blocked
// Gauss-Jordan elimination with partial (row) pivoting:
//   for every panel K of NB columns
//     (P) LU with partial pivoting of the active panel rows, then the Gauss-Jordan transform block
//         G' (n x NB):  G'[K] = inv(M[K,K]),  G'[R] = -M[R,K] inv(M[K,K])           (k_gj_panel)
//     (U) row swaps + trailing update of ALL other columns  M[:,J] += (G' - E_K) M[K,J]
//         as an FP64 tensor-core GEMM (DMMA m8n8k4), one CTA per column strip          (k_gj_update)
//   finally the column permutation that undoes the row swaps, written to the final storage (k_gj_gather).
// 2 n^3 flops per matrix, all but O(n^2 NB) of them in the DMMA update.
#include <cuda_runtime.h>

#include <algorithm>

#include "device.cuh"
#include "kernels.hpp"

namespace hymls {

static constexpr int GJ_NB = 32;       // panel width
static constexpr int GJ_PANEL_T = 512; // threads of the panel kernel
static constexpr int GJ_TJ = 64;       // column strip of the update kernel
static constexpr int GJ_CS = GJ_TJ / 32; // 8x8 tiles per warp along the strip (4 warps across)
static constexpr int GJ_TM = 32;       // row tile of the update kernel
static constexpr int GJ_UPD_T = 128;   // 4 warps side by side, each 4 x 2 DMMA tiles of 8x8; 4 CTAs per SM

// ---------------------------------------------------------------------------------------------
// identity in the padding rows/cols n..np-1 so the padded matrix stays invertible
// ---------------------------------------------------------------------------------------------
__global__ void k_pad_identity(double* __restrict__ W, const int64_t* __restrict__ off,
                               const int* __restrict__ nArr, const int* __restrict__ npArr, int count) {
  int mat = blockIdx.x;
  if (mat >= count) return;
  int n = nArr[mat], np = npArr[mat];
  double* M = W + off[mat];
  for (int i = n + threadIdx.x; i < np; i += blockDim.x) M[(int64_t)i * np + i] = 1.0;
}

// ---------------------------------------------------------------------------------------------
// panel kernel: one CTA per matrix
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GJ_PANEL_T)
k_gj_panel(double* __restrict__ W, const int64_t* __restrict__ off, const int* __restrict__ npArr,
           int* __restrict__ pivAll, int npMax, int k0, int* __restrict__ info, int rowsCap) {
  const int mat = blockIdx.x;
  const int np = npArr[mat];
  if (k0 >= np) return;
  if (np - k0 <= rowsCap) return;  // handled by the shared-memory kernel
  const int nb = min(GJ_NB, np - k0);
  double* M = W + off[mat];
  int* piv = pivAll + (int64_t)mat * npMax;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = GJ_PANEL_T / 32;

  __shared__ double sU[GJ_NB];
  __shared__ double sKK[GJ_NB][GJ_NB + 1];
  __shared__ double sLinv[GJ_NB][GJ_NB + 1];
  __shared__ double sDinv[GJ_NB][GJ_NB + 1];
  __shared__ double sRedV[NW];
  __shared__ int sRedI[NW];
  __shared__ int sPiv;
  __shared__ double sPivVal;

  // ---- (A) LU with partial pivoting on rows [k0, np) x cols [k0, k0+nb) ----
  for (int j = 0; j < nb; ++j) {
    const int c = k0 + j;
    double bestV = -1.0;
    int bestR = 0x7fffffff;
    for (int r = c + tid; r < np; r += GJ_PANEL_T) {
      double v = fabs(M[(int64_t)r * np + c]);
      if (v > bestV) {  // rows visited in increasing order per thread -> keeps the smallest row on ties
        bestV = v;
        bestR = r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_down_sync(0xffffffffu, bestV, o);
      int orow = __shfl_down_sync(0xffffffffu, bestR, o);
      if (ov > bestV || (ov == bestV && orow < bestR)) {
        bestV = ov;
        bestR = orow;
      }
    }
    if (lane == 0) {
      sRedV[wid] = bestV;
      sRedI[wid] = bestR;
    }
    __syncthreads();
    if (wid == 0) {
      bestV = lane < NW ? sRedV[lane] : -1.0;
      bestR = lane < NW ? sRedI[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, bestV, o);
        int orow = __shfl_down_sync(0xffffffffu, bestR, o);
        if (ov > bestV || (ov == bestV && orow < bestR)) {
          bestV = ov;
          bestR = orow;
        }
      }
      if (lane == 0) {
        if (bestR == 0x7fffffff) bestR = c;  // all-NaN column
        sPiv = bestR;
        piv[c] = bestR;
        if (!(bestV > 0.0)) atomicExch(info, mat + 1);  // exactly singular (or NaN) pivot column
      }
    }
    __syncthreads();
    const int p = sPiv;
    // swap rows c and p inside the panel; publish the pivot row
    if (tid < nb) {
      double a = M[(int64_t)c * np + k0 + tid];
      double b = M[(int64_t)p * np + k0 + tid];
      M[(int64_t)c * np + k0 + tid] = b;
      M[(int64_t)p * np + k0 + tid] = a;
      sU[tid] = b;
      if (tid == j) sPivVal = b;
    }
    __syncthreads();
    const double rp = 1.0 / sPivVal;
    for (int r = c + 1 + tid; r < np; r += GJ_PANEL_T) {
      double* row = M + (int64_t)r * np + k0;
      double l = row[j] * rp;
      row[j] = l;
      for (int q = j + 1; q < nb; ++q) row[q] -= l * sU[q];
    }
    __syncthreads();
  }
  // ---- (B) Linv = inv(L_KK) (unit lower), Dinv = inv(U) * Linv ----
  for (int e = tid; e < nb * nb; e += GJ_PANEL_T) sKK[e / nb][e % nb] = M[(int64_t)(k0 + e / nb) * np + k0 + e % nb];
  __syncthreads();
  if (tid < nb) {
    const int t = tid;  // column t of the inverses
    for (int i = 0; i < nb; ++i) {
      double x = (i == t) ? 1.0 : 0.0;
      for (int k = t; k < i; ++k) x -= sKK[i][k] * sLinv[k][t];
      sLinv[i][t] = (i < t) ? 0.0 : x;
    }
    for (int i = nb - 1; i >= 0; --i) {
      double x = sLinv[i][t];
      for (int k = i + 1; k < nb; ++k) x -= sKK[i][k] * sDinv[k][t];
      sDinv[i][t] = x / sKK[i][i];
    }
  }
  __syncthreads();
  // ---- (C) the Gauss-Jordan transform block G' overwrites the panel columns ----
  for (int r = tid; r < np; r += GJ_PANEL_T) {
    double* row = M + (int64_t)r * np + k0;
    if (r >= k0 && r < k0 + nb) {
      for (int q = 0; q < nb; ++q) row[q] = sDinv[r - k0][q];
      continue;
    }
    const bool above = r < k0;
    double a[GJ_NB];
#pragma unroll
    for (int q = 0; q < GJ_NB; ++q) a[q] = q < nb ? row[q] : 0.0;
    for (int q = 0; q < nb; ++q) {
      double s = 0.0;
      if (above) {
#pragma unroll
        for (int k = 0; k < GJ_NB; ++k) s += (k < nb) ? a[k] * sDinv[k][q] : 0.0;
      } else {
#pragma unroll
        for (int k = 0; k < GJ_NB; ++k) s += (k < nb) ? a[k] * sLinv[k][q] : 0.0;
      }
      row[q] = -s;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// panel kernel, shared-memory version: the 32-column panel is factored as four 8-column sub-panels
// that live in shared memory (column-major, [8][rows]) while the column-by-column pivot search / swap /
// rank-1 update runs; each finished sub-panel updates the rest of the panel in global memory (L2) once.
// Falls back to k_gj_panel when the active part of a matrix does not fit (rows > GJ_SMEM_ROWS).
// ---------------------------------------------------------------------------------------------
static constexpr int GJ_SW = 8;             // sub-panel width
static constexpr int GJ_SMEM_ROWS = 3072;   // 8 * 3072 * 8 B = 192 KB
static constexpr int GJ_SMEM_ROWS4 = 6144;  // 4-column sub-panels for taller panels (single large matrices)

template <int SW>
__global__ void __launch_bounds__(GJ_PANEL_T)
k_gj_panel_smem(double* __restrict__ W, const int64_t* __restrict__ off, const int* __restrict__ npArr,
                int* __restrict__ pivAll, int npMax, int k0, int* __restrict__ info, int rowsCap, int rowsMin) {
  const int mat = blockIdx.x;
  const int np = npArr[mat];
  if (k0 >= np) return;
  if (np - k0 > rowsCap) return;   // handled by the generic kernel
  if (np - k0 <= rowsMin) return;  // handled by the whole-panel kernel
  const int nb = min(GJ_NB, np - k0);
  double* M = W + off[mat];
  int* piv = pivAll + (int64_t)mat * npMax;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = GJ_PANEL_T / 32;

  extern __shared__ double sS[];  // [SW][rs] sub-panel, column major
  __shared__ double sU[SW];
  __shared__ double sUn[SW][GJ_NB];  // freshly solved U rows of the columns right of the sub-panel
  __shared__ double sKK[GJ_NB][GJ_NB + 1];
  __shared__ double sLinv[GJ_NB][GJ_NB + 1];
  __shared__ double sDinv[GJ_NB][GJ_NB + 1];
  __shared__ double sRedV[NW];
  __shared__ int sRedI[NW];
  __shared__ int sPiv;
  __shared__ double sPivVal;
  const int rs = (np - k0) | 1;  // odd stride: column-wise and row-wise accesses both conflict-light

  for (int c0 = k0; c0 < k0 + nb; c0 += SW) {
    const int w = min(SW, k0 + nb - c0);
    const int R = np - c0;  // active rows of this sub-panel (local row r <-> global row c0 + r)
    // load
    for (int e = tid; e < R * w; e += GJ_PANEL_T) {
      const int r = e / w, q = e % w;
      sS[q * rs + r] = M[(int64_t)(c0 + r) * np + c0 + q];
    }
    __syncthreads();
    for (int j = 0; j < w; ++j) {
      // pivot search in column j, local rows j..R-1
      double bestV = -1.0;
      int bestR = 0x7fffffff;
      for (int r = j + tid; r < R; r += GJ_PANEL_T) {
        const double v = fabs(sS[j * rs + r]);
        if (v > bestV) {
          bestV = v;
          bestR = r;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bestV, o);
        const int orow = __shfl_down_sync(0xffffffffu, bestR, o);
        if (ov > bestV || (ov == bestV && orow < bestR)) {
          bestV = ov;
          bestR = orow;
        }
      }
      if (lane == 0) {
        sRedV[wid] = bestV;
        sRedI[wid] = bestR;
      }
      __syncthreads();
      if (wid == 0) {
        bestV = lane < NW ? sRedV[lane] : -1.0;
        bestR = lane < NW ? sRedI[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_down_sync(0xffffffffu, bestV, o);
          const int orow = __shfl_down_sync(0xffffffffu, bestR, o);
          if (ov > bestV || (ov == bestV && orow < bestR)) {
            bestV = ov;
            bestR = orow;
          }
        }
        if (lane == 0) {
          if (bestR == 0x7fffffff) bestR = j;
          sPiv = bestR;
          piv[c0 + j] = c0 + bestR;
          if (!(bestV > 0.0)) atomicExch(info, mat + 1);
        }
      }
      __syncthreads();
      const int p = sPiv;
      if (tid < w) {
        const double a = sS[tid * rs + j];
        const double b = sS[tid * rs + p];
        sS[tid * rs + j] = b;
        sS[tid * rs + p] = a;
        sU[tid] = b;
        if (tid == j) sPivVal = b;
      }
      __syncthreads();
      const double rp = 1.0 / sPivVal;
      for (int r = j + 1 + tid; r < R; r += GJ_PANEL_T) {
        const double l = sS[j * rs + r] * rp;
        sS[j * rs + r] = l;
        for (int q = j + 1; q < w; ++q) sS[q * rs + r] -= l * sU[q];
      }
      __syncthreads();
    }
    // store the factored sub-panel
    for (int e = tid; e < R * w; e += GJ_PANEL_T) {
      const int r = e / w, q = e % w;
      M[(int64_t)(c0 + r) * np + c0 + q] = sS[q * rs + r];
    }
    // row swaps of this sub-panel applied to the other panel columns (one thread per column)
    const int nOther = nb - w;
    if (tid < nOther) {
      const int col = (tid < c0 - k0) ? k0 + tid : c0 + w + (tid - (c0 - k0));
      for (int j = 0; j < w; ++j) {
        const int a = c0 + j, b = piv[c0 + j];
        if (a != b) {
          const double x = M[(int64_t)a * np + col];
          const double y = M[(int64_t)b * np + col];
          M[(int64_t)a * np + col] = y;
          M[(int64_t)b * np + col] = x;
        }
      }
    }
    __syncthreads();
    // columns to the right of the sub-panel: U rows by forward substitution with the unit-lower w x w
    // block, then the rank-w update of all rows below
    const int nRight = k0 + nb - (c0 + w);
    if (nRight > 0) {
      if (tid < nRight) {
        const int col = c0 + w + tid;
        double u[SW];
#pragma unroll
        for (int i = 0; i < SW; ++i) {
          if (i < w) {
            double x = M[(int64_t)(c0 + i) * np + col];
            for (int t = 0; t < i; ++t) x -= sS[t * rs + i] * u[t];
            u[i] = x;
            M[(int64_t)(c0 + i) * np + col] = x;
            sUn[i][tid] = x;
          }
        }
      }
      __syncthreads();
      for (int r = w + tid; r < R; r += GJ_PANEL_T) {
        double l[SW];
#pragma unroll
        for (int t = 0; t < SW; ++t) l[t] = t < w ? sS[t * rs + r] : 0.0;
        double* row = M + (int64_t)(c0 + r) * np + c0 + w;
        for (int q = 0; q < nRight; ++q) {
          double x = row[q];
#pragma unroll
          for (int t = 0; t < SW; ++t) x -= l[t] * sUn[t][q];
          row[q] = x;
        }
      }
    }
    __syncthreads();
  }
  // ---- (B) Linv = inv(L_KK) (unit lower), Dinv = inv(U) * Linv ----
  for (int e = tid; e < nb * nb; e += GJ_PANEL_T) sKK[e / nb][e % nb] = M[(int64_t)(k0 + e / nb) * np + k0 + e % nb];
  __syncthreads();
  if (tid < nb) {
    const int t = tid;
    for (int i = 0; i < nb; ++i) {
      double x = (i == t) ? 1.0 : 0.0;
      for (int k = t; k < i; ++k) x -= sKK[i][k] * sLinv[k][t];
      sLinv[i][t] = (i < t) ? 0.0 : x;
    }
    for (int i = nb - 1; i >= 0; --i) {
      double x = sLinv[i][t];
      for (int k = i + 1; k < nb; ++k) x -= sKK[i][k] * sDinv[k][t];
      sDinv[i][t] = x / sKK[i][i];
    }
  }
  __syncthreads();
  // ---- (C) G' overwrites the panel columns ----
  for (int r = tid; r < np; r += GJ_PANEL_T) {
    double* row = M + (int64_t)r * np + k0;
    if (r >= k0 && r < k0 + nb) {
      for (int q = 0; q < nb; ++q) row[q] = sDinv[r - k0][q];
      continue;
    }
    const bool above = r < k0;
    double a[GJ_NB];
#pragma unroll
    for (int q = 0; q < GJ_NB; ++q) a[q] = q < nb ? row[q] : 0.0;
    for (int q = 0; q < nb; ++q) {
      double s = 0.0;
      if (above) {
#pragma unroll
        for (int k = 0; k < GJ_NB; ++k) s += (k < nb) ? a[k] * sDinv[k][q] : 0.0;
      } else {
#pragma unroll
        for (int k = 0; k < GJ_NB; ++k) s += (k < nb) ? a[k] * sLinv[k][q] : 0.0;
      }
      row[q] = -s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// panel kernel, whole-panel version: when the active part of the 32-column panel (R = np - k0 rows) fits in
// shared memory ([32][R] column major, R <= GJ_FULL_ROWS) the complete step runs out of it: pivot search,
// row swaps, the 8-column sub-panels with their rank-8 updates of the rest of the panel, the inverses of
// the triangular factors of the K x K block and the transform block G' of the active rows; global memory
// sees one coalesced read and one coalesced write of the panel.  The rows above the panel
// (G'[r] = -M[r,K] inv(M[K,K]), r < k0) stream through the same shared memory in tiles.
// ---------------------------------------------------------------------------------------------
static constexpr int GJ_FULL_ROWS = 824;   // 32 * 825 * 8 B = 206 KB dynamic + 20 KB static <= 227 KB
static constexpr int GJ_FULL_MINROWS = 256; // tile height of the rows above the panel is at least this
static constexpr int GJ_XS = GJ_NB + 4;    // row stride of the 32 x 32 factor inverses (16-byte aligned rows)

// rows [0, rows) of the tile sT ([nb][rsT] column major): row <- -(row * X), X = sX (nb x nb, stride GJ_XS)
__device__ __forceinline__ void gjTransformRows(double* __restrict__ sT, int rsT, int rFirst, int rows, int nb,
                                                const double* __restrict__ sX, int tid) {
  for (int r = rFirst + tid; r < rows; r += GJ_PANEL_T) {
    double a[GJ_NB];
#pragma unroll
    for (int k = 0; k < GJ_NB; ++k) a[k] = k < nb ? sT[k * rsT + r] : 0.0;
    for (int q = 0; q < nb; q += 4) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k < GJ_NB; ++k) {
        if (k < nb) {
          const double2 x01 = *reinterpret_cast<const double2*>(sX + k * GJ_XS + q);
          const double2 x23 = *reinterpret_cast<const double2*>(sX + k * GJ_XS + q + 2);
          s0 += a[k] * x01.x;
          s1 += a[k] * x01.y;
          s2 += a[k] * x23.x;
          s3 += a[k] * x23.y;
        }
      }
      sT[q * rsT + r] = -s0;
      sT[(q + 1) * rsT + r] = -s1;
      sT[(q + 2) * rsT + r] = -s2;
      sT[(q + 3) * rsT + r] = -s3;
    }
  }
}

__global__ void __launch_bounds__(GJ_PANEL_T)
k_gj_panel_full(double* __restrict__ W, const int64_t* __restrict__ off, const int* __restrict__ npArr,
                int* __restrict__ pivAll, int npMax, int k0, int* __restrict__ info, int rowsCap, int tileRows) {
  const int mat = blockIdx.x;
  const int np = npArr[mat];
  if (k0 >= np) return;
  const int R = np - k0;
  if (R > rowsCap) return;  // handled by the sub-panel kernels
  const int nb = min(GJ_NB, R);  // np is a multiple of 8, so nb is a multiple of 4
  double* M = W + off[mat];
  int* piv = pivAll + (int64_t)mat * npMax;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = GJ_PANEL_T / 32;

  extern __shared__ double sP[];  // [nb][rs] active panel / [nb][rsT] tiles of the rows above
  __shared__ __align__(16) double sLinv[GJ_NB * GJ_XS];
  __shared__ __align__(16) double sDinv[GJ_NB * GJ_XS];
  __shared__ double sU[GJ_SW];
  __shared__ double sUn[GJ_SW][GJ_NB];
  __shared__ double sRedV[NW];
  __shared__ int sRedI[NW];
  __shared__ int sPiv;
  __shared__ double sPivVal;
  const int rs = R | 1;

  for (int e = tid; e < R * nb; e += GJ_PANEL_T) {
    const int r = e / nb, q = e - r * nb;
    sP[q * rs + r] = M[(int64_t)(k0 + r) * np + k0 + q];
  }
  __syncthreads();
  for (int c0 = 0; c0 < nb; c0 += GJ_SW) {
    const int w = min(GJ_SW, nb - c0);
    for (int jj = 0; jj < w; ++jj) {
      const int j = c0 + jj;  // local column = local row of the pivot position
      double bestV = -1.0;
      int bestR = 0x7fffffff;
      for (int r = j + tid; r < R; r += GJ_PANEL_T) {
        const double v = fabs(sP[j * rs + r]);
        if (v > bestV) {
          bestV = v;
          bestR = r;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bestV, o);
        const int orow = __shfl_down_sync(0xffffffffu, bestR, o);
        if (ov > bestV || (ov == bestV && orow < bestR)) {
          bestV = ov;
          bestR = orow;
        }
      }
      if (lane == 0) {
        sRedV[wid] = bestV;
        sRedI[wid] = bestR;
      }
      __syncthreads();
      if (wid == 0) {
        bestV = lane < NW ? sRedV[lane] : -1.0;
        bestR = lane < NW ? sRedI[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_down_sync(0xffffffffu, bestV, o);
          const int orow = __shfl_down_sync(0xffffffffu, bestR, o);
          if (ov > bestV || (ov == bestV && orow < bestR)) {
            bestV = ov;
            bestR = orow;
          }
        }
        if (lane == 0) {
          if (bestR == 0x7fffffff) bestR = j;
          sPiv = bestR;
          piv[k0 + j] = k0 + bestR;
          if (!(bestV > 0.0)) atomicExch(info, mat + 1);
        }
      }
      __syncthreads();
      const int p = sPiv;
      if (tid < nb) {  // swap the two rows over the whole panel
        const double a = sP[tid * rs + j];
        const double b = sP[tid * rs + p];
        sP[tid * rs + j] = b;
        sP[tid * rs + p] = a;
        if (tid >= c0 && tid < c0 + w) sU[tid - c0] = b;
        if (tid == j) sPivVal = b;
      }
      __syncthreads();
      const double rp = 1.0 / sPivVal;
      for (int r = j + 1 + tid; r < R; r += GJ_PANEL_T) {
        const double l = sP[j * rs + r] * rp;
        sP[j * rs + r] = l;
        for (int q = jj + 1; q < w; ++q) sP[(c0 + q) * rs + r] -= l * sU[q];
      }
      __syncthreads();
    }
    const int nRight = nb - (c0 + w);
    if (nRight > 0) {
      if (tid < nRight) {
        const int col = c0 + w + tid;
        double u[GJ_SW];
#pragma unroll
        for (int i = 0; i < GJ_SW; ++i) {
          if (i < w) {
            double x = sP[col * rs + c0 + i];
            for (int t = 0; t < i; ++t) x -= sP[(c0 + t) * rs + c0 + i] * u[t];
            u[i] = x;
            sP[col * rs + c0 + i] = x;
            sUn[i][tid] = x;
          }
        }
      }
      __syncthreads();
      for (int r = c0 + w + tid; r < R; r += GJ_PANEL_T) {
        double l[GJ_SW];
#pragma unroll
        for (int t = 0; t < GJ_SW; ++t) l[t] = t < w ? sP[(c0 + t) * rs + r] : 0.0;
        for (int q = 0; q < nRight; ++q) {
          double x = sP[(c0 + w + q) * rs + r];
#pragma unroll
          for (int t = 0; t < GJ_SW; ++t) x -= l[t] * sUn[t][q];
          sP[(c0 + w + q) * rs + r] = x;
        }
      }
      __syncthreads();
    }
  }
  // Linv = inv(L_KK) (unit lower), Dinv = inv(U_KK) * Linv; the K x K block is the first nb rows of the panel
  if (tid < nb) {
    const int t = tid;
    for (int i = 0; i < nb; ++i) {
      double x = (i == t) ? 1.0 : 0.0;
      for (int k = t; k < i; ++k) x -= sP[k * rs + i] * sLinv[k * GJ_XS + t];
      sLinv[i * GJ_XS + t] = (i < t) ? 0.0 : x;
    }
    for (int i = nb - 1; i >= 0; --i) {
      double x = sLinv[i * GJ_XS + t];
      for (int k = i + 1; k < nb; ++k) x -= sP[k * rs + i] * sDinv[k * GJ_XS + t];
      sDinv[i * GJ_XS + t] = x / sP[i * rs + i];
    }
  }
  __syncthreads();
  // G' of the active rows, in place: rows below the block -L Linv, the block itself Dinv
  gjTransformRows(sP, rs, nb, R, nb, sLinv, tid);
  for (int e = tid; e < nb * nb; e += GJ_PANEL_T) {
    const int r = e / nb, q = e - r * nb;
    sP[q * rs + r] = sDinv[r * GJ_XS + q];
  }
  __syncthreads();
  for (int e = tid; e < R * nb; e += GJ_PANEL_T) {
    const int r = e / nb, q = e - r * nb;
    M[(int64_t)(k0 + r) * np + k0 + q] = sP[q * rs + r];
  }
  // rows above the panel, tile by tile
  const int rsT = tileRows | 1;
  for (int t0 = 0; t0 < k0; t0 += tileRows) {
    const int rows = min(tileRows, k0 - t0);
    __syncthreads();
    for (int e = tid; e < rows * nb; e += GJ_PANEL_T) {
      const int r = e / nb, q = e - r * nb;
      sP[q * rsT + r] = M[(int64_t)(t0 + r) * np + k0 + q];
    }
    __syncthreads();
    gjTransformRows(sP, rsT, 0, rows, nb, sDinv, tid);
    __syncthreads();
    for (int e = tid; e < rows * nb; e += GJ_PANEL_T) {
      const int r = e / nb, q = e - r * nb;
      M[(int64_t)(t0 + r) * np + k0 + q] = sP[q * rsT + r];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The 32 row interchanges of a panel composed into one gather list per matrix (one warp each):
// rowsT[i] = a row whose content changes (-1: unused), origT[i] = the row its new content comes from.
// Entries 0..31 are the panel rows k0..k0+31 themselves, 32..63 the pivot rows outside the panel.
// ---------------------------------------------------------------------------------------------
__global__ void k_gj_swaplist(const int* __restrict__ npArr, const int* __restrict__ pivAll, int npMax, int k0,
                              int count, int* __restrict__ rowsT, int* __restrict__ origT) {
  const int mat = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (mat >= count) return;
  const int np = npArr[mat];
  if (k0 >= np) return;
  const int lane = threadIdx.x & 31;
  const int nb = min(GJ_NB, np - k0);
  const int* piv = pivAll + (int64_t)mat * npMax;
  int origK = k0 + lane, outRow = -1, origOut = -1, nOut = 0;
  for (int j = 0; j < nb; ++j) {
    const int p = piv[k0 + j];
    if (p == k0 + j) continue;
    const int a = __shfl_sync(0xffffffffu, origK, j);
    if (p < k0 + nb) {
      const int b = __shfl_sync(0xffffffffu, origK, p - k0);
      if (lane == j) origK = b;
      if (lane == p - k0) origK = a;
    } else {
      const unsigned m = __ballot_sync(0xffffffffu, outRow == p);
      int slot;
      if (m) {
        slot = __ffs(m) - 1;
      } else {
        slot = nOut++;
        if (lane == slot) {
          outRow = p;
          origOut = p;
        }
      }
      const int b = __shfl_sync(0xffffffffu, origOut, slot);
      if (lane == j) origK = b;
      if (lane == slot) origOut = a;
    }
  }
  int* rT = rowsT + (int64_t)mat * 64;
  int* oT = origT + (int64_t)mat * 64;
  rT[lane] = lane < nb ? k0 + lane : -1;
  oT[lane] = origK;
  rT[32 + lane] = lane < nOut ? outRow : -1;
  oT[32 + lane] = origOut;
}

// ---------------------------------------------------------------------------------------------
// DMMA m8n8k4 (FP64 tensor core): D(8x8) += A(8x4, row) * B(4x8, col)
//   a : A[lane/4][lane%4]        b : B[lane%4][lane/4]        c0,c1 : C[lane/4][2*(lane%4) + {0,1}]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// update kernel: grid (column strips, matrices)
__global__ void __launch_bounds__(GJ_UPD_T, 4)
k_gj_update(double* __restrict__ W, const int64_t* __restrict__ off, const int* __restrict__ npArr,
            const int* __restrict__ rowsT, const int* __restrict__ origT, int k0) {
  const int mat = blockIdx.y;
  const int np = npArr[mat];
  if (k0 >= np) return;
  const int j0 = blockIdx.x * GJ_TJ;
  if (j0 >= np) return;
  const int nb = min(GJ_NB, np - k0);
  double* M = W + off[mat];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int jw = min(GJ_TJ, np - j0);  // np is a multiple of 8, so jw is too

  constexpr int SB = GJ_TJ + 4;  // stride = 4 mod 16 doubles -> conflict-free fragment loads
  constexpr int SA = GJ_NB + 4;
  extern __shared__ double gjSmem[];
  double* sB = gjSmem;                // GJ_NB x SB
  double* sA = gjSmem + GJ_NB * SB;   // GJ_TM x SA

  // (1) row interchanges of this strip from the composed gather list: every touched row is fetched from the
  //     row its new content comes from (coalesced, independent loads), staged in shared memory and written
  //     back; the first nb staged rows are B = M[K, strip] after the interchanges.
  __shared__ int sRow[64], sOrig[64];
  if (tid < 64) {
    sRow[tid] = rowsT[(int64_t)mat * 64 + tid];
    sOrig[tid] = origT[(int64_t)mat * 64 + tid];
  }
  __syncthreads();
  for (int e = tid; e < 64 * GJ_TJ; e += GJ_UPD_T) {
    const int i = e / GJ_TJ, c = e % GJ_TJ;
    const int row = sRow[i];
    gjSmem[i * SB + c] = (row >= 0 && c < jw) ? M[(int64_t)sOrig[i] * np + j0 + c] : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < 64 * GJ_TJ; e += GJ_UPD_T) {
    const int i = e / GJ_TJ, c = e % GJ_TJ;
    const int row = sRow[i], col = j0 + c;
    // the panel's own columns were interchanged by the panel kernel
    if (row >= 0 && sOrig[i] != row && c < jw && (col < k0 || col >= k0 + nb))
      M[(int64_t)row * np + col] = gjSmem[i * SB + c];
  }
  __syncthreads();
  // warp layout: 1 x 4 warps, each 4 x 2 tiles of 8 x 8 -> CTA tile 32 x 64
  const int wr = wid >> 2, wc = wid & 3;
  const int fr = lane >> 2, fk = lane & 3;

  // software pipeline over the row tiles: while tile i is multiplied, the G' rows of tile i+1 arrive in the
  // other shared-memory buffer (cp.async) and its accumulator values in registers
  auto stageA = [&](int r0, int buf) {
    double* dst = sA + buf * (GJ_TM * SA);
    for (int e = tid; e < GJ_TM * (GJ_NB / 2); e += GJ_UPD_T) {
      const int r = e / (GJ_NB / 2), k = (e % (GJ_NB / 2)) * 2;
      const bool ok = (r0 + r < np) && (k < nb);
      const double* src = ok ? M + (int64_t)(r0 + r) * np + k0 + k : M;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + r * SA + k);
      const int bytes = ok ? 16 : 0;  // src-size 0: zero fill
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(src), "r"(bytes));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  auto loadAcc = [&](int r0, double (&acc)[4][GJ_CS][2]) {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti) {
      const int row = r0 + (wr * 4 + ti) * 8 + fr;
      const bool inK = (row >= k0 && row < k0 + nb);  // rows of the panel are replaced, not accumulated
#pragma unroll
      for (int tj = 0; tj < GJ_CS; ++tj) {
        const int col = j0 + (wc * GJ_CS + tj) * 8 + 2 * fk;
        const bool ok = row < np && col < j0 + jw && !inK;
        double2 v = make_double2(0.0, 0.0);
        if (ok) v = *reinterpret_cast<const double2*>(M + (int64_t)row * np + col);
        acc[ti][tj][0] = v.x;
        acc[ti][tj][1] = v.y;
      }
    }
  };
  double accN[4][GJ_CS][2];
  stageA(0, 0);
  loadAcc(0, accN);
  int buf = 0;
  for (int r0 = 0; r0 < np; r0 += GJ_TM, buf ^= 1) {
    const bool more = r0 + GJ_TM < np;
    if (more) stageA(r0 + GJ_TM, buf ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    double acc[4][GJ_CS][2];
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < GJ_CS; ++tj) {
        acc[ti][tj][0] = accN[ti][tj][0];
        acc[ti][tj][1] = accN[ti][tj][1];
      }
    if (more) loadAcc(r0 + GJ_TM, accN);
    const double* cA = sA + buf * (GJ_TM * SA);
#pragma unroll
    for (int kk = 0; kk < GJ_NB / 4; ++kk) {
      double a[4], b[GJ_CS];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) a[ti] = cA[((wr * 4 + ti) * 8 + fr) * SA + kk * 4 + fk];
#pragma unroll
      for (int tj = 0; tj < GJ_CS; ++tj) b[tj] = sB[(kk * 4 + fk) * SB + (wc * GJ_CS + tj) * 8 + fr];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < GJ_CS; ++tj) dmma884(acc[ti][tj][0], acc[ti][tj][1], a[ti], b[tj]);
    }
#pragma unroll
    for (int ti = 0; ti < 4; ++ti) {
      const int row = r0 + (wr * 4 + ti) * 8 + fr;
#pragma unroll
      for (int tj = 0; tj < GJ_CS; ++tj) {
        const int col = j0 + (wc * GJ_CS + tj) * 8 + 2 * fk;
        // the panel's own columns hold G' and are left alone
        if (row < np && col < j0 + jw && (col < k0 || col >= k0 + nb))
          *reinterpret_cast<double2*>(M + (int64_t)row * np + col) = make_double2(acc[ti][tj][0], acc[ti][tj][1]);
      }
    }
    __syncthreads();  // all warps are done with this buffer before the next prefetch overwrites it
  }
}

// column permutation that undoes the row interchanges: idx = arange; for c = np-1..0 swap(idx[c], idx[piv[c]])
__global__ void k_gj_perm(const int* __restrict__ npArr, const int* __restrict__ pivAll, int* __restrict__ permAll,
                          int npMax, int count) {
  const int mat = blockIdx.x;
  if (mat >= count) return;
  extern __shared__ int sIdx[];
  const int np = npArr[mat];
  const int* piv = pivAll + (int64_t)mat * npMax;
  int* perm = permAll + (int64_t)mat * npMax;
  for (int i = threadIdx.x; i < np; i += blockDim.x) sIdx[i] = i;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = np - 1; c >= 0; --c) {
      int p = piv[c];
      if (p != c) {
        int t = sIdx[c];
        sIdx[c] = sIdx[p];
        sIdx[p] = t;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < np; i += blockDim.x) perm[i] = sIdx[i];
}

// F[r][j] = W[r][perm[j]]   grid (row tiles of 8, matrices)
__global__ void k_gj_gather(const double* __restrict__ W, double* __restrict__ F, const int64_t* __restrict__ off,
                            const int* __restrict__ npArr, const int* __restrict__ permAll, int npMax) {
  const int mat = blockIdx.y;
  const int np = npArr[mat];
  const int r0 = blockIdx.x * 8;
  if (r0 >= np) return;
  const int* perm = permAll + (int64_t)mat * npMax;
  const double* Wm = W + off[mat];
  double* Fm = F + off[mat];
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    const int pj = perm[j];
#pragma unroll
    for (int r = 0; r < 8; ++r) Fm[(int64_t)(r0 + r) * np + j] = Wm[(int64_t)(r0 + r) * np + pj];
  }
}

static constexpr size_t GJ_UPD_SMEM =
    (size_t)(GJ_NB * (GJ_TJ + 4) + (2 * GJ_TM * (GJ_NB + 4) > 32 * (GJ_TJ + 4) ? 2 * GJ_TM * (GJ_NB + 4) : 32 * (GJ_TJ + 4))) *
    sizeof(double);  // B tile + max(two G' buffers, rows 32..63 of the interchange staging)

void invertBatched(double* W, double* F, const int64_t* dOff, const int* dN, const int* dNp, int count, int npMax,
                   int* dPiv, int* dPerm, int* dSwap, int* dInfo, cudaStream_t s, int64_t* launches) {
  if (count == 0 || npMax == 0) return;
  static PerDeviceLimit attrLimit, permLimit;
  if (attrLimit.raise(GJ_UPD_SMEM + 48 * 1024)) {  // once per device
    HY_CUDA(cudaFuncSetAttribute(k_gj_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GJ_UPD_SMEM));
    HY_CUDA(cudaFuncSetAttribute(k_gj_panel_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)8 * (GJ_SMEM_ROWS | 1) * sizeof(double))));
    HY_CUDA(cudaFuncSetAttribute(k_gj_panel_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)4 * (GJ_SMEM_ROWS4 | 1) * sizeof(double))));
    HY_CUDA(cudaFuncSetAttribute(k_gj_panel_full, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)GJ_NB * (GJ_FULL_ROWS | 1) * sizeof(double))));
  }
  if ((size_t)npMax * sizeof(int) > 200 * 1024)
    throw Error(HYMLS_B200_ERR_UNSUPPORTED, "dense block larger than 51200 rows");
  int* rowsT = dSwap;                       // count x 64
  int* origT = dSwap + (size_t)count * 64;  // count x 64
  k_pad_identity<<<count, 64, 0, s>>>(W, dOff, dN, dNp, count);
  ++*launches;
  for (int k0 = 0; k0 < npMax; k0 += GJ_NB) {
    const int activeMax = npMax - k0;
    // panel: matrices whose active part fits in shared memory as a whole, then the two fallbacks
    {
      const int tileRows = std::max(std::min(activeMax, GJ_FULL_ROWS), GJ_FULL_MINROWS);
      const size_t fsm = (size_t)GJ_NB * (tileRows | 1) * sizeof(double);
      k_gj_panel_full<<<count, GJ_PANEL_T, fsm, s>>>(W, dOff, dNp, dPiv, npMax, k0, dInfo, GJ_FULL_ROWS, tileRows);
      ++*launches;
    }
    if (activeMax > GJ_FULL_ROWS) {
      const size_t psm = (size_t)8 * ((std::min(activeMax, GJ_SMEM_ROWS)) | 1) * sizeof(double);
      k_gj_panel_smem<8><<<count, GJ_PANEL_T, psm, s>>>(W, dOff, dNp, dPiv, npMax, k0, dInfo, GJ_SMEM_ROWS,
                                                        GJ_FULL_ROWS);
      ++*launches;
    }
    if (activeMax > GJ_SMEM_ROWS) {
      const size_t psm = (size_t)4 * ((std::min(activeMax, GJ_SMEM_ROWS4)) | 1) * sizeof(double);
      k_gj_panel_smem<4><<<count, GJ_PANEL_T, psm, s>>>(W, dOff, dNp, dPiv, npMax, k0, dInfo, GJ_SMEM_ROWS4,
                                                        GJ_SMEM_ROWS);
      ++*launches;
    }
    if (activeMax > GJ_SMEM_ROWS4) {
      k_gj_panel<<<count, GJ_PANEL_T, 0, s>>>(W, dOff, dNp, dPiv, npMax, k0, dInfo, GJ_SMEM_ROWS4);
      ++*launches;
    }
    k_gj_swaplist<<<(count + 3) / 4, 128, 0, s>>>(dNp, dPiv, npMax, k0, count, rowsT, origT);
    dim3 g((npMax + GJ_TJ - 1) / GJ_TJ, count);
    k_gj_update<<<g, GJ_UPD_T, GJ_UPD_SMEM, s>>>(W, dOff, dNp, rowsT, origT, k0);
    *launches += 2;
  }
  if ((size_t)npMax * sizeof(int) > 48 * 1024 && permLimit.raise(200 * 1024))
    HY_CUDA(cudaFuncSetAttribute(k_gj_perm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k_gj_perm<<<count, 256, npMax * sizeof(int), s>>>(dNp, dPiv, dPerm, npMax, count);
  dim3 g2((npMax + 7) / 8, count);
  k_gj_gather<<<g2, 256, 0, s>>>(W, F, dOff, dNp, dPerm, npMax);
  *launches += 2;
  HY_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// One Newton-Schulz step  X <- X + X (I - A X)  on every inverse of a batch.
// Gauss-Jordan elimination with partial pivoting is forward stable only up to cond(U): measured against an
// extended-precision ground truth (oracle/extended.py) its inverses of the badly scaled Stokes blocks carry
// up to 60x the error of a LAPACK getrf/getri inverse, and ApplyInverse up to 80x the error of an LU solve
// (profiles/r02_accuracy_before_refinement.json).  The residual R = I - A X is formed in FP64 from the
// ORIGINAL matrix, so one step brings X to the O(cond(A) eps) of a backward-stable factorization.
// Two batched FP64 tensor-core GEMMs (DMMA m8n8k4, cp.async double-buffered 32 x 64 x 32 tiles).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884r(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
static constexpr int RF_TM = 32, RF_TN = 64, RF_TK = 32, RF_T = 128;
static constexpr int RF_SA = RF_TK + 4, RF_SB = RF_TN + 4;  // strides = 4 mod 16 doubles: conflict-free fragments
// mode 0:  C = I - A B      mode 1:  C = A B        (all np x np, row major, the batch's offsets)
__global__ void __launch_bounds__(RF_T, 4)
k_refine_gemm(const double* __restrict__ Abase, const double* __restrict__ Bbase, double* __restrict__ Cbase,
              const int64_t* __restrict__ off, const int* __restrict__ npArr, int tilesN, int mode) {
  const int mat = blockIdx.y;
  const int np = npArr[mat];
  const int i0 = (blockIdx.x / tilesN) * RF_TM, j0 = (blockIdx.x % tilesN) * RF_TN;
  if (i0 >= np || j0 >= np) return;
  const double* __restrict__ A = Abase + off[mat];
  const double* __restrict__ B = Bbase + off[mat];
  double* __restrict__ Cm = Cbase + off[mat];
  extern __shared__ __align__(16) double rfSm[];
  constexpr int SZA = RF_TM * RF_SA, SZB = RF_TK * RF_SB;
  const int tid = threadIdx.x, lane = tid & 31, wc = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  auto stage = [&](int k0, int buf) {
    for (int e = tid; e < RF_TM * (RF_TK / 2); e += RF_T) {
      const int r = e / (RF_TK / 2), c = (e % (RF_TK / 2)) * 2;
      const bool ok = (i0 + r < np) && (k0 + c < np);
      const double* src = ok ? A + (int64_t)(i0 + r) * np + k0 + c : A;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(rfSm + buf * SZA + r * RF_SA + c);
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(src), "r"(bytes));
    }
    for (int e = tid; e < RF_TK * (RF_TN / 2); e += RF_T) {
      const int r = e / (RF_TN / 2), c = (e % (RF_TN / 2)) * 2;
      const bool ok = (k0 + r < np) && (j0 + c < np);
      const double* src = ok ? B + (int64_t)(k0 + r) * np + j0 + c : B;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(rfSm + 2 * SZA + buf * SZB + r * RF_SB + c);
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(src), "r"(bytes));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  double acc[4][2][2];
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) acc[ti][tj][0] = acc[ti][tj][1] = 0.0;
  const int nk = (np + RF_TK - 1) / RF_TK;
  stage(0, 0);
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    const bool more = kt + 1 < nk;
    if (more) stage((kt + 1) * RF_TK, buf ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    const double* cA = rfSm + buf * SZA;
    const double* cB = rfSm + 2 * SZA + buf * SZB;
#pragma unroll
    for (int kk = 0; kk < RF_TK / 4; ++kk) {
      double av[4], bv[2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) av[ti] = cA[(ti * 8 + fr) * RF_SA + kk * 4 + fk];
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) bv[tj] = cB[(kk * 4 + fk) * RF_SB + (wc * 2 + tj) * 8 + fr];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) dmma884r(acc[ti][tj][0], acc[ti][tj][1], av[ti], bv[tj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int ti = 0; ti < 4; ++ti) {
    const int row = i0 + ti * 8 + fr;
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) {
      const int col = j0 + (wc * 2 + tj) * 8 + 2 * fk;
      if (row < np && col < np) {
        double v0 = acc[ti][tj][0], v1 = acc[ti][tj][1];
        if (mode == 0) {
          v0 = (row == col ? 1.0 : 0.0) - v0;
          v1 = (row == col + 1 ? 1.0 : 0.0) - v1;
        }
        *reinterpret_cast<double2*>(Cm + (int64_t)row * np + col) = make_double2(v0, v1);
      }
    }
  }
}
// X += T for every matrix of the batch
__global__ void k_refine_add(double* __restrict__ Xbase, const double* __restrict__ Tbase,
                             const int64_t* __restrict__ off, const int* __restrict__ npArr) {
  const int mat = blockIdx.y;
  const int64_t len = (int64_t)npArr[mat] * npArr[mat];
  double* __restrict__ X = Xbase + off[mat];
  const double* __restrict__ T = Tbase + off[mat];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x)
    X[e] += T[e];
}

// R = I - A X for a batch whose ORIGINAL matrices are sparse and given as the dense-fill list of their entries
// (value index `src[e]` into `val`, dense position `dst[e]`, ascending, i.e. row by row): row r of R is e_r minus a
// combination of the few rows of X that row r of A touches.  One CTA per (matrix, row); instead of a 2 n^3 GEMM the
// residual costs (nnz/row + 1) n^2 memory accesses, most of them L2 hits.  listPtr[m], listPtr[m+1]: range of the
// list for matrix m; dstBase: offset subtracted from dst (chunk-relative layout like dOff).
__global__ void __launch_bounds__(128)
k_refine_residual_sparse(const double* __restrict__ val, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                         const int64_t* __restrict__ listPtr, int64_t dstBase, const double* __restrict__ X,
                         double* __restrict__ R, const int64_t* __restrict__ off, const int* __restrict__ nArr,
                         const int* __restrict__ npArr) {
  const int mat = blockIdx.y, r = blockIdx.x;
  const int n = nArr[mat], np = npArr[mat];
  if (r >= np) return;
  const int64_t o = off[mat];
  double* __restrict__ Rrow = R + o + (int64_t)r * np;
  const double* __restrict__ Xm = X + o;
  // entries of row r: binary search in the sorted destination list
  int64_t lo = listPtr[mat], hi = listPtr[mat + 1];
  const int64_t rowStart = dstBase + o + (int64_t)r * np, rowEnd = rowStart + np;
  int64_t a = lo, b = hi;
  while (a < b) { const int64_t mid = (a + b) >> 1; if (dst[mid] < rowStart) a = mid + 1; else b = mid; }
  const int64_t e0 = a;
  b = hi;
  while (a < b) { const int64_t mid = (a + b) >> 1; if (dst[mid] < rowEnd) a = mid + 1; else b = mid; }
  const int64_t e1 = a;
  for (int c = threadIdx.x; c < np; c += blockDim.x) {
    double t = (c == r) ? 1.0 : 0.0;
    if (r < n) {
      for (int64_t e = e0; e < e1; ++e) t -= val[src[e]] * Xm[(int64_t)(dst[e] - rowStart) * np + c];
    } else {
      t -= Xm[(int64_t)r * np + c];  // padding rows of A are rows of the identity
    }
    Rrow[c] = t;
  }
}

void refineInverseSparse(const double* val, const int64_t* src, const int64_t* dst, const int64_t* listPtr,
                         int64_t dstBase, double* T, double* F, double* R, const int64_t* dOff, const int* dN,
                         const int* dNp, int count, int npMax, cudaStream_t s, int64_t* launches) {
  if (count == 0 || npMax == 0) return;
  constexpr size_t smem = (size_t)(2 * RF_TM * RF_SA + 2 * RF_TK * RF_SB) * sizeof(double);
  static PerDeviceLimit limit;
  if (limit.raise(smem + 48 * 1024))
    HY_CUDA(cudaFuncSetAttribute(k_refine_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tilesM = (npMax + RF_TM - 1) / RF_TM, tilesN = (npMax + RF_TN - 1) / RF_TN;
  for (int c0 = 0; c0 < count; c0 += 32768) {
    const int cnt = std::min(count - c0, 32768);
    k_refine_residual_sparse<<<dim3((unsigned)npMax, (unsigned)cnt), 128, 0, s>>>(val, src, dst, listPtr + c0, dstBase, F, R,
                                                                                 dOff + c0, dN + c0, dNp + c0);
    dim3 g((unsigned)(tilesM * tilesN), (unsigned)cnt);
    k_refine_gemm<<<g, RF_T, smem, s>>>(F, R, T, dOff + c0, dNp + c0, tilesN, 1);  // T = X R
    const int blocks = std::max(1, std::min(64, (npMax * npMax + 1023) / 1024));
    k_refine_add<<<dim3((unsigned)blocks, (unsigned)cnt), 256, 0, s>>>(F, T, dOff + c0, dNp + c0);
    *launches += 3;
  }
  HY_CUDA(cudaGetLastError());
}

void refineInverseBatched(double* A, double* F, double* R, const int64_t* dOff, const int* dN, const int* dNp, int count,
                          int npMax, cudaStream_t s, int64_t* launches) {
  if (count == 0 || npMax == 0) return;
  k_pad_identity<<<count, 64, 0, s>>>(A, dOff, dN, dNp, count);  // the originals come without the padding identity
  ++*launches;
  constexpr size_t smem = (size_t)(2 * RF_TM * RF_SA + 2 * RF_TK * RF_SB) * sizeof(double);
  static PerDeviceLimit limit;
  if (limit.raise(smem + 48 * 1024))
    HY_CUDA(cudaFuncSetAttribute(k_refine_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tilesM = (npMax + RF_TM - 1) / RF_TM, tilesN = (npMax + RF_TN - 1) / RF_TN;
  for (int c0 = 0; c0 < count; c0 += 32768) {  // grid.y limit
    const int cnt = std::min(count - c0, 32768);
    dim3 g((unsigned)(tilesM * tilesN), (unsigned)cnt);
    k_refine_gemm<<<g, RF_T, smem, s>>>(A, F, R, dOff + c0, dNp + c0, tilesN, 0);  // R = I - A X
    k_refine_gemm<<<g, RF_T, smem, s>>>(F, R, A, dOff + c0, dNp + c0, tilesN, 1);  // T = X R  (A is dead: reused)
    const int blocks = std::max(1, std::min(64, (npMax * npMax + 1023) / 1024));
    k_refine_add<<<dim3((unsigned)blocks, (unsigned)cnt), 256, 0, s>>>(F, A, dOff + c0, dNp + c0);
    *launches += 3;
  }
  HY_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// General dense FP64 GEMM on the tensor cores:  C (M x N, ldc) = alpha * A (M x K, lda) * B (K x N, ldb) + beta * C
// (row major; same 32 x 64 x 32 DMMA tiling as the refinement kernel; out-of-range tiles read zeros).
// Used by the block-tridiagonal coarse factorization (coarse.cu).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RF_T, 4)
k_dense_gemm(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, double* __restrict__ C,
             int ldc, int M, int N, int K, double alpha, double beta, int tilesN) {
  const int i0 = (blockIdx.x / tilesN) * RF_TM, j0 = (blockIdx.x % tilesN) * RF_TN;
  if (i0 >= M || j0 >= N) return;
  extern __shared__ __align__(16) double rfSm[];
  constexpr int SZA = RF_TM * RF_SA, SZB = RF_TK * RF_SB;
  const int tid = threadIdx.x, lane = tid & 31, wc = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  auto stage = [&](int k0, int buf) {
    for (int e = tid; e < RF_TM * (RF_TK / 2); e += RF_T) {
      const int r = e / (RF_TK / 2), c = (e % (RF_TK / 2)) * 2;
      const bool ok = (i0 + r < M) && (k0 + c < K);
      const double* src = ok ? A + (int64_t)(i0 + r) * lda + k0 + c : A;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(rfSm + buf * SZA + r * RF_SA + c);
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(src), "r"(bytes));
    }
    for (int e = tid; e < RF_TK * (RF_TN / 2); e += RF_T) {
      const int r = e / (RF_TN / 2), c = (e % (RF_TN / 2)) * 2;
      const bool ok = (k0 + r < K) && (j0 + c < N);
      const double* src = ok ? B + (int64_t)(k0 + r) * ldb + j0 + c : B;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(rfSm + 2 * SZA + buf * SZB + r * RF_SB + c);
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(src), "r"(bytes));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  double acc[4][2][2];
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) acc[ti][tj][0] = acc[ti][tj][1] = 0.0;
  const int nk = (K + RF_TK - 1) / RF_TK;
  stage(0, 0);
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    const bool more = kt + 1 < nk;
    if (more) stage((kt + 1) * RF_TK, buf ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    const double* cA = rfSm + buf * SZA;
    const double* cB = rfSm + 2 * SZA + buf * SZB;
#pragma unroll
    for (int kk = 0; kk < RF_TK / 4; ++kk) {
      double av[4], bv[2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) av[ti] = cA[(ti * 8 + fr) * RF_SA + kk * 4 + fk];
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) bv[tj] = cB[(kk * 4 + fk) * RF_SB + (wc * 2 + tj) * 8 + fr];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) dmma884r(acc[ti][tj][0], acc[ti][tj][1], av[ti], bv[tj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int ti = 0; ti < 4; ++ti) {
    const int row = i0 + ti * 8 + fr;
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) {
      const int col = j0 + (wc * 2 + tj) * 8 + 2 * fk;
      if (row < M && col < N) {
        double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * ldc + col);
        double2 old = beta != 0.0 ? *dst : make_double2(0.0, 0.0);
        *dst = make_double2(alpha * acc[ti][tj][0] + beta * old.x, alpha * acc[ti][tj][1] + beta * old.y);
      }
    }
  }
}

// all dimensions and leading dimensions must be even (they are multiples of 8 here: padded blocks)
void denseGemm(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K, double alpha,
               double beta, cudaStream_t s, int64_t* launches) {
  if (M <= 0 || N <= 0) return;
  constexpr size_t smem = (size_t)(2 * RF_TM * RF_SA + 2 * RF_TK * RF_SB) * sizeof(double);
  static PerDeviceLimit limit;
  if (limit.raise(smem + 48 * 1024))
    HY_CUDA(cudaFuncSetAttribute(k_dense_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tilesM = (M + RF_TM - 1) / RF_TM, tilesN = (N + RF_TN - 1) / RF_TN;
  k_dense_gemm<<<(unsigned)(tilesM * tilesN), RF_T, smem, s>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, beta, tilesN);
  ++*launches;
  HY_CUDA(cudaGetLastError());
}

}  // namespace hymls
