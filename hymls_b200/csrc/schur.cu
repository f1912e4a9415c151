// Schur-complement assembly with on-the-fly orthogonal transformation and dropping.
//
// Reference: SchurComplement::Construct11/Construct22 (src/HYMLS_SchurComplement.cpp:131-306),
// SchurPreconditioner::AssembleTransformAndDrop / ConstructSCPart (src/HYMLS_SchurPreconditioner.cpp:698-986),
// RestrictedOT::Apply (src/HYMLS_RestrictedOT.hpp:21-36), Householder::Apply/ApplyR (src/HYMLS_Householder.cpp:38-126).
//
// The reference builds, per subdomain, the dense m x m matrix Sk (A22 part, then -A21 A11^-1 A12 with m
// right-hand sides), applies H_g from the left and right for every separator group g and keeps only
//   (a) the V-sum x V-sum entries (first node of every group)  and
//   (b) the non-V-sum x non-V-sum entries inside each linked set of groups.
// With the explicit inverse of A11 at hand, row i of Sk is a sparse combination of a few rows of A11^-1
// followed by a sparse column gather, and the kept entries of H Sk H are bilinear forms in
//   C = Sk W (m x G),  SV = Sk[:, first] (m x G),  and the diagonal blocks S_LL of the linked sets,
// where W = blockdiag(w_g) holds the normalised reflectors (H_g = s_g (2 w_g w_g' - I), s_g = -1 and
// w_g = 0 for the degenerate groups the reference leaves untouched):
//   (H Sk H)[i,j] = s_i s_j ( Sk[i,j] - 2 w_i (W'Sk)[g(i),j] - 2 (Sk W)[i,g(j)] w_j + 4 w_i w_j (W'Sk W)[g(i),g(j)] ).
// Pass 1 (A22 part) stores, pass 2 (-A21 A11^-1 A12 part) adds -- ReplaceGlobalValues / SumIntoGlobalValues.
#include <cuda_runtime.h>

#include "device.cuh"
#include "kernels.hpp"

namespace hymls {

// one CTA (128 threads) per local separator row of a subdomain
static constexpr int SCHUR_QPT = 9;  // register-resident entries of d per thread: subdomains up to 1152 unknowns
__global__ void __launch_bounds__(128)
k_schur_rows(SchurArgs a, int64_t R0, int pass, double* __restrict__ denseS, int64_t ldS,
             const int64_t* __restrict__ rowList) {
  const int64_t R = rowList ? rowList[R0 + blockIdx.x] : R0 + blockIdx.x;
  const int sd = a.rowSd[R];
  const int64_t base = a.sdRowPtr[sd];
  const int i = (int)(R - base);
  const int m = a.sdM[sd], n = a.sdN[sd], np = a.sdNp[sd];
  const int tid = threadIdx.x, T = blockDim.x;
  extern __shared__ double sm[];
  double* d = sm;            // n   : row i of A21(sd) * A11^-1
  double* sk = sm + a.dLen;  // m   : row i of Sk

  if (pass == 2 && a.SkD != nullptr) {
    // the whole row of -A21 A11^-1 A12 comes from the two dense GEMMs (schurGemm)
    const int mp = (m + 7) & ~7;
    const double* Srow = a.SkD + a.wsOffS[sd] + (int64_t)i * mp;
    for (int j = tid; j < m; j += T) sk[j] = Srow[j];
  } else if (pass == 2) {
    // d = sum_e A21[i, col_e] * Ainv[col_e, :]: every thread keeps its entries of d in registers while the
    // (few at level 0, ~100 at the coarser levels) rows of A11^-1 stream by, two rows in flight
    const double* Ainv = a.Ainv + a.a11Off[sd];
    const int64_t e0 = a.s21Ptr[R], e1 = a.s21Ptr[R + 1];
    if (a.D != nullptr) {  // row i of A21 A11^-1 was computed by the dense GEMM (schurGemm)
      const double* Drow = a.D + a.wsOffD[sd] + (int64_t)i * np;
      for (int q = tid; q < n; q += T) d[q] = Drow[q];
    } else if (n <= SCHUR_QPT * 128) {
      double acc[SCHUR_QPT];
#pragma unroll
      for (int j = 0; j < SCHUR_QPT; ++j) acc[j] = 0.0;
      int64_t e = e0;
      for (; e + 1 < e1; e += 2) {
        const double v0 = a.val[a.s21Src[e]], v1 = a.val[a.s21Src[e + 1]];
        if (v0 == 0.0 && v1 == 0.0) continue;  // stored zeros (dropped entries keep their slot): nothing to stream
        const double* r0p = Ainv + (int64_t)a.s21Col[e] * np;
        const double* r1p = Ainv + (int64_t)a.s21Col[e + 1] * np;
#pragma unroll
        for (int j = 0; j < SCHUR_QPT; ++j) {
          const int q = tid + j * 128;
          if (q < n) acc[j] += v0 * r0p[q] + v1 * r1p[q];
        }
      }
      if (e < e1) {
        const double v0 = a.val[a.s21Src[e]];
        const double* r0p = Ainv + (int64_t)a.s21Col[e] * np;
#pragma unroll
        for (int j = 0; j < SCHUR_QPT; ++j) {
          const int q = tid + j * 128;
          if (q < n) acc[j] += v0 * r0p[q];
        }
      }
#pragma unroll
      for (int j = 0; j < SCHUR_QPT; ++j) {
        const int q = tid + j * 128;
        if (q < n) d[q] = acc[j];
      }
    } else {
      for (int q = tid; q < n; q += T) d[q] = 0.0;
      __syncthreads();
      for (int64_t e = e0; e < e1; ++e) {
        const double v = a.val[a.s21Src[e]];
        const double* row = Ainv + (int64_t)a.s21Col[e] * np;
        for (int q = tid; q < n; q += T) d[q] += v * row[q];
      }
    }
    __syncthreads();
    for (int j = tid; j < m; j += T) {
      double s = 0.0;
      for (int64_t e = a.s12Ptr[base + j]; e < a.s12Ptr[base + j + 1]; ++e) s += d[a.s12Row[e]] * a.val[a.s12Src[e]];
      sk[j] = -s;
    }
  } else {
    for (int j = tid; j < m; j += T) sk[j] = 0.0;
    __syncthreads();
    for (int64_t e = a.s22Ptr[R] + tid; e < a.s22Ptr[R + 1]; e += T) sk[a.s22Col[e]] = a.val[a.s22Src[e]];
  }
  __syncthreads();

  if (denseS != nullptr) {  // Number of Levels = 0: the full Schur complement, dense (Construct :88-129)
    // plain adds: the rows of one launch belong to subdomains of one colour (no shared separator group),
    // so no two CTAs touch the same entry and the order of the sum over subdomains is fixed by the launches
    const int64_t prow = a.sdSep[base + i];
    for (int j = tid; j < m; j += T)
      if (sk[j] != 0.0) denseS[prow * ldS + a.sdSep[base + j]] += sk[j];
    return;
  }

  const int64_t ia = a.sdInstPtr[sd];
  const int G = (int)(a.sdInstPtr[sd + 1] - ia);
  double* C = a.wsC + a.wsOffC[sd] + (int64_t)i * G;
  double* SV = a.wsSV + a.wsOffC[sd] + (int64_t)i * G;
  for (int h = tid; h < G; h += T) {
    const int loc = a.instLoc[ia + h], len = a.instLen[ia + h];
    const double* w = a.wd + a.uniqStart[a.instUniq[ia + h]];
    double c = 0.0;
    for (int q = 0; q < len; ++q) c += sk[loc + q] * w[q];
    C[h] = c;
    SV[h] = sk[loc];
  }
  const int myLink = a.instLink[a.rowInst[R]];
  const int64_t lk = a.sdLinkPtr[sd] + myLink;
  const int lsz = a.lnkSize[lk];
  double* SLL = a.wsSLL + a.lnkOff[lk] + (int64_t)a.rowLinkPos[R] * lsz;
  for (int j = tid; j < m; j += T)
    if (a.instLink[a.rowInst[base + j]] == myLink) SLL[a.rowLinkPos[base + j]] = sk[j];
}

__device__ __forceinline__ int64_t findCol(const int* __restrict__ col, int64_t lo, int64_t hi, int c) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// (a) V-sum x V-sum entries: one CTA per subdomain
__global__ void __launch_bounds__(256)
k_schur_vsum(SchurArgs a, int sd0, int pass, const int* __restrict__ sdList) {
  const int sd = sdList ? sdList[sd0 + blockIdx.x] : sd0 + blockIdx.x;
  const int64_t ia = a.sdInstPtr[sd];
  const int G = (int)(a.sdInstPtr[sd + 1] - ia);
  const double* C = a.wsC + a.wsOffC[sd];
  const double* SV = a.wsSV + a.wsOffC[sd];
  for (int idx = threadIdx.x; idx < G * G; idx += blockDim.x) {
    const int g = idx / G, h = idx % G;
    const int locg = a.instLoc[ia + g], leng = a.instLen[ia + g], ug = a.instUniq[ia + g];
    const int uh = a.instUniq[ia + h];
    const double* wg = a.wd + a.uniqStart[ug];
    double Mgh = 0.0, RV = 0.0;
    for (int q = 0; q < leng; ++q) {
      Mgh += wg[q] * C[(int64_t)(locg + q) * G + h];
      RV += wg[q] * SV[(int64_t)(locg + q) * G + h];
    }
    const double w0g = wg[0], w0h = a.wd[a.uniqStart[uh]];
    double v = SV[(int64_t)locg * G + h] - 2.0 * w0g * RV - 2.0 * C[(int64_t)locg * G + h] * w0h +
               4.0 * w0g * w0h * Mgh;
    v *= a.usign[ug] * a.usign[uh];
    const int64_t pos = findCol(a.redCol, a.redPtr[ug], a.redPtr[ug + 1], uh);
    // pass 2 launches cover subdomains of ONE colour (no shared separator group): plain, race-free adds in a
    // fixed order instead of FP64 atomics -> Compute is bitwise reproducible
    if (pass == 1) a.redVal[pos] = v; else a.redVal[pos] += v;
  }
}

// (b) non-V-sum entries of one linked set: one CTA per (subdomain, linked set)
static constexpr int MAX_LINK_INST = 32;
__global__ void __launch_bounds__(256)
k_schur_blocks(SchurArgs a, int64_t lk0, int pass, const int64_t* __restrict__ lkList) {
  const int64_t lk = lkList ? lkList[lk0 + blockIdx.x] : lk0 + blockIdx.x;
  const int sd = a.lnkSd[lk];
  const int link = (int)(lk - a.sdLinkPtr[sd]);
  const int lsz = a.lnkSize[lk];
  const int64_t ia = a.sdInstPtr[sd];
  const int G = (int)(a.sdInstPtr[sd + 1] - ia);
  const double* SLL = a.wsSLL + a.lnkOff[lk];
  const int tid = threadIdx.x, T = blockDim.x;

  __shared__ int sStart[MAX_LINK_INST], sLen[MAX_LINK_INST], sU[MAX_LINK_INST], sW0[MAX_LINK_INST];
  __shared__ int sT;
  extern __shared__ double sm[];
  if (tid == 0) {
    int t = 0, off = 0;
    for (int g = 0; g < G; ++g)
      if (a.instLink[ia + g] == link) {
        if (t < MAX_LINK_INST) {
          sStart[t] = off;
          sLen[t] = a.instLen[ia + g];
          sU[t] = a.instUniq[ia + g];
          sW0[t] = a.uniqStart[sU[t]];
        }
        off += a.instLen[ia + g];
        ++t;
      }
    sT = t;
  }
  __syncthreads();
  const int nt = sT;
  if (nt > MAX_LINK_INST) { if (tid == 0) atomicExch(a.info, -7); return; }
  if (lsz <= nt) return;  // only V-sums in this set: nothing to keep
  double* CL = sm;                 // lsz x nt
  double* RL = sm + (int64_t)lsz * nt;  // nt x lsz
  double* ML = RL + (int64_t)lsz * nt;  // nt x nt
  for (int e = tid; e < lsz * nt; e += T) {
    const int i = e / nt, t = e % nt;
    const double* w = a.wd + sW0[t];
    double c = 0.0, r = 0.0;
    for (int q = 0; q < sLen[t]; ++q) {
      c += SLL[(int64_t)i * lsz + sStart[t] + q] * w[q];
      r += w[q] * SLL[(int64_t)(sStart[t] + q) * lsz + i];
    }
    CL[i * nt + t] = c;
    RL[t * lsz + i] = r;
  }
  __syncthreads();
  for (int e = tid; e < nt * nt; e += T) {
    const int t = e / nt, t2 = e % nt;
    const double* w = a.wd + sW0[t];
    double s = 0.0;
    for (int q = 0; q < sLen[t]; ++q) s += w[q] * CL[(sStart[t] + q) * nt + t2];
    ML[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < lsz * lsz; e += T) {
    const int i = e / lsz, j = e % lsz;
    int ti = 0, tj = 0;
    while (ti + 1 < nt && sStart[ti + 1] <= i) ++ti;
    while (tj + 1 < nt && sStart[tj + 1] <= j) ++tj;
    const int qi = i - sStart[ti], qj = j - sStart[tj];
    if (qi == 0 || qj == 0) continue;  // V-sum rows/cols are not part of the block
    const int ui = sU[ti], uj = sU[tj];
    const int b = a.uniqBlk[ui];
    if (b != a.uniqBlk[uj]) continue;  // pattern entry no block solver ever reads
    const double wi = a.wd[sW0[ti] + qi], wj = a.wd[sW0[tj] + qj];
    double v = SLL[(int64_t)i * lsz + j] - 2.0 * wi * RL[ti * lsz + j] - 2.0 * CL[i * nt + tj] * wj +
               4.0 * wi * wj * ML[ti * nt + tj];
    v *= a.usign[ui] * a.usign[uj];
    const int npb = a.blkNp[b];
    double* dst = a.blkW + a.blkOff[b] + (int64_t)(a.uniqBlkOff[ui] + qi - 1) * npb + (a.uniqBlkOff[uj] + qj - 1);
    if (pass == 1) *dst = v; else *dst += v;
  }
}

// ---------------------------------------------------------------------------------------------
// Dense path for the coarser levels, where a separator row couples to ~100 interior nodes: instead of
// streaming that many rows of A11^-1 per separator row, A21(sd) is densified and D = A21(sd) A11(sd)^-1 is one
// FP64 tensor-core GEMM per subdomain (mma.sync.m8n8k4, cp.async double-buffered 32 x 32 / 32 x 64 tiles).
// ---------------------------------------------------------------------------------------------
__global__ void k_schur_densify(SchurArgs a, int64_t R0, int64_t R1, const int64_t* __restrict__ rowList) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R1 - R0) return;
  const int64_t R = rowList ? rowList[R0 + t] : R0 + t;
  const int sd = a.rowSd[R];
  const int i = (int)(R - a.sdRowPtr[sd]);
  double* row = a.A21d + a.wsOffD[sd] + (int64_t)i * a.sdNp[sd];
  for (int64_t e = a.s21Ptr[R]; e < a.s21Ptr[R + 1]; ++e) row[a.s21Col[e]] = a.val[a.s21Src[e]];
}

// A12d[row, j] for the local separator column j = R - sdRowPtr[sd] (np x mp, row major)
__global__ void k_schur_densify12(SchurArgs a, int64_t R0, int64_t R1, const int64_t* __restrict__ rowList) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R1 - R0) return;
  const int64_t R = rowList ? rowList[R0 + t] : R0 + t;
  const int sd = a.rowSd[R];
  const int j = (int)(R - a.sdRowPtr[sd]);
  const int mp = (a.sdM[sd] + 7) & ~7;
  double* col = a.A12d + a.wsOffA[sd] + j;
  for (int64_t e = a.s12Ptr[R]; e < a.s12Ptr[R + 1]; ++e) col[(int64_t)a.s12Row[e] * mp] = a.val[a.s12Src[e]];
}

__device__ __forceinline__ void dmma884s(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

static constexpr int SG_TM = 32, SG_TN = 64, SG_TK = 32, SG_T = 128;
static constexpr int SG_SA = SG_TK + 4, SG_SB = SG_TN + 4;  // strides = 4 mod 16 doubles: conflict-free fragments
__global__ void __launch_bounds__(SG_T, 4)
k_schur_gemm(SchurArgs a, int sd0, const int* __restrict__ sdList, int tilesN, int which) {
  // which 0:  D (m x np)  =  A21d (m x np) * Ainv (np x np)
  // which 1:  SkD (m x mp) = -D (m x np) * A12d (np x mp)
  const int sd = sdList ? sdList[sd0 + blockIdx.y] : sd0 + blockIdx.y;
  const int m = a.sdM[sd], n = a.sdN[sd], np = a.sdNp[sd];
  if (n == 0 && !which) return;  // no interior: D is m x 0; SkD (which = 1) still has to be written (zeros)
  const int N = which ? ((m + 7) & ~7) : np;  // columns of B and C (= their leading dimension)
  const int i0 = (blockIdx.x / tilesN) * SG_TM, j0 = (blockIdx.x % tilesN) * SG_TN;
  if (i0 >= m || j0 >= N) return;
  const double* __restrict__ A = (which ? a.D : a.A21d) + a.wsOffD[sd];                        // m x np
  const double* __restrict__ B = which ? a.A12d + a.wsOffA[sd] : a.Ainv + a.a11Off[sd];        // np x N
  double* __restrict__ Dm = which ? a.SkD + a.wsOffS[sd] : a.D + a.wsOffD[sd];                 // m x N
  const double sgn = which ? -1.0 : 1.0;
  extern __shared__ __align__(16) double sgSm[];
  constexpr int SZA = SG_TM * SG_SA, SZB = SG_TK * SG_SB;  // sA[buf] = sgSm + buf*SZA, sB[buf] = sgSm + 2*SZA + buf*SZB
  const int tid = threadIdx.x, lane = tid & 31, wc = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  auto stage = [&](int k0, int buf) {
    for (int e = tid; e < SG_TM * (SG_TK / 2); e += SG_T) {
      const int r = e / (SG_TK / 2), c = (e % (SG_TK / 2)) * 2;
      const bool ok = (i0 + r < m) && (k0 + c < np);
      const double* src = ok ? A + (int64_t)(i0 + r) * np + k0 + c : A;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(sgSm + buf * SZA + r * SG_SA + c);
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(src), "r"(bytes));
    }
    for (int e = tid; e < SG_TK * (SG_TN / 2); e += SG_T) {
      const int r = e / (SG_TN / 2), c = (e % (SG_TN / 2)) * 2;
      const bool ok = (k0 + r < np) && (j0 + c < N);
      const double* src = ok ? B + (int64_t)(k0 + r) * N + j0 + c : B;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(sgSm + 2 * SZA + buf * SZB + r * SG_SB + c);
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
  const int nk = (np + SG_TK - 1) / SG_TK;
  if (nk > 0) stage(0, 0);
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    const bool more = kt + 1 < nk;
    if (more) stage((kt + 1) * SG_TK, buf ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    const double* cA = sgSm + buf * SZA;
    const double* cB = sgSm + 2 * SZA + buf * SZB;
#pragma unroll
    for (int kk = 0; kk < SG_TK / 4; ++kk) {
      double av[4], bv[2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) av[ti] = cA[(ti * 8 + fr) * SG_SA + kk * 4 + fk];
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) bv[tj] = cB[(kk * 4 + fk) * SG_SB + (wc * 2 + tj) * 8 + fr];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) dmma884s(acc[ti][tj][0], acc[ti][tj][1], av[ti], bv[tj]);
    }
    __syncthreads();  // the buffer is free for the prefetch after next
  }
#pragma unroll
  for (int ti = 0; ti < 4; ++ti) {
    const int row = i0 + ti * 8 + fr;
#pragma unroll
    for (int tj = 0; tj < 2; ++tj) {
      const int col = j0 + (wc * 2 + tj) * 8 + 2 * fk;
      if (row < m && col < N)
        *reinterpret_cast<double2*>(Dm + (int64_t)row * N + col) =
            make_double2(sgn * acc[ti][tj][0], sgn * acc[ti][tj][1]);
    }
  }
}

void schurGemm(const SchurArgs& a, int sd0, int sd1, int64_t R0, int64_t R1, int64_t dLen, int64_t aLen, int maxM,
               int maxNp, cudaStream_t s, int64_t* launches, const int* sdList, const int64_t* rowList) {
  if (sd1 <= sd0 || R1 <= R0 || dLen <= 0) return;
  constexpr size_t smem = (size_t)(2 * SG_TM * SG_SA + 2 * SG_TK * SG_SB) * sizeof(double);
  static PerDeviceLimit gemmLimit;
  if (gemmLimit.raise(smem))
    HY_CUDA(cudaFuncSetAttribute(k_schur_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned rowBlocks = (unsigned)((R1 - R0 + 127) / 128);
  const int tilesM = (maxM + SG_TM - 1) / SG_TM;
  HY_CUDA(cudaMemsetAsync(a.A21d, 0, (size_t)dLen * sizeof(double), s));
  k_schur_densify<<<rowBlocks, 128, 0, s>>>(a, R0, R1, rowList);
  {
    const int tilesN = (maxNp + SG_TN - 1) / SG_TN;
    dim3 g((unsigned)(tilesM * tilesN), (unsigned)(sd1 - sd0));
    k_schur_gemm<<<g, SG_T, smem, s>>>(a, sd0, sdList, tilesN, 0);
  }
  *launches += 2;
  if (a.SkD != nullptr) {
    HY_CUDA(cudaMemsetAsync(a.A12d, 0, (size_t)aLen * sizeof(double), s));
    k_schur_densify12<<<rowBlocks, 128, 0, s>>>(a, R0, R1, rowList);
    const int maxMp = (maxM + 7) & ~7;
    const int tilesN = (maxMp + SG_TN - 1) / SG_TN;
    dim3 g((unsigned)(tilesM * tilesN), (unsigned)(sd1 - sd0));
    k_schur_gemm<<<g, SG_T, smem, s>>>(a, sd0, sdList, tilesN, 1);
    *launches += 2;
  }
  HY_CUDA(cudaGetLastError());
}

// RelDropDiag / RelFullDiag value dropping (MatrixUtils::DropByValue :1010-1194) applied in place:
// entries |a_ij| <= 1e-14 * max(|a_ii|,|a_jj|) (or <= 1e-14) become exact zeros; the pattern is kept.
__global__ void k_drop_by_value(double* __restrict__ val, const int64_t* __restrict__ ptr, const int* __restrict__ col,
                                const double* __restrict__ diag, int n, double tol) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double di = fabs(diag[r]);
  for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
    const int c = col[e];
    const double v = fabs(val[e]);
    const double scal = (c == r) ? 1.0 : fmax(di, fabs(diag[c]));
    if (!(v > scal * tol && v > tol)) val[e] = 0.0;
  }
}
__global__ void k_extract_diag(const double* __restrict__ val, const int64_t* __restrict__ ptr,
                               const int* __restrict__ col, double* __restrict__ diag, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0.0;
  for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e)
    if (col[e] == r) d = val[e];
  diag[r] = d;
}

void dropByValue(double* val, const int64_t* ptr, const int* col, double* diagScratch, int n, double tol,
                 cudaStream_t s, int64_t* launches) {
  if (n == 0) return;
  k_extract_diag<<<(n + 255) / 256, 256, 0, s>>>(val, ptr, col, diagScratch, n);
  k_drop_by_value<<<(n + 255) / 256, 256, 0, s>>>(val, ptr, col, diagScratch, n, tol);
  *launches += 2;
}

// opt-in dynamic shared memory above 48 KB (grow-only, tracked per device: PerDeviceLimit)
static void ensureRowSmem(size_t rowSmem) {
  static PerDeviceLimit limit;
  if (rowSmem > 227 * 1024)
    throw Error(HYMLS_B200_ERR_UNSUPPORTED, "subdomain too large for the Schur row kernel (n + m > 29000)");
  if (limit.raise(rowSmem)) HY_CUDA(cudaFuncSetAttribute(k_schur_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowSmem));
}
static void ensureBlkSmem(size_t blkSmem) {
  static PerDeviceLimit limit;
  if (blkSmem > 200 * 1024)
    throw Error(HYMLS_B200_ERR_UNSUPPORTED, "linked separator set too large for the block kernel");
  if (limit.raise(blkSmem)) HY_CUDA(cudaFuncSetAttribute(k_schur_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blkSmem));
}

void schurRows(const SchurArgs& a, int64_t R0, int64_t R1, int pass, size_t rowSmem, cudaStream_t s,
               int64_t* launches, const int64_t* rowList) {
  if (R1 <= R0) return;
  ensureRowSmem(rowSmem);
  k_schur_rows<<<(unsigned)(R1 - R0), 128, rowSmem, s>>>(a, R0, pass, nullptr, 0, rowList);
  ++*launches;
  HY_CUDA(cudaGetLastError());
}

void schurScatter(const SchurArgs& a, int pass, int sd0, int sd1, int64_t lk0, int64_t lk1, size_t blkSmem,
                  cudaStream_t s, int64_t* launches, const int* sdList, const int64_t* lkList) {
  ensureBlkSmem(blkSmem);
  if (sd1 > sd0) {
    k_schur_vsum<<<sd1 - sd0, 256, 0, s>>>(a, sd0, pass, sdList);
    ++*launches;
  }
  if (lk1 > lk0) {
    k_schur_blocks<<<(unsigned)(lk1 - lk0), 256, blkSmem, s>>>(a, lk0, pass, lkList);
    ++*launches;
  }
  HY_CUDA(cudaGetLastError());
}

void schurDense(const SchurArgs& a, int64_t R0, int64_t R1, double* denseS, int64_t ldS, size_t rowSmem,
                cudaStream_t s, int64_t* launches, const int64_t* rowList) {
  if (R1 <= R0) return;
  ensureRowSmem(rowSmem);
  k_schur_rows<<<(unsigned)(R1 - R0), 128, rowSmem, s>>>(a, R0, 2, denseS, ldS, rowList);
  ++*launches;
  HY_CUDA(cudaGetLastError());
}

}  // namespace hymls
