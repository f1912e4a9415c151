// Microbenchmark: FP64 mma.sync shapes on sm_100a (throughput per shape + fragment-layout check).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE>
__global__ void __launch_bounds__(256) k_tp(double* out, int iters) {
  const int lane = threadIdx.x & 31;
  double acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (lane + i);
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (lane - i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0) { mma884(acc[i][0], acc[i][1], a[i & 7], b[i & 3]); mma884(acc[i][2], acc[i][3], a[(i + 1) & 7], b[(i + 1) & 3]); }
      if (SHAPE == 1) mma1684(acc[i], a + (i & 3) * 2, b[i & 3]);
      if (SHAPE == 2) mma1688(acc[i], a + (i & 1) * 4, b + (i & 1) * 2);
      if (SHAPE == 3) mma16816(acc[i], a, b);
      if (SHAPE == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// layout check: C(16x8) = A(16xK) * B(Kx8) with the assumed fragment layouts
template <int K>
__global__ void k_check(const double* A, const double* B, double* C) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double c[4] = {0, 0, 0, 0};
  if (K == 4) {
    double a[2] = {A[g * K + t], A[(g + 8) * K + t]};
    mma1684(c, a, B[t * 8 + g]);
  } else if (K == 8) {
    double a[4] = {A[g * K + t], A[(g + 8) * K + t], A[g * K + t + 4], A[(g + 8) * K + t + 4]};
    double b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
    mma1688(c, a, b);
  } else {
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
    for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
    mma16816(c, a, b);
  }
  C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1];
  C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
}

template <int K> void check() {
  std::vector<double> A(16 * K), B(K * 8), C(128), R(128, 0.0);
  for (int i = 0; i < 16 * K; ++i) A[i] = sin(0.37 * i + 1);
  for (int i = 0; i < K * 8; ++i) B[i] = cos(0.11 * i + 2);
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) for (int k = 0; k < K; ++k) R[i * 8 + j] += A[i * K + k] * B[k * 8 + j];
  double *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, 128 * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  k_check<K><<<1, 32>>>(dA, dB, dC);
  cudaMemcpy(C.data(), dC, 128 * 8, cudaMemcpyDeviceToHost);
  double e = 0; for (int i = 0; i < 128; ++i) e = fmax(e, fabs(C[i] - R[i]));
  printf("layout check m16n8k%d: max err %.3e (%s)\n", K, e, cudaGetErrorString(cudaGetLastError()));
}

template <int SHAPE> void run(const char* name, double flopsPerIterPerWarp) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_tp<SHAPE><<<148 * ctas, 256>>>(out, 100);
    cudaEventRecord(e0);
    k_tp<SHAPE><<<148 * ctas, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = (double)148 * ctas * 8 * iters * flopsPerIterPerWarp;
    printf("%-10s %d CTA/SM x 8 warps: %.2f TFLOP/s (%s)\n", name, ctas, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
}

int main() {
  check<4>(); check<8>(); check<16>();
  run<0>("m8n8k4", 16 * 512.0);
  run<1>("m16n8k4", 8 * 1024.0);
  run<2>("m16n8k8", 8 * 2048.0);
  run<3>("m16n8k16", 8 * 4096.0);
  run<4>("dfma", 8 * 4 * 2 * 32.0);
  return 0;
}
