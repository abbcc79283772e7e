// Latency microbenchmarks for the FP64 building blocks of the QR loop (sm_100a).  One warp / one CTA variants.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_ilp4(double* out, long long* cyc, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + b;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_add(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sqrt(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = sqrt(x) + a;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / x + a;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x) + a;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = nxt[p];
  long long t1 = clock64();
  out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc) {
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) __syncthreads();
  long long t1 = clock64();
  out[threadIdx.x] = 0; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_redux(double* out, long long* cyc) {
  unsigned x = threadIdx.x * 2654435761u;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __reduce_max_sync(0xffffffffu, x) + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: 16 warps x ILP-8 DFMA
__global__ void k_dfma_tp(double* out, long long* cyc, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = a + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fma(x[k], b, a);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192 * 8); cudaMalloc(&cyc, 8);
  long long h;
#define RUN(name, threads, ...) name<<<1, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); name<<<1, threads>>>(__VA_ARGS__); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-14s threads %4d  %.1f cycles/iter\n", #name, threads, (double)h / N);
  RUN(k_dfma, 32, out, cyc, 1.0000001, 0.9999999)
  RUN(k_dfma_ilp4, 32, out, cyc, 1.0000001, 0.9999999)
  RUN(k_dadd, 32, out, cyc, 1.0, 1e-9)
  RUN(k_shfl, 32, out, cyc, 1.0)
  RUN(k_shfl_add, 32, out, cyc, 1.0)
  RUN(k_sqrt, 32, out, cyc, 1.5)
  RUN(k_rcp, 32, out, cyc, 1.5)
  RUN(k_rsqrt, 32, out, cyc, 1.5)
  RUN(k_lds, 32, out, cyc)
  RUN(k_redux, 32, out, cyc)
  RUN(k_bar, 512, out, cyc)
  RUN(k_bar, 128, out, cyc)
  RUN(k_dfma_tp, 512, out, cyc, 1.0000001, 0.9999999)
  RUN(k_dfma_tp, 128, out, cyc, 1.0000001, 0.9999999)
  RUN(k_dfma, 512, out, cyc, 1.0000001, 0.9999999)
  RUN(k_sqrt, 512, out, cyc, 1.5)
  cudaError_t e = cudaGetLastError(); printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
