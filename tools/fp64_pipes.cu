// Microbenchmark: do DFMA (FP64 pipe) and DMMA m8n8k4 (tensor path) overlap on B200?  (debug tool)
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DFMA, odd warps DMMA
__global__ void k(double *out, int iters, int mode) {
  const int warp = threadIdx.x >> 5;
  double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999, c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  const bool use_dmma = mode == 1 || (mode == 2 && (warp & 1));
  if (use_dmma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out; cudaMalloc(&out, 148 * 4 * 512 * 8);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 2, 512>>>(out, 100, mode);
    cudaEventRecord(e0);
    k<<<148 * 2, 512>>>(out, iters, mode);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // FMA counts: DFMA warp-iter = 64 instr * 32 lanes = 2048 FMA; DMMA warp-iter = 8 * 256 = 2048 FMA
    const double fma = 148.0 * 2 * 16 * iters * 2048.0;
    printf("mode %d: %.3f ms  %.2f TFLOP/s (2*FMA)\n", mode, ms, 2 * fma / (ms * 1e-3) / 1e12);
  }
  return 0;
}
