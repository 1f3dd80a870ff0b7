// Microbenchmark (debug tool): FP64 DMMA m8n8k4 issue behaviour on B200 -- throughput against resident warps
// per SM and independent accumulator chains per warp, with operands from registers or from shared memory,
// plus single-warp latencies (dependent DMMA, DFMA, 64-bit shuffle, rsqrt+Newton).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dmma_probe tools/dmma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH, bool SMEM>
__global__ void k(double *out, int iters) {
  __shared__ double sh[64 * 20];
  for (int i = threadIdx.x; i < 64 * 20; i += blockDim.x) sh[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999, c[2 * CH];
  for (int i = 0; i < 2 * CH; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8 / CH; ++r)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        double av = a;
        if (SMEM) av = sh[(((it + r) & 15) * 4 + lc) * 20 + lr + (i & 1) * 8];
        dmma(c[2 * i], c[2 * i + 1], av, b);
      }
  }
  double s = 0;
  for (int i = 0; i < 2 * CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ double pivot_rsqrt(double dv) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(dv));
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-(dv * y), y, 1.0);
    y = fma(0.5 * y, e, y);
  }
  return y;
}
__global__ void lat(double *out, long long *cyc, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.9999999, c0 = 0.1, c1 = 0.2;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) dmma(c0, c1, a, b);
  long long t1 = clock64();
  double f = a;
  for (int it = 0; it < iters; ++it) f = fma(f, b, a);
  long long t2 = clock64();
  double sv = a;
  for (int it = 0; it < iters; ++it) sv = __shfl_sync(0xffffffffu, sv, (threadIdx.x + 1) & 31) + 1e-9;
  long long t3 = clock64();
  double r = 2.0 + a;
  for (int it = 0; it < iters; ++it) r = pivot_rsqrt(r) + 2.0;
  long long t4 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3;
  }
  out[threadIdx.x] = c0 + c1 + f + sv + r;
}
template <int CH, bool SMEM>
void run(double *out, int warps_per_sm) {
  const int iters = 4000;
  int ctas = 1, thr = warps_per_sm * 32;
  if (thr > 1024) { ctas = 2; thr /= 2; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH, SMEM><<<148 * ctas, thr>>>(out, 50);
  cudaEventRecord(e0);
  k<CH, SMEM><<<148 * ctas, thr>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 148.0 * warps_per_sm * (double)iters * 8 * 512.0;
  printf("warps/SM %2d chains %d smemA %d: %.2f TFLOP/s\n", warps_per_sm, CH, (int)SMEM, flop / (ms * 1e-3) / 1e12);
}
int main() {
  double *out; cudaMalloc(&out, 148 * 2048 * 8);
  long long *cyc; cudaMalloc(&cyc, 64);
  for (int w : {4, 8, 12, 16, 32, 64}) {
    run<1, false>(out, w); run<2, false>(out, w); run<4, false>(out, w); run<8, false>(out, w);
    run<4, true>(out, w); run<8, true>(out, w);
  }
  lat<<<1, 32>>>(out, cyc, 1000);
  long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("latency cycles: dependent DMMA %.1f, DFMA %.1f, shfl64+add %.1f, rsqrt+2 Newton(+add) %.1f\n", h[0] / 1000.0,
         h[1] / 1000.0, h[2] / 1000.0, h[3] / 1000.0);
  return 0;
}
