// Probe 3: toggle pieces of the raw-PTX TMA sequence (debug tool).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                        CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// flags: 1 = skip fence.mbarrier_init, 2 = expect_tx after copy, 4 = static smem, 8 = explicit .tile,
//        16 = syncthreads between init and copy, 32 = skip fence.proxy.async
__global__ void k(const __grid_constant__ CUtensorMap m, uint8_t *o, int x, int y, int bytes, int flags) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(128) uint8_t ssm[8192 + 64];
  uint8_t *sm = (flags & 4) ? ssm : dsm;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 8192);
  const uint32_t b = s32(bar), dst = s32(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    if (!(flags & 1)) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (!(flags & 32)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (flags & 16) __syncthreads();
  if (threadIdx.x == 0) {
    if (!(flags & 2)) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (flags & 8)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(&m), "r"(b), "r"(x), "r"(y) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(&m), "r"(b), "r"(x), "r"(y) : "memory");
    if (flags & 2) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n"
                 : "=r"(ok) : "r"(b), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) o[i] = sm[i];
}
int main(int argc, char **argv) {
  const int flags = argc > 1 ? atoi(argv[1]) : 0;
  const int x = argc > 2 ? atoi(argv[2]) : 37;
  const int W = 320, H = 240, TW = 64, TH = 32;
  std::vector<uint8_t> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, 16384);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m; memset(&m, 0, sizeof m);
  cuuint64_t dims[2] = {W, H}; cuuint64_t st[1] = {W}; cuuint32_t box[2] = {TW, TH}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((PFN)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  k<<<1, 128, 16384>>>(m, o, x, 8, TW * TH, flags);
  cudaError_t e = cudaDeviceSynchronize();
  int bad = -1;
  if (e == cudaSuccess) {
    std::vector<uint8_t> out(TW * TH); cudaMemcpy(out.data(), o, TW * TH, cudaMemcpyDeviceToHost);
    bad = 0;
    for (int r2 = 0; r2 < TH; ++r2) for (int c = 0; c < TW; ++c) {
      uint8_t exp = (x + c < W) ? h[(size_t)(8 + r2) * W + x + c] : 0;
      if (out[r2 * TW + c] != exp) ++bad;
    }
  }
  printf("flags %2d x %3d encode=%d sync: %s mismatches %d\n", flags, x, (int)r, cudaGetErrorString(e), bad);
  return 0;
}
