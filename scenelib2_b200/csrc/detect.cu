// detect.cu — Shi-Tomasi "best patch" detector on sm_100a.
//
// Replaces MonoSLAM::find_best_patch_inside_region + find_eigenvalues (monoslam.cpp:1070-1205):
// the smaller eigenvalue of the BOXSIZE x BOXSIZE gradient structure tensor, maximised over a
// region, first maximum in (v-major, u-minor) scan order (strict `>` against evbest = 0).
//
// The reference keeps running double sums of gx*gx etc. with gx = (I[c+1]-I[c-1])/2.0.  Those are
// multiples of 0.25 far below 2^53, so every running sum is EXACT whatever the order; the kernel
// therefore forms the exact int32 sums Sxx = sum dx^2 ... per position directly and converts with
// one exact multiply by 0.25.  The eigenvalue formula is evaluated op for op with never-fused
// __d*_rn (same bits as the x86-64 SSE2 build).  One CTA per region, grid-stride over positions.
#include "sl2_common.cuh"

namespace {

struct DBest {
  double ev;
  int idx;
};

__device__ __forceinline__ void dconsider(DBest &b, double ev, int idx) {
  // strict `eval2 > *evbest` in scan order: larger wins, equal keeps the earlier position
  if (ev > 0.0 && (ev > b.ev || (ev == b.ev && idx < b.idx))) {
    b.ev = ev;
    b.idx = idx;
  }
}

template <int BOX>
__global__ void __launch_bounds__(256) detect_kernel(const Sl2Dev d, int stream, int slot,
                                                     const int *regions, int *out_uv, double *out_ev) {
  constexpr int HALF = (BOX - 1) / 2;
  const int job = blockIdx.x;
  int us = regions[job * 4 + 0], vs = regions[job * 4 + 1], uf = regions[job * 4 + 2],
      vf = regions[job * 4 + 3];
  // monoslam.cpp:1083-1090
  if (us < HALF + 1) us = HALF + 1;
  if (uf > d.W - HALF - 1) uf = d.W - HALF - 1;
  if (vs < HALF + 1) vs = HALF + 1;
  if (vf > d.H - HALF - 1) vf = d.H - HALF - 1;
  if (vs >= vf || us >= uf) {  // :1093-1098
    if (threadIdx.x == 0) {
      out_uv[job * 2 + 0] = us;
      out_uv[job * 2 + 1] = vs;
      out_ev[job] = 0.0;
    }
    return;
  }
  const uint8_t *img = d.frames + ((size_t)slot * d.B + stream) * d.H * d.pitch;
  const int RW = uf - us, RH = vf - vs;
  DBest best = {0.0, 0x7fffffff};
  for (int p = threadIdx.x; p < RW * RH; p += blockDim.x) {
    const int v = vs + p / RW, u = us + p % RW;
    int Sxx = 0, Syy = 0, Sxy = 0;
    // rows v-HALF-1 .. v+HALF+1, columns u-HALF-1 .. u+HALF+1; three rows live in registers
    uint8_t prev[BOX + 2], cur[BOX + 2], next[BOX + 2];
    const uint8_t *r0 = img + (size_t)(v - HALF - 1) * d.pitch + (u - HALF - 1);
#pragma unroll
    for (int c = 0; c < BOX + 2; ++c) {
      prev[c] = __ldg(r0 + c);
      cur[c] = __ldg(r0 + d.pitch + c);
    }
#pragma unroll
    for (int r = 0; r < BOX; ++r) {
      const uint8_t *rn = r0 + (size_t)(r + 2) * d.pitch;
#pragma unroll
      for (int c = 0; c < BOX + 2; ++c) next[c] = __ldg(rn + c);
#pragma unroll
      for (int c = 1; c <= BOX; ++c) {
        const int dx = (int)cur[c + 1] - (int)cur[c - 1];
        const int dy = (int)next[c] - (int)prev[c];
        Sxx += dx * dx;
        Syy += dy * dy;
        Sxy += dx * dy;
      }
#pragma unroll
      for (int c = 0; c < BOX + 2; ++c) {
        prev[c] = cur[c];
        cur[c] = next[c];
      }
    }
    // TSgxsq = Sxx/4 etc. (exact); find_eigenvalues(A = TSgxsq, B = TSgxgy, C = TSgysq)
    const double A = mul_((double)Sxx, 0.25), B = mul_((double)Sxy, 0.25), C = mul_((double)Syy, 0.25);
    const double t1 = add_(A, C);
    const double BB = sqrt_(sub_(mul_(t1, t1), mul_(4.0, sub_(mul_(A, C), mul_(B, B)))));
    const double eval2 = div_(sub_(t1, BB), 2.0);
    dconsider(best, eval2, p);
  }
  __shared__ double s_ev[256];
  __shared__ int s_idx[256];
  s_ev[threadIdx.x] = best.ev;
  s_idx[threadIdx.x] = best.idx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      DBest a = {s_ev[threadIdx.x], s_idx[threadIdx.x]};
      dconsider(a, s_ev[threadIdx.x + o], s_idx[threadIdx.x + o]);
      s_ev[threadIdx.x] = a.ev;
      s_idx[threadIdx.x] = a.idx;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (s_idx[0] != 0x7fffffff) {
      out_uv[job * 2 + 0] = us + s_idx[0] % RW;
      out_uv[job * 2 + 1] = vs + s_idx[0] / RW;
    } else {
      out_uv[job * 2 + 0] = -1;  // nothing beat evbest = 0: the reference leaves *ubest/*vbest alone
      out_uv[job * 2 + 1] = -1;
    }
    out_ev[job] = s_ev[0];
  }
}

}  // namespace

cudaError_t sl2_launch_detect(const Sl2Dev &d, int stream, int slot, int n, const int *regions_dev,
                              int *out_uv_dev, double *out_ev_dev, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  switch (d.box) {
    case 11: detect_kernel<11><<<n, 256, 0, st>>>(d, stream, slot, regions_dev, out_uv_dev, out_ev_dev); break;
    case 15: detect_kernel<15><<<n, 256, 0, st>>>(d, stream, slot, regions_dev, out_uv_dev, out_ev_dev); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
