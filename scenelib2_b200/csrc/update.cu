// update.cu — EKF update on sm_100a as a pipeline of four kernels (all streams of a context per launch).
//
// Replaces, per camera stream:
//   Kalman::KalmanFilterUpdate             kalman.cpp:72-119  (+ gather/scatter monoslam.cpp:501-614)
//   MonoSLAM::normalise_state + symmetrise monoslam.cpp:616-637, 143-150
//   attempt / success bookkeeping          monoslam.cpp:479-496
//
// Mathematically the reference's  K = P H^T S^-1,  x += K nu,  P -= K S K^T  with S = H P H^T + R:
//   S = U^T U (Cholesky),  Y = U^-T (H P),  w = U^-T nu   =>   x += Y^T w,   P -= Y^T Y.
// H is structurally sparse (7 + 3 non-zero columns per row) and is never formed.
//
//   kernel           grid                      work per CTA
//   upd_factor       streams                   G = [ S | H P | nu ] (m x (m+n+1), row-major scratch); blocked
//                                              Cholesky of the S part only (16-row panels) -> U in place, and
//                                              W_pp = U_pp^-T of every panel.  The serial chain of the update
//                                              lives here and touches 200 x 200 numbers, not 200 x 514.
//   upd_solve        column slabs x streams    Y = U^-T [H P | nu]: each WARP owns 8 columns and keeps all m rows
//                                              of them in REGISTERS (DMMA B-fragment layout); per panel the
//                                              multipliers U(0:i0, panel) are staged once per CTA (cp.async) and
//                                              the product runs on the FP64 tensor path with B from registers.
//                                              No inter-warp or inter-CTA dependency: column slabs are independent.
//   upd_syrk         64x64 tiles x streams     P -= Y^T Y (upper tiles computed, lower mirrored; sub-tiles of a
//                                              diagonal tile below the diagonal are skipped); the nu column rides
//                                              along as column n of Y, so the tile row that holds it yields
//                                              x += Y^T w in its epilogue.
//   upd_finish       streams                   normalise_state, symmetrise, counters.
//
// Every re-read of the old single-kernel design (finished rows of G gathered from L2/HBM by every panel, Y slabs
// re-staged per tile by a CTA that owns the whole stream) is gone: G is written once and read once by upd_solve
// (registers), Y is written once and read by the tiles of the same stream, which run at the same time on
// neighbouring SMs (L2 hits).  The dense O(n^2 m) parts use ordinary FP64 FMAs / DMMA (tolerance 1e-5 relative,
// north star); nothing in this file decides which pixels are searched.
#include "sl2_common.cuh"

namespace {

struct UpdSmem {
  // upd_chol_global: carved from dynamic shared memory; sizes depend on Nmax
  double *mult;  // [mmax][UPD_MS] (negated) multipliers of the current panel
  double *dg;    // [NB][UPD_DS] diagonal block of the current panel (factor scratch)
  double *Wm;    // [NB][UPD_WS]  W = U_pp^-T of the current panel
  double *pan;   // panel buffer [NB][panw]
  int panw;
};

constexpr int UPD_THREADS = 256;
constexpr int UPD_NB = 16;   // Cholesky row-panel height (two DMMA M-tiles)
constexpr int UPD_WS = 20;   // row stride of the W table
constexpr int UPD_DS = 20;   // row stride of the diagonal-block scratch (conflict-free fragments)
constexpr int UPD_MS = 20;   // row stride of the multiplier table: 32 B (mod 128) => conflict-free A fragments
constexpr int UPD_KC = 32;   // k-chunk of the Y^T Y tiles
constexpr int UPD_YS = 68;   // padded row stride of a staged Y slab (doubles): conflict-free DMMA reads
constexpr int UPD_GB = 4;    // 8-column groups per warp iteration in the panel update
constexpr int SOLVE_MAX_WARPS = 8;   // warps (8-column groups) per upd_solve CTA: 2 per SM sub-partition, <= 255 registers

__host__ __device__ inline int upd_keven(int Nmax) { return (Nmax + 1) & ~1; }
__host__ __device__ inline int upd_panw(int Nmax) {
  // panel buffer of the factor kernel: S columns only; row stride = 2 (mod 16) doubles: the 8 rows of a DMMA
  // C fragment hit distinct banks
  return ((2 * upd_keven(Nmax) + 15) & ~15) + 2;
}
__host__ __device__ inline size_t upd_pan_doubles(int Nmax) { return (size_t)UPD_NB * upd_panw(Nmax); }

__device__ __forceinline__ UpdSmem carve(uint8_t *base, int Nmax) {
  UpdSmem u;
  const int K = upd_keven(Nmax), mmax = 2 * K;  // even counts keep every section 16 B aligned
  double *p = reinterpret_cast<double *>(base);
  u.mult = p;  p += (size_t)mmax * UPD_MS;
  u.dg = p;  p += UPD_NB * UPD_DS;
  u.Wm = p;  p += UPD_NB * UPD_WS;
  u.pan = p;
  u.panw = upd_panw(Nmax);
  return u;
}

// D(8x8) = A(8x4) * B(4x8) + C on the FP64 tensor path: lane holds A(lane/4, lane%4),
// B(lane%4, lane/4) and C(lane/4, 2*(lane%4) + {0,1}).
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// wait until at most n groups are pending; n folds to a constant in unrolled loops
__device__ __forceinline__ void cp_async_wait_n(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    case 7: cp_async_wait<7>(); break;
    case 8: cp_async_wait<8>(); break;
    case 9: cp_async_wait<9>(); break;
    case 10: cp_async_wait<10>(); break;
    case 11: cp_async_wait<11>(); break;
    case 12: cp_async_wait<12>(); break;
    case 13: cp_async_wait<13>(); break;
    case 14: cp_async_wait<14>(); break;
    default: cp_async_wait<15>(); break;
  }
}

// ---- mbarrier + bulk copy (cp.async.bulk: the TMA engine moves a contiguous run of bytes global -> shared and
//      reports completion as transaction bytes on an mbarrier; one instruction per row, no per-chunk index math,
//      and the consumers wait on the mbarrier instead of a CTA-wide barrier) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {
  }
}
// bytes: multiple of 16; dst / src 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}

// 1/sqrt(d) for a positive pivot: MUFU seed + two Newton steps (about 1 ulp); a handful of FP64
// instructions instead of the library routine -- this sits on the serial path of every panel.
__device__ __forceinline__ double pivot_rsqrt(double dv) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(dv));
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-(dv * y), y, 1.0);
    y = fma(0.5 * y, e, y);
  }
  return y;
}

// One warp: Cholesky of the 8x8 block at (o, o) of dg (upper triangle, U^T U = A) and W = U^-T into
// the same block of Wm.  Lane j (mod 8) holds column j in registers; pivots and multipliers travel
// by shuffles.  All 32 lanes must call.
__device__ __forceinline__ void chol8_inv(double *dg, double *Wm, int o, int lane) {
  const int j = lane & 7;
  double a[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (i <= j) ? dg[(o + i) * UPD_DS + o + j] : 0.0;
  double iud = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const double dv = __shfl_sync(0xffffffffu, a[r], r);
    const double iu = pivot_rsqrt(dv);
    const double urj = (j == r) ? dv * iu : a[r] * iu;
    a[r] = urj;
    if (j == r) iud = iu;
#pragma unroll
    for (int i = r + 1; i < 8; ++i) {
      const double uri = __shfl_sync(0xffffffffu, urj, i);
      a[i] -= uri * urj;
    }
  }
  // column j of W = U^-T (lower triangular): U^T W = I by forward substitution
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double iui = __shfl_sync(0xffffffffu, iud, i);
    double sacc = 0.0;
#pragma unroll
    for (int t = 0; t < i; ++t) {
      const double u = __shfl_sync(0xffffffffu, a[t], i);  // U(t, i), t < i
      sacc += u * w[t];
    }
    w[i] = (i == j) ? iui : ((i > j) ? -sacc * iui : 0.0);
  }
  __syncwarp();  // every lane has read its (mirrored) column before the block is overwritten
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i <= j) dg[(o + i) * UPD_DS + o + j] = a[i];
      Wm[(o + i) * UPD_WS + o + j] = w[i];
    }
  }
}

// =============================================================================================
// kernel 0: upd_hp — measurement list, G = [ S | H P | nu ]: one CTA per HP_ROWS measurement rows per stream
// =============================================================================================
// H P: the dense part H_xv (rows x 16, zero padded) * P(0:16, :) runs on DMMA tiles with H_xv k-major in shared
// memory; the 3 structural dh/dy columns of a row are added per element from P(:, pos_i + c) (16-byte loads,
// contiguous along the state index because P is symmetric).  The 13 dense columns of the CTA's rows of H P stay
// in shared memory for S = (H P) H^T + R (upper triangle): one warp per row, lane = measured feature (two columns
// of S), dense part from shared memory, structural part from the row of H P just written (L1/L2).
// The 7 CTAs of a stream (m = 200) share nothing but P; the measurement list is rebuilt by each of them with
// ballots (it is ~100 flag reads).
constexpr int HP_THREADS = 256;
constexpr int HP_ROWS = 32;   // measurement rows per CTA (4 DMMA M tiles)
constexpr int HP_HXS = 14;    // row stride of the CTA's H*P(:, 0:13) table
struct HpSmem {
  double *HxT;   // [16][hms]  H_xv transposed, k-major, zero padded (columns 13..15, rows >= m)
  double *Hy;    // [K][2][3]
  double *Rv;    // [K][3]  (R00, R01, R11)
  double *nu;    // [mmax]
  double *hpx;   // [HP_ROWS][HP_HXS]
  int *mfeat;    // [K]
  int *wcount;   // [8]
  int hms;
};
__host__ __device__ inline int hp_hms(int Nmax) {
  // k-major Hx table: row stride = 4 (mod 16) doubles => the 4 k-rows of a fragment are 32 B apart
  return ((2 * upd_keven(Nmax) + 15) & ~15) + 4;
}
__host__ __device__ inline size_t hp_smem_doubles(int Nmax) {
  const size_t K = upd_keven(Nmax);
  return (size_t)16 * hp_hms(Nmax) + K * 6 + K * 3 + (K & 1) + 2 * K + HP_ROWS * HP_HXS;
}
__device__ __forceinline__ HpSmem hp_carve(uint8_t *base, int Nmax) {
  HpSmem u;
  const int K = upd_keven(Nmax);
  double *p = reinterpret_cast<double *>(base);
  u.hms = hp_hms(Nmax);
  u.HxT = p;  p += (size_t)16 * u.hms;
  u.Hy = p;  p += (size_t)K * 6;
  u.Rv = p;  p += (size_t)K * 3 + (K & 1);
  u.nu = p;  p += 2 * K;
  u.hpx = p;  p += HP_ROWS * HP_HXS;
  u.mfeat = reinterpret_cast<int *>(p);
  u.wcount = u.mfeat + K;
  return u;
}

__global__ void __launch_bounds__(HP_THREADS, 2) upd_hp_kernel(
    const Sl2Dev d, int stream_lo, int staged_m, const int *st_feat, const double *st_Hxv,
    const double *st_Hy, const double *st_R, const double *st_nu) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const HpSmem sm = hp_carve(smem_raw, d.Nmax);
  const int s = stream_lo + blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;  // DMMA fragment coordinates
  const int nf = d.nfeat[s];
  const int n = SL2_NXV + 3 * nf;
  const int ld = d.ld, ldg = d.ldg;
  const double *__restrict__ P = d.P + (size_t)s * ld * ld;
  double *__restrict__ G = d.G + (size_t)s * d.mmax * ldg;
  const size_t fb = (size_t)s * d.Nmax;
  const int HMS = sm.hms;

  // ---- measurement list in selected order, successful only (monoslam.cpp:556-571) --------------
  int K;
  if (staged_m >= 0) {
    K = staged_m / 2;
    for (int k = tid; k < K; k += HP_THREADS) sm.mfeat[k] = st_feat[k];
  } else {
    const int nsel = d.nsel[s];
    int feat = -1;
    if (tid < d.Nmax && tid < nsel) {
      const int i = d.job_feat[fb + tid];
      if (i >= 0 && d.found[fb + i]) feat = i;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, feat >= 0);
    if (lane == 0) sm.wcount[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < (SL2_MAX_FEAT_SMEM + 31) / 32; ++w) {
      if (w < warp) base += sm.wcount[w];
      total += sm.wcount[w];
    }
    if (feat >= 0) sm.mfeat[base + __popc(bal & ((1u << lane) - 1u))] = feat;
    K = total;
  }
  const int m = 2 * K;
  if (blockIdx.x == 0 && tid == 0) {
    d.upd_m[s] = m;
    if (staged_m < 0) d.nmeas[s] = K;
  }
  if (HP_ROWS * (int)blockIdx.x >= m) return;
  for (int e = tid; e < 16 * HMS; e += HP_THREADS) sm.HxT[e] = 0.0;
  __syncthreads();
  // ---- H rows, R, nu of every measurement (flat loops: every CTA of the stream pays this prologue) -------
  if (staged_m >= 0) {
    for (int e = tid; e < m * 13; e += HP_THREADS) {
      const int i = e / 13, c = e - i * 13;
      sm.HxT[c * HMS + i] = st_Hxv[e];
    }
    for (int e = tid; e < K * 6; e += HP_THREADS) sm.Hy[e] = st_Hy[e];
    for (int e = tid; e < m; e += HP_THREADS) sm.nu[e] = st_nu[e];
    for (int k = tid; k < K; k += HP_THREADS) {
      // R_k 2x2 column-major (symmetric; the host entry point rejects R01 != R10)
      sm.Rv[k * 3 + 0] = st_R[k * 4 + 0];
      sm.Rv[k * 3 + 1] = st_R[k * 4 + 2];
      sm.Rv[k * 3 + 2] = st_R[k * 4 + 3];
    }
  } else {
    for (int e = tid; e < K * 14; e += HP_THREADS) {  // dh/dxv = [dh/dxp | 0] (motion_model.cpp:224-235)
      const int k = e / 14, q = e - k * 14, r = q >= 7;
      sm.HxT[(q - 7 * r) * HMS + 2 * k + r] = d.dh_dxp[(fb + sm.mfeat[k]) * 14 + q];
    }
    for (int e = tid; e < K * 6; e += HP_THREADS) {
      const int k = e / 6;
      sm.Hy[e] = d.dh_dy[(fb + sm.mfeat[k]) * 6 + (e - k * 6)];
    }
    for (int e = tid; e < m; e += HP_THREADS) {
      const size_t f = fb + sm.mfeat[e >> 1];
      // nu = z - h (full_feature_model.cpp:197-200), z = (double)(u,v) (monoslam.cpp:382-383)
      sm.nu[e] = (rd((double)d.z_uv[f * 2 + (e & 1)]) - rd(d.h[f * 2 + (e & 1)])).v;
    }
    for (int k = tid; k < K; k += HP_THREADS) {
      const double var = d.Rvar[fb + sm.mfeat[k]];  // R_i = var * I (camera.cpp:294-299)
      sm.Rv[k * 3 + 0] = var;
      sm.Rv[k * 3 + 1] = 0.0;
      sm.Rv[k * 3 + 2] = var;
    }
  }
  __syncthreads();

  // row blocks blockIdx.x, blockIdx.x + gridDim.x, ...: the measurement list and the H tables are built once
  for (int rb = blockIdx.x; HP_ROWS * rb < m; rb += gridDim.x) {
  const int row0 = HP_ROWS * rb, rows = min(HP_ROWS, m - row0);
  // ---- H*P for the CTA's rows -------------------------------------------------------------------------
  {
    constexpr int QB = 4;  // column groups per work item
    const int mt0 = row0 >> 3, mtiles = (rows + 7) >> 3, ngrp = (n + 7) >> 3;
    // work item = (QB column groups, a quarter of the M tiles): n = 313 gives 10 x 4 items = 5 per warp (items of
    // whole column passes left two warps with twice the work of the others: 21 % of the kernel at the barrier)
    const int npass = (ngrp + QB - 1) / QB, nq = min(4, mtiles), mtq = (mtiles + nq - 1) / nq;
    for (int item = warp; item < npass * nq; item += HP_THREADS / 32) {
      const int gq = (item / nq) * QB, mtlo = (item % nq) * mtq, mthi = min(mtiles, mtlo + mtq);
      double b[QB][4];
      int j0[QB];  // first of the two columns of this lane's C elements, per group (-1: none)
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        const int jb = (gq + q) * 8 + lr;  // column of this lane's B element
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          b[q][ks] = (gq + q < ngrp && jb < n) ? P[jb + (size_t)ld * (4 * ks + lc)] : 0.0;
        j0[q] = (gq + q) * 8 + 2 * lc;
        if (gq + q >= ngrp || j0[q] >= n) j0[q] = -1;
      }
      for (int mt = mtlo; mt < mthi; ++mt) {
        const int i = (mt0 + mt) * 8 + lr;
        const bool rv = i < m;
        const int k = rv ? (i >> 1) : 0;
        const int pos = SL2_NXV + 3 * sm.mfeat[k];
        const double *hy = sm.Hy + k * 6 + (i & 1) * 3;
        const double h0 = hy[0], h1 = hy[1], h2 = hy[2];
        // structural dh/dy columns: all loads of the pass first (independent, 16 B each)
        double2 pv[QB][3];
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            pv[q][c] = make_double2(0.0, 0.0);
            if (rv && j0[q] >= 0) {
              const double *src = P + j0[q] + (size_t)ld * (pos + c);
              if (j0[q] + 1 < n) pv[q][c] = *reinterpret_cast<const double2 *>(src);
              else pv[q][c].x = *src;
            }
          }
        double a[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) a[ks] = sm.HxT[(4 * ks + lc) * HMS + (mt0 + mt) * 8 + lr];
#pragma unroll
        for (int q = 0; q < QB; ++q) {
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) dmma884(c0, c1, a[ks], b[q][ks]);
          c0 += h0 * pv[q][0].x;
          c1 += h0 * pv[q][0].y;
          c0 += h1 * pv[q][1].x;
          c1 += h1 * pv[q][1].y;
          c0 += h2 * pv[q][2].x;
          c1 += h2 * pv[q][2].y;
          if (rv && j0[q] >= 0) {
            double *dst = G + (size_t)i * ldg + m + j0[q];
            if (j0[q] + 1 < n) *reinterpret_cast<double2 *>(dst) = make_double2(c0, c1);
            else *dst = c0;
            if (j0[q] < SL2_NXV) {  // dense 13 columns of H*P: kept in shared memory for S
              sm.hpx[(i - row0) * HP_HXS + j0[q]] = c0;
              if (j0[q] + 1 < SL2_NXV) sm.hpx[(i - row0) * HP_HXS + j0[q] + 1] = c1;
            }
          }
        }
      }
    }
  }
  for (int i = tid; i < rows; i += HP_THREADS) G[(size_t)(row0 + i) * ldg + m + n] = sm.nu[row0 + i];
  __syncthreads();
  // ---- S = (H P) H^T + R for the CTA's rows, columns from the row's own feature on --------------------
  {
    constexpr int SCH = 4;  // feature chunks of 32 per pass (covers K <= 128 in one pass)
    for (int il = warp; il < rows; il += HP_THREADS / 32) {
      const int i = row0 + il;
      const double *grow = G + (size_t)i * ldg + m;
      const int k0 = i >> 1;
      for (int kb = k0; kb < K; kb += 32 * SCH) {
        double hp[SCH][3];
#pragma unroll
        for (int t = 0; t < SCH; ++t) {  // all scattered loads of the pass first
          const int k = kb + 32 * t + lane;
          const int pos = SL2_NXV + 3 * sm.mfeat[k < K ? k : 0];
#pragma unroll
          for (int c = 0; c < 3; ++c) hp[t][c] = k < K ? grow[pos + c] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < SCH; ++t) {
          const int k = kb + 32 * t + lane;
          if (kb + 32 * t < K) {  // warp-uniform
            const int kk = k < K ? k : 0;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < 13; ++c) {
              const double hx = sm.hpx[il * HP_HXS + c];
              const double2 hv = *reinterpret_cast<const double2 *>(sm.HxT + c * HMS + 2 * kk);
              s0 += hx * hv.x;
              s1 += hx * hv.y;
            }
            const double *hy = sm.Hy + kk * 6;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              s0 += hp[t][c] * hy[c];
              s1 += hp[t][c] * hy[3 + c];
            }
            if (i == 2 * kk) {
              s0 += sm.Rv[kk * 3 + 0];
              s1 += sm.Rv[kk * 3 + 1];
            }
            if (i == 2 * kk + 1) s1 += sm.Rv[kk * 3 + 2];
            if (k < K) *reinterpret_cast<double2 *>(G + (size_t)i * ldg + 2 * k) = make_double2(s0, s1);
          }
        }
      }
    }
  }
  __syncthreads();  // hpx of this row block is rewritten by the next one
  }  // row blocks
}

// =============================================================================================
// kernel 1: upd_chol — Cholesky of S, finished rows of U in G (L2), two streams per SM
// =============================================================================================
// This kernel is a serial chain: 13 diagonal blocks per stream (m = 200), each factored by ONE warp (~7 k cycles:
// two 8x8 register/shuffle factorizations joined by 8x8 DMMA products), and nothing in the update can start
// before it ends.  What hides it is other streams on the same SM, so it is kept small enough for two CTAs per SM
// (all 296 streams of the benchmark resident at once).  Measured alternative: S / U resident in shared memory
// (193 KB, one CTA per SM, two waves) made every panel 1.7x faster and the kernel 17 % slower (0.159 vs 0.136 ms).
__global__ void __launch_bounds__(UPD_THREADS, 2) upd_chol_kernel(const Sl2Dev d, int stream_lo) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const UpdSmem sm = carve(smem_raw, d.Nmax);
  const int s = stream_lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int ldg = d.ldg;
  double *__restrict__ G = d.G + (size_t)s * d.mmax * ldg;
  double *__restrict__ Wp = d.Wp + (size_t)s * SL2_MAX_PANELS * 256;
  const int warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;  // DMMA fragment coordinates
  __shared__ int s_next;
  const int m = d.upd_m[s];
  if (m == 0) return;

  // ---- phase 2: left-looking Cholesky by row panels of 16 on the S part of G ------------------
  // Trailing update of a panel = C(16 x cols) - A(16 x i0) * B(i0 x cols) with A(r,k) = U(k,i0+r)
  // (multipliers, shared memory) and B = finished rows of U (global / L2): FP64 tensor-core
  // tiles (DMMA m8n8k4, two M tiles per B fragment), each warp owning groups of 8 columns; the B
  // fragments are software-pipelined three k-steps ahead.
  const int width = m;
  const int PW = sm.panw;
  for (int i0 = 0, pidx = 0; i0 < m; i0 += UPD_NB, ++pidx) {
    const int nbp = min(UPD_NB, m - i0);
    if (tid == 0) s_next = 1;  // batch 0 is reserved for warp 0
    // multipliers, negated so that D = (-A) * B + C
    for (int e = tid; e < i0 * UPD_NB; e += UPD_THREADS) {
      const int k = e / UPD_NB, r = e - k * UPD_NB;
      sm.mult[k * UPD_MS + r] = (r < nbp) ? -G[(size_t)k * ldg + i0 + r] : 0.0;
    }
    __syncthreads();
    const int ngroups = (width - i0 + 7) >> 3;
    const int nk = i0 >> 2;  // k-steps of 4 rows; i0 is a multiple of 16 so nk % 4 == 0
    const int nbatch = (ngroups + UPD_GB - 1) / UPD_GB;
    // Batches of UPD_GB column groups are handed out dynamically.  Warp 0 takes batch 0 (it holds
    // the 16 diagonal columns), factors the diagonal block straight away while the other warps
    // keep multiplying, and only then joins the pool again.
    bool first = true;
    for (;;) {
      int bt;
      if (warp == 0 && first) {
        bt = 0;
      } else {
        if (lane == 0) bt = atomicAdd(&s_next, 1);
        bt = __shfl_sync(0xffffffffu, bt, 0);
      }
      if (bt >= nbatch) break;
      const int g0 = bt * UPD_GB;
      double c[UPD_GB][2][2];
      int colb[UPD_GB];  // column of the B fragment element of this lane (-1: none)
#pragma unroll
      for (int q = 0; q < UPD_GB; ++q) {
        const int cbase = i0 + (g0 + q) * 8;
        colb[q] = (cbase + lr < width && g0 + q < ngroups) ? cbase + lr : -1;
        const int cc = cbase + 2 * lc;  // C fragment: rows lr / lr+8, columns cc, cc+1
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r = mt * 8 + lr;
          const bool rv = r < nbp && (g0 + q) < ngroups;
          c[q][mt][0] = (rv && cc < width) ? G[(size_t)(i0 + r) * ldg + cc] : 0.0;
          c[q][mt][1] = (rv && cc + 1 < width) ? G[(size_t)(i0 + r) * ldg + cc + 1] : 0.0;
        }
      }
      double b[4][UPD_GB];
      auto loadb = [&](int step, double *dst) {
        const double *gk = G + (size_t)(4 * step + lc) * ldg;
#pragma unroll
        for (int q = 0; q < UPD_GB; ++q) dst[q] = colb[q] >= 0 ? gk[colb[q]] : 0.0;
      };
      if (nk > 0) {
        loadb(0, b[0]);
        loadb(1, b[1]);
        loadb(2, b[2]);
      }
      for (int kb = 0; kb < nk; kb += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int st = kb + u;
          if (st + 3 < nk) loadb(st + 3, b[(u + 3) & 3]);
          const double a0 = sm.mult[(4 * st + lc) * UPD_MS + lr];
          const double a1 = sm.mult[(4 * st + lc) * UPD_MS + 8 + lr];
#pragma unroll
          for (int q = 0; q < UPD_GB; ++q) {
            dmma884(c[q][0][0], c[q][0][1], a0, b[u][q]);
            dmma884(c[q][1][0], c[q][1][1], a1, b[u][q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < UPD_GB; ++q) {
        if (g0 + q < ngroups) {
          const int pc = (g0 + q) * 8 + 2 * lc;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
            *reinterpret_cast<double2 *>(sm.pan + (size_t)(mt * 8 + lr) * PW + pc) =
                make_double2(c[q][mt][0], c[q][mt][1]);
        }
      }
      if (warp == 0 && first) {
        first = false;
        __syncwarp();
        // Factor the 16x16 diagonal block and form W = U_pp^-T (the panel is then finished with one
        // more DMMA product U_panel = W * C_panel).  This is the serial path of the panel, so it is
        // kept short: two 8x8 register/shuffle factorizations (chol8_inv) and 8x8 DMMA products
        //   U12 = W11 A12,  A22 -= U12^T U12,  W21 = -W22 (U12^T W11)
        // on a private copy of the block (identity padding for the ragged last panel).
        {
          double *dg = sm.dg;
          for (int e = lane; e < UPD_NB * UPD_NB; e += 32) {
            const int i = e >> 4, j = e & 15;
            dg[i * UPD_DS + j] = (i < nbp && j < nbp) ? sm.pan[(size_t)i * PW + j] : (i == j ? 1.0 : 0.0);
          }
          for (int e = lane; e < UPD_NB * UPD_WS; e += 32) sm.Wm[e] = 0.0;
          __syncwarp();
          chol8_inv(dg, sm.Wm, 0, lane);
          __syncwarp();
          {  // U12 = W11 * A12
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              dmma884(c0, c1, sm.Wm[lr * UPD_WS + 4 * ks + lc], dg[(4 * ks + lc) * UPD_DS + 8 + lr]);
            __syncwarp();
            *reinterpret_cast<double2 *>(dg + lr * UPD_DS + 8 + 2 * lc) = make_double2(c0, c1);
          }
          __syncwarp();
          {  // A22 -= U12^T U12   (A(i,k) = U12(k,i) and B(k,n) = U12(k,n): the same fragment)
            double2 cv = *reinterpret_cast<const double2 *>(dg + (8 + lr) * UPD_DS + 8 + 2 * lc);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const double v = dg[(4 * ks + lc) * UPD_DS + 8 + lr];
              dmma884(cv.x, cv.y, -v, v);
            }
            *reinterpret_cast<double2 *>(dg + (8 + lr) * UPD_DS + 8 + 2 * lc) = cv;
          }
          __syncwarp();
          chol8_inv(dg, sm.Wm, 8, lane);
          __syncwarp();
          {  // T = U12^T W11 (parked in the unused lower-left block of dg), W21 = -W22 T
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              dmma884(t0, t1, dg[(4 * ks + lc) * UPD_DS + 8 + lr], sm.Wm[(4 * ks + lc) * UPD_WS + lr]);
            *reinterpret_cast<double2 *>(dg + (8 + lr) * UPD_DS + 2 * lc) = make_double2(t0, t1);
            __syncwarp();
            double w0 = 0.0, w1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              dmma884(w0, w1, -sm.Wm[(8 + lr) * UPD_WS + 8 + 4 * ks + lc], dg[(8 + 4 * ks + lc) * UPD_DS + lr]);
            *reinterpret_cast<double2 *>(sm.Wm + (8 + lr) * UPD_WS + 2 * lc) = make_double2(w0, w1);
          }
          __syncwarp();
          // U back into the panel (upper triangle); rows / columns of the padding carry no W
          for (int e = lane; e < UPD_NB * UPD_NB; e += 32) {
            const int i = e >> 4, j = e & 15;
            if (i <= j && j < nbp) sm.pan[(size_t)i * PW + j] = dg[i * UPD_DS + j];
            if (i >= nbp || j >= nbp) sm.Wm[i * UPD_WS + j] = 0.0;
          }
        }
      }
    }
    __syncthreads();
    // finish the panel: rows of U for the 16 diagonal columns, U_panel = W * C_panel (DMMA) for
    // all other columns, written straight to G from the C fragments; W_pp goes to the solve kernel
    for (int e = tid; e < UPD_NB * UPD_NB; e += UPD_THREADS) {
      const int r = e / UPD_NB, cc = e - r * UPD_NB;
      if (r < nbp && cc < nbp && i0 + cc < width)
        G[(size_t)(i0 + r) * ldg + i0 + cc] = (cc >= r) ? sm.pan[(size_t)r * PW + cc] : 0.0;
      Wp[(size_t)pidx * 256 + e] = sm.Wm[r * UPD_WS + cc];
    }
    {
      double aw[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) aw[mt][ks] = sm.Wm[(mt * 8 + lr) * UPD_WS + 4 * ks + lc];
      const int ncol = width - i0;
      // FG column groups per iteration: independent DMMA chains.  Only full panels reach this loop
      // with columns to do (a ragged last panel has ncol == nbp: nothing right of the diagonal block).
      constexpr int FG = 4;
      for (int gb = 2 + warp; gb * 8 < ncol; gb += FG * (UPD_THREADS / 32)) {
        double c[FG][2][2];
#pragma unroll
        for (int f = 0; f < FG; ++f) c[f][0][0] = c[f][0][1] = c[f][1][0] = c[f][1][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          double bv[FG];
#pragma unroll
          for (int f = 0; f < FG; ++f) {
            const int cb = (gb + f * (UPD_THREADS / 32)) * 8 + lr;
            bv[f] = cb < ncol ? sm.pan[(size_t)(4 * ks + lc) * PW + cb] : 0.0;
          }
#pragma unroll
          for (int f = 0; f < FG; ++f) {
            dmma884(c[f][0][0], c[f][0][1], aw[0][ks], bv[f]);
            dmma884(c[f][1][0], c[f][1][1], aw[1][ks], bv[f]);
          }
        }
#pragma unroll
        for (int f = 0; f < FG; ++f) {
          const int cc = (gb + f * (UPD_THREADS / 32)) * 8 + 2 * lc;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const int r = mt * 8 + lr;
            if (r < nbp && cc < ncol) {
              double *dst = G + (size_t)(i0 + r) * ldg + i0 + cc;
              if (cc + 1 < ncol) *reinterpret_cast<double2 *>(dst) = make_double2(c[f][mt][0], c[f][mt][1]);
              else *dst = c[f][mt][0];
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// =============================================================================================
// kernel 2: upd_solve — Y = U^-T [H P | nu], 8 columns per warp held in registers
// =============================================================================================
// C fragments of a 16-row block (two M tiles: c[mt][e] = X(8 mt + lane/4, 2 (lane%4) + e)) -> B fragment of
// its k-step ks (rows 4 ks .. 4 ks + 3): lane (lr, lc) receives X(4 ks + lc, lr).
__device__ __forceinline__ double c_to_b(const double (&c)[2][2], int ks, int lane) {
  const int lr = lane >> 2, lc = lane & 3;
  const int src = ((4 * (ks & 1) + lc) << 2) | (lr >> 1);
  const double v0 = __shfl_sync(0xffffffffu, c[ks >> 1][0], src);
  const double v1 = __shfl_sync(0xffffffffu, c[ks >> 1][1], src);
  return (lr & 1) ? v1 : v0;
}

// NP = number of 16-row panels the instantiation covers (m <= 16 NP).  Column c of the slab space is
// column m + c of G: c < n is H P, c == n is nu.
//
// Right-looking, everything in registers: the warp holds all rows of its 8 columns as DMMA ACCUMULATORS
// (C layout: tile j = rows 8j..8j+7).  Per panel p: Y_p = W_pp C_p (8 DMMAs, C_p moved to the B layout by
// shuffles), the finished rows go straight to G as 16-byte stores, and every later row tile j gets
// acc_j -= U(panel, tile j)^T Y_p: 4 DMMAs per tile (k-step outer, tile inner: consecutive DMMAs never share an
// accumulator), A from the panel's rows of U in shared memory, B = Y_p from registers.
// Staging: ALL panels of U (the part right of the diagonal blocks, <= 166 KB at m = 208) and all W_pp are
// requested up front by warp 0 as bulk copies (one instruction per 16-row x row-segment / per W row, completion
// counted in bytes on one mbarrier per panel), so the only latency the kernel ever waits for is the first
// panel's, nobody computes a staging address, and there is NO CTA-wide barrier in the panel loop: a warp waits
// on the mbarrier of the panel it needs and otherwise runs at its own pace.  (Measured before: a 2-deep cp.async
// ring left the late, small panels bound by the L2 round trip of their own staging, and staging everything with
// per-thread cp.async cost 22 % of the kernel in index arithmetic.)  NP = 16 does not fit one SM that way: its
// panels >= SPLIT are requested at panel REUSE into the space of panels 0..REUSE-1 (one CTA barrier).
template <int NP> struct SolveLayout {
  static constexpr int SPLIT = NP > 13 ? 6 : NP;   // panels >= SPLIT are staged late (second generation)
  static constexpr int REUSE = NP > 13 ? 4 : NP;   // ... once panels < REUSE have been consumed
  __host__ __device__ static constexpr int pw(int p) { return 16 * (NP - 1 - p) + 4; }  // row stride (= 4 mod 16)
  __host__ __device__ static constexpr int off_lin(int p0, int p) {  // doubles before panel p when packing starts at p0
    int o = 0;
    for (int q = p0; q < p; ++q) o += 16 * pw(q);
    return o;
  }
  __host__ __device__ static constexpr int off(int p) { return p < SPLIT ? off_lin(0, p) : off_lin(SPLIT, p); }
  static constexpr int PAN_DOUBLES = off_lin(0, SPLIT < NP - 1 ? SPLIT : NP - 1);
  static constexpr int SMEM_DOUBLES = PAN_DOUBLES + NP * 16 * UPD_WS + NP + (NP & 1) +  // one mbarrier per panel
                                      2 * SOLVE_MAX_WARPS * 2 * 32 * 2;     // C-tile exchange of the warp pairs
  static_assert(SPLIT == NP || off_lin(SPLIT, NP - 1) <= off_lin(0, REUSE), "second generation must fit");
};

// Two warps share a group of 8 columns: warp (g, rho) owns the row tiles j = 2t + rho, so a thread keeps NP
// tiles (not 2 NP) and 16 warps fit one SM -- 4 per scheduler instead of 2, which is what hides the shared-memory
// latency in front of every DMMA (measured: 8 warps x 228 registers ran the FP64 pipe at 40 %).  Per panel the
// two warps swap their C tile through shared memory (one named barrier of 64 threads), both form Y_p = W_pp C_p
// (8 DMMAs, 4 of them redundant), each stores and keeps its own M tile.
template <int NP>
__global__ void __launch_bounds__(64 * SOLVE_MAX_WARPS, 1) upd_solve_kernel(const Sl2Dev d, int stream_lo) {
  using L = SolveLayout<NP>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  double *pan = reinterpret_cast<double *>(smem_raw);  // panel p at L::off(p): [16][L::pw(p)]  U(16p + r, 16p + 16 + c)
  double *Wm = pan + L::PAN_DOUBLES;                   // [NP][16][UPD_WS]
  const uint32_t bars = smem_u32(Wm + NP * 16 * UPD_WS);  // [NP] mbarriers
  double2 *xbuf = reinterpret_cast<double2 *>(Wm + NP * 16 * UPD_WS + NP + (NP & 1));  // [2][groups][2][32]
  const int s = stream_lo + blockIdx.y;
  const int m = d.upd_m[s];
  if (m == 0) return;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ngrp = nthr >> 6;  // column groups of this CTA
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int g = warp % ngrp, rho = warp / ngrp;
  const int n = SL2_NXV + 3 * d.nfeat[s];
  const int ncols = n + 1;
  if (blockIdx.x * ngrp * 8 >= ncols) return;  // whole CTA beyond the last column
  const int ldg = d.ldg;
  double *__restrict__ G = d.G + (size_t)s * d.mmax * ldg;
  const double *__restrict__ Wp = d.Wp + (size_t)s * SL2_MAX_PANELS * 256;
  const int m8 = (m + 7) & ~7;

  if (tid == 0) {
    for (int p = 0; p < NP; ++p) mbar_init(bars + 8 * p, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // the columns m .. m8-1 of a staged panel are read (rows of the last tile that do not exist) but not copied
  auto zero_pads = [&](int plo, int phi) {
    for (int e = tid; e < (phi - plo) * 16; e += nthr) {
      const int p = plo + (e >> 4), r = e & 15;
      const int cfirst = 16 * p + 16;
      if (16 * p < m)
        for (int c = max(m, cfirst); c < m8; ++c) pan[L::off(p) + r * L::pw(p) + (c - cfirst)] = 0.0;
    }
  };
  zero_pads(0, L::SPLIT);
  __syncthreads();
  // warp 0: lane r < 16 requests row r of U(panel p) right of the diagonal block, lane 16 + r row r of W_pp
  auto request = [&](int p) {
    if (16 * p < m) {
      const int cfirst = 16 * p + 16;
      const int nrow = min(16, m - 16 * p);
      const uint32_t ubytes = m > cfirst ? (uint32_t)(m - cfirst) * 8u : 0u;
      const uint32_t bar = bars + 8 * p;
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)nrow * ubytes + 16u * 128u);
      __syncwarp();
      if (lane < 16) {
        if (lane < nrow && ubytes)
          bulk_g2s(pan + L::off(p) + lane * L::pw(p), G + (size_t)(16 * p + lane) * ldg + cfirst, ubytes, bar);
      } else {
        bulk_g2s(Wm + ((size_t)p * 16 + (lane - 16)) * UPD_WS, Wp + (size_t)p * 256 + (lane - 16) * 16, 128u, bar);
      }
    }
  };
  if (warp == 0) {
#pragma unroll 1
    for (int p = 0; p < L::SPLIT; ++p) request(p);
  }

  // The warp pair walks the column groups g, g + gridDim.x * ngrp, ...: with one CTA per stream (the batched
  // launch) U is staged ONCE for all of the stream's columns and the pairs drift apart, so one pair's reload of
  // its accumulators hides behind the other pairs' DMMAs; nothing below synchronises the CTA (NP <= 13).
  int xpar = 0;
  for (int grp = blockIdx.x * ngrp + g; L::SPLIT < NP ? grp == (int)(blockIdx.x * ngrp + g) : 8 * grp < ncols;
       grp += gridDim.x * ngrp) {
  const int c0 = 8 * grp;        // first column of this warp pair
  const bool wact = c0 < ncols;  // warp pairs past the last column have nothing to do
  const int cc = c0 + 2 * lc;    // this lane's two columns (C layout)
  const int cval = wact ? min(2, ncols - cc) : 0;  // how many of them exist (<= 0: none)
  double *gcol = G + m + cc;
  double acc[NP][2];  // tile j = 2 t + rho: rows 8 j + lr, columns cc, cc + 1
#pragma unroll
  for (int t = 0; t < NP; ++t) {
    const int row = 8 * (2 * t + rho) + lr;
    acc[t][0] = acc[t][1] = 0.0;
    if (row < m) {
      if (cval >= 2) {
        const double2 v = *reinterpret_cast<const double2 *>(gcol + (size_t)row * ldg);
        acc[t][0] = v.x;
        acc[t][1] = v.y;
      } else if (cval == 1) {
        acc[t][0] = gcol[(size_t)row * ldg];
      }
    }
  }

#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (16 * p < m) {  // uniform over the CTA
      if (L::SPLIT < NP && p == L::REUSE) {
        __syncthreads();  // every warp is done with panels < REUSE: their space takes the second generation
        zero_pads(L::SPLIT, NP);
        __syncthreads();
        if (warp == 0) {
          fence_proxy_async();  // generic-proxy reads of the old panels are ordered before the bulk writes
#pragma unroll 1
          for (int q = L::SPLIT; q < NP; ++q) request(q);
        }
      }
      if (!wact) continue;
      // swap the C tiles of panel p with the partner warp
      xpar ^= 1;  // alternates over every panel this pair executes (also across column groups)
      double2 *xb = xbuf + ((size_t)(xpar * ngrp + g) * 2) * 32;
      xb[rho * 32 + lane] = make_double2(acc[p][0], acc[p][1]);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
      const double2 other = xb[(rho ^ 1) * 32 + lane];
      const double cp2[2][2] = {{rho ? other.x : acc[p][0], rho ? other.y : acc[p][1]},
                                {rho ? acc[p][0] : other.x, rho ? acc[p][1] : other.y}};
      mbar_wait(bars + 8 * p, 0);
      const double *pb = pan + L::off(p);
      const int PW = L::pw(p);
      const double *wb = Wm + (size_t)p * 16 * UPD_WS;
      // Y_p = W_pp * C_p
      double cb[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) cb[ks] = c_to_b(cp2, ks, lane);
      double dd[2][2];
      dd[0][0] = dd[0][1] = dd[1][0] = dd[1][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        dmma884(dd[0][0], dd[0][1], wb[lr * UPD_WS + 4 * ks + lc], cb[ks]);
        dmma884(dd[1][0], dd[1][1], wb[(8 + lr) * UPD_WS + 4 * ks + lc], cb[ks]);
      }
      // this warp's finished M tile: 16-byte stores from the C fragments
      {
        const int row = 16 * p + 8 * rho + lr;
        const double v0 = rho ? dd[1][0] : dd[0][0], v1 = rho ? dd[1][1] : dd[0][1];
        if (row < m) {
          if (cval >= 2) *reinterpret_cast<double2 *>(gcol + (size_t)row * ldg) = make_double2(v0, v1);
          else if (cval == 1) gcol[(size_t)row * ldg] = v0;
        }
      }
      // this warp's later row tiles: acc_j -= U(panel, tile j)^T Y_p
      double yb[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) yb[ks] = -c_to_b(dd, ks, lane);
      const double *pbr = pb + 8 * rho + lr;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int t = p + 1; t < NP; ++t) {
          if (8 * (2 * t + rho) < m)  // warp-uniform; a tile past m is never touched (its columns are not staged)
            dmma884(acc[t][0], acc[t][1], pbr[(4 * ks + lc) * PW + 16 * (t - p - 1)], yb[ks]);
        }
      }
    }
  }
  }  // column groups of this warp pair
}

// =============================================================================================
// kernel 3: upd_syrk — P -= Y^T Y on 64x64 tiles, x += Y^T w from the column that carries nu
// =============================================================================================
// One 64x64 tile T = sum_{k < kr} A(k, :)^T B(k, :), A / B = 64-column slabs of the row-major matrix Gm (row
// stride ldg) starting at columns colA / colB.  FP64 DMMA tiles; the slabs are staged by cp.async (LDGSTS)
// into a double-buffered, conflict-free (stride UPD_YS) shared tile.  Warp w owns rows 16*(w%4).. and columns
// 32*(w/4).. of the tile: acc[i][j][e] = T(16*(w%4) + 8*i + lane/4, 32*(w/4) + 8*j + 2*(lane%4) + e).
// Columns >= lim and rows >= kr are zero-filled.  Warps with `skip` stage and synchronise but issue no DMMA.
__device__ __forceinline__ void tile_product(double *stage_buf, const double *__restrict__ Gm, int ldg, int kr,
                                             int colA, int colB, int lim, bool skip, double (&acc)[2][4][2]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int wa = (warp & 3) * 16, wb = (warp >> 2) * 32;
  const int nchunk = (kr + UPD_KC - 1) / UPD_KC;
  const int cA = colA + 2 * lane, cB = colB + 2 * lane;
  const int bytesA = cA + 1 < lim ? 16 : (cA < lim ? 8 : 0);
  const int bytesB = cB + 1 < lim ? 16 : (cB < lim ? 8 : 0);
  const double *srcA = Gm + (bytesA ? cA : 0);
  const double *srcB = Gm + (bytesB ? cB : 0);
  // stage loader: 2 slabs x KC rows x 64 columns; thread = (row warp + 8*j, 16-byte segment `lane`)
  auto stage = [&](int chunk, int buf) {
    double *dst = stage_buf + (size_t)buf * (2 * UPD_KC * UPD_YS) + 2 * lane;
#pragma unroll
    for (int j = 0; j < UPD_KC / 8; ++j) {
      const int kk = warp + 8 * j;
      const int k = chunk * UPD_KC + kk;
      const bool kv = k < kr;
      const size_t ro = (size_t)(kv ? k : 0) * ldg;
      cp_async16(dst + kk * UPD_YS, srcA + ro, kv ? bytesA : 0);
      cp_async16(dst + UPD_KC * UPD_YS + kk * UPD_YS, srcB + ro, kv ? bytesB : 0);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  stage(0, 0);
  for (int ch = 0; ch < nchunk; ++ch) {
    if (ch + 1 < nchunk) {
      stage(ch + 1, (ch + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (!skip) {
      const double *Ya = stage_buf + (size_t)(ch & 1) * (2 * UPD_KC * UPD_YS);
      const double *Yb = Ya + UPD_KC * UPD_YS;
      auto kstep = [&](int kk) {
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = Ya[(kk + lc) * UPD_YS + wa + i * 8 + lr];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Yb[(kk + lc) * UPD_YS + wb + j * 8 + lr];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      };
      const int krem = kr - ch * UPD_KC;  // rows of this chunk that exist (the rest is zero fill)
      if (krem >= UPD_KC) {
#pragma unroll
        for (int kk = 0; kk < UPD_KC; kk += 4) kstep(kk);
      } else {
#pragma unroll 2
        for (int kk = 0; kk < krem; kk += 4) kstep(kk);
      }
    }
    __syncthreads();  // buffer (ch & 1) may be refilled by the stage issued in the next iteration
  }
}

__global__ void __launch_bounds__(UPD_THREADS, 3) upd_syrk_kernel(const Sl2Dev d, int stream_lo) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  double *stage_buf = reinterpret_cast<double *>(smem_raw);
  const int s = stream_lo + blockIdx.y;
  const int m = d.upd_m[s];
  if (m == 0) return;
  const int n = SL2_NXV + 3 * d.nfeat[s];
  // tiles in (tb outer, ta <= tb inner) order
  int tb = 0;
  const int t = blockIdx.x;
  while ((tb + 1) * (tb + 2) / 2 <= t) ++tb;
  const int ta = t - tb * (tb + 1) / 2;
  if (tb * 64 >= n + 1) return;  // this stream's map is smaller than the capacity the grid was sized for
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int wa = (warp & 3) * 16, wb = (warp >> 2) * 32;
  const int ld = d.ld, ldg = d.ldg;
  double *__restrict__ P = d.P + (size_t)s * ld * ld;
  double *__restrict__ x = d.x + (size_t)s * ld;
  const double *__restrict__ G = d.G + (size_t)s * d.mmax * ldg;
  const bool diag = ta == tb;
  const bool skip = diag && wa >= wb + 32;      // sub-tile strictly below the diagonal: mirrored instead
  const bool mirror = !diag || wa + 16 <= wb;   // sub-tile strictly above the diagonal
  double acc[2][4][2];
  tile_product(stage_buf, G, ldg, m, m + ta * 64, m + tb * 64, m + n + 1, skip, acc);
  if (skip) return;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int a = ta * 64 + wa + i * 8 + lr;
    double pold[4][2];  // the loads of one row group first (independent), then the subtraction
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int bq = tb * 64 + wb + j * 8 + 2 * lc + e;
        pold[j][e] = (a < n && bq < n) ? P[a + (size_t)ld * bq] : 0.0;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int bq = tb * 64 + wb + j * 8 + 2 * lc;
      const double v0 = pold[j][0] - acc[i][j][0], v1 = pold[j][1] - acc[i][j][1];
      if (a < n && bq < n) {
        P[a + (size_t)ld * bq] = v0;
        if (bq + 1 < n) P[a + (size_t)ld * (bq + 1)] = v1;
        if (mirror) {  // lower counterpart: rows = b range (contiguous in P), column a
          double *dst = P + bq + (size_t)ld * a;
          if (bq + 1 < n) *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
          else *dst = v0;
        }
      }
      // column n of Y is w = U^-T nu: (Y^T Y)(a, n) = (Y^T w)(a)  =>  x += Y^T w   (kalman.cpp:112)
      if (a < n) {
        if (bq == n) x[a] += acc[i][j][0];
        if (bq + 1 == n) x[a] += acc[i][j][1];
      }
    }
  }
}

// =============================================================================================
// kernel 4: upd_finish — normalise_state, symmetrise, counters
// =============================================================================================
__global__ void __launch_bounds__(UPD_THREADS) upd_finish_kernel(const Sl2Dev d, int stream_lo, int staged,
                                                                  int only_normalise) {
  const int s = stream_lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int nf = d.nfeat[s];
  const int n = SL2_NXV + 3 * nf;
  const int ld = d.ld;
  double *__restrict__ P = d.P + (size_t)s * ld * ld;
  const double *__restrict__ x = d.x + (size_t)s * ld;
  const size_t fb = (size_t)s * d.Nmax;
  const int m = only_normalise ? 0 : d.upd_m[s];
  __shared__ int s_cull;
  // ---- normalise_state (monoslam.cpp:616-637): P <- J P J^T, J = diag(I3, dqnorm, I6, I) ---------
  if (m > 0 || only_normalise) {
    __shared__ double J4[16];
    if (tid == 0) {
      const rd q[4] = {rd(x[3]), rd(x[4]), rd(x[5]), rd(x[6])};
      const rd qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)  // motion_model.cpp:371-380 (quirk Q2)
          J4[i * 4 + j] = (i == j) ? ((rd(1.0) - q[i] * q[i] / (qq * qq)) / qq).v
                                   : ((-q[i]) * q[j] / (qq * qq * qq)).v;
    }
    __syncthreads();
    // rows 3..6 of every column: P(3:7, j) = J4 * P(3:7, j)
    for (int j = tid; j < n; j += UPD_THREADS) {
      double v[4], o[4];
      for (int k = 0; k < 4; ++k) v[k] = P[(3 + k) + (size_t)ld * j];
      for (int i = 0; i < 4; ++i) {
        rd a(0.0);
        for (int k = 0; k < 4; ++k) a = a + rd(J4[i * 4 + k]) * rd(v[k]);
        o[i] = a.v;
      }
      for (int k = 0; k < 4; ++k) P[(3 + k) + (size_t)ld * j] = o[k];
    }
    __syncthreads();
    // columns 3..6: Pxx part gets (J Pxx) J^T; rows >= 13 are the mirror of the updated Pxy
    for (int i = tid; i < n; i += UPD_THREADS) {
      if (i < SL2_NXV) {
        double v[4], o[4];
        for (int k = 0; k < 4; ++k) v[k] = P[i + (size_t)ld * (3 + k)];
        for (int c = 0; c < 4; ++c) {
          rd a(0.0);
          for (int k = 0; k < 4; ++k) a = a + rd(v[k]) * rd(J4[c * 4 + k]);
          o[c] = a.v;
        }
        for (int k = 0; k < 4; ++k) P[i + (size_t)ld * (3 + k)] = o[k];
      } else {
        for (int k = 0; k < 4; ++k) P[i + (size_t)ld * (3 + k)] = P[(3 + k) + (size_t)ld * i];
      }
    }
    __syncthreads();
  }
  // ---- symmetrise (monoslam.cpp:143-150): only the Pxx block can be asymmetric here ---------
  {
    const int i = tid % 13, j = (tid / 13) % 13;
    const double a = P[i + (size_t)ld * j], b = P[j + (size_t)ld * i];
    const double v = (rd(a) * rd(0.5) + rd(b) * rd(0.5)).v;
    __syncthreads();
    if (tid < 169) P[i + (size_t)ld * j] = v;
    __syncthreads();
  }
  // ---- bookkeeping: attempt / success counters (monoslam.cpp:479-496) ------------------------
  if (!staged && !only_normalise) {
    if (tid == 0) s_cull = 0;  // number of features delete_bad_features would cull
    __syncthreads();
    for (int i = tid; i < nf; i += UPD_THREADS) {
      int att = d.attempted[fb + i], suc = d.successful[fb + i];
      if (d.sel_rank[fb + i] >= 0) {
        att += 1;
        if (d.found[fb + i]) suc += 1;
        d.attempted[fb + i] = att;
        d.successful[fb + i] = suc;
      }
      // monoslam.cpp:650-653; lets the cull kernel of the fused step return at once when idle
      if (att >= d.min_attempts && (double)suc / (double)att < d.match_fraction) atomicAdd(&s_cull, 1);
    }
    __syncthreads();
    if (tid == 0) d.ncull[s] = s_cull;
  }
}

// ---- host-side shapes ---------------------------------------------------------------------------
inline int solve_panels(int Nmax) { return (2 * upd_keven(Nmax) + 15) / 16; }
inline int solve_np(int Nmax) {  // instantiation that covers the capacity
  const int p = solve_panels(Nmax);
  return p <= 4 ? 4 : (p <= 7 ? 7 : (p <= 10 ? 10 : (p <= 13 ? 13 : 16)));
}
inline size_t solve_smem(int np) {
  switch (np) {
    case 4: return SolveLayout<4>::SMEM_DOUBLES * sizeof(double);
    case 7: return SolveLayout<7>::SMEM_DOUBLES * sizeof(double);
    case 10: return SolveLayout<10>::SMEM_DOUBLES * sizeof(double);
    case 13: return SolveLayout<13>::SMEM_DOUBLES * sizeof(double);
    default: return SolveLayout<16>::SMEM_DOUBLES * sizeof(double);
  }
}
inline void solve_shape(int Nmax, int &nslab, int &warps) {
  const int ngroups = (SL2_NXV + 3 * Nmax + 1 + 7) / 8;
  nslab = (ngroups + SOLVE_MAX_WARPS - 1) / SOLVE_MAX_WARPS;
  warps = (ngroups + nslab - 1) / nslab;
}
constexpr size_t SYRK_SMEM = (size_t)2 * 2 * UPD_KC * UPD_YS * sizeof(double);

}  // namespace

size_t sl2_update_smem_bytes(const Sl2Dev &d) {  // upd_chol
  const size_t K = upd_keven(d.Nmax), mmax = 2 * K;
  return (mmax * UPD_MS + UPD_NB * UPD_DS + UPD_NB * UPD_WS + upd_pan_doubles(d.Nmax)) * sizeof(double);
}

static size_t hp_smem_bytes(const Sl2Dev &d) {
  return hp_smem_doubles(d.Nmax) * sizeof(double) + (upd_keven(d.Nmax) + 8) * sizeof(int);
}

cudaError_t sl2_configure_update(const Sl2Dev &d) {
  cudaError_t e = cudaFuncSetAttribute(upd_hp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)hp_smem_bytes(d));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upd_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sl2_update_smem_bytes(d));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upd_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  if (e != cudaSuccess) return e;
  const int np = solve_np(d.Nmax);
  const int smem = (int)solve_smem(np);
  switch (np) {
    case 4: return cudaFuncSetAttribute(upd_solve_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    case 7: return cudaFuncSetAttribute(upd_solve_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    case 10: return cudaFuncSetAttribute(upd_solve_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    case 13: return cudaFuncSetAttribute(upd_solve_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    default: return cudaFuncSetAttribute(upd_solve_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  }
}

// ev6 (optional): 6 events recorded around the 5 kernels (hp, chol, solve, syrk, finish)
cudaError_t sl2_launch_update(const Sl2Dev &d, int stream_lo, int stream_cnt, int staged_m,
                              const int *st_feat, const double *st_Hxv, const double *st_Hy,
                              const double *st_R, const double *st_nu, int only_normalise,
                              cudaStream_t st, cudaEvent_t *ev6, int *launches) {
  if (stream_cnt <= 0) return cudaSuccess;
  cudaError_t e;
  int nl = 0;
  auto mark = [&](int i) { return ev6 ? cudaEventRecord(ev6[i], st) : cudaSuccess; };
  if ((e = mark(0)) != cudaSuccess) return e;
  // row blocks of H P / S per stream: spread over CTAs unless the batch already fills the GPU (measured at 296
  // streams: 1 CTA per stream 0.156 ms, 2: 0.167, 7: 0.200 -- every CTA rebuilds the measurement list and H tables)
  const int hp_all = (2 * upd_keven(d.Nmax) + HP_ROWS - 1) / HP_ROWS;
  const int hp_blocks = stream_cnt >= 2 * 148 ? 1 : hp_all;
  if (!only_normalise) {
    upd_hp_kernel<<<dim3(hp_blocks, stream_cnt), HP_THREADS, hp_smem_bytes(d), st>>>(d, stream_lo, staged_m, st_feat,
                                                                                st_Hxv, st_Hy, st_R, st_nu);
    ++nl;
  }
  if ((e = mark(1)) != cudaSuccess) return e;
  if (!only_normalise) {
    upd_chol_kernel<<<stream_cnt, UPD_THREADS, sl2_update_smem_bytes(d), st>>>(d, stream_lo);
    ++nl;
  }
  if ((e = mark(2)) != cudaSuccess) return e;
  if (!only_normalise) {
    int nslab, warps;
    solve_shape(d.Nmax, nslab, warps);
    const int np = solve_np(d.Nmax);
    const size_t smem = solve_smem(np);
    // two warps per 8-column group; a batch that fills the GPU runs one CTA per stream (U staged once per
    // stream, groups walked inside), a small one spreads a stream over nslab CTAs (latency)
    const bool walk = np <= 13 && stream_cnt >= 148;  // measured at 296 streams: 0.219 ms against 0.270 ms
    if (walk) warps = SOLVE_MAX_WARPS;
    const dim3 grid(walk ? 1 : nslab, stream_cnt), block(64 * warps);
    switch (np) {
      case 4: upd_solve_kernel<4><<<grid, block, smem, st>>>(d, stream_lo); break;
      case 7: upd_solve_kernel<7><<<grid, block, smem, st>>>(d, stream_lo); break;
      case 10: upd_solve_kernel<10><<<grid, block, smem, st>>>(d, stream_lo); break;
      case 13: upd_solve_kernel<13><<<grid, block, smem, st>>>(d, stream_lo); break;
      default: upd_solve_kernel<16><<<grid, block, smem, st>>>(d, stream_lo); break;
    }
    ++nl;
  }
  if ((e = mark(3)) != cudaSuccess) return e;
  if (!only_normalise) {
    // one 64x64 tile per CTA (measured: CTAs that walk several tiles with cross-tile prefetch were slower, 0.29-0.31
    // against 0.264 ms, because they cost the third resident CTA per SM)
    const int nt = (SL2_NXV + 3 * d.Nmax + 1 + 63) / 64;
    upd_syrk_kernel<<<dim3(nt * (nt + 1) / 2, stream_cnt), UPD_THREADS, SYRK_SMEM, st>>>(d, stream_lo);
    ++nl;
  }
  if ((e = mark(4)) != cudaSuccess) return e;
  upd_finish_kernel<<<stream_cnt, UPD_THREADS, 0, st>>>(d, stream_lo, staged_m >= 0, only_normalise);
  ++nl;
  if ((e = mark(5)) != cudaSuccess) return e;
  if (launches) *launches += nl;
  return cudaGetLastError();
}
