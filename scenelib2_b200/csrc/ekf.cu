// ekf.cu — EKF predict / measurement prediction + selection / update / cull on sm_100a.
//
// Replaces, per camera stream (one CTA per stream, all streams of a context in one launch):
//   Kalman::KalmanFilterPredict            kalman.cpp:50-69   (+ motion_model.cpp:84-217)
//   MonoSLAM::auto_select_n_features       monoslam.cpp:187-254 (+ :289-308,
//                                          full_feature_model.cpp:67-195, camera.cpp:90-300)
//   Kalman::KalmanFilterUpdate             kalman.cpp:72-119  (+ gather/scatter monoslam.cpp:501-614)
//   MonoSLAM::normalise_state + symmetrise monoslam.cpp:616-637, 143-150
//   MonoSLAM::delete_bad_features          monoslam.cpp:644-703, 770-812
//
// State layout in HBM: ONE dense column-major P (ld x ld) per stream in the order of
// construct_total_covariance (monoslam.cpp:518-546): [xv(13) | y_0 | y_1 | ...], with both
// triangles kept bit-consistent (the reference rebuilds the lower triangle from the upper blocks
// on every gather, so P is exactly block-symmetric whenever it is read).
//
// Update algorithm (mathematically the reference's K = P H^T S^-1, P -= K S K^T):
//   G = [ S | H P | nu ]  (m x (m+n+1), row-major scratch),  S = H P H^T + R
//   left-looking blocked Cholesky by row panels of SL2_NB rows applied to the whole of G
//   => G = [ U | Y | w ] with U^T U = S, Y = U^-T H P, w = U^-T nu
//   x += Y^T w ;  P -= Y^T Y  (upper 64x64 tiles computed, mirrored to the lower triangle)
// H is structurally sparse (13 + 3 non-zero columns per row) and is never formed.
// Small bit-critical prologue math (everything that decides WHICH pixels are searched: S_i,
// Sinv, h_i) uses never-fused __d*_rn ops in the oracle's evaluation order; the dense O(n^2 m)
// parts use ordinary FP64 FMAs (tolerance 1e-5 relative, north star).
#include "sl2_common.cuh"

namespace {

// never-fused FP64 scalar with natural operator syntax
struct rd {
  double v;
  __device__ __forceinline__ rd() : v(0.0) {}
  __device__ __forceinline__ rd(double x) : v(x) {}
};
__device__ __forceinline__ rd operator+(rd a, rd b) { return rd(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ rd operator-(rd a, rd b) { return rd(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ rd operator*(rd a, rd b) { return rd(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ rd operator/(rd a, rd b) { return rd(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ rd operator-(rd a) { return rd(-a.v); }
__device__ __forceinline__ rd rsqrt_(rd a) { return rd(__dsqrt_rn(a.v)); }

struct Quat {
  rd w, x, y, z;
};

__device__ Quat quat_mul(const Quat &a, const Quat &b) {
  Quat q;
  q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  q.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  q.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  q.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return q;
}

__device__ Quat quat_inverse(const Quat &q) {
  const rd n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  Quat r;
  if (n2.v > 0.0) {
    r.w = q.w / n2;
    r.x = (-q.x) / n2;
    r.y = (-q.y) / n2;
    r.z = (-q.z) / n2;
  }
  return r;
}

__device__ void quat_to_R(const Quat &q, rd R[3][3]) {
  const rd two(2.0), one(1.0);
  const rd tx = two * q.x, ty = two * q.y, tz = two * q.z;
  const rd twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const rd txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const rd tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = one - (tyy + tzz);
  R[0][1] = txy - twz;
  R[0][2] = txz + twy;
  R[1][0] = txy + twz;
  R[1][1] = one - (txx + tzz);
  R[1][2] = tyz - twx;
  R[2][0] = txz - twy;
  R[2][1] = tyz + twx;
  R[2][2] = one - (txx + tyy);
}

// ---------------------------------------------------------------------------------------------
// motion model on one thread: fv, F (13x13 col-major), G (13x6 col-major) -- motion_model.cpp
// ---------------------------------------------------------------------------------------------
__device__ void dqomegadt_by_domega(const rd om[3], rd dt, rd m[4][3]) {
  const rd omega = rsqrt_(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const rd two(2.0), one(1.0);
  const double sn = sin((omega * dt / two).v), cs = cos((omega * dt / two).v);
  const rd s(sn), c(cs);
  // motion_model.cpp:318-349
  auto dq0 = [&](rd a) { return ((-dt) / two) * (a / omega) * s; };
  auto dqA_A = [&](rd a) {
    return (dt / two) * a * a / (omega * omega) * c +
           (one / omega) * (one - a * a / (omega * omega)) * s;
  };
  auto dqA_B = [&](rd a, rd b) {
    return (a * b / (omega * omega)) * ((dt / two) * c - (one / omega) * s);
  };
  m[0][0] = dq0(om[0]);
  m[0][1] = dq0(om[1]);
  m[0][2] = dq0(om[2]);
  m[1][0] = dqA_A(om[0]);
  m[1][1] = dqA_B(om[0], om[1]);
  m[1][2] = dqA_B(om[0], om[2]);
  m[2][0] = dqA_B(om[1], om[0]);
  m[2][1] = dqA_A(om[1]);
  m[2][2] = dqA_B(om[1], om[2]);
  m[3][0] = dqA_B(om[2], om[0]);
  m[3][1] = dqA_B(om[2], om[1]);
  m[3][2] = dqA_A(om[2]);
}

__device__ void dq3_by_dq1(const Quat &q, rd m[4][4]) {  // math_util.cpp:82-97
  const rd x = q.x, y = q.y, z = q.z, w = q.w;
  const rd v[16] = {w, -x, -y, -z, x, w, -z, y, y, z, w, -x, z, -y, x, w};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) m[i][j] = v[i * 4 + j];
}
__device__ void dq3_by_dq2(const Quat &q, rd m[4][4]) {  // math_util.cpp:99-114
  const rd x = q.x, y = q.y, z = q.z, w = q.w;
  const rd v[16] = {w, -x, -y, -z, x, w, z, -y, y, -z, w, x, z, y, -x, w};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) m[i][j] = v[i * 4 + j];
}

// F and Gn are shared-memory col-major arrays (13x13, 13x6); fv 13.
__device__ void motion_model(const double *xv, const double *u3, double dt_, double *fv, double *F,
                             double *Gn) {
  const rd dt(dt_);
  const Quat qold = {rd(xv[3]), rd(xv[4]), rd(xv[5]), rd(xv[6])};
  const rd om[3] = {rd(xv[10]), rd(xv[11]), rd(xv[12])};
  // QuaternionFromAngularVelocity(omega * dt), math_util.cpp:61-80
  const rd av[3] = {om[0] * dt, om[1] * dt, om[2] * dt};
  const rd angle = rsqrt_(av[0] * av[0] + av[1] * av[1] + av[2] * av[2]);
  Quat qwt;
  if (angle.v > 0.0) {
    const rd sn(sin((angle / rd(2.0)).v)), cs(cos((angle / rd(2.0)).v));
    const rd s = sn / angle;
    qwt.x = s * av[0];
    qwt.y = s * av[1];
    qwt.z = s * av[2];
    qwt.w = cs;
  } else {
    qwt.w = rd(1.0);
  }
  const Quat qnew = quat_mul(qold, qwt);
  for (int i = 0; i < 3; ++i) fv[i] = (rd(xv[i]) + rd(xv[7 + i]) * dt).v;
  fv[3] = qnew.w.v;
  fv[4] = qnew.x.v;
  fv[5] = qnew.y.v;
  fv[6] = qnew.z.v;
  for (int i = 0; i < 3; ++i) fv[7 + i] = (rd(xv[7 + i]) + rd(u3 ? u3[i] : 0.0) * dt).v;
  for (int i = 0; i < 3; ++i) fv[10 + i] = om[i].v;

  for (int i = 0; i < 169; ++i) F[i] = 0.0;
  for (int i = 0; i < 13; ++i) F[i + 13 * i] = 1.0;
  for (int i = 0; i < 3; ++i) F[i + 13 * (7 + i)] = (rd(1.0) * dt).v;
  rd m44[4][4], m43[4][3], t44[4][4];
  dq3_by_dq2(qwt, m44);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) F[(3 + i) + 13 * (3 + j)] = m44[i][j].v;
  dq3_by_dq1(qold, t44);
  dqomegadt_by_domega(om, dt, m43);
  for (int i = 0; i < 78; ++i) Gn[i] = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) {
      rd sacc(0.0);
      for (int k = 0; k < 4; ++k) sacc = sacc + t44[i][k] * m43[k][j];
      F[(3 + i) + 13 * (10 + j)] = sacc.v;
      Gn[(3 + i) + 13 * (3 + j)] = sacc.v;  // same product in func_Q (motion_model.cpp:202-213)
    }
  for (int i = 0; i < 3; ++i) {
    Gn[(7 + i) + 13 * i] = 1.0;
    Gn[(10 + i) + 13 * (3 + i)] = 1.0;
    Gn[i + 13 * i] = (rd(1.0) * dt).v;
  }
}

// ---------------------------------------------------------------------------------------------
// per-feature measurement prediction (one thread per feature)
// ---------------------------------------------------------------------------------------------
struct FeatPred {
  rd h[2];
  rd dxp[2][7];
  rd dy[2][3];
  rd var;
  rd S[2][2];
};

__device__ void zeroedyi(const rd yi[3], const double *xp, rd z[3], rd dz_dxp[3][7], rd RRW[3][3]) {
  const rd d[3] = {yi[0] - rd(xp[0]), yi[1] - rd(xp[1]), yi[2] - rd(xp[2])};
  const Quat q = {rd(xp[3]), rd(xp[4]), rd(xp[5]), rd(xp[6])};
  const Quat qi = quat_inverse(q);
  quat_to_R(qi, RRW);
  for (int i = 0; i < 3; ++i) {
    rd s(0.0);
    for (int k = 0; k < 3; ++k) s = s + RRW[i][k] * d[k];
    z[i] = s;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) dz_dxp[i][j] = RRW[i][j] * rd(-1.0);
  // feature_model.cpp:187-238 applied to (qRW, d); then * dqbar_by_dq = diag(1,-1,-1,-1)
  const rd two(2.0);
  const rd w2 = two * qi.w, x2 = two * qi.x, y2 = two * qi.y, z2 = two * qi.z;
  const rd m0[9] = {w2, -z2, y2, z2, w2, -x2, -y2, x2, w2};
  const rd mx[9] = {x2, y2, z2, y2, -x2, -w2, z2, w2, -x2};
  const rd my[9] = {-y2, x2, w2, x2, y2, z2, -w2, z2, -y2};
  const rd mz[9] = {-z2, -w2, x2, w2, -z2, y2, x2, y2, z2};
  for (int i = 0; i < 3; ++i) {
    rd s0(0.0), s1(0.0), s2(0.0), s3(0.0);
    for (int k = 0; k < 3; ++k) {
      s0 = s0 + m0[i * 3 + k] * d[k];
      s1 = s1 + mx[i * 3 + k] * d[k];
      s2 = s2 + my[i * 3 + k] * d[k];
      s3 = s3 + mz[i * 3 + k] * d[k];
    }
    dz_dxp[i][3] = s0;
    dz_dxp[i][4] = -s1;
    dz_dxp[i][5] = -s2;
    dz_dxp[i][6] = -s3;
  }
}

// Pxx: shared 13x13 col-major; Pcol: global pointer to P(0, pos) (column-major, ld)
__device__ void predict_feature(const double *cam, const double *xv, const rd yi[3],
                                const double *Pxx, const double *Pcol, int ld, int pos,
                                FeatPred &o) {
  rd z[3], dz_dxp[3][7], RRW[3][3];
  zeroedyi(yi, xv, z, dz_dxp, RRW);
  const rd fku(cam[2]), fkv(cam[3]), u0(cam[4]), v0(cam[5]), kd1(cam[6]), sd(cam[7]);
  const rd one(1.0), two(2.0);
  // Camera::Project, camera.cpp:90-114
  const rd uc = (-fku) * z[0] / z[2];
  const rd vc = (-fkv) * z[1] / z[2];
  const rd radius2 = uc * uc + vc * vc;
  const rd factor = rsqrt_(one + two * kd1 * radius2);
  o.h[0] = uc / factor + u0;
  o.h[1] = vc / factor + v0;
  // Camera::ProjectionJacobian, camera.cpp:183-215
  const rd fku_yz = fku / z[2], fkv_yz = fkv / z[2];
  rd du[2][3];
  du[0][0] = -fku_yz;
  du[0][1] = rd(0.0);
  du[0][2] = fku_yz * z[0] / z[2];
  du[1][0] = rd(0.0);
  du[1][1] = -fkv_yz;
  du[1][2] = fkv_yz * z[1] / z[2];
  rd dh[2][2];
  dh[0][0] = uc * uc;
  dh[0][1] = uc * vc;
  dh[1][0] = vc * uc;
  dh[1][1] = vc * vc;
  const rd r2 = dh[0][0] + dh[1][1];
  const rd distor = one + two * kd1 * r2;
  const rd distor1_2 = rsqrt_(distor);
  const rd distor3_2 = distor1_2 * distor;
  const rd scale = rd(-2.0) * kd1 / distor3_2;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) dh[i][j] = dh[i][j] * scale;
  dh[0][0] = dh[0][0] + (one / distor1_2);
  dh[1][1] = dh[1][1] + (one / distor1_2);
  rd dhid[2][3];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      rd s(0.0);
      for (int k = 0; k < 2; ++k) s = s + dh[i][k] * du[k][j];
      dhid[i][j] = s;
    }
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 7; ++j) {
      rd s(0.0);
      for (int k = 0; k < 3; ++k) s = s + dhid[i][k] * dz_dxp[k][j];
      o.dxp[i][j] = s;
    }
    for (int j = 0; j < 3; ++j) {
      rd s(0.0);
      for (int k = 0; k < 3; ++k) s = s + dhid[i][k] * RRW[k][j];
      o.dy[i][j] = s;
    }
  }
  // Camera::MeasurementNoise, camera.cpp:282-300
  const rd dx = o.h[0] - u0, dyv = o.h[1] - v0;
  const rd distance = rsqrt_(dx * dx + dyv * dyv);
  const rd max_distance = rsqrt_(u0 * u0 + v0 * v0);
  const rd ratio = distance / max_distance;
  const rd sd_use = sd * (one + ratio);
  o.var = one * (sd_use * sd_use);
  // FeatureModel::func_Si, feature_model.cpp:99-116.  dh_by_dxv = [dh_by_dxp | 0(2x6)]
  // (motion_model.cpp:224-235), so terms with k >= 7 are exact zeros and are skipped.
  rd A[2][7], Bm[2][3], Cm[2][3];
  for (int r = 0; r < 2; ++r) {
    for (int j = 0; j < 7; ++j) {
      rd s(0.0);
      for (int k = 0; k < 7; ++k) s = s + o.dxp[r][k] * rd(Pxx[k + 13 * j]);
      A[r][j] = s;
    }
    for (int j = 0; j < 3; ++j) {
      rd s(0.0);
      for (int k = 0; k < 7; ++k) s = s + o.dxp[r][k] * rd(Pcol[k + (size_t)ld * j]);
      Bm[r][j] = s;
      rd t(0.0);
      for (int k = 0; k < 3; ++k) t = t + o.dy[r][k] * rd(Pcol[(pos + k) + (size_t)ld * j]);
      Cm[r][j] = t;
    }
  }
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) {
      rd s1(0.0), t1(0.0), t1t(0.0), s4(0.0);
      for (int k = 0; k < 7; ++k) s1 = s1 + A[r][k] * o.dxp[c][k];
      for (int k = 0; k < 3; ++k) t1 = t1 + Bm[r][k] * o.dy[c][k];
      for (int k = 0; k < 3; ++k) t1t = t1t + Bm[c][k] * o.dy[r][k];
      for (int k = 0; k < 3; ++k) s4 = s4 + Cm[r][k] * o.dy[c][k];
      rd S = rd(0.0) + s1;
      S = S + t1;
      S = S + t1t;
      S = S + s4;
      S = S + (r == c ? o.var : rd(0.0));
      o.S[r][c] = S;
    }
}

__device__ int visibility_test(const double *cam, const double *xp, const rd yi[3],
                               const double *xp_orig, const rd h[2]) {
  int cant = 0;
  const double bound = 20.0;  // kImageSearchBoundary_, full_feature_model.cpp:51
  if (h[0].v < 0.0 + bound || h[0].v > (double)((int)cam[0] - 1 - bound)) cant |= 1;
  if (h[1].v < 0.0 + bound || h[1].v > (double)((int)cam[1] - 1 - bound)) cant |= 2;
  rd z[3], t[3][7], R1[3][3], RWR[3][3];
  zeroedyi(yi, xp, z, t, R1);
  if (z[2].v <= 0) cant |= 16;
  rd a[3], b[3];
  quat_to_R(Quat{rd(xp[3]), rd(xp[4]), rd(xp[5]), rd(xp[6])}, RWR);
  for (int i = 0; i < 3; ++i) {
    rd s(0.0);
    for (int k = 0; k < 3; ++k) s = s + RWR[i][k] * z[k];
    a[i] = s;
  }
  zeroedyi(yi, xp_orig, z, t, R1);
  quat_to_R(Quat{rd(xp_orig[3]), rd(xp_orig[4]), rd(xp_orig[5]), rd(xp_orig[6])}, RWR);
  for (int i = 0; i < 3; ++i) {
    rd s(0.0);
    for (int k = 0; k < 3; ++k) s = s + RWR[i][k] * z[k];
    b[i] = s;
  }
  const rd ma = rsqrt_(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const rd mb = rsqrt_(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  const rd ratio = ma / mb;
  if (ratio.v > 2.0 || ratio.v < (1.0 / 2.0)) cant |= 4;
  const rd dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  double angle = acos((dot / (ma * mb)).v);
  angle = (angle >= 0.0 ? angle : -angle);
  if (angle > 3.14159265358979323846 * 45.0 / 180.0) cant |= 8;
  return cant;
}

// ---------------------------------------------------------------------------------------------
// kernel 1: predict (kalman.cpp:50-69) + measurement prediction / selection (monoslam.cpp:187-254)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) predict_kernel(const Sl2Dev d, int stream_lo,
                                                      const double *u3, int do_predict,
                                                      int do_measure) {
  const int s = stream_lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int nf = d.nfeat[s];
  const int n = SL2_NXV + 3 * nf;
  const int ld = d.ld;
  double *P = d.P + (size_t)s * ld * ld;
  double *x = d.x + (size_t)s * ld;

  __shared__ double F[169], Gn[78], Pxx[169], TT[169], Qm[169], xv[13], fv[13];
  __shared__ double score[SL2_MAX_FEAT_SMEM];
  __shared__ int vis[SL2_MAX_FEAT_SMEM];
  __shared__ int s_nvis, s_r0;

  if (tid < 13) xv[tid] = x[tid];
  for (int e = tid; e < 169; e += blockDim.x) Pxx[e] = P[(e % 13) + (size_t)ld * (e / 13)];
  __syncthreads();

  if (do_predict) {
    if (tid == 0) motion_model(xv, u3, d.dt, fv, F, Gn);
    __syncthreads();
    // Q = (G * Pnn) * G^T, Pnn = diag(lin x3, ang x3)   (motion_model.cpp:157-216)
    // TT = F * Pxx
    for (int e = tid; e < 169; e += blockDim.x) {
      const int i = e % 13, j = e / 13;
      const rd dt(d.dt);
      const rd lin = rd(4.0) * rd(4.0) * dt * dt, ang = rd(6.0) * rd(6.0) * dt * dt;
      rd q(0.0), t(0.0);
      for (int k = 0; k < 6; ++k) {
        const rd gp = rd(0.0) + rd(Gn[i + 13 * k]) * (k < 3 ? lin : ang);  // (G*Pnn)(i,k)
        q = q + gp * rd(Gn[j + 13 * k]);
      }
      for (int k = 0; k < 13; ++k) t = t + rd(F[i + 13 * k]) * rd(Pxx[k + 13 * j]);
      Qm[e] = q.v;
      TT[e] = t.v;
    }
    __syncthreads();
    // Pxx = TT * F^T + Q
    for (int e = tid; e < 169; e += blockDim.x) {
      const int i = e % 13, j = e / 13;
      rd a(0.0);
      for (int k = 0; k < 13; ++k) a = a + rd(TT[i + 13 * k]) * rd(F[j + 13 * k]);
      const double v = (a + rd(Qm[e])).v;
      Pxx[e] = v;  // own element only: no hazard with TT/F readers
      P[i + (size_t)ld * j] = v;
    }
    // Pxy_i = F * Pxy_i  (one thread per column of the 13 x 3N panel), mirrored below the diagonal
    for (int c = SL2_NXV + tid; c < n; c += blockDim.x) {
      double col[13], out[13];
      for (int k = 0; k < 13; ++k) col[k] = P[k + (size_t)ld * c];
      for (int i = 0; i < 13; ++i) {
        rd a(0.0);
        for (int k = 0; k < 13; ++k) a = a + rd(F[i + 13 * k]) * rd(col[k]);
        out[i] = a.v;
      }
      for (int i = 0; i < 13; ++i) {
        P[i + (size_t)ld * c] = out[i];
        P[c + (size_t)ld * i] = out[i];
      }
    }
    if (tid < 13) {
      x[tid] = fv[tid];
      xv[tid] = fv[tid];
    }
    __syncthreads();
  }
  if (!do_measure) return;

  // ---- per-feature prediction, visibility, score ---------------------------------------------
  const size_t fb = (size_t)s * d.Nmax;
  for (int i0 = 0; i0 < d.Nmax; i0 += blockDim.x) {
    const int i = i0 + tid;
    if (i < d.Nmax) {
      vis[i] = 0;
      score[i] = 0.0;
      d.sel_rank[fb + i] = -1;
      // d.found keeps its previous value for features that are not measured this frame, like
      // Feature::successful_measurement_flag_ (only written by make_measurements, monoslam.cpp:479-496)
      d.job_feat[fb + i] = -1;
    }
    if (i < nf) {
      const int pos = SL2_NXV + 3 * i;
      const rd yi[3] = {rd(x[pos]), rd(x[pos + 1]), rd(x[pos + 2])};
      FeatPred fp;
      predict_feature(d.cam, xv, yi, Pxx, P + (size_t)ld * pos, ld, pos, fp);
      d.h[(fb + i) * 2 + 0] = fp.h[0].v;
      d.h[(fb + i) * 2 + 1] = fp.h[1].v;
      for (int r = 0; r < 2; ++r) {
        for (int j = 0; j < 7; ++j) d.dh_dxp[(fb + i) * 14 + r * 7 + j] = fp.dxp[r][j].v;
        for (int j = 0; j < 3; ++j) d.dh_dy[(fb + i) * 6 + r * 3 + j] = fp.dy[r][j].v;
      }
      d.Rvar[fb + i] = fp.var.v;
      d.S[(fb + i) * 4 + 0] = fp.S[0][0].v;
      d.S[(fb + i) * 4 + 1] = fp.S[1][0].v;
      d.S[(fb + i) * 4 + 2] = fp.S[0][1].v;
      d.S[(fb + i) * 4 + 3] = fp.S[1][1].v;
      const int cant = visibility_test(d.cam, xv, yi, d.xp_org + (fb + i) * 7, fp.h);
      vis[i] = (cant == 0);
      score[i] = (fp.S[0][0] + fp.S[1][1]).v;  // trace, full_feature_model.cpp:172-176
    }
  }
  __syncthreads();
  // ---- insertion sort of monoslam.cpp:211-230 as a rank: strictly larger scores first, ties in
  //      feature order; selection stops at the first zero score or after n_select (:241-249)
  if (tid == 0) {
    s_nvis = 0;
    s_r0 = 1 << 30;
  }
  __syncthreads();
  for (int i = tid; i < nf; i += blockDim.x) {
    if (vis[i]) {
      int rank = 0;
      const double si = score[i];
      for (int j = 0; j < nf; ++j)
        if (vis[j] && (score[j] > si || (j < i && !(si > score[j])))) ++rank;
      d.sel_rank[fb + i] = rank;  // provisional: rank among visible
      atomicAdd(&s_nvis, 1);
      if (si == 0.0) atomicMin(&s_r0, rank);
    }
  }
  __syncthreads();
  const int nsel = min(min(d.n_select, s_r0), s_nvis);
  for (int i = tid; i < nf; i += blockDim.x) {
    int rank = d.sel_rank[fb + i];
    if (rank >= nsel) rank = -1;
    d.sel_rank[fb + i] = rank;
    if (rank >= 0) {
      d.job_feat[fb + rank] = i;
      d.job_centre[(fb + rank) * 2 + 0] = d.h[(fb + i) * 2 + 0];
      d.job_centre[(fb + rank) * 2 + 1] = d.h[(fb + i) * 2 + 1];
      double p00, p01, p11;
      if (d.ovr[0] > 0.0) {
        p00 = d.ovr[0];
        p01 = d.ovr[1];
        p11 = d.ovr[2];
      } else {
        // monoslam.cpp:371-374: LLT, inverse of L, Sinv = Linv^T Linv (closed form, see oracle)
        const rd s00(d.S[(fb + i) * 4 + 0]), s10(d.S[(fb + i) * 4 + 1]), s11(d.S[(fb + i) * 4 + 3]);
        const rd l00 = rsqrt_(s00);
        const rd l10 = s10 / l00;
        const rd l11 = rsqrt_(s11 - l10 * l10);
        const rd x00 = rd(1.0) / l00;
        const rd x10 = (rd(0.0) - l10 * x00) / l11;
        const rd x11 = rd(1.0) / l11;
        p00 = (x00 * x00 + x10 * x10).v;
        p01 = (x10 * x11).v;
        p11 = (x11 * x11).v;
      }
      d.job_puinv[(fb + rank) * 3 + 0] = p00;
      d.job_puinv[(fb + rank) * 3 + 1] = p01;
      d.job_puinv[(fb + rank) * 3 + 2] = p11;
    }
  }
  if (tid == 0) {
    d.nsel[s] = nsel;
    d.nvisible[s] = s_nvis;
    d.nmeas[s] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// kernel 2: EKF update (kalman.cpp:72-119) + normalise (monoslam.cpp:616-637) + bookkeeping
// ---------------------------------------------------------------------------------------------
struct UpdSmem {
  // carved from dynamic shared memory; sizes depend on Nmax
  double *wv;    // [mmax]  nu (copied into the last column of G)
  int *mfeat;    // [K]
  double *Rv;    // [K]
  double *mult;  // [mmax][UPD_MS] (negated) multipliers of the current panel; phases 1a/1b: H*P(:, 0:13)
  double *dg;    // [NB][UPD_DS] diagonal block of the current panel (factor scratch)
  double *Wm;    // [NB][UPD_WS]  W = U_pp^-T of the current panel
  double *xacc;  // [ld]  Y^T w accumulated panel by panel
  double *pan;   // phase 1: HxT / Hy (aliased); phase 2: panel buffer [NB][panw];
                 // phase 4: Y slabs of tile_products (2 stages x 2 slabs)
  double *HxT;   // = pan          [16][hms]  Hx transposed, k-major, zero padded (cols 13..15, rows >= m)
  double *Hy;    // = pan + 16*hms [K][2][3]
  int panw, hms;
};

constexpr int UPD_THREADS = 256;
constexpr int UPD_NB = 16;   // Cholesky row-panel height (two DMMA M-tiles)
constexpr int UPD_WS = 20;   // row stride of the W table
constexpr int UPD_HXS = 14;  // row stride of the H*P(:, 0:13) table phase 1a leaves in sm.mult for phase 1b
constexpr int UPD_DS = 20;   // row stride of the diagonal-block scratch (conflict-free fragments)
constexpr int UPD_MS = 20;   // row stride of the multiplier table: 32 B (mod 128) => conflict-free A fragments
constexpr int UPD_KC = 32;   // k-chunk of the Y^T Y tiles
constexpr int UPD_YS = 68;   // padded row stride of a staged Y slab (doubles): conflict-free DMMA reads
constexpr int UPD_GB = 4;    // 8-column groups per warp iteration in the panel update

__host__ __device__ inline int upd_keven(int Nmax) { return (Nmax + 1) & ~1; }
__host__ __device__ inline int upd_panw(int Nmax) {
  // row stride = 2 (mod 16) doubles: the 8 rows of a DMMA C fragment hit distinct banks
  return ((2 * upd_keven(Nmax) + SL2_NXV + 3 * Nmax + 1 + 15) & ~15) + 2;
}
__host__ __device__ inline int upd_hms(int Nmax) {
  // k-major Hx table: row stride = 4 (mod 16) doubles => the 4 k-rows of a fragment are 32 B apart
  return ((2 * upd_keven(Nmax) + 15) & ~15) + 4;
}
__host__ __device__ inline size_t upd_pan_doubles(int Nmax) {
  size_t a = (size_t)UPD_NB * upd_panw(Nmax);
  const size_t b = 2 * 2 * UPD_KC * UPD_YS, c = 64 * 65;
  const size_t h = (size_t)16 * upd_hms(Nmax) + (size_t)upd_keven(Nmax) * 6;
  if (b > a) a = b;
  if (c > a) a = c;
  if (h > a) a = h;
  return (a + 1) & ~(size_t)1;
}

__device__ __forceinline__ UpdSmem carve(uint8_t *base, int Nmax) {
  UpdSmem u;
  const int K = upd_keven(Nmax), mmax = 2 * K;  // even counts keep every section 16 B aligned
  double *p = reinterpret_cast<double *>(base);
  u.wv = p;  p += mmax;
  u.Rv = p;  p += K;
  u.mult = p;  p += (size_t)mmax * UPD_MS;
  u.dg = p;  p += UPD_NB * UPD_DS;
  u.Wm = p;  p += UPD_NB * UPD_WS;
  u.xacc = p;  p += (SL2_NXV + 3 * Nmax + 7) & ~7;
  u.pan = p;  p += upd_pan_doubles(Nmax);
  u.panw = upd_panw(Nmax);
  u.hms = upd_hms(Nmax);
  u.HxT = u.pan;
  u.Hy = u.pan + (size_t)16 * u.hms;
  u.mfeat = reinterpret_cast<int *>(p);
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

// 64x64 tile products T_t = sum_{k < kr} A_t(k, :)^T B_t(k, :) for a list of tiles, A_t / B_t = 64-column
// slabs of the row-major matrix Gm (row stride ldg) starting at the columns tile_cols(t) returns.  FP64
// DMMA tiles; the slabs are staged by cp.async (LDGSTS) into a double-buffered, conflict-free (stride
// UPD_YS) shared tile; the stage sequence is flattened over (tile, k-chunk) so the first chunk of the
// next tile is in flight while the epilogue of the current one runs.  Warp w owns rows 16*(w%4).. and
// columns 32*(w/4).. of the tile; epi(t, acc) gets the DMMA C fragments:
// acc[i][j][e] = T(16*(w%4) + 8*i + lane/4, 32*(w/4) + 8*j + 2*(lane%4) + e).
// Columns >= limA / limB and rows >= kr are zero-filled.  `stage_buf` = 2*2*UPD_KC*UPD_YS doubles.
template <class ColsFn, class EpiFn>
__device__ __forceinline__ void tile_products(double *stage_buf, const double *__restrict__ Gm, int ldg,
                                              int kr, int ntiles, int limA, int limB, ColsFn tile_cols,
                                              EpiFn epi) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int wa = (warp & 3) * 16, wb = (warp >> 2) * 32;
  const int nchunk = (kr + UPD_KC - 1) / UPD_KC;
  const int total = ntiles * nchunk;
  if (total <= 0) return;
  // stage loader: 2 slabs x KC rows x 64 columns; thread = (row warp + 8*j, 16-byte segment `lane`)
  int st_tile = 0, st_chunk = 0;  // the (tile, chunk) the next stage() call loads
  auto stage = [&](int buf) {
    int colA, colB;
    tile_cols(st_tile, colA, colB);
    colA += 2 * lane;
    colB += 2 * lane;
    const int bytesA = colA + 1 < limA ? 16 : (colA < limA ? 8 : 0);
    const int bytesB = colB + 1 < limB ? 16 : (colB < limB ? 8 : 0);
    const double *srcA = Gm + (bytesA ? colA : 0);
    const double *srcB = Gm + (bytesB ? colB : 0);
    double *dst = stage_buf + (size_t)buf * (2 * UPD_KC * UPD_YS) + 2 * lane;
#pragma unroll
    for (int j = 0; j < UPD_KC / 8; ++j) {
      const int kk = warp + 8 * j;
      const int k = st_chunk * UPD_KC + kk;
      const bool kv = k < kr;
      const size_t ro = (size_t)(kv ? k : 0) * ldg;
      cp_async16(dst + kk * UPD_YS, srcA + ro, kv ? bytesA : 0);
      cp_async16(dst + UPD_KC * UPD_YS + kk * UPD_YS, srcB + ro, kv ? bytesB : 0);
    }
    cp_async_commit();
    if (++st_chunk == nchunk) {
      st_chunk = 0;
      ++st_tile;
    }
  };
  double acc[2][4][2];
  __syncthreads();  // previous users of the staging area are done
  stage(0);
  int tile = 0, ch = 0;
  for (int sidx = 0; sidx < total; ++sidx) {
    if (sidx + 1 < total) {
      stage((sidx + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (ch == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    const double *Ya = stage_buf + (size_t)(sidx & 1) * (2 * UPD_KC * UPD_YS);
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
    __syncthreads();  // buffer (sidx & 1) may be refilled by the stage issued in the next iteration
    if (++ch == nchunk) {
      epi(tile, acc);
      ch = 0;
      ++tile;
    }
  }
}

__global__ void __launch_bounds__(UPD_THREADS, 2) update_kernel(
    const Sl2Dev d, int stream_lo, int staged_m, const int *st_feat, const double *st_Hxv,
    const double *st_Hy, const double *st_R, const double *st_nu, int only_normalise) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const UpdSmem sm = carve(smem_raw, d.Nmax);
  const int s = stream_lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int nf = d.nfeat[s];
  const int n = SL2_NXV + 3 * nf;
  const int ld = d.ld, ldg = d.ldg;
  double *__restrict__ P = d.P + (size_t)s * ld * ld;
  double *__restrict__ x = d.x + (size_t)s * ld;
  double *__restrict__ G = d.G + (size_t)s * d.mmax * ldg;
  const size_t fb = (size_t)s * d.Nmax;
  const int warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;  // DMMA fragment coordinates
  __shared__ int s_m, s_next;
#ifdef SL2_PHASE_STAMPS  // clock64 stamps of CTA 0 (tools / bench SL2_PHASES=1); off in product builds
#define PH(i) do { if (blockIdx.x == 0 && tid == 0) d.dbg[(i)] = clock64(); } while (0)
#define PHQ(stmt) do { if (blockIdx.x == 0 && tid == 32) { stmt; } } while (0)
#else
#define PH(i) do { } while (0)
#define PHQ(stmt) do { } while (0)
#endif
  PH(0);
  PHQ(d.dbg[16] = d.dbg[17] = d.dbg[18] = d.dbg[19] = d.dbg[20] = 0);

  // ---- phase 0: measurement list in selected order, successful only (monoslam.cpp:556-571) ---
  if (tid == 0) {
    int k = 0;
    if (only_normalise) {
      k = 0;
    } else if (staged_m >= 0) {
      k = staged_m / 2;
    } else {
      const int nsel = d.nsel[s];
      for (int r = 0; r < nsel; ++r) {
        const int i = d.job_feat[fb + r];
        if (i >= 0 && d.found[fb + i]) sm.mfeat[k++] = i;
      }
      d.nmeas[s] = k;
    }
    s_m = 2 * k;
  }
  __syncthreads();
  const int m = s_m;
  const int K = m / 2;
  if (m > 0) {
    const int HMS = sm.hms;
    // HxT[c][i] = H_xv(i, c) (k-major, zero padded), Hy[k][r][c], Rv[k], wv = nu
    for (int e = tid; e < 16 * HMS; e += UPD_THREADS) sm.HxT[e] = 0.0;
    __syncthreads();
    if (staged_m >= 0) {
      for (int k = tid; k < K; k += UPD_THREADS) {
        sm.mfeat[k] = st_feat[k];
        sm.Rv[k] = st_R[k * 4];  // R_i = var * I (camera.cpp:294-299); off-diagonals ignored
        sm.wv[2 * k] = st_nu[2 * k];
        sm.wv[2 * k + 1] = st_nu[2 * k + 1];
      }
      for (int e = tid; e < m * 13; e += UPD_THREADS) {
        const int i = e / 13, c = e - i * 13;
        sm.HxT[c * HMS + i] = st_Hxv[e];
      }
      for (int e = tid; e < K * 6; e += UPD_THREADS) sm.Hy[e] = st_Hy[e];
    } else {
      for (int k = tid; k < K; k += UPD_THREADS) {
        const int i = sm.mfeat[k];
        sm.Rv[k] = d.Rvar[fb + i];
        // nu = z - h (full_feature_model.cpp:197-200), z = (double)(u,v) (monoslam.cpp:382-383)
        sm.wv[2 * k] = (rd((double)d.z_uv[(fb + i) * 2]) - rd(d.h[(fb + i) * 2])).v;
        sm.wv[2 * k + 1] = (rd((double)d.z_uv[(fb + i) * 2 + 1]) - rd(d.h[(fb + i) * 2 + 1])).v;
        for (int r = 0; r < 2; ++r) {
          for (int c = 0; c < 7; ++c) sm.HxT[c * HMS + 2 * k + r] = d.dh_dxp[(fb + i) * 14 + r * 7 + c];
          for (int c = 0; c < 3; ++c) sm.Hy[k * 6 + r * 3 + c] = d.dh_dy[(fb + i) * 6 + r * 3 + c];
        }
      }
    }
    __syncthreads();

    PH(1);
    // ---- phase 1a: H*P.  Dense part H_xv (m x 16) * P(0:16, :) on DMMA tiles; the 3 structural
    //      columns of dh/dy are added per element; nu goes into the last column.
    {
      constexpr int QB = 4;  // column groups per warp pass
      const int mtiles = (m + 7) >> 3, ngrp = (n + 7) >> 3;
      for (int gq = warp * QB; gq < ngrp; gq += (UPD_THREADS / 32) * QB) {
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
        for (int mt = 0; mt < mtiles; ++mt) {
          const int i = mt * 8 + lr;
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
          for (int ks = 0; ks < 4; ++ks) a[ks] = sm.HxT[(4 * ks + lc) * HMS + mt * 8 + lr];
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
              if (j0[q] < SL2_NXV) {  // dense 13 columns of H*P: kept in shared memory for phase 1b
                sm.mult[i * UPD_HXS + j0[q]] = c0;
                if (j0[q] + 1 < SL2_NXV) sm.mult[i * UPD_HXS + j0[q] + 1] = c1;
              }
            }
          }
        }
      }
    }
    for (int i = tid; i < m; i += UPD_THREADS) G[(size_t)i * ldg + m + n] = sm.wv[i];
    __syncthreads();
    PH(2);
    // ---- phase 1b: S = (H P) H^T + R, upper triangle, one warp per row.  The dense 13 columns of
    //      the row of H*P come from shared memory (written by phase 1a; broadcast reads), lane = measured feature
    //      (two columns of S); the 3 structural dh/dy columns and R are added per element.
    {
      constexpr int HXS = UPD_HXS;  // H*P(:, 0:13), left in sm.mult by phase 1a
      const double *hpx = sm.mult;
      constexpr int SCH = 4;  // feature chunks of 32 per pass (covers K <= 128 in one pass)
      for (int i = warp; i < m; i += UPD_THREADS / 32) {
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
                const double hx = hpx[i * HXS + c];
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
              if (i == 2 * kk) s0 += sm.Rv[kk];
              if (i == 2 * kk + 1) s1 += sm.Rv[kk];
              if (k < K) *reinterpret_cast<double2 *>(G + (size_t)i * ldg + 2 * k) = make_double2(s0, s1);
            }
          }
        }
      }
    }
    __syncthreads();

    PH(3);
    // ---- phase 2: left-looking Cholesky by row panels of 16 on G = [S | HP | nu] ---------------
    // Trailing update of a panel = C(16 x cols) - A(16 x i0) * B(i0 x cols) with A(r,k) = U(k,i0+r)
    // (multipliers, shared memory) and B = finished rows of G (global / L2): FP64 tensor-core
    // tiles (DMMA m8n8k4, two M tiles per B fragment), each warp owning groups of 8 columns; the B
    // fragments are software-pipelined three k-steps ahead.
    const int width = m + n + 1;
    const int PW = sm.panw;
    for (int j = tid; j < n; j += UPD_THREADS) sm.xacc[j] = 0.0;
    for (int i0 = 0; i0 < m; i0 += UPD_NB) {
      const int nbp = min(UPD_NB, m - i0);
#ifdef SL2_PHASE_STAMPS
      long long tq0 = 0, tq1 = 0, tq2 = 0, tq3 = 0, tq35 = 0;
#endif
      PHQ(tq0 = clock64());
      if (tid == 0) s_next = 1;  // batch 0 is reserved for warp 0
      // multipliers, negated so that D = (-A) * B + C
      for (int e = tid; e < i0 * UPD_NB; e += UPD_THREADS) {
        const int k = e / UPD_NB, r = e - k * UPD_NB;
        sm.mult[k * UPD_MS + r] = (r < nbp) ? -G[(size_t)k * ldg + i0 + r] : 0.0;
      }
      __syncthreads();
      PHQ(tq1 = clock64());
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
          // more DMMA product Y_panel = W * C_panel).  This is the serial path of the panel, so it is
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
      PHQ(tq2 = clock64());
      __syncthreads();
      PHQ(tq3 = clock64());
      // finish the panel: rows of U for the 16 diagonal columns, Y_panel = W * C_panel (DMMA) for
      // all other columns, written straight to G from the C fragments
      for (int e = tid; e < UPD_NB * UPD_NB; e += UPD_THREADS) {
        const int r = e / UPD_NB, cc = e - r * UPD_NB;
        if (r < nbp && cc < nbp && i0 + cc < width)
          G[(size_t)(i0 + r) * ldg + i0 + cc] = (cc >= r) ? sm.pan[(size_t)r * PW + cc] : 0.0;
      }
      {
        double aw[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) aw[mt][ks] = sm.Wm[(mt * 8 + lr) * UPD_WS + 4 * ks + lc];
        const int ncol = width - i0;
        // w_panel = W * nu_panel for this lane's two rows (the nu column is the last one)
        double w_lo = 0.0, w_hi = 0.0;
#pragma unroll
        for (int k = 0; k < UPD_NB; ++k) {
          const double nuk = sm.pan[(size_t)k * PW + ncol - 1];
          w_lo += sm.Wm[lr * UPD_WS + k] * nuk;
          w_hi += sm.Wm[(8 + lr) * UPD_WS + k] * nuk;
        }
        // FG column groups per iteration: independent DMMA chains; a column belongs to exactly one
        // warp within a panel and panels are separated by barriers, so xacc needs no atomics
        constexpr int FG = 4;
        for (int gb = (nbp >> 3) + warp; gb * 8 < ncol; gb += FG * (UPD_THREADS / 32)) {
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
          double p0[FG], p1[FG];
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
            // x += Y^T w, fused: the 16 rows of a column live in the 8 lanes sharing lc
            p0[f] = c[f][0][0] * w_lo + c[f][1][0] * w_hi;
            p1[f] = c[f][0][1] * w_lo + c[f][1][1] * w_hi;
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1)
#pragma unroll
            for (int f = 0; f < FG; ++f) {
              p0[f] += __shfl_xor_sync(0xffffffffu, p0[f], o);
              p1[f] += __shfl_xor_sync(0xffffffffu, p1[f], o);
            }
          if (lr == 0) {
#pragma unroll
            for (int f = 0; f < FG; ++f) {
              const int j = i0 + (gb + f * (UPD_THREADS / 32)) * 8 + 2 * lc - m;  // column of Y
              if (j >= 0 && j < n) sm.xacc[j] += p0[f];
              if (j + 1 >= 0 && j + 1 < n) sm.xacc[j + 1] += p1[f];
            }
          }
        }
      }
      PHQ(tq35 = clock64());
      __syncthreads();
      PHQ(d.dbg[16] += tq1 - tq0;      // multipliers + barrier
          d.dbg[17] += tq2 - tq1;      // this warp's DMMA batches
          d.dbg[18] += tq3 - tq2;      // waiting for the other warps / the diagonal factor
          d.dbg[19] += clock64() - tq3;  // finishing the panel + barrier
          d.dbg[20] += tq35 - tq3);    // finishing work of this warp alone
    }

    PH(4);
    // ---- phase 3: x += Y^T w (accumulated panel by panel above) --------------------------------
    for (int j = tid; j < n; j += UPD_THREADS) x[j] += sm.xacc[j];

    PH(5);
    // ---- phase 4: P -= Y^T Y on 64x64 tiles (upper triangle computed, lower mirrored), written back
    //      straight from the DMMA fragments (64-byte row segments either way, P is column-major)
    {
      const int nt = (n + 63) / 64;
      const int wa = (warp & 3) * 16, wb = (warp >> 2) * 32;
      auto unrank = [&](int t, int &ta, int &tb) {  // tiles in (tb outer, ta <= tb inner) order
        tb = 0;
        while ((tb + 1) * (tb + 2) / 2 <= t) ++tb;
        ta = t - tb * (tb + 1) / 2;
      };
      tile_products(
          sm.pan, G, ldg, m, nt * (nt + 1) / 2, m + n, m + n,
          [&](int t, int &colA, int &colB) {
            int ta, tb;
            unrank(t, ta, tb);
            colA = m + ta * 64;
            colB = m + tb * 64;
          },
          [&](int t, double (&acc)[2][4][2]) {
            int ta, tb;
            unrank(t, ta, tb);
            const bool mirror = ta != tb;
            double pold[2][4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int a = ta * 64 + wa + i * 8 + lr, bq = tb * 64 + wb + j * 8 + 2 * lc + e;
                  pold[i][j][e] = (a < n && bq < n) ? P[a + (size_t)ld * bq] : 0.0;
                }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int a = ta * 64 + wa + i * 8 + lr, bq = tb * 64 + wb + j * 8 + 2 * lc;
                const double v0 = pold[i][j][0] - acc[i][j][0], v1 = pold[i][j][1] - acc[i][j][1];
                if (a < n && bq < n) {
                  P[a + (size_t)ld * bq] = v0;
                  if (bq + 1 < n) P[a + (size_t)ld * (bq + 1)] = v1;
                  if (mirror) {  // lower tile: rows = b range (contiguous in P), column a
                    double *dst = P + bq + (size_t)ld * a;
                    if (bq + 1 < n) *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
                    else *dst = v0;
                  }
                }
              }
          });
    }
    __syncthreads();
  }

  PH(6);
  // ---- phase 5: normalise_state (monoslam.cpp:616-637): P <- J P J^T, J = diag(I3, dqnorm, I6, I)
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

  PH(7);
  // ---- bookkeeping: attempt / success counters (monoslam.cpp:479-496) ------------------------
  if (staged_m < 0 && !only_normalise) {
    if (tid == 0) s_next = 0;  // reused: number of features delete_bad_features would cull
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
      if (att >= d.min_attempts && (double)suc / (double)att < d.match_fraction) atomicAdd(&s_next, 1);
    }
    __syncthreads();
    if (tid == 0) d.ncull[s] = s_next;
  }
}

// ---------------------------------------------------------------------------------------------
// kernel 3: delete_bad_features (monoslam.cpp:644-703) / delete_feature (:770-812)
// removes the rows/columns of the culled features from x and P in place.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cull_kernel(const Sl2Dev d, int stream_lo, int force_index) {
  const int s = stream_lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int nf = d.nfeat[s];
  const int ld = d.ld;
  const size_t fb = (size_t)s * d.Nmax;
  if (force_index < 0 && d.ncull[s] == 0) return;  // nothing to cull (decided by update_kernel)
  __shared__ int keep[SL2_MAX_FEAT_SMEM];  // new index of feature i or -1
  __shared__ int s_new;
  if (tid == 0) {
    int k = 0;
    for (int i = 0; i < nf; ++i) {
      bool kill;
      if (force_index >= 0) {
        kill = (i == force_index);
      } else {
        const int att = d.attempted[fb + i], suc = d.successful[fb + i];
        kill = att >= d.min_attempts && (double)suc / (double)att < d.match_fraction;
      }
      keep[i] = kill ? -1 : k++;
    }
    s_new = k;
  }
  __syncthreads();
  const int nk = s_new;
  if (nk == nf) return;
  double *P = d.P + (size_t)s * ld * ld;
  double *x = d.x + (size_t)s * ld;
  double *scr = d.G + (size_t)s * d.mmax * d.ldg;  // scratch >= ld*ld? no: compact column by column
  const int n = SL2_NXV + 3 * nf;
  // destination indices are never larger than source indices, so walking columns in increasing
  // order with a per-column staging buffer in scratch is race-free inside one CTA.
  for (int c = 0; c < n; ++c) {
    int cn;
    if (c < SL2_NXV) {
      cn = c;
    } else {
      const int f = (c - SL2_NXV) / 3;
      cn = keep[f] < 0 ? -1 : SL2_NXV + 3 * keep[f] + (c - SL2_NXV) % 3;
    }
    if (cn < 0) continue;  // uniform across the CTA
    for (int r = tid; r < n; r += blockDim.x) scr[r] = P[r + (size_t)ld * c];
    __syncthreads();
    for (int r = tid; r < n; r += blockDim.x) {
      int rn;
      if (r < SL2_NXV) {
        rn = r;
      } else {
        const int f = (r - SL2_NXV) / 3;
        rn = keep[f] < 0 ? -1 : SL2_NXV + 3 * keep[f] + (r - SL2_NXV) % 3;
      }
      if (rn >= 0) P[rn + (size_t)ld * cn] = scr[r];
    }
    __syncthreads();
  }
  // state vector and per-feature records
  for (int r = tid; r < n; r += blockDim.x) scr[r] = x[r];
  __syncthreads();
  for (int r = SL2_NXV + tid; r < n; r += blockDim.x) {
    const int f = (r - SL2_NXV) / 3;
    if (keep[f] >= 0) x[SL2_NXV + 3 * keep[f] + (r - SL2_NXV) % 3] = scr[r];
  }
  __syncthreads();
  if (tid == 0) {
    // serial compaction of the small per-feature records (rare path)
    // selected_feature_list_.erase (monoslam.cpp:258-281 via :797-798): later entries move up
    for (int i = 0; i < nf; ++i) {
      const int r = d.sel_rank[fb + i];
      if (keep[i] < 0 && r >= 0) {
        for (int j = 0; j < nf; ++j)
          if (d.sel_rank[fb + j] > r) d.sel_rank[fb + j] -= 1;
        d.sel_rank[fb + i] = -1;
      }
    }
    const int box16 = d.box * 16;
    for (int i = 0; i < nf; ++i) {
      const int k = keep[i];
      if (k < 0 || k == i) continue;
      for (int e = 0; e < 7; ++e) d.xp_org[(fb + k) * 7 + e] = d.xp_org[(fb + i) * 7 + e];
      d.attempted[fb + k] = d.attempted[fb + i];
      d.successful[fb + k] = d.successful[fb + i];
      for (int e = 0; e < box16; ++e)
        d.patches[(fb + k) * box16 + e] = d.patches[(fb + i) * box16 + e];
      for (int e = 0; e < 2; ++e) {
        d.h[(fb + k) * 2 + e] = d.h[(fb + i) * 2 + e];
        d.z_uv[(fb + k) * 2 + e] = d.z_uv[(fb + i) * 2 + e];
      }
      for (int e = 0; e < 4; ++e) d.S[(fb + k) * 4 + e] = d.S[(fb + i) * 4 + e];
      d.sel_rank[fb + k] = d.sel_rank[fb + i];
      d.found[fb + k] = d.found[fb + i];
    }
    d.nfeat[s] = nk;
  }
}

}  // namespace

size_t sl2_update_smem_bytes(const Sl2Dev &d) {
  const size_t K = upd_keven(d.Nmax), mmax = 2 * K;
  const size_t doubles = mmax + K + mmax * UPD_MS + UPD_NB * UPD_DS + UPD_NB * UPD_WS +
                         ((SL2_NXV + 3 * d.Nmax + 7) & ~7) + upd_pan_doubles(d.Nmax);
  return doubles * 8 + K * 4 + 16;
}

cudaError_t sl2_configure_update(const Sl2Dev &d) {
  return cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sl2_update_smem_bytes(d));
}

cudaError_t sl2_launch_predict(const Sl2Dev &d, int stream_lo, int stream_cnt, const double *u3_dev,
                               int do_predict, int do_measure, cudaStream_t st) {
  if (stream_cnt <= 0) return cudaSuccess;
  // 128 threads = one per feature (SL2_MAX_FEATURES); the kernel needs ~255 registers per thread,
  // so 128-thread CTAs are what lets two streams share an SM
  predict_kernel<<<stream_cnt, 128, 0, st>>>(d, stream_lo, u3_dev, do_predict, do_measure);
  return cudaGetLastError();
}

cudaError_t sl2_launch_update(const Sl2Dev &d, int stream_lo, int stream_cnt, int staged_m,
                              const int *st_feat, const double *st_Hxv, const double *st_Hy,
                              const double *st_R, const double *st_nu, int only_normalise,
                              cudaStream_t st) {
  if (stream_cnt <= 0) return cudaSuccess;
  const size_t smem = sl2_update_smem_bytes(d);
  update_kernel<<<stream_cnt, UPD_THREADS, smem, st>>>(d, stream_lo, staged_m, st_feat, st_Hxv,
                                                        st_Hy, st_R, st_nu, only_normalise);
  return cudaGetLastError();
}

cudaError_t sl2_launch_cull(const Sl2Dev &d, int stream_lo, int stream_cnt, int force_index,
                            cudaStream_t st) {
  if (stream_cnt <= 0) return cudaSuccess;
  cull_kernel<<<stream_cnt, 256, 0, st>>>(d, stream_lo, force_index);
  return cudaGetLastError();
}
