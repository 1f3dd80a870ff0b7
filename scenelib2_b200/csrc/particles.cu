// Depth particles of partially-initialised features, device side of sl2_measure_partial_features /
// sl2_measure_particles.
//   particle_predict_kernel   MonoSLAM::predict_partially_initialised_feature_measurements (monoslam.cpp:1347-1400):
//                             per particle h_pi (PartFeatureModel, part_feature_model.cpp:80-143, 231-265), R_i,
//                             S_i (feature_model.cpp:99-116), S_i^-1 and det S_i (Particle::set_S,
//                             feature_init_info.cpp:57-65); one CTA per feature, one thread per particle.
//   particle_kernel           MonoSLAM::update_partially_initialised_feature_probabilities
// (monoslam.cpp:1447-1493, body for one FeatureInitInfo) + FeatureInitInfo::normalise_particle_vector_and_
// calculate_cumulative / prune_particle_vector / calculate_mean_and_covariance (feature_init_info.cpp:95-172).
//                             after the SMOE kernels (smoe.cu), on the match positions they left in device memory;
//                             one CTA per feature.
#include "sl2_common.cuh"

namespace {

// One CTA.  The likelihoods are independent (one thread per particle); the normalisation sums are
// order dependent in FP64, so one thread adds them in particle order like the reference's loops.
__global__ void __launch_bounds__(128) particle_kernel(int Kmax, const int *__restrict__ Kf,
                                                       const double *__restrict__ h,
                                                       const double *__restrict__ sinv3,
                                                       const double *__restrict__ detS,
                                                       const double *__restrict__ lambda,
                                                       const int *__restrict__ z_uv,
                                                       const uint8_t *__restrict__ found, double prune_threshold,
                                                       double *prob, uint8_t *keep, double *cumulative,
                                                       double *mean_var, int *left_out) {
  {  // feature blockIdx.x: its K particles at offset blockIdx.x * Kmax of every array
    const size_t o = (size_t)blockIdx.x * Kmax;
    h += 2 * o, sinv3 += 3 * o, detS += o, lambda += o, z_uv += 2 * o, found += o;
    prob += o, keep += o, cumulative += o, mean_var += 2 * blockIdx.x, left_out += blockIdx.x;
  }
  const int K = Kf[blockIdx.x];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {  // monoslam.cpp:1456-1478
    double likelihood = 0.0;
    if (found[k]) {
      const double nu0 = sub_((double)z_uv[2 * k], h[2 * k]), nu1 = sub_((double)z_uv[2 * k + 1], h[2 * k + 1]);
      const double r0 = add_(mul_(sinv3[3 * k], nu0), mul_(sinv3[3 * k + 1], nu1));
      const double r1 = add_(mul_(sinv3[3 * k + 1], nu0), mul_(sinv3[3 * k + 2], nu1));
      const double q = add_(mul_(nu0, r0), mul_(nu1, r1));
      // 1 / sqrt(2 pi det S) * exp(-q / 2); device exp is within 1 ulp of the correctly rounded value
      likelihood = mul_(div_(1.0, sqrt_(mul_(6.283185307179586476925286766559, detS[k]))), exp(mul_(-0.5, q)));
    }
    prob[k] = mul_(prob[k], likelihood);
    keep[k] = 1;
    cumulative[k] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  auto normalise = [&]() -> bool {  // feature_init_info.cpp:95-119
    double total = 0.0;
    for (int k = 0; k < K; ++k)
      if (keep[k]) total = add_(total, prob[k]);
    if (total == 0.0) return false;
    double cum = 0.0;
    for (int k = 0; k < K; ++k)
      if (keep[k]) {
        const double p = div_(prob[k], total);
        prob[k] = p;
        cumulative[k] = add_(cum, p);
        cum = add_(cum, p);
      }
    return true;
  };
  mean_var[0] = mean_var[1] = 0.0;
  if (!normalise()) {  // every match failed: the reference deletes the feature (monoslam.cpp:1486-1490)
    for (int k = 0; k < K; ++k) keep[k] = 0;
    *left_out = 0;
    return;
  }
  const double thr = div_(prune_threshold, (double)K);  // feature_init_info.cpp:128
  int left = 0;
  for (int k = 0; k < K; ++k) {
    if (prob[k] < thr) {
      keep[k] = 0;
      cumulative[k] = 0.0;
    } else {
      ++left;
    }
  }
  normalise();
  double mean = 0.0, e2 = 0.0;  // feature_init_info.cpp:152-172, scalar lambda
  for (int k = 0; k < K; ++k)
    if (keep[k]) {
      mean = add_(mean, mul_(prob[k], lambda[k]));
      e2 = add_(e2, mul_(prob[k], mul_(lambda[k], lambda[k])));
    }
  mean_var[0] = mean;
  mean_var[1] = sub_(e2, mul_(mean, mean));
  *left_out = left;
}

struct PQuat {
  rd w, x, y, z;
};

// Thread k: particle k of feature blockIdx.x.  Operation order = oracle/models.hpp PartFeatureModel (every matrix
// entry summed over k ascending from 0.0, terms that are exact zeros dropped), never-fused arithmetic.
__global__ void __launch_bounds__(128) particle_predict_kernel(const Sl2Dev d, int s, int Kmax,
                                                               const int *__restrict__ Kf,
                                                               const double *__restrict__ ypi_all,
                                                               const double *__restrict__ Pxy_all,
                                                               const double *__restrict__ Pyy_all,
                                                               const double *__restrict__ lambda_all,
                                                               double *__restrict__ h_out, double *__restrict__ sinv3_out,
                                                               double *__restrict__ detS_out) {
  const int f = blockIdx.x, K = Kf[f];
  const double *ypi = ypi_all + 6 * f, *Pxy = Pxy_all + 78 * f, *Pyy = Pyy_all + 36 * f;
  const double *xv = d.x + (size_t)s * d.ld;
  const double *P = d.P + (size_t)s * d.ld * d.ld;  // Pxx = P(0:13, 0:13), column-major with stride ld
  const int ld = d.ld;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const size_t o = (size_t)f * Kmax + k;
    const rd lam(lambda_all[o]);
    // ---- func_zeroedyi_and_dzeroedyi_by_dxp_and_dzeroedyi_by_dyi, part_feature_model.cpp:80-143 ----------
    const rd dv[3] = {rd(ypi[0]) - rd(xv[0]), rd(ypi[1]) - rd(xv[1]), rd(ypi[2]) - rd(xv[2])};
    const rd hh[3] = {rd(ypi[3]), rd(ypi[4]), rd(ypi[5])};
    PQuat qi;
    {  // Eigen::Quaterniond::inverse(): conjugate / squaredNorm
      const rd w(xv[3]), x(xv[4]), y(xv[5]), z(xv[6]);
      const rd n2 = w * w + x * x + y * y + z * z;
      qi.w = w / n2;
      qi.x = (-x) / n2;
      qi.y = (-y) / n2;
      qi.z = (-z) / n2;
    }
    rd R[3][3];
    {  // Eigen::Quaterniond::toRotationMatrix()
      const rd two(2.0), one(1.0);
      const rd tx = two * qi.x, ty = two * qi.y, tz = two * qi.z;
      const rd twx = tx * qi.w, twy = ty * qi.w, twz = tz * qi.w;
      const rd txx = tx * qi.x, txy = ty * qi.x, txz = tz * qi.x;
      const rd tyy = ty * qi.y, tyz = tz * qi.y, tzz = tz * qi.z;
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
    rd zr[3], zh[3];
    for (int i = 0; i < 3; ++i) {
      rd a(0.0), b(0.0);
      for (int c = 0; c < 3; ++c) {
        a = a + R[i][c] * dv[c];
        b = b + R[i][c] * hh[c];
      }
      zr[i] = a;
      zh[i] = b;
    }
    // dRq_times_a_by_dq(qRW, a) * dqbar_by_dq (feature_model.cpp:187-238, 152-162) for a = d and a = hhat
    rd Dr[3][4], Dh[3][4];
    {
      const rd two(2.0);
      const rd w2 = two * qi.w, x2 = two * qi.x, y2 = two * qi.y, z2 = two * qi.z;
      const rd m0[9] = {w2, -z2, y2, z2, w2, -x2, -y2, x2, w2};
      const rd mx[9] = {x2, y2, z2, y2, -x2, -w2, z2, w2, -x2};
      const rd my[9] = {-y2, x2, w2, x2, y2, z2, -w2, z2, -y2};
      const rd mz[9] = {-z2, -w2, x2, w2, -z2, y2, x2, y2, z2};
      for (int i = 0; i < 3; ++i) {
        rd a0(0.0), a1(0.0), a2(0.0), a3(0.0), b0(0.0), b1(0.0), b2(0.0), b3(0.0);
        for (int c = 0; c < 3; ++c) {
          a0 = a0 + m0[i * 3 + c] * dv[c];
          a1 = a1 + mx[i * 3 + c] * dv[c];
          a2 = a2 + my[i * 3 + c] * dv[c];
          a3 = a3 + mz[i * 3 + c] * dv[c];
          b0 = b0 + m0[i * 3 + c] * hh[c];
          b1 = b1 + mx[i * 3 + c] * hh[c];
          b2 = b2 + my[i * 3 + c] * hh[c];
          b3 = b3 + mz[i * 3 + c] * hh[c];
        }
        Dr[i][0] = a0, Dr[i][1] = -a1, Dr[i][2] = -a2, Dr[i][3] = -a3;
        Dh[i][0] = b0, Dh[i][1] = -b1, Dh[i][2] = -b2, Dh[i][3] = -b3;
      }
    }
    // ---- hLR = zeroedri + lambda zeroedhhati; Camera::Project / ProjectionJacobian (camera.cpp:90-114, 183-215)
    const rd c3[3] = {zr[0] + lam * zh[0], zr[1] + lam * zh[1], zr[2] + lam * zh[2]};
    const rd fku(d.cam[2]), fkv(d.cam[3]), u0(d.cam[4]), v0(d.cam[5]), kd1(d.cam[6]), sd(d.cam[7]);
    const rd one(1.0), two(2.0);
    const rd uc = (-fku) * c3[0] / c3[2];
    const rd vc = (-fkv) * c3[1] / c3[2];
    const rd radius2 = uc * uc + vc * vc;
    const rd factor = rsqrt_(one + two * kd1 * radius2);
    const rd h0 = uc / factor + u0, h1 = vc / factor + v0;
    const rd fku_yz = fku / c3[2], fkv_yz = fkv / c3[2];
    rd du[2][3];
    du[0][0] = -fku_yz;
    du[0][1] = rd(0.0);
    du[0][2] = fku_yz * c3[0] / c3[2];
    du[1][0] = rd(0.0);
    du[1][1] = -fkv_yz;
    du[1][2] = fkv_yz * c3[1] / c3[2];
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
    rd J[2][3], Jl[2][3];  // dhpi_by_dhLRi and its product with lambda (the lambda block of dhLRi_by_dzeroedyi)
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 3; ++j) {
        rd a(0.0);
        for (int c = 0; c < 2; ++c) a = a + dh[i][c] * du[c][j];
        J[i][j] = a;
        Jl[i][j] = a * lam;
      }
    // ---- dhpi_by_dxp (2x7) and dhpi_by_dyi (2x6), part_feature_model.cpp:262-264 --------------------------
    rd dxp[2][7], dy[2][6];
    for (int i = 0; i < 2; ++i) {
      for (int j = 0; j < 3; ++j) {
        rd a(0.0), b(0.0), c(0.0);
        for (int q = 0; q < 3; ++q) {
          a = a + J[i][q] * (R[q][j] * rd(-1.0));
          b = b + J[i][q] * R[q][j];
          c = c + Jl[i][q] * R[q][j];
        }
        dxp[i][j] = a;
        dy[i][j] = b;
        dy[i][3 + j] = c;
      }
      for (int j = 0; j < 4; ++j) {
        rd a(0.0);
        for (int q = 0; q < 3; ++q) a = a + J[i][q] * Dr[q][j];
        for (int q = 0; q < 3; ++q) a = a + Jl[i][q] * Dh[q][j];
        dxp[i][3 + j] = a;
      }
    }
    // ---- Camera::MeasurementNoise (camera.cpp:282-300), FeatureModel::func_Si (feature_model.cpp:99-116) ------
    const rd ddx = h0 - u0, ddy = h1 - v0;
    const rd distance = rsqrt_(ddx * ddx + ddy * ddy);
    const rd max_distance = rsqrt_(u0 * u0 + v0 * v0);
    const rd ratio = distance / max_distance;
    const rd sd_use = sd * (one + ratio);
    const rd var = one * (sd_use * sd_use);
    rd A[2][7], Bm[2][6], Cm[2][6];
    for (int r = 0; r < 2; ++r) {
      for (int j = 0; j < 7; ++j) {
        rd a(0.0);
        for (int q = 0; q < 7; ++q) a = a + dxp[r][q] * rd(P[q + (size_t)ld * j]);
        A[r][j] = a;
      }
      for (int j = 0; j < 6; ++j) {
        rd a(0.0), b(0.0);
        for (int q = 0; q < 7; ++q) a = a + dxp[r][q] * rd(Pxy[q + 13 * j]);
        for (int q = 0; q < 6; ++q) b = b + dy[r][q] * rd(Pyy[q + 6 * j]);
        Bm[r][j] = a;
        Cm[r][j] = b;
      }
    }
    rd S[2][2];
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 2; ++c) {
        rd s1(0.0), t1(0.0), t1t(0.0), s4(0.0);
        for (int q = 0; q < 7; ++q) s1 = s1 + A[r][q] * dxp[c][q];
        for (int q = 0; q < 6; ++q) t1 = t1 + Bm[r][q] * dy[c][q];
        for (int q = 0; q < 6; ++q) t1t = t1t + Bm[c][q] * dy[r][q];
        for (int q = 0; q < 6; ++q) s4 = s4 + Cm[r][q] * dy[c][q];
        rd v = rd(0.0) + s1;
        v = v + t1;
        v = v + t1t;
        v = v + s4;
        v = v + (r == c ? var : rd(0.0));
        S[r][c] = v;
      }
    // ---- Particle::set_S (feature_init_info.cpp:57-65): LLT, L^-1 in closed form, L^-T L^-1, determinant -------
    const rd l00 = rsqrt_(S[0][0]);
    const rd l10 = S[1][0] / l00;
    const rd l11 = rsqrt_(S[1][1] - l10 * l10);
    const rd x00 = one / l00;
    const rd x10 = (rd(0.0) - l10 * x00) / l11;
    const rd x11 = one / l11;
    h_out[2 * o] = h0.v;
    h_out[2 * o + 1] = h1.v;
    sinv3_out[3 * o] = (x00 * x00 + x10 * x10).v;
    sinv3_out[3 * o + 1] = (x10 * x11).v;
    sinv3_out[3 * o + 2] = (x11 * x11).v;
    detS_out[o] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]).v;
  }
}

}  // namespace

// F features, Kmax = stride between features in every per-particle array, K_dev[f] particles used
cudaError_t sl2_launch_particles(int F, int Kmax, const int *K_dev, const double *h, const double *sinv3,
                                 const double *detS, const double *lambda, const int *z_uv, const uint8_t *found,
                                 double prune_threshold, double *prob, uint8_t *keep, double *cumulative,
                                 double *mean_var, int *left_out, cudaStream_t st) {
  if (F <= 0) return cudaSuccess;
  particle_kernel<<<F, 128, 0, st>>>(Kmax, K_dev, h, sinv3, detS, lambda, z_uv, found, prune_threshold, prob, keep,
                                     cumulative, mean_var, left_out);
  return cudaGetLastError();
}

cudaError_t sl2_launch_particle_predict(const Sl2Dev &d, int s, int F, int Kmax, const int *K_dev,
                                        const double *ypi, const double *Pxy, const double *Pyy,
                                        const double *lambda, double *h, double *sinv3, double *detS,
                                        cudaStream_t st) {
  if (F <= 0) return cudaSuccess;
  particle_predict_kernel<<<F, 128, 0, st>>>(d, s, Kmax, K_dev, ypi, Pxy, Pyy, lambda, h, sinv3, detS);
  return cudaGetLastError();
}
