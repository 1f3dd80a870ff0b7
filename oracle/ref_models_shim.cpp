// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points into the REFERENCE'S OWN model classes
// (MotionModel, Camera, FullFeatureModel), compiled unmodified from /root/reference/scenelib2 against the
// arithmetic stand-in of oracle/stubs_arith (see its header for what is and is not pinned).  Signatures
// mirror the oracle's orc_* functions so tests can compare the two directly.
#include <cmath>
#include <cstring>
#include <vector>

#include "camera.h"
#include "feature_init_info.h"
#include "full_feature_model.h"
#include "motion_model.h"

using namespace SceneLib2;

namespace {
Eigen::VectorXd vec(const double *p, int n) {
  Eigen::VectorXd v(n);
  for (int i = 0; i < n; ++i) v(i) = p[i];
  return v;
}
Eigen::MatrixXd mat(const double *p, int r, int c) {  // column-major in, column-major stored
  Eigen::MatrixXd m(r, c);
  std::memcpy(m.data(), p, sizeof(double) * r * c);
  return m;
}
void out(const Eigen::MatrixXd &m, double *p) { std::memcpy(p, m.data(), sizeof(double) * m.size()); }
void set_camera(Camera &cam, const double *c8) {
  cam.SetCameraParameters((int)c8[0], (int)c8[1], c8[2], c8[3], c8[4], c8[5], c8[6], (int)c8[7]);
}
}  // namespace

extern "C" {

// motion_model.cpp:84-217
void ref_motion(const double *xv, const double *u, double delta_t, double *fv, double *F, double *Q) {
  MotionModel mm;
  const Eigen::VectorXd x = vec(xv, 13), uu = vec(u, 3);
  mm.func_fv_and_dfv_by_dxv(x, uu, delta_t);
  mm.func_Q(x, uu, delta_t);
  out(mm.fvRES_, fv);
  out(mm.dfv_by_dxvRES_, F);
  out(mm.QxRES_, Q);
}

// motion_model.cpp:237-263
void ref_dxvnorm_by_dxv(const double *xv, double *J, double *xvnorm) {
  MotionModel mm;
  mm.func_xvnorm_and_dxvnorm_by_dxv(vec(xv, 13));
  out(mm.dxvnorm_by_dxvRES_, J);
  out(mm.xvnormRES_, xvnorm);
}

// the body of MonoSLAM::predict_single_feature_measurements (monoslam.cpp:289-308) on the reference's
// own model objects: h_i, dh_i/dx_v, dh_i/dy_i, R_i, S_i
void ref_predict_feature(const double *cam8, const double *xv, const double *y, const double *Pxx,
                         const double *Pxy, const double *Pyy, double *h, double *dh_by_dxv,
                         double *dh_by_dy, double *R, double *S) {
  Camera cam;
  set_camera(cam, cam8);
  MotionModel mm;
  FullFeatureModel fm(2, 3, 3, &cam, &mm);
  const Eigen::VectorXd x = vec(xv, 13);
  mm.func_xp(x);
  const Eigen::VectorXd xp = mm.xpRES_;
  fm.func_hi_and_dhi_by_dxp_and_dhi_by_dyi(vec(y, 3), xp);
  mm.func_dxp_by_dxv(x);
  const Eigen::MatrixXd dh_dxv = fm.dhi_by_dxpRES_ * mm.dxp_by_dxvRES_;
  fm.func_Ri(fm.hiRES_);
  fm.func_Si(mat(Pxx, 13, 13), mat(Pxy, 13, 3), mat(Pyy, 3, 3), dh_dxv, fm.dhi_by_dyiRES_, fm.RiRES_);
  out(fm.hiRES_, h);
  out(dh_dxv, dh_by_dxv);
  out(fm.dhi_by_dyiRES_, dh_by_dy);
  out(fm.RiRES_, R);
  out(fm.SiRES_, S);
}

// full_feature_model.cpp:103-170
int ref_visibility_test(const double *cam8, const double *xp, const double *y, const double *xp_org,
                        const double *h) {
  Camera cam;
  set_camera(cam, cam8);
  MotionModel mm;
  FullFeatureModel fm(2, 3, 3, &cam, &mm);
  return fm.visibility_test(vec(xp, 7), vec(y, 3), vec(xp_org, 7), vec(h, 2));
}

// camera.cpp:90-170
void ref_project(const double *cam8, const double *c3, double *h2, double *J23) {
  Camera cam;
  set_camera(cam, cam8);
  const Eigen::Vector2d hh = cam.Project(Eigen::Vector3d(c3[0], c3[1], c3[2]));
  out(hh, h2);
  out(cam.ProjectionJacobian(), J23);
}
void ref_unproject(const double *cam8, const double *h2, double *c3, double *J32) {
  Camera cam;
  set_camera(cam, cam8);
  const Eigen::Vector3d cc = cam.Unproject(Eigen::Vector2d(h2[0], h2[1]));
  out(cc, c3);
  out(cam.UnprojectionJacobian(), J32);
}

// Particle::set_S (feature_init_info.cpp:55-63): S (2x2, column-major) -> SInv (2x2), det S.  The same
// LLT -> matrixL -> inverse -> L^-T L^-1 sequence as MonoSLAM::measure_feature (monoslam.cpp:371-374).
void ref_particle_set_S(const double *S4, double *Sinv4, double *detS) {
  Eigen::VectorXd l(1);
  Particle p(l, 1.0, 2);
  p.set_S(mat(S4, 2, 2));
  out(p.m_SInv_, Sinv4);
  *detS = p.m_detS_;
}

// One FeatureInitInfo through the particle branch of MonoSLAM::update_partially_initialised_feature_
// probabilities (monoslam.cpp:1447-1493).  That loop lives in monoslam.cpp (not compilable here) and is
// transcribed below; normalise_particle_vector_and_calculate_cumulative, prune_particle_vector and
// calculate_mean_and_covariance are the reference's own feature_init_info.cpp.
// Outputs are per INPUT particle: keep[k] = 0 for pruned ones.
int ref_particle_update(int K, const double *h, const double *Sinv3, const double *detS, const double *lambda,
                        const int *z_uv, const unsigned char *found, double prune_probability_threshold,
                        double *prob, unsigned char *keep, double *cumulative, double *mean_var) {
  FeatureInitInfo feat(nullptr, 1, 2);
  for (int k = 0; k < K; ++k) {
    Eigen::VectorXd l(1);
    l(0) = lambda[k];
    feat.add_particle(l, prob[k]);
    Particle &p = feat.particle_vector_.back();
    p.m_h_(0) = h[2 * k];
    p.m_h_(1) = h[2 * k + 1];
    p.m_SInv_(0, 0) = Sinv3[3 * k];
    p.m_SInv_(0, 1) = p.m_SInv_(1, 0) = Sinv3[3 * k + 1];
    p.m_SInv_(1, 1) = Sinv3[3 * k + 2];
    p.m_detS_ = detS[k];
    p.m_z_(0) = z_uv[2 * k];
    p.m_z_(1) = z_uv[2 * k + 1];
    p.m_successful_measurement_flag_ = found[k] != 0;
    p.cumulative_probability_ = -(double)(k + 1);  // tag: survives erase(), identifies the input particle
  }
  for (vector<Particle>::iterator it = feat.particle_vector_.begin(); it != feat.particle_vector_.end(); ++it) {
    double likelihood;
    if (it->m_successful_measurement_flag_ == true) {
      Eigen::VectorXd nu = it->m_z_ - it->m_h_;
      Eigen::VectorXd SInv_times_nu = it->m_SInv_ * nu;
      double nuT_Sinv_nu = nu.dot(SInv_times_nu);
      double pux = (1.0 / (sqrt(2.0 * M_PI * it->m_detS_))) * exp(-0.5 * nuT_Sinv_nu);
      likelihood = pux;
    } else {
      likelihood = 0.0;
    }
    it->probability_ = it->probability_ * likelihood;
  }
  for (int k = 0; k < K; ++k) {
    keep[k] = 0;
    cumulative[k] = 0.0;
  }
  mean_var[0] = mean_var[1] = 0.0;
  // tags are overwritten by normalise...(); remember the input index in lambda order instead
  std::vector<double> lam_in(lambda, lambda + K);
  if (!feat.normalise_particle_vector_and_calculate_cumulative()) {
    for (int k = 0; k < K; ++k) prob[k] = feat.particle_vector_[k].probability_;
    return 0;
  }
  feat.prune_particle_vector(prune_probability_threshold);
  feat.calculate_mean_and_covariance();
  for (int k = 0; k < K; ++k) prob[k] = 0.0;
  size_t j = 0;  // survivors keep their order: match them to the inputs by position of equal lambda
  for (int k = 0; k < K && j < feat.particle_vector_.size(); ++k) {
    if (feat.particle_vector_[j].lambda_(0) == lam_in[k]) {
      keep[k] = 1;
      prob[k] = feat.particle_vector_[j].probability_;
      cumulative[k] = feat.particle_vector_[j].cumulative_probability_;
      ++j;
    }
  }
  mean_var[0] = feat.mean_(0);
  mean_var[1] = feat.covariance_(0, 0);
  return (int)feat.particle_vector_.size();
}

}  // extern "C"
