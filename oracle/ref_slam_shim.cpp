// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points into the REFERENCE'S OWN MonoSLAM tracking step:
// monoslam.cpp, kalman.cpp, feature.cpp, part_feature_model.cpp, the model files, support/*.cpp and improc/*.cpp
// are compiled UNMODIFIED where they lie under /root/reference against the stand-ins of oracle/stubs_arith
// (matrix class with plain-loop arithmetic instead of Eigen, pixel container instead of OpenCV, "key = value;"
// parser instead of Pangolin; GUI and frame grabber reduced to the empty definitions below).
// So GoOneStep, KalmanFilterPredict/Update, auto_select_n_features, make_measurements, measure_feature,
// elliptical_search, construct_total_*, fill_*, normalise_state, delete_bad_features ... executed by the
// functions below ARE the reference's code; only Eigen's internal summation order is not reproduced.
// Signatures mirror the oracle's orc_slam_* API (oracle/sl2_oracle.h).
#include <cstdint>
#include <cstring>
#include <string>

#include "kalman.h"
#include "monoslam.h"

namespace SceneLib2 {
// the two collaborators MonoSLAM::Init creates that are out of scope here (monoslam.cpp:1961-1963)
GraphicTool::GraphicTool(MonoSLAM *)
    : kQR0_(0.0, 0.0, 1.0, 0.0), kMoveClippingPlaneFactor_(0.0), kSemiInfiniteLineLength_(0.0),
      kCovariancesNumberOfSigma_(0.0), kDrawNOverlappingEllipses_(0) {}
GraphicTool::~GraphicTool() {}
FrameGrabber::FrameGrabber() : file_grabber_(nullptr), usb_cam_grabber_(nullptr) {}
FrameGrabber::~FrameGrabber() {}
void FrameGrabber::Init(const std::string &, const bool) {}
}  // namespace SceneLib2

using namespace SceneLib2;

extern "C" {

// MonoSLAM::Init (monoslam.cpp:1576-1969): camera, models, initial state and the four known features of the cfg
void *ref_slam_create(const char *cfg_path) {
  MonoSLAM *m = new MonoSLAM();
  m->Init(cfg_path);
  return m;
}
void ref_slam_destroy(void *p) { delete static_cast<MonoSLAM *>(p); }

// MonoSLAM::AddNewKnownFeature (monoslam.cpp:1278-1289); the template is read by cv::imread (feature.cpp:119)
void ref_slam_add_feature(void *p, const double *y, const double *xp_org, const char *patch_path) {
  MonoSLAM *m = static_cast<MonoSLAM *>(p);
  Eigen::VectorXd yy(3), xp(7);
  for (int i = 0; i < 3; ++i) yy(i) = y[i];
  for (int i = 0; i < 7; ++i) xp(i) = xp_org[i];
  m->AddNewKnownFeature(yy, xp, patch_path);
}
void ref_slam_set_params(void *p, int n_select, int min_attempts, double match_fraction) {
  MonoSLAM *m = static_cast<MonoSLAM *>(p);
  if (n_select >= 0) m->kNumberOfFeaturesToSelect_ = n_select;
  if (min_attempts >= 0) m->minimum_attempted_measurements_of_feature_ = min_attempts;
  if (match_fraction >= 0.0) m->successful_match_fraction_ = match_fraction;
}
int32_t ref_slam_num_features(void *p) { return (int32_t) static_cast<MonoSLAM *>(p)->feature_list_.size(); }
int32_t ref_slam_state_size(void *p) { return static_cast<MonoSLAM *>(p)->total_state_size_; }

// fill_states / fill_covariances (monoslam.cpp:574-614); P column-major n x n
void ref_slam_set_state(void *p, const double *x, const double *P) {
  MonoSLAM *m = static_cast<MonoSLAM *>(p);
  const int n = m->total_state_size_;
  Eigen::VectorXd V(n);
  Eigen::MatrixXd M(n, n);
  std::memcpy(V.data(), x, sizeof(double) * n);
  std::memcpy(M.data(), P, sizeof(double) * n * n);
  m->fill_states(V);
  m->fill_covariances(M);
}
// construct_total_state / construct_total_covariance (monoslam.cpp:501-546)
void ref_slam_get_state(void *p, double *x, double *P) {
  MonoSLAM *m = static_cast<MonoSLAM *>(p);
  const int n = m->total_state_size_;
  Eigen::VectorXd V(n);
  Eigen::MatrixXd M(n, n);
  V.setZero();
  M.setZero();
  m->construct_total_state(V);
  m->construct_total_covariance(M);
  std::memcpy(x, V.data(), sizeof(double) * n);
  std::memcpy(P, M.data(), sizeof(double) * n * n);
}
// MonoSLAM::GoOneStep (monoslam.cpp:108-180), tracking only
void ref_slam_step(void *p, const uint8_t *frame, int32_t width, int32_t height) {
  MonoSLAM *m = static_cast<MonoSLAM *>(p);
  cv::Mat f(height, width, CV_8UC1, const_cast<uint8_t *>(frame));
  m->GoOneStep(f, false, false);
}
void ref_slam_get_features(void *p, int32_t *label, double *h, double *z, double *S, uint8_t *flags,
                           int32_t *attempted, int32_t *successful, int32_t *select_rank) {
  MonoSLAM *m = static_cast<MonoSLAM *>(p);
  int i = 0;
  for (Feature *f : m->feature_list_) {
    label[i] = f->label_;
    for (int k = 0; k < 2; ++k) {
      h[2 * i + k] = f->h_.size() == 2 ? f->h_(k) : 0.0;
      z[2 * i + k] = f->z_.size() == 2 ? f->z_(k) : 0.0;
    }
    for (int k = 0; k < 4; ++k) S[4 * i + k] = f->S_.size() == 4 ? f->S_.data()[k] : 0.0;
    // Feature::Initialise (feature.cpp:150-160) leaves successful_measurement_flag_ uninitialised; it only has
    // a defined value once the feature has been measured at least once
    const bool succ = f->attempted_measurements_of_feature_ > 0 && f->successful_measurement_flag_;
    flags[i] = (f->selected_flag_ ? 1 : 0) | (succ ? 2 : 0);
    attempted[i] = f->attempted_measurements_of_feature_;
    successful[i] = f->successful_measurements_of_feature_;
    select_rank[i] = -1;
    int r = 0;
    for (Feature *s : m->selected_feature_list_) {
      if (s == f) select_rank[i] = r;
      ++r;
    }
    ++i;
  }
}

}  // extern "C"
