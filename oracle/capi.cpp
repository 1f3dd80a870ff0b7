// ORACLE — TEST INFRASTRUCTURE ONLY (see dense.hpp header).  extern "C" surface of liboracle.so.
#include <chrono>
#include <pthread.h>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "sl2_oracle.h"
#include "slam.hpp"
#include "particles.hpp"

using namespace sl2o;

struct orc_slam {
  Slam s;
  explicit orc_slam(const SlamConfig &c) : s(c) {}
};

static SlamConfig to_cfg(const orc_config *c) {
  SlamConfig k;
  k.width = c->width;
  k.height = c->height;
  k.fku = c->fku;
  k.fkv = c->fkv;
  k.u0 = c->u0;
  k.v0 = c->v0;
  k.kd1 = c->kd1;
  k.sd = c->sd;
  k.delta_t = c->delta_t;
  k.number_of_features_to_select = c->number_of_features_to_select;
  k.boxsize = c->boxsize;
  for (int i = 0; i < 3; ++i) k.search_override[i] = c->search_override[i];
  k.minimum_attempted_measurements_of_feature = c->minimum_attempted_measurements_of_feature;
  k.successful_match_fraction = c->successful_match_fraction;
  return k;
}

static Camera cam_from8(const double *c) {
  Camera cam;
  cam.width = (int)c[0];
  cam.height = (int)c[1];
  cam.fku = c[2];
  cam.fkv = c[3];
  cam.u0 = c[4];
  cam.v0 = c[5];
  cam.kd1 = c[6];
  cam.sd = c[7];
  return cam;
}

static Mat mat_from(const double *p, int r, int c) {
  Mat m(r, c);
  std::memcpy(m.a.data(), p, sizeof(double) * (size_t)r * c);
  return m;
}

extern "C" {

double orc_correlate2_warning(const uint8_t *patch, int32_t patch_width, int32_t x0lim,
                              int32_t y0lim, const uint8_t *image, int32_t image_width, int32_t x1,
                              int32_t y1, double *sd0, double *sd1) {
  return correlate2_warning(0, 0, x0lim, y0lim, x1, y1, patch, patch_width, image, image_width,
                            sd0, sd1);
}

int32_t orc_elliptical_search(const uint8_t *image, int32_t width, int32_t height,
                              const uint8_t *patch, int32_t boxsize, const double *centre,
                              const double *P, int32_t *u, int32_t *v, double *best) {
  int uu = *u, vv = *v;
  const bool ok =
      elliptical_search(image, width, height, patch, centre, P[0], P[1], P[2], &uu, &vv, boxsize, best);
  *u = uu;
  *v = vv;
  return ok ? 1 : 0;
}

void orc_elliptical_search_batch(const uint8_t *image, int32_t width, int32_t height,
                                 const uint8_t *patches, int32_t boxsize, int32_t n,
                                 const double *centres, const double *P, int32_t *u, int32_t *v,
                                 uint8_t *found, double *best) {
  for (int i = 0; i < n; ++i) {
    int uu = -1, vv = -1;
    double b = 0;
    const bool ok = elliptical_search(image, width, height, patches + (size_t)i * boxsize * boxsize,
                                      centres + 2 * i, P[3 * i], P[3 * i + 1], P[3 * i + 2], &uu,
                                      &vv, boxsize, &b);
    u[i] = uu;
    v[i] = vv;
    found[i] = ok ? 1 : 0;
    if (best) best[i] = b;
  }
}

void orc_search_box(int32_t width, int32_t height, int32_t BOXSIZE, const double *centre,
                    const double *P, int32_t *box6) {
  const double P00 = P[0], P01 = P[1], P11 = P[2];
  const int halfwidth = (int)(kNoSigma / std::sqrt(P00 - P01 * P01 / P11));
  const int halfheight = (int)(kNoSigma / std::sqrt(P11 - P01 * P01 / P00));
  const int ucentre = int(centre[0] + 0.5);
  const int vcentre = int(centre[1] + 0.5);
  int urelstart = -halfwidth, urelfinish = halfwidth;
  int vrelstart = -halfheight, vrelfinish = halfheight;
  if (ucentre + urelstart - (BOXSIZE - 1) / 2 < 0) urelstart = (BOXSIZE - 1) / 2 - ucentre;
  if (ucentre + urelfinish - (BOXSIZE - 1) / 2 > width - BOXSIZE)
    urelfinish = width - BOXSIZE - ucentre + (BOXSIZE - 1) / 2;
  if (vcentre + vrelstart - (BOXSIZE - 1) / 2 < 0) vrelstart = (BOXSIZE - 1) / 2 - vcentre;
  if (vcentre + vrelfinish - (BOXSIZE - 1) / 2 > height - BOXSIZE)
    vrelfinish = height - BOXSIZE - vcentre + (BOXSIZE - 1) / 2;
  box6[0] = urelstart;
  box6[1] = urelfinish;
  box6[2] = vrelstart;
  box6[3] = vrelfinish;
  box6[4] = ucentre;
  box6[5] = vcentre;
}

void orc_score_map(const uint8_t *image, int32_t width, int32_t height, const uint8_t *patch,
                   int32_t B, const double *centre, const double *P, double *corr,
                   double *sd_image, uint8_t *inside) {
  int32_t box[6];
  orc_search_box(width, height, B, centre, P, box);
  const int nv = box[3] - box[2] + 1;
  for (int urel = box[0]; urel <= box[1]; ++urel)
    for (int vrel = box[2]; vrel <= box[3]; ++vrel) {
      const size_t idx = (size_t)(urel - box[0]) * nv + (vrel - box[2]);
      double sd0, sd1;
      corr[idx] = correlate2_warning(0, 0, B, B, box[4] + urel - (B - 1) / 2,
                                     box[5] + vrel - (B - 1) / 2, patch, B, image, width, &sd0, &sd1);
      sd_image[idx] = sd1;
      inside[idx] =
          (P[0] * urel * urel + 2 * P[1] * urel * vrel + P[2] * vrel * vrel < kNoSigma * kNoSigma)
              ? 1
              : 0;
    }
}

void orc_puinv_from_S(const double *S, double *P3) {
  double P[4];
  puinv_from_S(S, P);
  P3[0] = P[0];
  P3[1] = P[2];
  P3[2] = P[3];
}

void orc_smoe_search(const uint8_t *image, int32_t width, int32_t height, const uint8_t *patch,
                     int32_t boxsize, int32_t K, const double *PuInv3, const double *centres,
                     int32_t *res_u, int32_t *res_v, uint8_t *res_flag, double *res_best) {
  smoe_search(image, width, height, patch, boxsize, K, PuInv3, centres, res_u, res_v, res_flag,
              res_best);
}

void orc_find_best_patch(const uint8_t *image, int32_t width, int32_t height, int32_t boxsize,
                         const int32_t *r, int32_t *ubest, int32_t *vbest, double *evbest) {
  int u = *ubest, v = *vbest;
  find_best_patch_inside_region(image, width, height, &u, &v, evbest, boxsize, r[0], r[1], r[2], r[3]);
  *ubest = u;
  *vbest = v;
}

int32_t orc_particle_update(int32_t K, const double *h, const double *Sinv3, const double *detS,
                            const double *lambda, const int32_t *z_uv, const uint8_t *found,
                            double prune_probability_threshold, double *prob, uint8_t *keep,
                            double *cumulative, double *mean_var) {
  return particle_update(K, h, Sinv3, detS, lambda, z_uv, found, prune_probability_threshold, prob, keep,
                         cumulative, mean_var);
}

void orc_motion(const double *xv, const double *u, double delta_t, double *fv, double *F,
                double *Q) {
  Mat Fm;
  MotionModel::fv_and_dfv_by_dxv(xv, u, delta_t, fv, Fm);
  const Mat Qm = MotionModel::Q(xv, delta_t);
  std::memcpy(F, Fm.a.data(), sizeof(double) * 169);
  std::memcpy(Q, Qm.a.data(), sizeof(double) * 169);
}

void orc_dxvnorm_by_dxv(const double *xv, double *J) {
  const Mat Jm = MotionModel::dxvnorm_by_dxv(xv);
  std::memcpy(J, Jm.a.data(), sizeof(double) * 169);
}

void orc_predict_feature(const double *cam8, const double *xv, const double *y, const double *Pxx,
                         const double *Pxy, const double *Pyy, double *h, double *dh_by_dxv,
                         double *dh_by_dy, double *R, double *S) {
  Camera cam = cam_from8(cam8);
  FeaturePrediction p;
  FullFeatureModel::predict(cam, xv, y, mat_from(Pxx, 13, 13), mat_from(Pxy, 13, 3),
                            mat_from(Pyy, 3, 3), p);
  h[0] = p.h[0];
  h[1] = p.h[1];
  std::memcpy(dh_by_dxv, p.dh_by_dxv.a.data(), sizeof(double) * 26);
  std::memcpy(dh_by_dy, p.dh_by_dy.a.data(), sizeof(double) * 6);
  std::memcpy(R, p.R.a.data(), sizeof(double) * 4);
  std::memcpy(S, p.S.a.data(), sizeof(double) * 4);
}

void orc_predict_particles(const double *cam8, const double *xv, const double *ypi, int32_t K,
                           const double *lambda, const double *Pxx, const double *Pxy, const double *Pyy,
                           double *h, double *S4, double *Sinv3, double *detS) {
  Camera cam = cam_from8(cam8);
  const Mat mPxx = mat_from(Pxx, 13, 13), mPxy = mat_from(Pxy, 13, 6), mPyy = mat_from(Pyy, 6, 6);
  for (int k = 0; k < K; ++k) {
    ParticlePrediction p;
    PartFeatureModel::predict_particle(cam, xv, ypi, lambda[k], mPxx, mPxy, mPyy, p);
    h[2 * k] = p.h[0];
    h[2 * k + 1] = p.h[1];
    if (S4) std::memcpy(S4 + 4 * k, p.S.a.data(), sizeof(double) * 4);
    Sinv3[3 * k] = p.Sinv[0];
    Sinv3[3 * k + 1] = p.Sinv[1];
    Sinv3[3 * k + 2] = p.Sinv[3];
    detS[k] = p.detS;
  }
}

int32_t orc_visibility_test(const double *cam8, const double *xp, const double *y,
                            const double *xp_org, const double *h) {
  const Camera cam = cam_from8(cam8);
  return FullFeatureModel::visibility_test(cam, xp, y, xp_org, h);
}

void orc_kalman_update_dense(int32_t n, int32_t m, double *x, double *P, const double *H,
                             const double *R, const double *nu) {
  Vec xv(x, x + n), nuv(nu, nu + m);
  Mat Pm = mat_from(P, n, n);
  Slam::kalman_update_dense(xv, Pm, mat_from(H, m, n), mat_from(R, m, m), nuv);
  std::memcpy(x, xv.data(), sizeof(double) * n);
  std::memcpy(P, Pm.a.data(), sizeof(double) * (size_t)n * n);
}

orc_slam *orc_slam_create(const orc_config *cfg) { return new orc_slam(to_cfg(cfg)); }
void orc_slam_destroy(orc_slam *s) { delete s; }
void orc_slam_add_feature(orc_slam *s, const double *y, const double *xp_org,
                          const uint8_t *patch) {
  s->s.add_known_feature(y, xp_org, patch);
}
int32_t orc_slam_num_features(const orc_slam *s) { return (int32_t)s->s.feature_list.size(); }
int32_t orc_slam_state_size(const orc_slam *s) { return s->s.total_state_size; }
void orc_slam_set_state(orc_slam *s, const double *x, const double *P) {
  const int n = s->s.total_state_size;
  s->s.fill_states(Vec(x, x + n));
  s->s.fill_covariances(mat_from(P, n, n));
}
void orc_slam_get_state(const orc_slam *s, double *x, double *P) {
  const int n = s->s.total_state_size;
  Vec xv((size_t)n, 0.0);
  s->s.construct_total_state(xv);
  std::memcpy(x, xv.data(), sizeof(double) * n);
  const Mat Pm = s->s.dense_P();
  std::memcpy(P, Pm.a.data(), sizeof(double) * (size_t)n * n);
}
void orc_slam_step(orc_slam *s, const uint8_t *frame) { s->s.go_one_step(frame); }
void orc_slam_predict(orc_slam *s) {
  const double u[3] = {0, 0, 0};
  s->s.kalman_predict(u);
}
int32_t orc_slam_select(orc_slam *s) {
  s->s.number_of_visible_features =
      s->s.auto_select_n_features(s->s.cfg.number_of_features_to_select);
  return s->s.number_of_visible_features;
}
int32_t orc_slam_measure(orc_slam *s, const uint8_t *frame) {
  return s->s.make_measurements(frame);
}
void orc_slam_update(orc_slam *s) {
  if (!s->s.selected_feature_list.empty() && s->s.successful_measurement_vector_size != 0)
    s->s.kalman_update();
}
void orc_slam_normalise(orc_slam *s) { s->s.normalise_state(); }
void orc_slam_finish(orc_slam *s) {
  s->s.delete_bad_features();
  Mat P = s->s.dense_P();
  const Mat PT = transpose(P);
  for (size_t i = 0; i < P.a.size(); ++i) P.a[i] = P.a[i] * 0.5 + PT.a[i] * 0.5;
  s->s.fill_covariances(P);
}
void orc_slam_get_features(const orc_slam *s, int32_t *label, double *h, double *z, double *S,
                           uint8_t *flags, int32_t *attempted, int32_t *successful,
                           int32_t *select_rank) {
  const auto &fl = s->s.feature_list;
  for (size_t i = 0; i < fl.size(); ++i) {
    const Feature &f = *fl[i];
    label[i] = f.label;
    h[2 * i] = f.h[0];
    h[2 * i + 1] = f.h[1];
    z[2 * i] = f.z[0];
    z[2 * i + 1] = f.z[1];
    for (int k = 0; k < 4; ++k) S[4 * i + k] = f.S.a.size() == 4 ? f.S.a[k] : 0.0;
    flags[i] = (uint8_t)((f.selected_flag ? 1 : 0) | (f.successful_measurement_flag ? 2 : 0));
    attempted[i] = f.attempted_measurements_of_feature;
    successful[i] = f.successful_measurements_of_feature;
    select_rank[i] = -1;
  }
  for (size_t r = 0; r < s->s.selected_feature_list.size(); ++r)
    select_rank[s->s.selected_feature_list[r]->position_in_list] = (int32_t)r;
}

// CPUs this process may really use: the scheduler affinity mask capped by the cgroup CPU quota
// (std::thread::hardware_concurrency() reports the machine, not the container: round 1's "128 cores").
static std::vector<int> usable_cpu_list() {
  std::vector<int> cpus;
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0)
    for (int c = 0; c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &set)) cpus.push_back(c);
  if (cpus.empty()) cpus.push_back(0);
  double quota = -1.0;
  if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2: "<quota|max> <period>"
    char q[64];
    long long period = 0;
    if (fscanf(f, "%63s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) quota = atof(q) / (double)period;
    fclose(f);
  } else if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {  // cgroup v1
    long long qv = -1, pv = 0;
    if (fscanf(g, "%lld", &qv) != 1) qv = -1;
    fclose(g);
    if (FILE *h = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
      if (fscanf(h, "%lld", &pv) != 1) pv = 0;
      fclose(h);
    }
    if (qv > 0 && pv > 0) quota = (double)qv / (double)pv;
  }
  if (quota > 0.0) {
    const size_t cap = (size_t)(quota + 0.999);
    if (cap >= 1 && cap < cpus.size()) cpus.resize(cap);
  }
  return cpus;
}

double orc_slam_run_pinned(orc_slam **slams, int32_t nslam, const uint8_t *const *frames, int32_t nframes,
                           int32_t nsteps, int32_t nthreads, int32_t pin, double *gather_scatter_seconds) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nslam) nthreads = nslam;
  const std::vector<int> cpus = usable_cpu_list();
  for (int i = 0; i < nslam; ++i) slams[i]->s.gather_scatter_seconds = 0.0;
  auto worker = [&](int t) {
    if (pin) {  // one stream per core: thread t stays on the t-th usable CPU
      cpu_set_t set;
      CPU_ZERO(&set);
      CPU_SET(cpus[(size_t)t % cpus.size()], &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    for (int i = t; i < nslam; i += nthreads) {
      Slam &s = slams[i]->s;
      const size_t fsz = (size_t)s.cfg.width * s.cfg.height;
      for (int k = 0; k < nsteps; ++k) s.go_one_step(frames[i] + fsz * (size_t)(k % nframes));
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
  {  // the calling thread works too, and gets its affinity back afterwards
    cpu_set_t old;
    const bool have = pin && pthread_getaffinity_np(pthread_self(), sizeof(old), &old) == 0;
    worker(0);
    if (have) pthread_setaffinity_np(pthread_self(), sizeof(old), &old);
  }
  for (auto &x : th) x.join();
  const auto t1 = std::chrono::steady_clock::now();
  if (gather_scatter_seconds) {
    double g = 0.0;
    for (int i = 0; i < nslam; ++i) g += slams[i]->s.gather_scatter_seconds;
    *gather_scatter_seconds = g;  // summed over streams (= over threads when one stream runs per thread)
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

double orc_slam_run(orc_slam **slams, int32_t nslam, const uint8_t *const *frames, int32_t nframes,
                    int32_t nsteps, int32_t nthreads) {
  return orc_slam_run_pinned(slams, nslam, frames, nframes, nsteps, nthreads, 0, nullptr);
}

int32_t orc_hardware_threads(void) { return (int32_t)std::thread::hardware_concurrency(); }
int32_t orc_usable_cpus(void) { return (int32_t)usable_cpu_list().size(); }

}  // extern "C"
