/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * C ABI of the CPU restatement (liboracle.so), loaded through ctypes by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 * Nothing under scenelib2_b200/ may include or link this.
 * All matrices are column-major FP64; images are row-major u8 with stride = width.
 */
#ifndef SL2_ORACLE_H
#define SL2_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_config {
  int32_t width, height;
  double fku, fkv, u0, v0, kd1, sd;
  double delta_t;
  int32_t number_of_features_to_select;
  int32_t boxsize;
  double search_override[3]; /* (P00,P01,P11); P00 <= 0 disables */
  int32_t minimum_attempted_measurements_of_feature;
  double successful_match_fraction;
} orc_config;

/* A1  improc.cpp:55-134 */
double orc_correlate2_warning(const uint8_t *patch, int32_t patch_width, int32_t x0lim,
                              int32_t y0lim, const uint8_t *image, int32_t image_width, int32_t x1,
                              int32_t y1, double *sd0, double *sd1);
/* A2  monoslam.cpp:401-477 ; PuInv3 = (P00,P01,P11) ; returns 1 on success */
int32_t orc_elliptical_search(const uint8_t *image, int32_t width, int32_t height,
                              const uint8_t *patch, int32_t boxsize, const double *centre,
                              const double *PuInv3, int32_t *u, int32_t *v, double *best);
/* batched convenience: n features, patches n*B*B, centres n*2, PuInv3 n*3 */
void orc_elliptical_search_batch(const uint8_t *image, int32_t width, int32_t height,
                                 const uint8_t *patches, int32_t boxsize, int32_t n,
                                 const double *centres, const double *PuInv3, int32_t *u,
                                 int32_t *v, uint8_t *found, double *best);
/* score map of one feature over the clamped bounding box (for bit-level score parity):
 * box = (urelstart, urelfinish, vrelstart, vrelfinish, ucentre, vcentre); arrays are
 * [(urelfinish-urelstart+1) x (vrelfinish-vrelstart+1)], urel-major; inside = ellipse test. */
void orc_search_box(int32_t width, int32_t height, int32_t boxsize, const double *centre,
                    const double *PuInv3, int32_t *box6);
void orc_score_map(const uint8_t *image, int32_t width, int32_t height, const uint8_t *patch,
                   int32_t boxsize, const double *centre, const double *PuInv3, double *corr,
                   double *sd_image, uint8_t *inside);
/* A3  monoslam.cpp:371-374 ; S col-major 2x2 -> (P00,P01,P11) */
void orc_puinv_from_S(const double *S, double *PuInv3);
/* A11 search_multiple_overlapping_ellipses.cpp:106-196 */
void orc_smoe_search(const uint8_t *image, int32_t width, int32_t height, const uint8_t *patch,
                     int32_t boxsize, int32_t K, const double *PuInv3, const double *centres,
                     int32_t *res_u, int32_t *res_v, uint8_t *res_flag, double *res_best);

/* N3  monoslam.cpp:1070-1194 ; region4 = (ustart,vstart,ufinish,vfinish); ubest/vbest in-out */
void orc_find_best_patch(const uint8_t *image, int32_t width, int32_t height, int32_t boxsize,
                         const int32_t *region4, int32_t *ubest, int32_t *vbest, double *evbest);

/* N2  monoslam.cpp:1447-1493 (one feature) + feature_init_info.cpp:95-172: Bayes re-weighting of K depth
 * particles from their SMOE matches, normalise, prune below thr/K, re-normalise, mean / variance of
 * lambda.  Returns survivors (0: all probabilities zero, the reference deletes the feature). */
int32_t orc_particle_update(int32_t K, const double *h, const double *Sinv3, const double *detS,
                            const double *lambda, const int32_t *z_uv, const uint8_t *found,
                            double prune_probability_threshold, double *prob, uint8_t *keep,
                            double *cumulative, double *mean_var);

/* A5  motion_model.cpp:84-217 ; F,Q 13x13 col-major */
void orc_motion(const double *xv, const double *u, double delta_t, double *fv, double *F,
                double *Q);
/* A8  motion_model.cpp:237-263 ; J 13x13 */
void orc_dxvnorm_by_dxv(const double *xv, double *J);
/* N1  monoslam.cpp:289-308 ; cam8 = (width,height,fku,fkv,u0,v0,kd1,sd) */
/* N2 prediction: the K depth particles of one partially-initialised feature (ypi = (r, hhat), 6 numbers),
 * MonoSLAM::predict_partially_initialised_feature_measurements (monoslam.cpp:1347-1400, body per particle):
 * h (K x 2), S (K x 2x2 col-major, may be NULL), Sinv3 (K x (00, 01, 11)), detS (K).
 * Pxx 13x13, Pxy 13x6, Pyy 6x6 column-major. */
void orc_predict_particles(const double *cam8, const double *xv, const double *ypi, int32_t K,
                           const double *lambda, const double *Pxx, const double *Pxy, const double *Pyy,
                           double *h, double *S4, double *Sinv3, double *detS);

void orc_predict_feature(const double *cam8, const double *xv, const double *y, const double *Pxx,
                         const double *Pxy, const double *Pyy, double *h, double *dh_by_dxv,
                         double *dh_by_dy, double *R, double *S);
int32_t orc_visibility_test(const double *cam8, const double *xp, const double *y,
                            const double *xp_org, const double *h);
/* A6  kalman.cpp:100-115, dense as written; x (n), P (n x n), H (m x n), R (m x m), nu (m) */
void orc_kalman_update_dense(int32_t n, int32_t m, double *x, double *P, const double *H,
                             const double *R, const double *nu);

/* whole-step object (monoslam.cpp:108-180, tracking only) */
typedef struct orc_slam orc_slam;
orc_slam *orc_slam_create(const orc_config *cfg);
void orc_slam_destroy(orc_slam *s);
void orc_slam_add_feature(orc_slam *s, const double *y, const double *xp_org,
                          const uint8_t *patch);
int32_t orc_slam_num_features(const orc_slam *s);
int32_t orc_slam_state_size(const orc_slam *s);
void orc_slam_set_state(orc_slam *s, const double *x, const double *P);
void orc_slam_get_state(const orc_slam *s, double *x, double *P);
void orc_slam_step(orc_slam *s, const uint8_t *frame);
/* staged pieces, same order as GoOneStep */
void orc_slam_predict(orc_slam *s);
int32_t orc_slam_select(orc_slam *s);
int32_t orc_slam_measure(orc_slam *s, const uint8_t *frame);
void orc_slam_update(orc_slam *s);
void orc_slam_normalise(orc_slam *s);
void orc_slam_finish(orc_slam *s); /* delete_bad_features + symmetrise */
/* per-feature read-back, arrays sized num_features: labels, h(2), z(2), S(4), flags
 * (bit0 selected, bit1 successful), attempted, successful, selection order (-1 if not selected) */
void orc_slam_get_features(const orc_slam *s, int32_t *label, double *h, double *z, double *S,
                           uint8_t *flags, int32_t *attempted, int32_t *successful,
                           int32_t *select_rank);
/* CPU baseline timing: nslam independent streams, each stepped nsteps times over its own
 * ring of nframes frames (frames[i] -> nframes*width*height bytes), spread over nthreads
 * std::threads.  Returns wall seconds for the stepped region. */
double orc_slam_run(orc_slam **slams, int32_t nslam, const uint8_t *const *frames, int32_t nframes,
                    int32_t nsteps, int32_t nthreads);
int32_t orc_hardware_threads(void);
/* the same with one thread pinned per usable CPU (pin != 0) and the seconds the streams spent in the reference's
 * gather / scatter passes (monoslam.cpp:518-614), summed over streams (may be NULL) */
double orc_slam_run_pinned(orc_slam **slams, int32_t nslam, const uint8_t *const *frames, int32_t nframes,
                           int32_t nsteps, int32_t nthreads, int32_t pin, double *gather_scatter_seconds);
/* CPUs this process can really use: affinity mask capped by the cgroup CPU quota */
int32_t orc_usable_cpus(void);

#ifdef __cplusplus
}
#endif
#endif
