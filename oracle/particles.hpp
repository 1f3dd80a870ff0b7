// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the particle re-weighting of one
// partially-initialised feature: MonoSLAM::update_partially_initialised_feature_probabilities
// (monoslam.cpp:1447-1493, body for one FeatureInitInfo) with
// FeatureInitInfo::normalise_particle_vector_and_calculate_cumulative (feature_init_info.cpp:95-119),
// prune_particle_vector (:126-141) and calculate_mean_and_covariance (:152-172); lambda is scalar
// (kParticleDimension_ = 1 in MonoSLAM).  Pinned: bit-exact against the reference's own feature_init_info.cpp
// (normalise / prune / mean-covariance) and against the reference's whole particle cycle in monoslam.cpp
// (likelihood loop of :1456-1478 included; _ref/libsl2refmodels.so, tests/test_oracle_ref.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace sl2o {

// h, z: K x 2; Sinv3: K x (S00, S01, S11); returns the number of surviving particles, 0 when every
// probability is zero (the reference then deletes the feature: monoslam.cpp:1486-1490; prob is left
// un-normalised like there).  keep[k] = 1 for survivors; cumulative[k] of survivors in order, 0 otherwise.
inline int particle_update(int K, const double *h, const double *Sinv3, const double *detS,
                           const double *lambda, const int32_t *z_uv, const uint8_t *found,
                           double prune_probability_threshold, double *prob, uint8_t *keep,
                           double *cumulative, double *mean_var) {
  for (int k = 0; k < K; ++k) {  // monoslam.cpp:1456-1478
    double likelihood = 0.0;
    if (found[k]) {
      const double nu0 = (double)z_uv[2 * k] - h[2 * k], nu1 = (double)z_uv[2 * k + 1] - h[2 * k + 1];
      const double r0 = Sinv3[3 * k] * nu0 + Sinv3[3 * k + 1] * nu1;      // SInv * nu
      const double r1 = Sinv3[3 * k + 1] * nu0 + Sinv3[3 * k + 2] * nu1;
      const double q = nu0 * r0 + nu1 * r1;                                // nu . (SInv nu)
      likelihood = (1.0 / (std::sqrt(2.0 * M_PI * detS[k]))) * std::exp(-0.5 * q);
    }
    prob[k] = prob[k] * likelihood;
    keep[k] = 1;
    cumulative[k] = 0.0;
  }
  auto normalise = [&]() -> bool {  // feature_init_info.cpp:95-119, over the particles still kept
    double total = 0.0;
    for (int k = 0; k < K; ++k)
      if (keep[k]) total += prob[k];
    if (total == 0.0) return false;
    double cum = 0.0;
    for (int k = 0; k < K; ++k)
      if (keep[k]) {
        prob[k] = prob[k] / total;
        cumulative[k] = cum + prob[k];
        cum += prob[k];
      }
    return true;
  };
  mean_var[0] = mean_var[1] = 0.0;
  if (!normalise()) {
    for (int k = 0; k < K; ++k) keep[k] = 0;
    return 0;
  }
  const double thr = prune_probability_threshold / double(K);  // :128
  int left = 0;
  for (int k = 0; k < K; ++k) {
    if (prob[k] < thr) {
      keep[k] = 0;
      cumulative[k] = 0.0;
    } else {
      ++left;
    }
  }
  normalise();  // :140 (return value ignored there too)
  double mean = 0.0, e2 = 0.0;  // :152-172
  for (int k = 0; k < K; ++k)
    if (keep[k]) {
      mean += prob[k] * lambda[k];
      e2 += prob[k] * (lambda[k] * lambda[k]);
    }
  mean_var[0] = mean;
  mean_var[1] = e2 - (mean * mean);
  return left;
}

}  // namespace sl2o
