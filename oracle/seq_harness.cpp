/*
 * seq_harness.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Drives rfs::RBPHDFilter<MotionModel_Odometry2d, StaticProcessModel<Landmark2d>,
 * MeasurementModel_RngBrg, KalmanFilter_RngBrg> through its PUBLIC API only — predict(),
 * setParticlePose(), update(), getGMSize(), getLandmark(), particle weights — for a scripted
 * sequence of steps, and dumps the final state.  The same source is compiled twice by
 * oracle/Makefile:
 *   _ref/libseq_ref.so   against the reference's own RBPHDFilter.hpp            (entry seq_run_ref)
 *   _ref/libseq_b200.so  against include/rfs_b200/RBPHDFilter.hpp + librfsb200  (entry seq_run_b200)
 * so the drop-in header is checked against the class it replaces on identical inputs, including
 * birth Gaussians, the landmark process noise and resampling.
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <list>
#include <memory>
#include <queue>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
#include <sys/times.h>
#include <unistd.h>

#include "ProcessModel_Odometry2D.hpp"
#include "RBPHDFilter.hpp"
#include "KalmanFilter_RngBrg.hpp"

using namespace rfs;

typedef RBPHDFilter<MotionModel_Odometry2d, StaticProcessModel<Landmark2d>, MeasurementModel_RngBrg, KalmanFilter_RngBrg> Filter;

struct seq_io {
  int32_t N, n_steps, nZ_max;
  const double* poses;     /* [n_steps][N][3] pose given to every particle before each update */
  const double* pose_cov;  /* [6] upper triangle, shared */
  const double* Z;         /* [n_steps][nZ_max][2] */
  const int32_t* nZ;       /* [n_steps] */
  const double* R;         /* [4] */
  const double* Q_lmk;     /* [4] or NULL */
  double Pd, clutter, range_min, range_max, range_buffer, thr_r, thr_b;
  double birth_w, gate, eval_w, wl_gate, merge_t, merge_f, prune_t;
  int32_t n_eval, use_sc, min_updates_before_resample;
  double neff_threshold;
  int32_t precision;       /* b200 only */
  uint32_t seed48;
  /* out */
  int64_t cap_total;
  int32_t* count_out;      /* [N] */
  double* mean_out;        /* [..][2] */
  double* cov_out;         /* [..][3] */
  double* w_out;
  double* weight_out;      /* [N] */
  int32_t* n_resampled;    /* [1] number of updates after which a resampling happened */
  int32_t* gm_size_trace;  /* [n_steps] getGMSize(0) after each update */
  /* candidate-list births (include/RBPHDFilter.hpp:1023-1080): birth_count_thr == 1 is the direct form */
  uint32_t birth_count_thr, birth_check_thr, birth_cur_thr, birth_reserved;
  double birth_support_dist;
};

#ifdef SEQ_B200
#define SEQ_ENTRY seq_run_b200
#else
#define SEQ_ENTRY seq_run_ref
#endif

extern "C" int SEQ_ENTRY(seq_io* io) {
  const int N = io->N;
  Filter f(N);
#ifdef SEQ_B200
  f.deviceConfig.precision = io->precision;
  f.deviceConfig.gmCapacity = 128;
  f.deviceConfig.zCapacity = 32;
  f.deviceConfig.workCapacity = 256;
#else
  f.config.importanceWeightingEvalPointGuassianWeight_ = 0.75;
  f.config.useClusterProcess_ = false;
#endif
  Eigen::Matrix2d R;
  R << io->R[0], io->R[1], io->R[2], io->R[3];
  f.getMeasurementModel()->setNoise(R);
  f.getMeasurementModel()->config.probabilityOfDetection_ = io->Pd;
  f.getMeasurementModel()->config.uniformClutterIntensity_ = io->clutter;
  f.getMeasurementModel()->config.rangeLimMax_ = io->range_max;
  f.getMeasurementModel()->config.rangeLimMin_ = io->range_min;
  f.getMeasurementModel()->config.rangeLimBuffer_ = io->range_buffer;
  f.getKalmanFilter()->config.rangeInnovationThreshold_ = io->thr_r;
  f.getKalmanFilter()->config.bearingInnovationThreshold_ = io->thr_b;
  if (io->Q_lmk) {
    Landmark2d::Mat Q;
    Q << io->Q_lmk[0], io->Q_lmk[1], io->Q_lmk[2], io->Q_lmk[3];
    f.getLmkProcessModel()->setNoise(Q);
  }
  f.config.birthGaussianWeight_ = io->birth_w;
  if (io->birth_count_thr > 0) {
    f.config.birthGaussianMeasurementCountThreshold_ = io->birth_count_thr;
    f.config.birthGaussianMeasurementCheckThreshold_ = io->birth_check_thr;
    f.config.birthGaussianCurrentMeasurementCountThreshold_ = io->birth_cur_thr;
    f.config.birthGaussianMeasurementSupportDist_ = io->birth_support_dist;
  }
  f.config.newGaussianCreateInnovMDThreshold_ = io->gate;
  f.config.importanceWeightingEvalPointCount_ = io->n_eval;
  f.config.importanceWeightingEvalPointGuassianWeight_ = io->eval_w;
  f.config.importanceWeightingMeasurementLikelihoodMDThreshold_ = io->wl_gate;
  f.config.gaussianMergingThreshold_ = io->merge_t;
  f.config.gaussianMergingCovarianceInflationFactor_ = io->merge_f;
  f.config.gaussianPruningThreshold_ = io->prune_t;
  f.config.useClusterProcess_ = io->use_sc != 0;
  f.config.minUpdatesBeforeResample_ = io->min_updates_before_resample;
  f.config.minMeasurementsBeforeResample_ = 0;
  f.setEffectiveParticleCountThreshold(io->neff_threshold);
  srand48(io->seed48);

  Pose2d::Mat Sx;
  const double* s = io->pose_cov;
  Sx << s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5];
  MotionModel_Odometry2d::TInput u;   /* zero odometry; the poses are set explicitly below */
  Odometry2d::Vec uv;
  uv << 0, 0, 0;
  Odometry2d::Mat uS;
  uS.setZero();
  u.set(uv, uS);
  TimeStamp dT(0.1);
  double prev_w0 = -1;
  int n_res = 0;
  try {
    for (int k = 0; k < io->n_steps; k++) {
      f.predict(u, dT, false, false, true);
      for (int i = 0; i < N; i++) {
        Pose2d::Vec x;
        const double* pp = io->poses + ((size_t)k * N + i) * 3;
        x << pp[0], pp[1], pp[2];
        Pose2d p(x, Sx);
        f.setParticlePose(i, p);
      }
      std::vector<Measurement2d> Z;
      for (int z = 0; z < io->nZ[k]; z++) {
        Measurement2d::Vec zv;
        zv << io->Z[((size_t)k * io->nZ_max + z) * 2], io->Z[((size_t)k * io->nZ_max + z) * 2 + 1];
        Z.push_back(Measurement2d(zv, R));
      }
      f.update(Z);
      /* a resampling leaves every weight at exactly 1 */
      bool all_one = true;
      for (int i = 0; i < N; i++) all_one = all_one && (f.getParticleSet()->at(i)->getWeight() == 1.0);
      if (all_one && io->nZ[k] > 0) n_res++;
      if (io->gm_size_trace) io->gm_size_trace[k] = f.getGMSize(0);
      (void)prev_w0;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "seq harness: %s\n", e.what());
    return -2;
  }
  if (io->n_resampled) io->n_resampled[0] = n_res;
  int64_t k = 0;
  for (int i = 0; i < N; i++) {
    const int n = f.getGMSize(i);
    io->count_out[i] = n;
    for (int m = 0; m < n; m++, k++) {
      if (k >= io->cap_total) return -4;
      Landmark2d::Vec lx;
      Landmark2d::Mat lS;
      double w;
      f.getLandmark(i, m, lx, lS, w);
      io->mean_out[2 * k] = lx(0);
      io->mean_out[2 * k + 1] = lx(1);
      io->cov_out[3 * k] = lS(0, 0);
      io->cov_out[3 * k + 1] = lS(0, 1);
      io->cov_out[3 * k + 2] = lS(1, 1);
      io->w_out[k] = w;
    }
    io->weight_out[i] = f.getParticleSet()->at(i)->getWeight();
  }
  return 0;
}
