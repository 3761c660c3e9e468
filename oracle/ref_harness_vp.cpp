/*
 * ref_harness_vp.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The Victoria Park instantiation of the reference's OWN filter,
 *   rfs::RBPHDFilter<MotionModel_Ackerman2d, StaticProcessModel<Landmark3d>,
 *                    MeasurementModel_VictoriaPark, KalmanFilter_VictoriaPark>
 * (src/rbphdslam_VictoriaPark.cpp:61-64), compiled unmodified from /root/reference against
 * oracle/compat and driven through the flat entry point of phd_oracle.h.  Same role as
 * ref_harness.cpp for the 2-D range-bearing model: pins the 3-D restatement (phd_oracle_vp.cpp),
 * produces the golden fixtures and is the timed CPU baseline of the VP workload.
 * For this model phd_io carries mean [..][3], cov [..][6] (xx,xy,xz,yy,yz,zz), Z [nZ][3].
 */
#include <algorithm>
#include <chrono>
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
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
#include <sys/times.h>
#include <unistd.h>

#include "Eigen/Core"
#include "boost/shared_ptr.hpp"
#include "boost/shared_array.hpp"
#include "boost/multi_array.hpp"
#include "boost/timer/timer.hpp"
#include "boost/lexical_cast.hpp"
#include "boost/random/mersenne_twister.hpp"
#include "boost/random/normal_distribution.hpp"
#include "boost/random/variate_generator.hpp"

#define private public
#define protected public
#include "RBPHDFilter.hpp"
#include "KalmanFilter_VictoriaPark.hpp"
#include "MeasurementModel_VictoriaPark.hpp"
#include "ProcessModel_Ackerman2D.hpp"
#undef private
#undef protected

#include "phd_oracle.h"

using namespace rfs;

typedef RBPHDFilter<MotionModel_Ackerman2d, StaticProcessModel<Landmark3d>, MeasurementModel_VictoriaPark,
                    KalmanFilter_VictoriaPark>
    FilterVP;

extern "C" int phd_ref_update_vp(phd_io* io) {
  if (!io || !io->model || !io->cfg || io->N <= 0) return -1;
  if (io->model->model_id != RFSB200_MODEL_VICTORIAPARK) return -5;
  const int N = io->N;
  if (io->n_threads > 0) omp_set_num_threads(io->n_threads);
  FilterVP* f = new FilterVP(N);

  // ---- plugin configuration as src/rbphdslam_VictoriaPark.cpp:366-398 does it ------------------
  const rfsb200_model_desc& md = *io->model;
  MeasurementModel_VictoriaPark* mm = f->getMeasurementModel();
  Eigen::Matrix3d R;
  R << md.R[0], md.R[1], md.R[2], md.R[3], md.R[4], md.R[5], md.R[6], md.R[7], md.R[8];
  mm->setNoise(R, md.Slb);
  mm->config.probabilityOfDetection_.assign(md.pd_table, md.pd_table + md.pd_table_n);
  mm->config.expectedClutterNumber_ = md.clutter_integral;
  mm->config.rangeLimMax_ = md.range_max;
  mm->config.rangeLimMin_ = md.range_min;
  mm->config.bearingLimitMax_ = md.bearing_max;
  mm->config.bearingLimitMin_ = md.bearing_min;
  mm->config.bufferZonePd_ = md.buffer_zone_pd;
  std::vector<double> scan(md.scan, md.scan + md.scan_n);
  mm->setLaserScan(scan);
  {  // the descriptor's clutter intensity must be what the plugin itself derives from the scan
    Measurement3d dummy;
    const double ci = mm->clutterIntensity(dummy, io->nZ);
    if (!(fabs(ci - md.clutter_intensity) <= 1e-12 * fabs(ci))) {
      fprintf(stderr, "phd_ref_update_vp: clutter intensity %g in the descriptor, plugin says %g\n", md.clutter_intensity, ci);
      delete f;
      return -1;
    }
  }
  f->getKalmanFilter()->config.rangeInnovationThreshold_ = md.innov_thr_range;
  f->getKalmanFilter()->config.bearingInnovationThreshold_ = md.innov_thr_bearing;

  const rfsb200_filter_cfg& fc = *io->cfg;
  f->config.birthGaussianWeight_ = fc.birth_gaussian_weight;
  f->config.newGaussianCreateInnovMDThreshold_ = fc.new_gaussian_create_innov_md_threshold;
  f->config.importanceWeightingEvalPointCount_ = fc.eval_point_count;
  f->config.importanceWeightingEvalPointGuassianWeight_ = fc.eval_point_gaussian_weight;
  f->config.importanceWeightingMeasurementLikelihoodMDThreshold_ = fc.meas_likelihood_md_threshold;
  f->config.gaussianMergingThreshold_ = fc.merging_threshold;
  f->config.gaussianMergingCovarianceInflationFactor_ = fc.merging_cov_inflation_factor;
  f->config.gaussianPruningThreshold_ = fc.pruning_threshold;
  f->config.useClusterProcess_ = fc.use_cluster_process != 0;
  f->config.minUpdatesBeforeResample_ = INT_MAX;
  f->config.minMeasurementsBeforeResample_ = INT_MAX;
  f->resampleOccured_ = false;

  // ---- inject the particle state --------------------------------------------------------------
  int64_t k = 0;
  for (int i = 0; i < N; i++) {
    Pose2d::Vec x;
    x << io->pose[3 * i], io->pose[3 * i + 1], io->pose[3 * i + 2];
    Pose2d::Mat Sx;
    Sx.setZero();
    const double* s = NULL;
    if (io->pose_cov_mode == 1) s = io->pose_cov;
    if (io->pose_cov_mode == 2) s = io->pose_cov + 6 * (size_t)i;
    if (s) Sx << s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5];
    Pose2d p(x, Sx);
    f->setParticlePose(i, p);
    f->getParticleSet()->at(i)->setWeight(io->weight_in[i]);
    for (int m = 0; m < io->count_in[i]; m++, k++) {
      Landmark3d::Vec lx;
      lx << io->mean_in[3 * k], io->mean_in[3 * k + 1], io->mean_in[3 * k + 2];
      const double* c = io->cov_in + 6 * k;
      Landmark3d::Mat lS;
      lS << c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5];
      Landmark3d lm(lx, lS);
      f->getParticle(i)->getData()->addGaussian(&lm, io->w_in[k], true);
    }
  }

  std::vector<Measurement3d> Z;
  for (int z = 0; z < io->nZ; z++) {
    Measurement3d::Vec zv;
    zv << io->Z[3 * z], io->Z[3 * z + 1], io->Z[3 * z + 2];
    Z.push_back(Measurement3d(zv, R));
  }

  // ---- run ------------------------------------------------------------------------------------------
  auto t0 = std::chrono::steady_clock::now();
  if (io->stage >= PHD_STAGE_FULL + 1) {
    f->update(Z);
  } else if (io->nZ > 0) {
    // same sequence as RBPHDFilter::update() (include/RBPHDFilter.hpp:444-520), stage by stage
    f->nUpdatesSinceResample_++;
    f->setMeasurements(Z);
    if (f->nThreads_ > 1)
      for (int j = 1; j < f->nThreads_; j++) f->kfs_[j] = f->kfs_[0];
#pragma omp parallel
    {
#pragma omp for
      for (int i = 0; i < N; i++) f->updateMap(i);
      if (!f->config.useClusterProcess_ && io->stage >= PHD_STAGE_WEIGHTING) {
#pragma omp for
        for (int i = 0; i < N; i++) f->importanceWeighting(i);
      }
      if (io->stage >= PHD_STAGE_MERGE) {
#pragma omp for
        for (int i = 0; i < N; i++)
          f->particleSet_[i]->getData()->merge(f->config.gaussianMergingThreshold_,
                                              f->config.gaussianMergingCovarianceInflationFactor_);
      }
      if (io->stage >= PHD_STAGE_FULL) {
#pragma omp for
        for (int i = 0; i < N; i++) f->particleSet_[i]->getData()->prune(f->config.gaussianPruningThreshold_);
      }
    }
  }
  io->elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  // ---- dump ---------------------------------------------------------------------------------------------
  k = 0;
  int rc = 0;
  for (int i = 0; i < N && rc == 0; i++) {
    FilterVP::TGM* gm = f->getParticle(i)->getData().get();
    int cnt = 0;
    for (size_t m = 0; m < gm->gList_.size(); m++) {
      if (gm->gList_[m].landmark == NULL) continue;
      if (k >= io->cap_total) { rc = -4; break; }
      Landmark3d::Vec lx;
      Landmark3d::Mat lS;
      gm->gList_[m].landmark->get(lx, lS);
      for (int d = 0; d < 3; d++) io->mean_out[3 * k + d] = lx(d);
      double* c = io->cov_out + 6 * k;
      c[0] = lS(0, 0); c[1] = lS(0, 1); c[2] = lS(0, 2); c[3] = lS(1, 1); c[4] = lS(1, 2); c[5] = lS(2, 2);
      io->w_out[k] = gm->gList_[m].weight;
      if (io->wprev_out) io->wprev_out[k] = gm->gList_[m].weight_prev;
      k++;
      cnt++;
    }
    io->count_out[i] = cnt;
    io->weight_out[i] = f->getParticleSet()->at(i)->getWeight();
    if (io->unused_mask) {
      uint64_t mk = 0;
      for (size_t u = 0; u < f->unused_measurements_[i].size(); u++) mk |= (1ull << f->unused_measurements_[i][u]);
      io->unused_mask[i] = mk;
    }
    if (io->n_in_fov) io->n_in_fov[i] = (int32_t)f->nLandmarksInFOV_[i];
    if (io->flags) io->flags[i] = 0;
  }
  delete f;
  return rc;
}

/* MeasurementModel_VictoriaPark::probabilityOfDetection on one (pose, landmark) with the descriptor's
 * configuration: a stand-alone probe used to pin the restatement of the P_D geometry. */
extern "C" double phd_ref_vp_pd(const rfsb200_model_desc* md, const double* pose, const double* lx, const double* lcov6,
                                int* close_out) {
  MeasurementModel_VictoriaPark mm;
  Eigen::Matrix3d R;
  R << md->R[0], md->R[1], md->R[2], md->R[3], md->R[4], md->R[5], md->R[6], md->R[7], md->R[8];
  mm.setNoise(R, md->Slb);
  mm.config.probabilityOfDetection_.assign(md->pd_table, md->pd_table + md->pd_table_n);
  mm.config.expectedClutterNumber_ = md->clutter_integral;
  mm.config.rangeLimMax_ = md->range_max;
  mm.config.rangeLimMin_ = md->range_min;
  mm.config.bearingLimitMax_ = md->bearing_max;
  mm.config.bearingLimitMin_ = md->bearing_min;
  mm.config.bufferZonePd_ = md->buffer_zone_pd;
  std::vector<double> scan(md->scan, md->scan + md->scan_n);
  mm.setLaserScan(scan);
  Pose2d::Vec x;
  x << pose[0], pose[1], pose[2];
  Pose2d p(x, Pose2d::Mat::Zero());
  Landmark3d::Vec m;
  m << lx[0], lx[1], lx[2];
  Landmark3d::Mat S;
  S << lcov6[0], lcov6[1], lcov6[2], lcov6[1], lcov6[3], lcov6[4], lcov6[2], lcov6[4], lcov6[5];
  Landmark3d lm(m, S);
  bool close = false;
  const double pd = mm.probabilityOfDetection(p, lm, close);
  if (close_out) *close_out = close ? 1 : 0;
  return pd;
}

/* MotionModel_Ackerman2d::step (src/ProcessModel_Ackerman2D.cpp:49-77): params = (h, l, poi dx, poi dy), u = (v, steering) */
extern "C" void phd_ref_ackerman2d_step(const double* params, const double* pose, const double* u, double dt, double* out) {
  MotionModel_Ackerman2d mm(params[0], params[1], params[2], params[3]);
  Pose2d::Vec x;
  x << pose[0], pose[1], pose[2];
  Pose2d s_km(x, Pose2d::Mat::Zero()), s_k;
  AckermanInput::Vec uv;
  uv << u[0], u[1];
  AckermanInput in(uv, AckermanInput::Mat::Zero());
  TimeStamp dT(dt);
  mm.step(s_k, s_km, in, dT);
  for (int k = 0; k < 3; k++) out[k] = s_k.get(k);
}

/* The reference's own addBirthGaussians() (include/RBPHDFilter.hpp:1000-1080), Victoria Park plugin set. */
#include "ref_births.hpp"
extern "C" int phd_ref_birth_candidates_vp(phd_birth_io* io) {
  if (!io || !io->model || io->N <= 0 || io->model->model_id != RFSB200_MODEL_VICTORIAPARK) return -1;
  FilterVP* f = new FilterVP(io->N);
  const rfsb200_model_desc& md = *io->model;
  Eigen::Matrix3d R;
  R << md.R[0], md.R[1], md.R[2], md.R[3], md.R[4], md.R[5], md.R[6], md.R[7], md.R[8];
  f->getMeasurementModel()->setNoise(R, md.Slb);
  f->getKalmanFilter()->config.rangeInnovationThreshold_ = md.innov_thr_range;
  f->getKalmanFilter()->config.bearingInnovationThreshold_ = md.innov_thr_bearing;
  const int rc = ref_birth_candidates_run<FilterVP, 3>(f, io, R);
  delete f;
  return rc;
}
