/*
 * ref_harness.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Wraps the reference's OWN rfs::RBPHDFilter (headers and TUs compiled unmodified from
 * /root/reference against oracle/compat shims) behind the flat entry point of phd_oracle.h, so
 * that the restatement in phd_oracle.cpp can be pinned against the real implementation on
 * identical inputs, and so that the reference's OpenMP update() can be timed as the CPU baseline.
 * Built by `make -C oracle ref` into oracle/_ref/libphd_ref.so (git-ignored, travels with gpurun).
 *
 * Private members of the reference classes are reached with -Dprivate=public on this TU only
 * (after the standard headers have been included).
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
#include "KalmanFilter_RngBrg.hpp"
#include "MeasurementModel_RngBrg.hpp"
#include "ProcessModel_Odometry2D.hpp"
#include "MatrixPermanent.hpp"
#include "PermutationLexicographic.hpp"
#undef private
#undef protected

#include "phd_oracle.h"

using namespace rfs;

typedef RBPHDFilter<MotionModel_Odometry2d, StaticProcessModel<Landmark2d>, MeasurementModel_RngBrg,
                    KalmanFilter_RngBrg>
    Filter;

extern "C" int phd_ref_update(phd_io* io) {
  if (!io || !io->model || !io->cfg || io->N <= 0) return -1;
  if (io->model->model_id != RFSB200_MODEL_RNGBRG) return -5;
  const int N = io->N;
  if (io->n_threads > 0) omp_set_num_threads(io->n_threads);
  Filter* f = new Filter(N);

  // ---- plugin configuration through the reference's public config structs -----------------
  const rfsb200_model_desc& md = *io->model;
  MeasurementModel_RngBrg* mm = f->getMeasurementModel();
  Eigen::Matrix2d R;
  R << md.R[0], md.R[1], md.R[2], md.R[3];
  mm->setNoise(R);
  mm->config.probabilityOfDetection_ = md.Pd;
  mm->config.uniformClutterIntensity_ = md.clutter_intensity;
  mm->config.rangeLimMax_ = md.range_max;
  mm->config.rangeLimMin_ = md.range_min;
  mm->config.rangeLimBuffer_ = md.range_buffer;
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
  f->config.minUpdatesBeforeResample_ = INT_MAX;  // never resample: update() ends in normalizeWeights()
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
      Landmark2d::Vec lx;
      lx << io->mean_in[2 * k], io->mean_in[2 * k + 1];
      Landmark2d::Mat lS;
      lS << io->cov_in[3 * k], io->cov_in[3 * k + 1], io->cov_in[3 * k + 1], io->cov_in[3 * k + 2];
      Landmark2d lm(lx, lS);
      f->getParticle(i)->getData()->addGaussian(&lm, io->w_in[k], true);
    }
  }

  std::vector<Measurement2d> Z;
  for (int z = 0; z < io->nZ; z++) {
    Measurement2d::Vec zv;
    zv << io->Z[2 * z], io->Z[2 * z + 1];
    Z.push_back(Measurement2d(zv, R));
  }

  // ---- run ------------------------------------------------------------------------------------------
  auto t0 = std::chrono::steady_clock::now();
  if (io->stage >= PHD_STAGE_FULL + 1) {
    // the reference's public entry point, timed as the CPU baseline (ends in normalizeWeights())
    f->update(Z);
  } else if (io->nZ > 0) {
    // same sequence as RBPHDFilter::update() (include/RBPHDFilter.hpp:444-520), stage by stage,
    // without the final resample / normalise
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
    Filter::TGM* gm = f->getParticle(i)->getData().get();
    int cnt = 0;
    for (size_t m = 0; m < gm->gList_.size(); m++) {
      if (gm->gList_[m].landmark == NULL) continue;
      if (k >= io->cap_total) { rc = -4; break; }
      Landmark2d::Vec lx;
      Landmark2d::Mat lS;
      gm->gList_[m].landmark->get(lx, lS);
      io->mean_out[2 * k] = lx(0);
      io->mean_out[2 * k + 1] = lx(1);
      io->cov_out[3 * k] = lS(0, 0);
      io->cov_out[3 * k + 1] = lS(0, 1);
      io->cov_out[3 * k + 2] = lS(1, 1);
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

/* MotionModel_Odometry2d::step (src/ProcessModel_Odometry2D.cpp:41-89): pose [3], u = (dx, dy, dtheta) */
extern "C" void phd_ref_odometry2d_step(const double* pose, const double* u, double* out) {
  MotionModel_Odometry2d mm;
  Pose2d::Vec x;
  x << pose[0], pose[1], pose[2];
  Pose2d s_km(x, Pose2d::Mat::Zero()), s_k;
  Odometry2d::Vec uv;
  uv << u[0], u[1], u[2];
  Odometry2d in(uv, Odometry2d::Mat::Zero());
  TimeStamp dT(0.1);
  mm.step(s_k, s_km, in, dT);
  for (int k = 0; k < 3; k++) out[k] = s_k.get(k);
}

/* rfs::MatPerm::calc on a row-major n x n matrix */
extern "C" double phd_ref_permanent(const double* A, int n) {
  Eigen::MatrixXd M(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) M(i, j) = A[i * n + j];
  return MatPerm::calc(M);
}

/* number of assignments PermutationLexicographic visits */
extern "C" int64_t phd_ref_lexi_count(int nM, int nZ) {
  PermutationLexicographic pl(nM, nZ, true);
  std::vector<unsigned> o(nM + nZ + 1);
  int64_t c = 0;
  while (pl.next(o.data()) != 0) c++;
  return c;
}

/* CostMatrixGeneral::partition() on a dense table: returns nP and writes the partition sizes the
 * reference's caller would see for p < nP (Q6) */
extern "C" int phd_ref_partition(const double* L, int nR, int nC, int* nRows, int* nCols, int* isZero) {
  double** C;
  CostMatrixGeneral cm(C, nR, nC);
  for (int i = 0; i < nR; i++)
    for (int j = 0; j < nC; j++) C[i][j] = L[i * nC + j];
  int nP = cm.partition();
  for (int p = 0; p < nP; p++) {
    unsigned r, c;
    bool nz = cm.getPartitionSize(p, r, c);
    nRows[p] = r;
    nCols[p] = c;
    isZero[p] = nz ? 0 : 1;
  }
  return nP;
}

/* the Murty branch of rfsMeasurementLikelihood (include/RBPHDFilter.hpp:920-959) on one partition */
extern "C" double phd_ref_murty_sum(const double* Lp, int nR, int nC, const double* rowPd, const double* colClutter) {
  const double BIG_NEG_NUM = -1000;
  int n = nR + nC;
  double** Cp = new double*[n];
  for (int i = 0; i < n; i++) Cp[i] = new double[n];
  for (int r = 0; r < nR; r++)
    for (int c = 0; c < nC; c++) {
      double v = Lp[r * nC + c];
      if (v == 0) v = BIG_NEG_NUM;
      else { v = log(v); if (v < BIG_NEG_NUM) v = BIG_NEG_NUM; }
      Cp[r][c] = v;
    }
  for (int r = 0; r < nR; r++)
    for (int c = nC; c < n; c++) Cp[r][c] = (r == c - nC) ? log(1 - rowPd[r]) : BIG_NEG_NUM;
  for (int r = nR; r < n; r++)
    for (int c = 0; c < nC; c++) Cp[r][c] = (r - nR == c) ? log(colClutter[c]) : BIG_NEG_NUM;
  for (int r = nR; r < n; r++)
    for (int c = nC; c < n; c++) Cp[r][c] = 0;
  double pl = 0;
  {
    Murty murtyAlgo(Cp, n);
    Murty::Assignment a;
    double score = 0;
    murtyAlgo.setRealAssignmentBlock(nR, nC);
    for (int k = 0; k < 200; k++) {
      int rank = murtyAlgo.findNextBest(a, score);
      if (rank == -1 || score < BIG_NEG_NUM) break;
      pl += exp(score);
    }
  }
  for (int i = 0; i < n; i++) delete[] Cp[i];
  delete[] Cp;
  return pl;
}

/* The reference's own addBirthGaussians() (include/RBPHDFilter.hpp:1000-1080), range-bearing plugin set. */
#include "ref_births.hpp"
extern "C" int phd_ref_birth_candidates(phd_birth_io* io) {
  if (!io || !io->model || io->N <= 0 || io->model->model_id != RFSB200_MODEL_RNGBRG) return -1;
  Filter* f = new Filter(io->N);
  const rfsb200_model_desc& md = *io->model;
  Eigen::Matrix2d R;
  R << md.R[0], md.R[1], md.R[2], md.R[3];
  MeasurementModel_RngBrg* mm = f->getMeasurementModel();
  mm->setNoise(R);
  mm->config.rangeLimMax_ = md.range_max;
  mm->config.rangeLimMin_ = md.range_min;
  f->getKalmanFilter()->config.rangeInnovationThreshold_ = md.innov_thr_range;
  f->getKalmanFilter()->config.bearingInnovationThreshold_ = md.innov_thr_bearing;
  const int rc = ref_birth_candidates_run<Filter, 2>(f, io, R);
  delete f;
  return rc;
}
