/*
 * RBPHDFilter.hpp (B200 drop-in) — rfs::RBPHDFilter<RobotProcessModel, LmkProcessModel,
 * MeasurementModel, KalmanFilter> with the reference's public surface
 * (reference include/RBPHDFilter.hpp:72-251), whose update() runs on the GPU through the C ABI of
 * include/rfsb200.h instead of the OpenMP region of the reference (:469-520).
 *
 * How to use: put this directory BEFORE the reference's include/ on the include path
 * (-I<repo>/include/rfs_b200 -I<reference>/include) and link librfsb200.so.  Every other header
 * (ParticleFilter.hpp, GaussianMixture.hpp, the plugin classes ...) is the reference's own,
 * unmodified; drivers such as src/rbphdslam2dSim.cpp compile unchanged.
 *
 * What lives where
 *   host (unchanged reference code): particle poses / weights / trajectories (ParticleFilter),
 *     particle propagation (ProcessModel::sample), the resampling decision (one drand48()).
 *   device (librfsb200.so): every particle's Gaussian mixture, resident in HBM between calls;
 *     predict()'s map part (birth Gaussians + landmark process noise), update() (map update,
 *     particle weighting, merge, prune, weight sums), the data movement of resampling.
 *   The host-side GaussianMixture objects of the particles stay EMPTY; getGMSize / getLandmark
 *     read the device state (one particle's planes per call, cached until the next mutation).
 *
 * Plugin objects cannot be called from the device, so before every update the live plugin
 * objects are read into a POD (rfs::b200::ModelTraits<MeasurementModel, KalmanFilter>::describe).
 * Traits specialisations exist for MeasurementModel_RngBrg + KalmanFilter_RngBrg (2-D landmarks) and
 * for MeasurementModel_VictoriaPark + KalmanFilter_VictoriaPark (3-D landmarks, P_D from the lidar
 * scan); other plugin types do not compile against this header (static_assert) — they keep using the
 * reference header.
 *
 * Birth Gaussians: with birthGaussianMeasurementCountThreshold_ == 1 (the 2-D simulator) they are
 * created on the device; otherwise (Victoria Park configuration) the candidate lists of
 * addBirthGaussians() (reference :1000-1080) are kept on the host, evaluated with the live plugin
 * objects exactly as the reference does, and the promoted candidates are appended to the device maps.
 */
#ifndef RBPHDFILTER_HPP
#define RBPHDFILTER_HPP

#include <Eigen/Core>
#include <math.h>
#include <stdio.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "GaussianMixture.hpp"
#include "KalmanFilter.hpp"
#include "ParticleFilter.hpp"
#include "Timer.hpp"
#include "KalmanFilter_VictoriaPark.hpp"   /* reference header (also brings MeasurementModel_VictoriaPark.hpp) */
#include "ProcessModel_Odometry2D.hpp"     /* reference headers: the two motion models rfsb200_propagate implements */
#include "ProcessModel_Ackerman2D.hpp"

#include "../rfsb200.h"

namespace rfs {
namespace b200 {
namespace detail {
/* Host array of doubles in page-locked memory (rfsb200_host_alloc): with such buffers rfsb200_update_host stages
 * nothing — one conversion kernel reads the inputs over PCIe and the update kernel stores the results into them.
 * Falls back to pageable memory (and the staged copies) if pinning is not possible.  Contents are not initialised
 * and not preserved by a growing resize(): every user fills the array completely before it is read. */
class PinnedDoubles {
 public:
  PinnedDoubles() : p_(NULL), n_(0), cap_(0), pinned_(false) {}
  ~PinnedDoubles() { release(); }
  void resize(size_t n) {
    if (n > cap_) {
      release();
      void* q = NULL;
      if (rfsb200_host_alloc(&q, (uint64_t)n * sizeof(double)) == RFSB200_OK && q) {
        p_ = static_cast<double*>(q);
        pinned_ = true;
      } else {
        p_ = static_cast<double*>(malloc(n * sizeof(double)));
        pinned_ = false;
        if (!p_) throw std::bad_alloc();
      }
      cap_ = n;
    }
    n_ = n;
  }
  double* data() { return p_; }
  const double* data() const { return p_; }
  double& operator[](size_t i) { return p_[i]; }
  const double& operator[](size_t i) const { return p_[i]; }
  size_t size() const { return n_; }

 private:
  PinnedDoubles(const PinnedDoubles&);              /* not copyable */
  PinnedDoubles& operator=(const PinnedDoubles&);
  void release() {
    if (p_) {
      if (pinned_) rfsb200_host_free(p_);
      else free(p_);
    }
    p_ = NULL;
    n_ = cap_ = 0;
  }
  double* p_;
  size_t n_, cap_;
  bool pinned_;
};
}  // namespace detail
}  // namespace b200
}  // namespace rfs

namespace rfs {

namespace b200 {

/** Reads the live plugin objects into the device descriptor.  Specialise per plugin pair. */
template <class MeasurementModel, class KalmanFilter>
struct ModelTraits {
  static const bool supported = false;
};

}  // namespace b200
}  // namespace rfs

/* The RngBrg specialisation is only compiled when the plugin headers are in the translation unit
 * (the reference's drivers include KalmanFilter_RngBrg.hpp themselves, after this header, so the
 * specialisation is declared against forward declarations and defined where both are complete). */
namespace rfs {
class MeasurementModel_RngBrg;
class KalmanFilter_RngBrg;

namespace b200 {
template <>
struct ModelTraits<MeasurementModel_RngBrg, KalmanFilter_RngBrg> {
  static const bool supported = true;
  static const int lmk_dim = 2, meas_dim = 2, pose_dim = 3;
  /* defined as a template so that the plugin classes only need to be complete at the call site */
  template <class MM, class KF>
  static void describe(MM& mm, KF& kf, unsigned nZ, rfsb200_model_desc& d) {
    d = rfsb200_model_desc();
    d.model_id = RFSB200_MODEL_RNGBRG;
    typename MM::TMeasurement::Mat R;
    mm.getNoise(R);
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) d.R[i * 2 + j] = R(i, j);
    d.Pd = mm.config.probabilityOfDetection_;
    typename MM::TMeasurement z;   /* uniform clutter: the value does not depend on z */
    d.clutter_intensity = mm.clutterIntensity(z, (int)nZ);
    d.clutter_integral = mm.clutterIntensityIntegral((int)nZ);
    d.range_min = mm.config.rangeLimMin_;
    d.range_max = mm.config.rangeLimMax_;
    d.range_buffer = mm.config.rangeLimBuffer_;
    d.innov_thr_range = kf.config.rangeInnovationThreshold_;
    d.innov_thr_bearing = kf.config.bearingInnovationThreshold_;
  }
};
}  // namespace b200

/* ---- Victoria Park plugin pair -------------------------------------------------------------------
 * The lidar scan (setLaserScan) and the beam-angle variance (setNoise(R, Slb)) are private members of
 * the reference class without getters.  The class is reused UNMODIFIED, so they are read through
 * member pointers obtained by explicit template instantiation (access checking does not apply to the
 * arguments of an explicit instantiation, [temp.spec]). */
namespace b200 {
namespace detail {
/* Read access to private members of reference classes that have no getter (the class is reused unmodified): the address
 * of a private member may be named in an explicit instantiation (access checking does not apply to its arguments), and
 * the instantiation defines a friend function that hands the member pointer out.  The pointer is a constant expression:
 * nothing is initialised at run time, so there is no static-initialisation order to get wrong.  The tags live in an
 * unnamed namespace, i.e. every translation unit that includes this header instantiates ITS OWN specialisations and no
 * explicit instantiation definition appears twice in a program. */
namespace {
template <class Tag, typename Tag::type P>
struct MemberAccess {
  friend typename Tag::type member_ptr(Tag) { return P; }
};
struct VPScanTag {
  typedef std::vector<double> MeasurementModel_VictoriaPark::*type;
  friend type member_ptr(VPScanTag);
};
struct VPSlbTag {
  typedef double MeasurementModel_VictoriaPark::*type;
  friend type member_ptr(VPSlbTag);
};
template struct MemberAccess<VPScanTag, &MeasurementModel_VictoriaPark::laserscan_>;
template struct MemberAccess<VPSlbTag, &MeasurementModel_VictoriaPark::Slb_>;
}  // namespace
}  // namespace detail

template <>
struct ModelTraits<MeasurementModel_VictoriaPark, KalmanFilter_VictoriaPark> {
  static const bool supported = true;
  static const int lmk_dim = 3, meas_dim = 3, pose_dim = 3;
  template <class MM, class KF>
  static void describe(MM& mm, KF& kf, unsigned nZ, rfsb200_model_desc& d) {
    d = rfsb200_model_desc();
    d.model_id = RFSB200_MODEL_VICTORIAPARK;
    typename MM::TMeasurement::Mat R;
    mm.getNoise(R);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) d.R[i * 3 + j] = R(i, j);
    d.Slb = mm.*member_ptr(detail::VPSlbTag());
    const std::vector<double>& tab = mm.config.probabilityOfDetection_;
    if (tab.empty() || tab.size() > 16)
      throw std::runtime_error("rfs::RBPHDFilter (B200): config.probabilityOfDetection_ must hold 1..16 entries");
    d.pd_table_n = (int32_t)tab.size();
    for (size_t k = 0; k < tab.size(); k++) d.pd_table[k] = tab[k];
    typename MM::TMeasurement z;   /* uniform clutter: the value does not depend on z */
    d.clutter_intensity = mm.clutterIntensity(z, (int)nZ);
    d.clutter_integral = mm.clutterIntensityIntegral((int)nZ);
    d.range_min = mm.config.rangeLimMin_;
    d.range_max = mm.config.rangeLimMax_;
    d.bearing_min = mm.config.bearingLimitMin_;
    d.bearing_max = mm.config.bearingLimitMax_;
    d.buffer_zone_pd = mm.config.bufferZonePd_;
    const std::vector<double>& scan = mm.*member_ptr(detail::VPScanTag());
    d.scan = scan.empty() ? NULL : &scan[0];
    d.scan_n = (int32_t)(scan.size() > 720 ? 720 : scan.size());
    d.innov_thr_range = kf.config.rangeInnovationThreshold_;
    d.innov_thr_bearing = kf.config.bearingInnovationThreshold_;
  }
};
}  // namespace b200

/* ---- motion models for the OPTIONAL device-side propagation (env RFSB200_DEVICE_PROPAGATE=1) -----------------
 * ParticleFilter::propagate() stays host code by default so that unchanged drivers reproduce the reference's random
 * stream; with the switch the poses are propagated by rfsb200_propagate (Philox stream, statistical parity only). */
namespace b200 {
template <class RobotProcessModel>
struct MotionTraits {
  static const bool supported = false;
  template <class PM, class U>
  static void describe(PM&, U&, const TimeStamp&, rfsb200_motion_desc&) {}
};
template <>
struct MotionTraits<MotionModel_Odometry2d> {
  static const bool supported = true;
  template <class PM, class U>
  static void describe(PM& pm, U& u, const TimeStamp& dT, rfsb200_motion_desc& d) {
    d.model_id = RFSB200_MOTION_ODOMETRY2D;
    typename U::Vec uv;
    typename U::Mat uS;
    u.get(uv, uS);
    for (int i = 0; i < 3; i++) d.input[i] = uv(i);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) d.input_cov[i * 3 + j] = uS(i, j);
    d.dt = dT.getTimeAsDouble();
    (void)pm;
  }
};
namespace detail {
namespace {
struct AckHTag { typedef double MotionModel_Ackerman2d::*type; friend type member_ptr(AckHTag); };
struct AckLTag { typedef double MotionModel_Ackerman2d::*type; friend type member_ptr(AckLTag); };
struct AckXTag { typedef double MotionModel_Ackerman2d::*type; friend type member_ptr(AckXTag); };
struct AckYTag { typedef double MotionModel_Ackerman2d::*type; friend type member_ptr(AckYTag); };
template struct MemberAccess<AckHTag, &MotionModel_Ackerman2d::h_>;
template struct MemberAccess<AckLTag, &MotionModel_Ackerman2d::l_>;
template struct MemberAccess<AckXTag, &MotionModel_Ackerman2d::poi_offset_x_>;
template struct MemberAccess<AckYTag, &MotionModel_Ackerman2d::poi_offset_y_>;
}  // namespace
}  // namespace detail
template <>
struct MotionTraits<MotionModel_Ackerman2d> {
  static const bool supported = true;
  template <class PM, class U>
  static void describe(PM& pm, U& u, const TimeStamp& dT, rfsb200_motion_desc& d) {
    d.model_id = RFSB200_MOTION_ACKERMAN2D;
    typename U::Vec uv;
    typename U::Mat uS;
    u.get(uv, uS);
    for (int i = 0; i < 2; i++) d.input[i] = uv(i);
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) d.input_cov[i * 2 + j] = uS(i, j);
    d.dt = dT.getTimeAsDouble();
    d.ackerman_h = pm.*member_ptr(detail::AckHTag());
    d.ackerman_l = pm.*member_ptr(detail::AckLTag());
    d.ackerman_dx = pm.*member_ptr(detail::AckXTag());
    d.ackerman_dy = pm.*member_ptr(detail::AckYTag());
  }
};
}  // namespace b200

template <class RobotProcessModel, class LmkProcessModel, class MeasurementModel, class KalmanFilter>
class RBPHDFilter : public ParticleFilter<RobotProcessModel, MeasurementModel,
                                          GaussianMixture<typename MeasurementModel::TLandmark> > {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;

  typedef typename RobotProcessModel::TState TPose;
  typedef typename RobotProcessModel::TInput TInput;
  typedef typename MeasurementModel::TLandmark TLandmark;
  typedef typename MeasurementModel::TMeasurement TMeasurement;
  typedef GaussianMixture<TLandmark> TGM;
  typedef typename TGM::Gaussian TGaussian;
  typedef b200::ModelTraits<MeasurementModel, KalmanFilter> Traits;
  static const int LD = Traits::lmk_dim;          /**< landmark / measurement dimension */
  static const int NC = LD * (LD + 1) / 2;        /**< unique covariance entries (upper triangle, row-major) */

  /** Same fields, same names as the reference (include/RBPHDFilter.hpp:90-146). */
  struct Config {
    double birthGaussianWeight_;
    uint birthGaussianMeasurementCountThreshold_;
    uint birthGaussianMeasurementCheckThreshold_;
    double birthGaussianMeasurementSupportDist_;
    uint birthGaussianCurrentMeasurementCountThreshold_;
    double newGaussianCreateInnovMDThreshold_;
    int importanceWeightingEvalPointCount_;
    double importanceWeightingEvalPointGuassianWeight_;
    double importanceWeightingMeasurementLikelihoodMDThreshold_;
    double gaussianMergingThreshold_;
    double gaussianMergingCovarianceInflationFactor_;
    double gaussianPruningThreshold_;
    int minUpdatesBeforeResample_;
    int minMeasurementsBeforeResample_;
    bool useClusterProcess_;
  } config;

  /** include/RBPHDFilter.hpp:152-167; the *_wall fields are filled from CUDA events / the device timers (ns).  The update
   *  is ONE fused kernel, so with deviceConfig.stageTiming its wall time is attributed to the phases by the warp cycles
   *  spent in them (rfsb200_get_stage_times); mapUpdate_wall is always the whole device step. */
  struct TimingInfo {
    long long predict_wall, predict_cpu;
    long long mapUpdate_wall, mapUpdate_cpu;
    long long mapUpdate_kf_wall, mapUpdate_kf_cpu;
    long long particleWeighting_wall, particleWeighting_cpu;
    long long mapMerge_wall, mapMerge_cpu;
    long long mapPrune_wall, mapPrune_cpu;
    long long particleResample_wall, particleResample_cpu;
  } timingInfo_;

  class BirthGaussianCandidate : public TLandmark {
   public:
    uint nSupportingMeasurements;
    uint nChecks;
  };

  /** Device-side capacities; change before the first predict()/update() if the defaults do not fit. */
  struct DeviceConfig {
    int gmCapacity;     /**< Gaussians per particle kept between steps */
    int workCapacity;   /**< Gaussians per particle inside one update (inputs + created) */
    int zCapacity;      /**< measurements per update (<= 64) */
    int device;         /**< CUDA device ordinal */
    int precision;      /**< 32 (product) or 64 (verification build) */
    int stageTiming;    /**< != 0 (or RFSB200_STAGE_TIMES=1 in the environment): every update runs the stage-timing build of
                             the kernel (RFSB200_UPDATE_STAGE_TIMES, a few percent slower) and getTimingInfo() fills
                             mapUpdate_kf / particleWeighting / mapMerge / mapPrune as the reference's single-thread run
                             does (include/RBPHDFilter.hpp:1219-1232); 0: the whole step is booked under mapUpdate */
    int hostBirthCandidates;  /**< != 0 (or RFSB200_HOST_BIRTHS=1 in the environment): the candidate lists of addBirthGaussians()
                                   (birthGaussianMeasurementCountThreshold_ != 1) are kept on the host and evaluated with the
                                   live plugin objects (verification); 0: rfsb200_birth_candidates keeps them on the device */
  } deviceConfig;

  RBPHDFilter(int n);
  ~RBPHDFilter();

  LmkProcessModel* getLmkProcessModel() { return lmkModelPtr_; }
  void predict(TInput u, TimeStamp const& dT, bool useModelNoise = true, bool useInputNoise = false,
               bool birthGaussianCheck = true);
  void update(std::vector<TMeasurement>& Z);
  int getGMSize(int i);
  bool getLandmark(const int i, const int m, typename TLandmark::Vec& u, typename TLandmark::Mat& S, double& w);
  KalmanFilter* getKalmanFilter() { return &kf_; }
  void setParticlePose(int i, TPose& p) { *(this->particleSet_[i]) = p; }
  TimingInfo* getTimingInfo();

  /** Hides ParticleFilter::resample (same decisions, same single drand48()); the particle copies on
   *  the host carry poses / trajectories / ids, the maps are moved on the device. */
  bool resample(unsigned int n = 0, bool forceResample = false);

  /** Diagnostics of the last update (n_overflow, n_murty, device time ...). */
  const rfsb200_step_out& lastStep() const { return lastStep_; }
  const char* lastError() const { return ctx_ ? rfsb200_last_error(ctx_) : rfsb200_last_error(NULL); }

 private:
  static_assert(Traits::supported,
                "rfs::b200 has no device descriptor for this MeasurementModel/KalmanFilter pair: "
                "use the reference RBPHDFilter.hpp for it");

  rfsb200_ctx* ctx_;
  int nAlloc_;                 /* particle count the ctx was created for */
  KalmanFilter kf_;            /* the reference keeps one per thread; getKalmanFilter() returns [0] */
  LmkProcessModel* lmkModelPtr_;
  unsigned int nUpdatesSinceResample_;
  unsigned int nMeasurementsSinceResample_;
  bool resampleOccured_;
  std::vector<int> pendingAuxSrc_;   /* Q: parent-id lookup of addBirthGaussians after a resample */
  /* candidate-list births (birthGaussianMeasurementCountThreshold_ != 1), host side as in the reference */
  std::vector<std::list<BirthGaussianCandidate> > birthGaussians_;
  std::vector<int> birthParent_;     /* slot whose candidate list slot i takes over after a resample (or i) */
  bool unusedFresh_;                 /* the device holds unused-measurement masks nobody consumed yet */
  bool overflowWarned_;
  int nUpdateCalls_;                 /* update() calls with a non-empty measurement set (diagnostics) */
  void addBirthGaussiansHost();
  void addBirthGaussiansDevice();
  /* optional device-side ParticleFilter::propagate (RFSB200_DEVICE_PROPAGATE=1; RFSB200_SEED selects the stream) */
  bool devicePropagate_;
  unsigned long long propagateSeed_, propagateCount_;
  void propagateOnDevice(TInput& u, TimeStamp const& dT, bool useModelNoise, bool useInputNoise);
  rfsb200_step_out lastStep_;
  Timer timer_predict_, timer_particleResample_;
  long long ns_update_;
  long long ns_kf_ = 0, ns_weighting_ = 0, ns_merge_ = 0, ns_prune_ = 0;   /* deviceConfig.stageTiming */
  /* one-particle cache for getLandmark */
  int cacheIdx_;
  std::vector<double> cMean_, cCov_, cW_;
  int cN_;
  /* page-locked (rfsb200_host_alloc): rfsb200_update_host lets the kernels read / write such buffers directly */
  b200::detail::PinnedDoubles hPose_, hPoseCov_, hW_, hWout_;

  void ensureCtx();
  void check(int rc, const char* what) {
    if (rc != RFSB200_OK) {
      std::string msg = std::string("rfs::RBPHDFilter (B200): ") + what + ": " + lastError();
      throw std::runtime_error(msg);
    }
  }
  bool gatherPoses();   /* particleSet_ -> hPose_, hPoseCov_, hW_; true if some particle carries a pose covariance */
  void invalidateCache() { cacheIdx_ = -1; }
  /* ParticleFilter declares this pure virtual; the weighting of every particle happens inside
   * rfsb200_update (reference :728-819), so the per-particle host hook has nothing to do */
  void importanceWeighting(const uint) {}
};

////////// Implementation //////////

template <class R, class L, class M, class K>
RBPHDFilter<R, L, M, K>::RBPHDFilter(int n)
    : ParticleFilter<R, M, GaussianMixture<typename M::TLandmark> >(n),
      ctx_(NULL), nAlloc_(0), kf_(), lmkModelPtr_(new L), nUpdatesSinceResample_(0), nMeasurementsSinceResample_(0),
      resampleOccured_(false), unusedFresh_(false), overflowWarned_(false), nUpdateCalls_(0), ns_update_(0), cacheIdx_(-1), cN_(0) {
  kf_ = K(lmkModelPtr_, this->getMeasurementModel());
  for (int i = 0; i < n; i++) this->particleSet_[i]->setData(boost::shared_ptr<TGM>(new TGM()));
  birthGaussians_.resize(n);
  birthParent_.resize(n);
  for (int i = 0; i < n; i++) birthParent_[i] = i;
  /* reference defaults (include/RBPHDFilter.hpp:370-382); the two it leaves uninitialised get
   * defined values */
  config.birthGaussianWeight_ = 0.25;
  config.birthGaussianMeasurementCountThreshold_ = 1;
  config.birthGaussianMeasurementCheckThreshold_ = 1;
  config.birthGaussianMeasurementSupportDist_ = 1;
  config.birthGaussianCurrentMeasurementCountThreshold_ = 1;
  config.gaussianMergingThreshold_ = 0.5;
  config.gaussianMergingCovarianceInflationFactor_ = 1.5;
  config.gaussianPruningThreshold_ = 0.2;
  config.importanceWeightingEvalPointCount_ = 8;
  config.importanceWeightingEvalPointGuassianWeight_ = 0.75;
  config.importanceWeightingMeasurementLikelihoodMDThreshold_ = 3.0;
  config.newGaussianCreateInnovMDThreshold_ = 0.2;
  config.minUpdatesBeforeResample_ = 1;
  config.minMeasurementsBeforeResample_ = 1;
  config.useClusterProcess_ = false;
  deviceConfig.gmCapacity = 192;
  deviceConfig.workCapacity = 256;
  deviceConfig.zCapacity = 64;
  deviceConfig.device = 0;
  deviceConfig.precision = 32;
  deviceConfig.stageTiming = 0;
  deviceConfig.hostBirthCandidates = getenv("RFSB200_HOST_BIRTHS") != NULL ? atoi(getenv("RFSB200_HOST_BIRTHS")) : 0;
  /* unchanged drivers cannot reach deviceConfig: the environment can (verification runs) */
  if (const char* e = getenv("RFSB200_PRECISION")) deviceConfig.precision = atoi(e);
  if (const char* e = getenv("RFSB200_DEVICE")) deviceConfig.device = atoi(e);
  if (const char* e = getenv("RFSB200_GM_CAPACITY")) deviceConfig.gmCapacity = atoi(e);
  if (const char* e = getenv("RFSB200_WORK_CAPACITY")) deviceConfig.workCapacity = atoi(e);
  devicePropagate_ = false;
  if (const char* e = getenv("RFSB200_DEVICE_PROPAGATE")) devicePropagate_ = (atoi(e) != 0) && b200::MotionTraits<R>::supported;
  propagateSeed_ = 1;
  if (const char* e = getenv("RFSB200_SEED")) propagateSeed_ = strtoull(e, NULL, 10);
  propagateCount_ = 0;
  lastStep_ = rfsb200_step_out();
  timingInfo_ = TimingInfo();
}

template <class R, class L, class M, class K>
RBPHDFilter<R, L, M, K>::~RBPHDFilter() {
  for (int i = 0; i < this->nParticles_; i++) this->particleSet_[i]->deleteData();
  delete lmkModelPtr_;
  if (ctx_) rfsb200_destroy(ctx_);
}

template <class R, class L, class M, class K>
void RBPHDFilter<R, L, M, K>::ensureCtx() {
  if (ctx_ && nAlloc_ == this->nParticles_) return;
  if (ctx_) throw std::runtime_error("rfs::RBPHDFilter (B200): the particle count changed after the device state was created");
  rfsb200_dims d = rfsb200_dims();
  d.n_particles = this->nParticles_;
  d.gm_capacity = deviceConfig.gmCapacity;
  d.work_capacity = deviceConfig.workCapacity;
  d.z_capacity = deviceConfig.zCapacity;
  d.lmk_dim = Traits::lmk_dim;
  d.meas_dim = Traits::meas_dim;
  d.pose_dim = Traits::pose_dim;
  d.device = deviceConfig.device;
  d.precision = deviceConfig.precision;
  int rc = rfsb200_create(&ctx_, &d);
  if (rc != RFSB200_OK) {
    ctx_ = NULL;
    check(rc, "rfsb200_create");
  }
  nAlloc_ = this->nParticles_;
  /* empty maps */
  std::vector<int32_t> cnt(nAlloc_, 0);
  check(rfsb200_upload_maps(ctx_, cnt.data(), NULL, NULL, NULL), "rfsb200_upload_maps");
  check(rfsb200_synchronize(ctx_), "rfsb200_synchronize");
}

template <class R, class L, class M, class K>
bool RBPHDFilter<R, L, M, K>::gatherPoses() {
  const int N = this->nParticles_;
  hPose_.resize((size_t)N * 3);
  hPoseCov_.resize((size_t)N * 6);
  hW_.resize(N);
  hWout_.resize(N);
  bool anyCov = false;
  for (int i = 0; i < N; i++) {
    typename TPose::Vec x;
    typename TPose::Mat S;
    this->particleSet_[i]->get(x, S);
    for (int k = 0; k < 3; k++) hPose_[3 * i + k] = x(k);
    double* c = &hPoseCov_[6 * i];
    c[0] = S(0, 0); c[1] = S(0, 1); c[2] = S(0, 2); c[3] = S(1, 1); c[4] = S(1, 2); c[5] = S(2, 2);
    for (int k = 0; k < 6; k++) anyCov = anyCov || (c[k] != 0.0);
    hW_[i] = this->particleSet_[i]->getWeight();
  }
  return anyCov;
}

template <class R, class L, class M, class K>
void RBPHDFilter<R, L, M, K>::predict(TInput u, TimeStamp const& dT, bool useModelNoise, bool useInputNoise,
                                      bool birthGaussianCheck) {
  timer_predict_.resume();
  ensureCtx();
  invalidateCache();
  /* candidate-list form of addBirthGaussians (:1023-1080); "hostBirths" = the births are not rfsb200_predict_maps' */
  const bool hostBirths = birthGaussianCheck && config.birthGaussianMeasurementCountThreshold_ != 1;
  if (hostBirths) {
    if (deviceConfig.hostBirthCandidates) addBirthGaussiansHost();
    else addBirthGaussiansDevice();
  }
  /* landmark process noise: StaticProcessModel::step adds Q only if it was set (ProcessModel.hpp:198) */
  typename TLandmark::Mat Q;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int i = 0; i < Q.rows(); i++)
    for (int j = 0; j < Q.cols(); j++) Q(i, j) = nan;
  lmkModelPtr_->getNoise(Q);
  const bool haveQ = (Q(0, 0) == Q(0, 0));
  double q[NC];
  for (int r = 0, k = 0; r < LD; r++)
    for (int c = r; c < LD; c++, k++) q[k] = haveQ ? Q(r, c) : 0.0;
  /* births use the pose BEFORE the propagation = the pose of the last update, still on the device
   * (addBirthGaussians runs first in the reference too, :425-427) */
  check(rfsb200_predict_maps(ctx_, haveQ ? q : NULL, (birthGaussianCheck && !hostBirths) ? 1 : 0, config.birthGaussianWeight_),
        "rfsb200_predict_maps");
  if (birthGaussianCheck && !hostBirths) unusedFresh_ = false;
  if (devicePropagate_) propagateOnDevice(u, dT, useModelNoise, useInputNoise);
  else this->propagate(u, dT, useModelNoise, useInputNoise, true);
  timer_predict_.stop();
}

template <class R, class L, class M, class K>
void RBPHDFilter<R, L, M, K>::update(std::vector<TMeasurement>& Z) {
  nUpdatesSinceResample_++;
  this->setMeasurements(Z);   /* Z is cleared, as in the reference */
  const unsigned nZ = this->measurements_.size();
  if (nZ == 0) return;        /* include/RBPHDFilter.hpp:451-452 */
  nMeasurementsSinceResample_ += nZ;
  ensureCtx();
  invalidateCache();

  rfsb200_model_desc md;
  Traits::describe(*this->pMeasurementModel_, kf_, nZ, md);
  check(rfsb200_set_model(ctx_, &md), "rfsb200_set_model");
  rfsb200_filter_cfg fc = rfsb200_filter_cfg();
  fc.birth_gaussian_weight = config.birthGaussianWeight_;
  fc.new_gaussian_create_innov_md_threshold = config.newGaussianCreateInnovMDThreshold_;
  fc.eval_point_gaussian_weight = config.importanceWeightingEvalPointGuassianWeight_;
  fc.meas_likelihood_md_threshold = config.importanceWeightingMeasurementLikelihoodMDThreshold_;
  fc.merging_threshold = config.gaussianMergingThreshold_;
  fc.merging_cov_inflation_factor = config.gaussianMergingCovarianceInflationFactor_;
  fc.pruning_threshold = config.gaussianPruningThreshold_;
  fc.eval_point_count = config.importanceWeightingEvalPointCount_ < 0 ? 32 : config.importanceWeightingEvalPointCount_;  /* Q14 */
  /* (the device evaluates at most 32 eval points, MAX_EVAL: the reference's -1 means "as many as qualify"; a count above 32
   *  makes rfsb200_set_filter_cfg fail with RFSB200_EUNSUPPORTED rather than truncate silently) */
  fc.use_cluster_process = config.useClusterProcess_ ? 1 : 0;
  /* quirk Q7: partitions with nR + nC > 8 contribute the sum of Murty's 200 best assignments, as in the reference
   * (include/RBPHDFilter.hpp:904-959); RFSB200_EXACT_SUMS=1 keeps the device's exact sums (such particles are flagged) */
  fc.murty_compat = getenv("RFSB200_EXACT_SUMS") ? 0 : 1;
  check(rfsb200_set_filter_cfg(ctx_, &fc), "rfsb200_set_filter_cfg");

  const bool anyCov = gatherPoses();
  std::vector<double> z((size_t)nZ * LD);
  for (unsigned k = 0; k < nZ; k++) {
    typename TMeasurement::Vec v;
    this->measurements_[k].get(v);
    for (int d = 0; d < LD; d++) z[LD * k + d] = v(d);
  }
  nUpdateCalls_++;
  if (const char* e = getenv("RFSB200_DUMP_UPDATE")) {
    /* diagnostics for unchanged drivers: the complete input of the k-th update() goes to a file that
     * tools/replay_dump.py feeds to the oracle and to the device (layout documented there) */
    if (atoi(e) == nUpdateCalls_) {
      const int N = this->nParticles_;
      std::vector<int32_t> cnt(N);
      const int64_t capTotal = (int64_t)N * (deviceConfig.gmCapacity + 8);
      std::vector<double> mean((size_t)capTotal * LD), cov((size_t)capTotal * NC), w((size_t)capTotal);
      check(rfsb200_download_maps(ctx_, 0, capTotal, cnt.data(), mean.data(), cov.data(), w.data()), "rfsb200_download_maps");
      int64_t total = 0;
      for (int i = 0; i < N; i++) total += cnt[i];
      char name[64];
      snprintf(name, sizeof(name), "rfsb200_dump_%d.bin", nUpdateCalls_);
      FILE* f = fopen(name, "wb");
      if (f) {
        int32_t hdr[6] = {0x52465342, N, LD, (int32_t)nZ, (int32_t)sizeof(md), (int32_t)sizeof(fc)};
        fwrite(hdr, 4, 6, f);
        fwrite(cnt.data(), 4, N, f);
        fwrite(&total, 8, 1, f);
        fwrite(mean.data(), 8, (size_t)total * LD, f);
        fwrite(cov.data(), 8, (size_t)total * NC, f);
        fwrite(w.data(), 8, (size_t)total, f);
        fwrite(hPose_.data(), 8, (size_t)N * 3, f);
        fwrite(hPoseCov_.data(), 8, (size_t)N * 6, f);
        fwrite(hW_.data(), 8, N, f);
        fwrite(z.data(), 8, z.size(), f);
        fwrite(&md, sizeof(md), 1, f);
        int32_t sn = md.scan ? md.scan_n : 0;
        fwrite(&sn, 4, 1, f);
        if (sn) fwrite(md.scan, 8, sn, f);
        fwrite(&fc, sizeof(fc), 1, f);
        fclose(f);
      }
    }
  }
  /* all particles: map update, weighting, merge, prune, weight sums — no normalisation yet, the
   * resampling gate needs the unnormalised weights exactly like the reference */
  /* one ABI call, one synchronisation: poses / pose covariances / weights / Z in, unnormalised weights out */
  const bool stageTiming = (deviceConfig.stageTiming != 0 || getenv("RFSB200_STAGE_TIMES") != NULL) && deviceConfig.precision == 32 &&
                           md.model_id == RFSB200_MODEL_RNGBRG;
  check(rfsb200_update_host(ctx_, hPose_.data(), anyCov ? hPoseCov_.data() : NULL, anyCov ? 2 : 0, hW_.data(), z.data(),
                            (int32_t)nZ, RFSB200_UPDATE_NO_NORMALIZE | (stageTiming ? RFSB200_UPDATE_STAGE_TIMES : 0u),
                            hWout_.data(), NULL, NULL, &lastStep_),
        "rfsb200_update_host");
  ns_update_ += (long long)(lastStep_.elapsed_us * 1000.0);
  if (stageTiming) {
    rfsb200_stage_times st;
    check(rfsb200_get_stage_times(ctx_, &st), "rfsb200_get_stage_times");
    const double loop_ns = st.particles_us * 1000.0;
    ns_kf_ += (long long)(loop_ns * st.share_map_update_kf);
    ns_weighting_ += (long long)(loop_ns * st.share_weighting);
    ns_merge_ += (long long)(loop_ns * st.share_merge);
    ns_prune_ += (long long)(loop_ns * st.share_prune);
  }
  unusedFresh_ = true;
  if (lastStep_.n_overflow > 0 && !overflowWarned_) {
    /* a map outgrew the device capacities: Gaussians were dropped, results differ from the reference from here on */
    fprintf(stderr, "[rfsb200] WARNING: %d particle(s) exceeded the device capacities (gm %d / work %d Gaussians per particle); "
                    "raise RFSB200_GM_CAPACITY / RFSB200_WORK_CAPACITY (<= 1024)\n", lastStep_.n_overflow,
            deviceConfig.gmCapacity, deviceConfig.workCapacity);
    overflowWarned_ = true;
  }
  if (getenv("RFSB200_TRACE"))   /* diagnostics for unchanged drivers: one line per update on stderr */
    fprintf(stderr, "[rfsb200] update nZ=%u gm_in=%lld gm_out=%lld max_out=%d overflow=%d murty=%d device_us=%.1f\n", nZ,
            (long long)lastStep_.gm_total_in, (long long)lastStep_.gm_total_out, lastStep_.gm_max_out, lastStep_.n_overflow,
            lastStep_.n_murty, lastStep_.elapsed_us);
  for (int i = 0; i < this->nParticles_; i++) this->particleSet_[i]->setWeight(hWout_[i]);

  timer_particleResample_.resume();
  resampleOccured_ = false;
  if (nUpdatesSinceResample_ >= (unsigned)config.minUpdatesBeforeResample_ &&
      nMeasurementsSinceResample_ >= (unsigned)config.minMeasurementsBeforeResample_) {
    resampleOccured_ = resample();
  }
  if (resampleOccured_) {
    nUpdatesSinceResample_ = 0;
    nMeasurementsSinceResample_ = 0;
  } else {
    this->normalizeWeights();   /* host copy; the device copy is refreshed by the next update() */
  }
  timer_particleResample_.stop();
}

template <class R, class L, class M, class K>
bool RBPHDFilter<R, L, M, K>::resample(unsigned int n, bool forceResample) {
  /* Restates ParticleFilter::resample (include/ParticleFilter.hpp:399-492) so that the slot each new
   * particle was copied from is known; arithmetic and the single drand48() are the same. */
  const int N = this->nParticles_;
  this->normalizeWeights();
  if (!forceResample) {
    double s2 = 0;
    for (int i = 0; i < N; i++) {
      const double w = this->particleSet_[i]->getWeight();
      s2 += w * w;
    }
    const double nEff = 1.0 / s2;
    if (nEff > this->effNParticles_t_ && nEff / N > this->effNParticles_t_percent_) return false;
  }
  if (n == 0 || n > (unsigned)N) n = N;
  if ((int)n != N) throw std::runtime_error("rfs::RBPHDFilter (B200): resampling to a smaller particle set is not supported");
  ensureCtx();
  invalidateCache();
  const double r01 = drand48();
  const double interval = 1.0 / double(n);
  double samplePoint = interval * r01;
  unsigned idx = 0;
  double cumulative = this->particleSet_[idx]->getWeight();
  std::vector<char> sampled(N, 0);
  std::vector<unsigned> sampledIdx(n, 0);
  for (unsigned i = 0; i < n; i++) {
    while (samplePoint > cumulative) {
      idx++;
      cumulative += this->particleSet_[idx]->getWeight();
    }
    sampledIdx[i] = idx;
    sampled[idx] = 1;
    samplePoint += interval;
  }
  std::vector<int> mapSrc(N);
  for (int i = 0; i < N; i++) mapSrc[i] = i;
  unsigned idxPrev = 0, nextFree = 0;
  for (unsigned i = 0; i < n; i++) {
    idx = sampledIdx[i];
    const bool first = !(i > 0 && idx == idxPrev);
    idxPrev = idx;
    if (idx < n && first) {
      this->particleSet_[idx]->setParentId(this->particleSet_[idx]->getId());
    } else {
      while (nextFree < (unsigned)N && sampled[nextFree] == 1) nextFree++;
      this->particleSet_[nextFree] = this->particleSet_[idx]->copy();   /* host part: pose, trajectory, ids (empty GM) */
      this->particleSet_[nextFree]->setParentId(this->particleSet_[idx]->getId());
      mapSrc[nextFree] = (int)idx;
      nextFree++;
    }
  }
  /* addBirthGaussians (:1001-1011) looks the unused measurements of particle i up through
   * getParentId() INSIDE the loop that also consumes them, in ascending i: a parent slot below i has
   * already been emptied when i copies it (no births for i), a parent slot above i still holds its
   * list.  The ids are those the particle copies carry (not slot numbers any more after the first
   * resampling).  Reproduced as is: -1 = empty. */
  std::vector<int> auxSrc(N);
  for (int i = 0; i < N; i++) {
    const unsigned parent = this->particleSet_[i]->getParentId();
    if (parent == (unsigned)i || parent >= (unsigned)N) auxSrc[i] = i;
    else auxSrc[i] = (parent > (unsigned)i) ? (int)parent : -1;
  }
  for (int i = 0; i < N; i++) {
    const unsigned parent = this->particleSet_[i]->getParentId();
    birthParent_[i] = (parent == (unsigned)i || parent >= (unsigned)N) ? i : (int)parent;
  }
  for (int i = 0; i < N; i++) this->particleSet_[i]->setWeight(1);
  const double one = 1.0;
  check(rfsb200_resample(ctx_, mapSrc.data(), auxSrc.data(), &one), "rfsb200_resample");
  return true;
}

template <class R, class L, class M, class K>
int RBPHDFilter<R, L, M, K>::getGMSize(int i) {
  if (i < 0 || i >= this->nParticles_) return -1;
  ensureCtx();
  if (cacheIdx_ != i) {
    typename TLandmark::Vec u;
    typename TLandmark::Mat S;
    double w;
    getLandmark(i, 0, u, S, w);   /* fills the cache */
  }
  return cN_;
}

template <class R, class L, class M, class K>
bool RBPHDFilter<R, L, M, K>::getLandmark(const int i, const int m, typename TLandmark::Vec& u,
                                          typename TLandmark::Mat& S, double& w) {
  if (i < 0 || i >= this->nParticles_) return false;
  ensureCtx();
  if (cacheIdx_ != i) {
    const int cap = deviceConfig.gmCapacity + 8;
    cMean_.resize((size_t)cap * LD);
    cCov_.resize((size_t)cap * NC);
    cW_.resize(cap);
    int32_t n = 0;
    check(rfsb200_get_map(ctx_, 0, i, cap, &n, cMean_.data(), cCov_.data(), cW_.data()), "rfsb200_get_map");
    cN_ = n;
    cacheIdx_ = i;
  }
  if (m < 0 || m >= cN_) return false;
  for (int d = 0; d < LD; d++) u(d) = cMean_[LD * m + d];
  for (int r = 0, k = 0; r < LD; r++)
    for (int c = r; c < LD; c++, k++) S(r, c) = S(c, r) = cCov_[NC * m + k];
  w = cW_[m];
  return true;
}

/* ParticleFilter::propagate(u, dT, ..., maintainTrajectory = true) (include/ParticleFilter.hpp:322-341) with
 * ProcessModel::sample() evaluated on the device: the host poses go up (drivers may have overwritten them through
 * setParticlePose), rfsb200_propagate draws the noise and steps the motion model for all particles, the poses come
 * back and the trajectory chain is extended exactly as the reference does. */
template <class R, class L, class M, class K>
void RBPHDFilter<R, L, M, K>::propagateOnDevice(TInput& u, TimeStamp const& dT, bool useModelNoise, bool useInputNoise) {
  const int N = this->nParticles_;
  gatherPoses();
  check(rfsb200_set_poses(ctx_, hPose_.data(), NULL, 0, NULL), "rfsb200_set_poses");
  rfsb200_motion_desc md = rfsb200_motion_desc();
  b200::MotionTraits<R>::describe(*this->pProcessModel_, u, dT, md);
  typename TPose::Mat Q = TPose::Mat::Zero();
  this->pProcessModel_->getNoise(Q);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) md.Q[i * 3 + j] = Q(i, j);
  md.use_model_noise = useModelNoise ? 1 : 0;
  md.use_input_noise = useInputNoise ? 1 : 0;
  md.seed = propagateSeed_;
  md.step_counter = propagateCount_++;
  check(rfsb200_propagate(ctx_, &md), "rfsb200_propagate");
  check(rfsb200_get_poses(ctx_, hPose_.data()), "rfsb200_get_poses");
  const bool withQ = useModelNoise && (Q != TPose::Mat::Zero());
  for (int i = 0; i < N; i++) {
    typename ParticleFilter<R, M, TGM>::pParticle p_km(new Particle<TPose, TGM>());
    *p_km = *(this->particleSet_[i]);
    this->particleSet_[i]->prev = p_km;
    typename TPose::Vec x;
    for (int k = 0; k < 3; k++) x(k) = hPose_[3 * i + k];
    TPose x_k;
    TimeStamp t_k = this->particleSet_[i]->getTime() + dT;
    x_k.set(x, t_k);
    if (withQ) x_k.setCov(Q);
    *(this->particleSet_[i]) = x_k;
  }
}

/* addBirthGaussians() of the reference (include/RBPHDFilter.hpp:1000-1080) for the candidate-list
 * configuration: the unused measurements and nLandmarksInFOV_ of the last update come from the device,
 * the candidate lists live here, the plugin objects are the live host ones, and the Gaussians that
 * become real go to the device maps in the order the reference would have appended them. */
template <class R, class L, class M, class K>
void RBPHDFilter<R, L, M, K>::addBirthGaussiansHost() {
  const int N = this->nParticles_;
  std::vector<uint64_t> mask(N, 0);
  std::vector<int32_t> nfov(N, 0);
  if (unusedFresh_) check(rfsb200_get_unused(ctx_, mask.data(), nfov.data()), "rfsb200_get_unused");
  else check(rfsb200_get_unused(ctx_, NULL, nfov.data()), "rfsb200_get_unused");
  std::vector<int32_t> addCount(N, 0);
  std::vector<double> aMean, aCov, aW;
  const unsigned nZ = this->measurements_.size();
  for (int i = 0; i < N; i++) {
    if (resampleOccured_ && birthParent_[i] != i) birthGaussians_[i] = birthGaussians_[birthParent_[i]];   /* :1005-1011 */
    std::list<BirthGaussianCandidate>& cand = birthGaussians_[i];
    TPose x = *(this->particleSet_[i]);
    auto addReal = [&](TLandmark& lm) {
      typename TLandmark::Vec u;
      typename TLandmark::Mat S;
      lm.get(u, S);
      for (int d = 0; d < LD; d++) aMean.push_back(u(d));
      for (int r = 0; r < LD; r++)
        for (int c = r; c < LD; c++) aCov.push_back(S(r, c));
      aW.push_back(config.birthGaussianWeight_);
      addCount[i]++;
    };
    for (int zi = (int)nZ - 1; zi >= 0; zi--) {   /* the reference pops the list from the back: descending index */
      if (!((mask[i] >> zi) & 1ull)) continue;
      TMeasurement unused_z = this->measurements_[zi];
      bool isNew = true;
      for (typename std::list<BirthGaussianCandidate>::iterator it = cand.begin(); it != cand.end(); it++) {
        TMeasurement z_exp;
        this->pMeasurementModel_->measure(x, *it, z_exp);
        const double d2 = z_exp.mahalanobisDist2(unused_z);
        if (d2 <= config.birthGaussianMeasurementSupportDist_ * config.birthGaussianMeasurementSupportDist_) {
          kf_.correct(x, unused_z, *it, *it);
          (it->nSupportingMeasurements)++;
          isNew = false;
          break;
        }
      }
      if (isNew) {
        BirthGaussianCandidate c;
        c.nSupportingMeasurements = 1;
        c.nChecks = 0;
        this->pMeasurementModel_->inverseMeasure(x, unused_z, c);
        if (config.birthGaussianMeasurementCountThreshold_ == 1 ||
            (unsigned)nfov[i] <= config.birthGaussianCurrentMeasurementCountThreshold_)
          addReal(c);
        else
          cand.push_back(c);
      }
    }
    /* :1056-1075.  When the LAST candidate of the list is erased the reference leaves its while loop with
     * it == end() and then increments it in the for statement; with libstdc++'s circular list that lands on
     * begin() again, i.e. the remaining candidates get another pass (nChecks++ each).  Reproduced as built. */
    typename std::list<BirthGaussianCandidate>::iterator it = cand.begin();
    while (it != cand.end()) {
      it->nChecks++;
      bool wrapped = false;
      while (it->nSupportingMeasurements >= config.birthGaussianMeasurementCountThreshold_ ||
             it->nChecks > config.birthGaussianMeasurementCheckThreshold_ ||
             (unsigned)nfov[i] <= config.birthGaussianCurrentMeasurementCountThreshold_) {
        if (it->nSupportingMeasurements >= config.birthGaussianMeasurementCountThreshold_) addReal(*it);
        else if ((unsigned)nfov[i] <= config.birthGaussianCurrentMeasurementCountThreshold_) addReal(*it);
        it = cand.erase(it);
        if (it != cand.end()) it->nChecks++;
        else { wrapped = true; break; }
      }
      if (wrapped) it = cand.begin();
      else ++it;
    }
  }
  check(rfsb200_append_gaussians(ctx_, addCount.data(), aMean.empty() ? NULL : &aMean[0], aCov.empty() ? NULL : &aCov[0],
                                 aW.empty() ? NULL : &aW[0]),
        "rfsb200_append_gaussians");
  if (unusedFresh_) check(rfsb200_predict_maps(ctx_, NULL, -1, 0.0), "rfsb200_predict_maps(clear)");   /* masks consumed */
  unusedFresh_ = false;
}

/* The same on the device (rfsb200_birth_candidates): the candidate lists, the unused-measurement masks,
 * nLandmarksInFOV_ and the maps stay where the update left them; only the poses (which a driver may have
 * overwritten through setParticlePose since the update) and, after a resampling, the parent slots go up. */
template <class R, class L, class M, class K>
void RBPHDFilter<R, L, M, K>::addBirthGaussiansDevice() {
  if (unusedFresh_) {   /* (with every mask consumed only the check pass runs, which does not look at the poses) */
    const bool anyCov = gatherPoses();
    check(rfsb200_set_poses(ctx_, &hPose_[0], anyCov ? &hPoseCov_[0] : NULL, anyCov ? 2 : 0, NULL), "rfsb200_set_poses");
  }
  rfsb200_birth_cfg b;
  memset(&b, 0, sizeof(b));
  b.birth_weight = config.birthGaussianWeight_;
  b.support_dist = config.birthGaussianMeasurementSupportDist_;
  b.count_threshold = config.birthGaussianMeasurementCountThreshold_;
  b.check_threshold = config.birthGaussianMeasurementCheckThreshold_;
  b.current_count_threshold = config.birthGaussianCurrentMeasurementCountThreshold_;
  std::vector<int32_t> parent;
  if (resampleOccured_) parent.assign(birthParent_.begin(), birthParent_.end());
  check(rfsb200_birth_candidates(ctx_, &b, resampleOccured_ ? parent.data() : NULL), "rfsb200_birth_candidates");
  unusedFresh_ = false;
}

template <class R, class L, class M, class K>
typename RBPHDFilter<R, L, M, K>::TimingInfo* RBPHDFilter<R, L, M, K>::getTimingInfo() {
  timer_predict_.elapsed(timingInfo_.predict_wall, timingInfo_.predict_cpu);
  timer_particleResample_.elapsed(timingInfo_.particleResample_wall, timingInfo_.particleResample_cpu);
  /* the whole device step is booked under mapUpdate, like the reference does when it runs with more
   * than one thread (include/RBPHDFilter.hpp:466-468,521-523) */
  timingInfo_.mapUpdate_wall = ns_update_;
  timingInfo_.mapUpdate_cpu = 0;
  timingInfo_.mapUpdate_kf_wall = ns_kf_;            /* 0 unless deviceConfig.stageTiming */
  timingInfo_.particleWeighting_wall = ns_weighting_;
  timingInfo_.mapMerge_wall = ns_merge_;
  timingInfo_.mapPrune_wall = ns_prune_;
  return &timingInfo_;
}

}  // namespace rfs

#endif
