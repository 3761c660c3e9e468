/*
 * rfsb200.h — C ABI of the B200-native PHD measurement-update path.
 *
 * This is the drop-in boundary for ONE path of kykleung/RFS-SLAM: the body of
 * rfs::RBPHDFilter<...>::update() (reference include/RBPHDFilter.hpp:444-541), i.e.
 * the "#pragma omp parallel" region :469-520 (updateMap :543-725, importanceWeighting
 * :728-819 with rfsMeasurementLikelihood :821-997, GaussianMixture::merge / prune
 * include/GaussianMixture.hpp:394-521) plus the weight normalisation that follows it
 * (include/ParticleFilter.hpp:352-363, ESS :406-411).
 *
 * Everything crossing this boundary is plain C: pointers, sizes, PODs.  Host-side
 * numbers are fp64 (the reference's arithmetic type); the device keeps fp32
 * particle-major SoA (precision = 32, the product) or fp64 (precision = 64, the
 * verification build used for structural parity).  No torch / Eigen / Boost types.
 *
 * Conventions
 *   - All functions return 0 (RFSB200_OK) or a negative RFSB200_E* code and never
 *     throw.  rfsb200_last_error() gives a human-readable message.
 *   - The caller owns every host buffer.  The ctx owns device memory and its stream.
 *   - One ctx = one GPU = one shard of particles; a ctx is driven by one host thread
 *     at a time (the reference calls update() from one thread; its OpenMP team is
 *     what the GPU replaces).
 *   - Symmetric covariances are passed as their upper triangle, row-major:
 *     2-D: (xx, xy, yy); 3-D: (xx, xy, xz, yy, yz, zz).
 *   - Gaussian mixtures are passed "packed": count[i] Gaussians of particle i follow
 *     those of particle i-1 with no padding.
 */
#ifndef RFSB200_H
#define RFSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFSB200_ABI_VERSION 2

/* ---- error codes -------------------------------------------------------------- */
#define RFSB200_OK            0
#define RFSB200_EINVAL       -1  /* bad argument / NULL pointer / size out of range   */
#define RFSB200_ECUDA        -2  /* a CUDA runtime call failed (see last_error)       */
#define RFSB200_ENOMEM       -3  /* host or device allocation failed                  */
#define RFSB200_ECAPACITY    -4  /* an input does not fit the capacities of the ctx   */
#define RFSB200_EUNSUPPORTED -5  /* model / configuration not implemented on device   */
#define RFSB200_ESTATE       -6  /* call sequence error (e.g. update before set_model)*/
#define RFSB200_ENODEVICE    -7  /* no CUDA device: there is NO CPU fallback          */

/* ---- measurement-model ids (plugin classes of the reference) ---------------- */
#define RFSB200_MODEL_RNGBRG 1   /* rfs::MeasurementModel_RngBrg + KalmanFilter_RngBrg: 2-D landmarks (x, y),
                                    2-D measurements (range, bearing)                              */
#define RFSB200_MODEL_VICTORIAPARK 2 /* rfs::MeasurementModel_VictoriaPark + KalmanFilter_VictoriaPark
                                    (BASELINE config 5): 3-D landmarks (x, y, diameter), 3-D
                                    measurements (range, bearing, diameter); P_D from the lidar scan */

/* ---- update flags -------------------------------------------------------------- */
#define RFSB200_UPDATE_DEFAULT    0u
#define RFSB200_UPDATE_NO_COMMIT  1u  /* compute the step into the back buffers but keep the
                                         current state as the state (benchmark / replay aid) */
#define RFSB200_UPDATE_FUSED_ALLREDUCE 4u /* after rfsb200_comm_connect: the update kernel itself adds the
                                         [sum w, sum w^2] pairs of all ranks through peer memory (NVLink)
                                         and normalises, in the same launch; every rank must call
                                         rfsb200_update with this flag the same number of times      */
#define RFSB200_UPDATE_STAGE_TIMES 8u   /* run the stage-timing build of the update kernel (same results up to the last
                                         bits of fp32: another instantiation of the same source): fills what
                                         rfsb200_get_stage_times() returns — the reference's per-phase TimingInfo
                                         (include/RBPHDFilter.hpp:152-167,1219-1232).  fp32 build; a measurement
                                         aid, a few percent slower than the product kernel                   */
#define RFSB200_UPDATE_DEFER_NORMALIZE 16u /* with RFSB200_UPDATE_FUSED_ALLREDUCE on more than one rank: the kernel sends
                                         its [sum w, sum w^2] pair to the peers and ends WITHOUT waiting for theirs; the
                                         particle weights stay unnormalised until somebody needs them: the next
                                         rfsb200_update picks the pairs up from its mailbox during set-up
                                         — they arrived a step ago — and divides the weights by the total while it loads
                                         them (same division, same operands: bit-identical to the eager step); any other
                                         consumer (rfsb200_get_weights, _update_host, _resample, _export_particles, ...)
                                         or rfsb200_comm_resolve() runs the open normalisation first.  No rank waits for
                                         the slowest one inside a step any more; ranks may drift by one step.  Meant for
                                         the steps between two resampling decisions (minUpdatesBeforeResample_), where the
                                         host does not look at the weights.  rfsb200_step_out of such a step holds the
                                         LOCAL sums.  With RFSB200_UPDATE_NO_COMMIT the open weights are those of the
                                         back buffer: a reader of that result closes them, the next update overwrites
                                         them and only picks the pairs up.                                            */
#define RFSB200_UPDATE_NO_NORMALIZE 2u /* stop after the local [sum w, sum w^2] reduction so the
                                         caller can all-reduce rfsb200_weight_sums_device() across
                                         GPUs and then call rfsb200_normalize()               */

typedef struct rfsb200_ctx rfsb200_ctx;

/* Sizes fixed for the life of a ctx. */
typedef struct rfsb200_dims {
  int32_t n_particles;    /* particles in THIS shard                                        */
  int32_t gm_capacity;    /* max Gaussians per particle held in HBM between steps (<= 1024) */
  int32_t work_capacity;  /* max Gaussians per particle inside a step: inputs + Gaussians
                             created by the corrector (>= gm_capacity, <= 1024)            */
  int32_t z_capacity;     /* max measurements per update (<= 64)                            */
  int32_t lmk_dim;        /* 2 (RngBrg) or 3 (VictoriaPark)                                 */
  int32_t meas_dim;       /* = lmk_dim                                                      */
  int32_t pose_dim;       /* 3                                                              */
  int32_t device;         /* CUDA device ordinal                                            */
  int32_t precision;      /* 32 = fp32 SoA (product), 64 = fp64 SoA (verification)          */
  int32_t reserved[7];
} rfsb200_dims;

/*
 * POD mirror of the live plugin objects, refreshed by the host before every update()
 * because the reference's configs are mutable public structs
 * (MeasurementModel_RngBrg::config include/MeasurementModel_RngBrg.hpp:65-71,
 *  KalmanFilter_RngBrg::config include/KalmanFilter_RngBrg.hpp:55-60,
 *  MeasurementModel::R_ via getNoise include/MeasurementModel.hpp).
 */
typedef struct rfsb200_model_desc {
  int32_t model_id;            /* RFSB200_MODEL_*                                               */
  int32_t reserved0;
  double  R[9];                /* measurement noise, row-major meas_dim x meas_dim              */
  double  Pd;                  /* config.probabilityOfDetection_                                */
  double  clutter_intensity;   /* clutterIntensity(z, nZ): uniform (RngBrg: config.uniformClutterIntensity_) */
  double  clutter_integral;    /* clutterIntensityIntegral(nZ) as evaluated by the host plugin  */
  double  range_min;           /* config.rangeLimMin_                                           */
  double  range_max;           /* config.rangeLimMax_                                           */
  double  range_buffer;        /* config.rangeLimBuffer_                                        */
  double  innov_thr_range;     /* KalmanFilter_RngBrg config.rangeInnovationThreshold_  (<=0 off)*/
  double  innov_thr_bearing;   /* KalmanFilter_RngBrg config.bearingInnovationThreshold_ (<=0 off)*/
  /* ---- RFSB200_MODEL_VICTORIAPARK only (MeasurementModel_VictoriaPark::Config,
   *      include/MeasurementModel_VictoriaPark.hpp:148-156; setNoise(R, Slb) src/...VictoriaPark.cpp:66-72;
   *      setLaserScan :268-281).  For this model R is the full 3x3, Pd / range_buffer are unused,
   *      clutter_intensity = expectedClutterNumber_ / FoVArea(scan) and clutter_integral =
   *      expectedClutterNumber_, both as the host plugin evaluates them. ---- */
  double  bearing_min;         /* config.bearingLimitMin_ (rad)                                 */
  double  bearing_max;         /* config.bearingLimitMax_ (rad)                                 */
  double  Slb;                 /* laser bearing variance: S_dd = P_dd + R_dd + r^2 * Slb        */
  double  buffer_zone_pd;      /* config.bufferZonePd_                                          */
  double  pd_table[16];        /* config.probabilityOfDetection_[k], k = lidar points on the trunk */
  int32_t pd_table_n;          /* entries of pd_table (1..16)                                   */
  int32_t scan_n;              /* entries of scan (<= 720)                                      */
  const double* scan;          /* laserscan_ (host pointer, copied by rfsb200_set_model); ranges
                                  at half-degree steps, 0 = no return                           */
  double  reserved[4];
} rfsb200_model_desc;

/* Mirror of rfs::RBPHDFilter::Config (include/RBPHDFilter.hpp:90-146), update-path fields only. */
typedef struct rfsb200_filter_cfg {
  double  birth_gaussian_weight;                 /* birthGaussianWeight_ (used by updateMap :692) */
  double  new_gaussian_create_innov_md_threshold;/* newGaussianCreateInnovMDThreshold_            */
  double  eval_point_gaussian_weight;            /* importanceWeightingEvalPointGuassianWeight_   */
  double  meas_likelihood_md_threshold;          /* importanceWeightingMeasurementLikelihoodMDThreshold_ */
  double  merging_threshold;                     /* gaussianMergingThreshold_                     */
  double  merging_cov_inflation_factor;          /* gaussianMergingCovarianceInflationFactor_     */
  double  pruning_threshold;                     /* gaussianPruningThreshold_ (must be > 0)       */
  int32_t eval_point_count;                      /* importanceWeightingEvalPointCount_ (0..32)    */
  int32_t use_cluster_process;                   /* useClusterProcess_                            */
  int32_t assignment_sum_method;                 /* 0 = enumeration order of the reference
                                                    (PermutationLexicographic / Murty-200),
                                                    1 = matrix-permanent identity (MatPerm path) for
                                                    partitions of up to 11 members whose Ryser-type sum is
                                                    numerically safe (a-posteriori cancellation bound),
                                                    the subset DP otherwise                           */
  int32_t murty_compat;                          /* != 0 (multi-feature weighting): reproduce quirk Q7 — a partition with
                                                    nR + nC > 8 contributes the sum of Murty's 200 best assignments
                                                    (include/RBPHDFilter.hpp:904-959) instead of the exact sum: the device
                                                    writes such partitions out, the host replaces their sums (own k-best
                                                    enumeration, csrc/murty_compat.hpp) before the weights are added up.
                                                    The update is synchronous then, and not available together with
                                                    RFSB200_UPDATE_FUSED_ALLREDUCE on more than one rank.  0: exact sums,
                                                    flag bit 2 marks the particles whose weight the reference truncates */
  int32_t reserved_i[2];
  double  reserved[6];
} rfsb200_filter_cfg;

/* Scalars that come back from one update (filled only when the pointer is non-NULL,
 * which makes the call synchronous). */
typedef struct rfsb200_step_out {
  double  sum_w;              /* sum_i w_i of the unnormalised weights of this shard (after all-reduce if the caller did one) */
  double  sum_w2;             /* sum_i w_i^2                                                                             */
  double  n_eff;              /* (sum_w)^2 / sum_w2 : effective particle count (include/ParticleFilter.hpp:406-411)        */
  int64_t gm_total_in;        /* sum_i nM_in(i)                                                                          */
  int64_t gm_total_out;       /* sum_i nM_out(i)                                                                         */
  int32_t gm_max_out;         /* max_i nM_out(i)                                                                         */
  int32_t n_overflow;         /* particles whose work set exceeded work_capacity or output exceeded gm_capacity           */
  int32_t n_murty;            /* particles that took the k-best (Murty, nR+nC>8) branch of rfsMeasurementLikelihood        */
  int32_t n_launches;         /* kernels launched by this call                                                           */
  float   elapsed_us;         /* device time of the step (CUDA events on the ctx stream)                                  */
  int32_t n_merge_redo;       /* particles recomputed with the exhaustive merge (interacting merge clusters)              */
  int32_t reserved[6];
} rfsb200_step_out;

/* ---- lifetime ------------------------------------------------------------------ */
int rfsb200_abi_version(void);
int rfsb200_device_count(void);
int rfsb200_create(rfsb200_ctx** out, const rfsb200_dims* dims);
int rfsb200_destroy(rfsb200_ctx* ctx);
const char* rfsb200_last_error(const rfsb200_ctx* ctx); /* ctx may be NULL: last creation error */

/* external != 0: run on the caller's CUDA stream (cudaStream_t as void*; NULL is the legacy
 * default stream), e.g. torch's current stream, so that the caller's collectives and events
 * are ordered with the kernels.  external == 0 restores the ctx-owned stream. */
int rfsb200_set_stream(rfsb200_ctx* ctx, void* cuda_stream, int external);
int rfsb200_synchronize(rfsb200_ctx* ctx);

/* ---- configuration (replaces the plugin virtual calls inside updateMap) --------- */
int rfsb200_set_model(rfsb200_ctx* ctx, const rfsb200_model_desc* desc);
int rfsb200_set_filter_cfg(rfsb200_ctx* ctx, const rfsb200_filter_cfg* cfg);

/* ---- state in --------------------------------------------------------------------
 * Replaces the host AoS-of-pointers maps (GaussianMixture::gList_,
 * include/GaussianMixture.hpp:60-64,191) by device SoA.  Host pointers may be pageable
 * or pinned (rfsb200_host_alloc); copies run on the ctx stream. */
int rfsb200_upload_maps(rfsb200_ctx* ctx, const int32_t* count /*[N]*/,
                        const double* mean /*[sum count][lmk_dim]*/,
                        const double* cov  /*[sum count][lmk_dim*(lmk_dim+1)/2]*/,
                        const double* w    /*[sum count]*/);
/* pose_cov_mode: 0 = none (zero pose covariance), 1 = one shared [6], 2 = per particle [N][6]
 * (Q1: the reference adds Hx*Sigma_x*Hx^T to S, src/MeasurementModel_RngBrg.cpp:102). */
int rfsb200_set_poses(rfsb200_ctx* ctx, const double* pose /*[N][pose_dim]*/,
                      const double* pose_cov, int pose_cov_mode,
                      const double* weight /*[N] or NULL = keep*/);

/* ---- the hot path ----------------------------------------------------------------
 * One call = RBPHDFilter::update() for all particles of the shard
 * (include/RBPHDFilter.hpp:444-541 minus the resampling itself):
 * map update, particle weighting (SC-PHD or multi-feature), merge, prune,
 * [sum w, sum w^2] reduction and (unless NO_NORMALIZE) weight normalisation.
 * nZ == 0 returns immediately and changes nothing (reference :451-452).
 * out == NULL: fully asynchronous on the ctx stream. */
int rfsb200_update(rfsb200_ctx* ctx, const double* Z /*[nZ][meas_dim]*/, int32_t nZ,
                   uint32_t flags, rfsb200_step_out* out);

/* One host-facing step = rfsb200_set_poses + rfsb200_update + rfsb200_get_weights + rfsb200_get_unused in ONE
 * call with ONE synchronisation: the copies in, the kernels and the copies out are queued back to back on the
 * ctx stream.  This is what RBPHDFilter::update() moves per step when the maps stay resident: poses / pose
 * covariances / particle weights / Z in, particle weights (normalised unless NO_NORMALIZE), the
 * unused-measurement masks and nLandmarksInFOV_ out.  w_out / unused_out / n_in_fov_out / out may be NULL.
 * Page-locked buffers (rfsb200_host_alloc, cudaMallocHost, cudaHostRegister): nothing is copied at all — one
 * conversion kernel reads pose / pose_cov / weight from the caller's memory over PCIe (Z travels as a kernel
 * argument) and the update kernel stores the results straight into w_out / unused_out / n_in_fov_out, two launches
 * and one synchronisation per step.  If ANY of the large buffers is pageable (or RFSB200_ZERO_COPY=0 is set when the
 * ctx is created) the same bytes are staged through copies; the results are bit-identical either way. */
int rfsb200_update_host(rfsb200_ctx* ctx, const double* pose /*[N][3]*/, const double* pose_cov, int pose_cov_mode,
                        const double* weight /*[N] or NULL = keep*/, const double* Z, int32_t nZ, uint32_t flags,
                        double* w_out /*[N]*/, uint64_t* unused_out /*[N]*/, int32_t* n_in_fov_out /*[N]*/,
                        rfsb200_step_out* out);

/* ---- the callers either side of the hot path -------------
 * rfsb200_predict_maps = the map part of RBPHDFilter::predict() (include/RBPHDFilter.hpp:415-442):
 *   add_births != 0: addBirthGaussians() in its direct form (:1013-1052 with
 *     birthGaussianMeasurementCountThreshold_ == 1, the 2-D simulator's setting): for every particle,
 *     every measurement of the LAST update that no Gaussian used (unused_measurements_, consumed from
 *     the back, i.e. descending index) becomes a Gaussian at inverseMeasure(pose of the last update, z)
 *     (src/MeasurementModel_RngBrg.cpp:117-136) with weight birth_weight; the masks are cleared.
 *   Q_lmk != NULL: StaticProcessModel::staticStep on every Gaussian incl. the births: P += Q
 *     (include/ProcessModel.hpp:195-219); Q_lmk = upper triangle (xx, xy, yy), already scaled by the
 *     caller exactly as it scales the plugin's Q.
 *   add_births < 0: no births, but the unused-measurement masks are cleared (the host consumed them, see
 *     rfsb200_append_gaussians).
 * For RFSB200_MODEL_VICTORIAPARK the births are at MeasurementModel_VictoriaPark::inverseMeasure
 * (src/MeasurementModel_VictoriaPark.cpp:75-102) and Q_lmk has 6 entries (xx, xy, xz, yy, yz, zz).
 * Runs on the committed state, in place.  Births that do not fit gm_capacity set flag bits 1 and 8 of the particle;
 * the next update carries the drop into its own result: flag bit 1 and rfsb200_step_out::n_overflow. */
int rfsb200_predict_maps(rfsb200_ctx* ctx, const double* Q_lmk /*[3] / [6] or NULL*/, int32_t add_births,
                         double birth_weight);

/* Appends Gaussians to the committed maps: count[i] Gaussians (packed like rfsb200_upload_maps) go behind the
 * existing ones of particle i.  This is GaussianMixture::addGaussian (include/GaussianMixture.hpp:267-284)
 * for the birth Gaussians the HOST decides on — the candidate-list form of addBirthGaussians()
 * (include/RBPHDFilter.hpp:1023-1080, used when birthGaussianMeasurementCountThreshold_ != 1, e.g. the
 * Victoria Park configuration) keeps its per-particle candidate lists on the host.  Gaussians that do not
 * fit gm_capacity are dropped and set flag bits 1 and 8 of the particle (counted by the next update, see above). */
int rfsb200_append_gaussians(rfsb200_ctx* ctx, const int32_t* count /*[N]*/, const double* mean,
                             const double* cov, const double* w);

/* addBirthGaussians() in its candidate-list form ON THE DEVICE (include/RBPHDFilter.hpp:1000-1080; the setting of
 * cfg/rbphdslam_VictoriaPark_artificialClutter.xml:71-77, birthGaussianMeasurementCountThreshold_ != 1) for the two
 * built-in plugin sets.  The ctx keeps birthGaussians_[i] (up to RFSB200_BIRTH_CAND_CAP candidates per particle: mean,
 * covariance, nSupportingMeasurements, nChecks).  One call = one addBirthGaussians(): per particle, the unused
 * measurements of the last update (descending index) support the first candidate within support_dist (Mahalanobis
 * distance of measure(pose, candidate), which then takes KalmanFilter::correct) or open a new candidate at
 * inverseMeasure(pose, z) — a real Gaussian straight away if count_threshold == 1 or nLandmarksInFOV_[i] <=
 * current_count_threshold; then every candidate is checked (nChecks++) and leaves the list with enough support or few
 * landmarks in view (appended to the map with birth_weight) or when nChecks > check_threshold (dropped), including the
 * second pass the reference makes when the last list element is erased.  The masks are cleared.
 * parent: NULL, or after a resampling (resampleOccured_) the parent slot of every particle (:1005-1011): a copy takes
 *   the list its parent slot holds when the reference's ascending loop reaches it — the parent's list as it was for a
 *   higher slot, the parent's list after its own turn for a lower one (which may itself be a copy from a still lower
 *   slot: parent ids are not slot numbers after the first resampling; chains are resolved level by level, one launch
 *   per level).  (The masks were routed by rfsb200_resample's aux_src.)  rfsb200_export/import_particles do not carry
 *   candidate lists.
 * Gaussians beyond gm_capacity set flag bits 1 and 8, a candidate beyond RFSB200_BIRTH_CAND_CAP is dropped with flag
 * bit 32; like bit 8 it is carried into the next update's result (flag bit 1, rfsb200_step_out::n_overflow).  Arithmetic is fp64 whatever the ctx precision; the appended Gaussians are rounded to the map's type. */
#define RFSB200_BIRTH_CAND_CAP 64
typedef struct rfsb200_birth_cfg {
  double   birth_weight;             /* birthGaussianWeight_                              */
  double   support_dist;             /* birthGaussianMeasurementSupportDist_              */
  uint32_t count_threshold;          /* birthGaussianMeasurementCountThreshold_           */
  uint32_t check_threshold;          /* birthGaussianMeasurementCheckThreshold_           */
  uint32_t current_count_threshold;  /* birthGaussianCurrentMeasurementCountThreshold_    */
  uint32_t reserved;
} rfsb200_birth_cfg;
int rfsb200_birth_candidates(rfsb200_ctx* ctx, const rfsb200_birth_cfg* cfg, const int32_t* parent /*[N] or NULL*/);
/* The candidate lists, [N][RFSB200_BIRTH_CAND_CAP] slots: mean [..][landmark_dim], cov [..][upper triangle],
 * support / checks [..]; n[i] = length of the list of particle i.  Arrays other than n may be NULL in the getter. */
int rfsb200_get_birth_candidates(rfsb200_ctx* ctx, int32_t* n /*[N]*/, double* mean, double* cov, int32_t* support,
                                 int32_t* checks);
int rfsb200_set_birth_candidates(rfsb200_ctx* ctx, const int32_t* n /*[N]*/, const double* mean, const double* cov,
                                 const int32_t* support, const int32_t* checks);

/* ---- particle propagation ------------------------------------------------
 * ParticleFilter::propagate() (include/ParticleFilter.hpp:322-341) = ProcessModel::sample() for every particle
 * (include/ProcessModel.hpp:125-150) on the device, in place on the poses of the ctx (fp64 copy kept next to the
 * T copy the update kernel reads):
 *   use_input_noise: the input is drawn from N(input, input_cov) per particle (RandomVec::sample, Cholesky factor);
 *   step():          MotionModel_Odometry2d (src/ProcessModel_Odometry2D.cpp:41-89) or MotionModel_Ackerman2d
 *                    (src/ProcessModel_Ackerman2D.cpp:49-77);
 *   use_model_noise and Q != 0: N(0, Q) is added and the pose covariance of every particle becomes Q (Q1), otherwise
 *                    the poses carry no covariance afterwards (the reference's x_k is default-constructed).
 * Random numbers: Philox4x32-10 keyed by `seed`, counter (particle, step_counter): the stream does not depend on the
 * launch shape.  The reference draws from one host mt19937 seeded by an unseeded rand(), so parity with it is
 * statistical (moments) for the noise and exact for step(). */
#define RFSB200_MOTION_ODOMETRY2D 1   /* input = (dx, dy, dtheta) in the frame of the previous pose          */
#define RFSB200_MOTION_ACKERMAN2D 2   /* input = (velocity, steering angle)                                  */
typedef struct rfsb200_motion_desc {
  int32_t model_id;          /* RFSB200_MOTION_*                                                             */
  int32_t use_model_noise;   /* useAdditiveWhiteGaussianNoise                                                */
  int32_t use_input_noise;   /* useInputWhiteGaussianNoise                                                   */
  int32_t reserved_i;
  double  Q[9];              /* ProcessModel::Q_, 3x3 row-major (already scaled by the caller as it scales Q_) */
  double  input[3];          /* Odometry2d: dx, dy, dtheta; Ackerman2d: velocity, steering, unused           */
  double  input_cov[9];      /* covariance of the input, row-major n_in x n_in (3x3 / 2x2)                   */
  double  dt;                /* dT (Ackerman2d only)                                                         */
  double  ackerman_h, ackerman_l, ackerman_dx, ackerman_dy;   /* h_, l_, poi_offset_x_, poi_offset_y_          */
  uint64_t seed;
  uint64_t step_counter;
  double  reserved[4];
} rfsb200_motion_desc;
int rfsb200_propagate(rfsb200_ctx* ctx, const rfsb200_motion_desc* m);
int rfsb200_get_poses(rfsb200_ctx* ctx, double* pose /*[N][3]*/);

/* The data movement of ParticleFilter::resample() (include/ParticleFilter.hpp:446-479): particle i
 * of the new set takes the map of particle map_src[i] and the unused-measurement mask / in-FOV count of
 * particle aux_src[i] (NULL = map_src, -1 = none; the reference looks those up through
 * Particle::getParentId() while it consumes them, include/RBPHDFilter.hpp:1001-1011).  weight != NULL sets every particle weight to *weight (the
 * reference resets them to 1, :486-488).  The sampling itself (one drand48()) stays with the caller. */
int rfsb200_resample(rfsb200_ctx* ctx, const int32_t* map_src /*[N]*/, const int32_t* aux_src /*[N] or NULL*/,
                     const double* weight /*scalar or NULL*/);

/* ---- cross-GPU particle exchange for resampling -----------------------------------
 * A particle travels as one fixed-size record in DEVICE memory: its map planes, Gaussian count, pose (fp64 and
 * device precision), pose covariance, unused-measurement mask and in-FOV count.  rfsb200_particle_record_bytes gives
 * the record size of this ctx (identical on every rank with the same dims).  rfsb200_export_particles packs the
 * particles idx[0..n) of the COMMITTED state into dev_buf (n records; an index may repeat and n may exceed the particle
 * count: a heavy shard feeds many slots elsewhere); rfsb200_import_particles unpacks n <= N records into
 * the slots slot[0..n) of the committed state (weights are set to `weight`).  dev_buf is a device pointer owned by the
 * caller (e.g. the send / receive buffer of an all-to-all); both calls are queued on the ctx stream.  The resampling
 * plan itself (which particle goes where) is host logic: rfs-slam_b200/dist.py:global_resample_plan. */
int rfsb200_particle_record_bytes(rfsb200_ctx* ctx, int64_t* bytes);
int rfsb200_export_particles(rfsb200_ctx* ctx, const int32_t* idx /*[n] host*/, int32_t n, void* dev_buf);
int rfsb200_import_particles(rfsb200_ctx* ctx, const int32_t* slot /*[n] host*/, int32_t n, const void* dev_buf,
                             double weight);

/* ---- fused cross-GPU weight sum (one process per GPU on one node) ---------------------------------
 * rfsb200_comm_export writes a 64-byte CUDA IPC handle of this ctx's mailbox; the caller exchanges the
 * handles of all ranks (any transport: MPI, torch.distributed, files) and passes them, in rank
 * order, to rfsb200_comm_connect, which maps the peers' mailboxes (world <= 8).  From then on
 * RFSB200_UPDATE_FUSED_ALLREDUCE replaces the caller's all-reduce + rfsb200_normalize(). */
int rfsb200_comm_export(rfsb200_ctx* ctx, void* handle64);
int rfsb200_comm_connect(rfsb200_ctx* ctx, int32_t rank, int32_t world, const void* handles /*[world][64]*/);
/* The same for contexts that live in ONE process (a host thread per GPU, or several shards on one device): the peers'
 * mailboxes are reached by plain device pointers (peer access is enabled between different devices).  peers[rank] must be
 * ctx; call it on every ctx of the group before the first fused update of any of them. */
int rfsb200_comm_connect_local(rfsb200_ctx* ctx, int32_t rank, int32_t world, rfsb200_ctx* const* peers /*[world]*/);
/* Runs the normalisation a RFSB200_UPDATE_DEFER_NORMALIZE step left open (one small kernel on the ctx stream; no-op if
 * nothing is open).  Afterwards rfsb200_weight_sums_device() holds the global pair of that step. */
int rfsb200_comm_resolve(rfsb200_ctx* ctx);
/* Queues a barrier of the connected ranks on the ctx stream (one tiny kernel, flags through the peer mailboxes): every
 * rank leaves it within an NVLink round trip of the last one to arrive.  No-op for a single rank. */
int rfsb200_comm_barrier(rfsb200_ctx* ctx);
/* 1 if a peer did not arrive in time (2 s, or RFSB200_COMM_TIMEOUT_MS when the ctx was created) in some fused update or
 * barrier since the last call (the sums of that update are NaN) */
int rfsb200_comm_error(rfsb200_ctx* ctx, int32_t* flag);

/* Device address of the double[2] {sum w, sum w^2} of the last update, for the caller's
 * cross-GPU all-reduce (NCCL / torch.distributed) on the ctx stream; then normalize. */
int rfsb200_weight_sums_device(rfsb200_ctx* ctx, void** dev_ptr);
int rfsb200_normalize(rfsb200_ctx* ctx);
/* After NO_COMMIT: results live in the back buffers; which=0 reads the committed state,
 * which=1 the back buffers of the last update. */

/* ---- state out --------------------------------------------------------------------
 * Replace getGMSize / getLandmark (include/RBPHDFilter.hpp:1153-1178) and
 * Particle::getWeight; "which" = 0 committed state, 1 = result of the last NO_COMMIT update. */
int rfsb200_get_weights(rfsb200_ctx* ctx, int which, double* w /*[N]*/);
int rfsb200_get_gm_sizes(rfsb200_ctx* ctx, int which, int32_t* n /*[N]*/);
int rfsb200_get_map(rfsb200_ctx* ctx, int which, int32_t i, int32_t cap, int32_t* n,
                    double* mean, double* cov, double* w);
int rfsb200_download_maps(rfsb200_ctx* ctx, int which, int64_t cap_total, int32_t* count /*[N]*/,
                          double* mean, double* cov, double* w);
/* unused_measurements_[i] and nLandmarksInFOV_[i] (include/RBPHDFilter.hpp:269-272,709-720),
 * consumed by addBirthGaussians on the host. unused_mask bit z set = measurement z unused. */
int rfsb200_get_unused(rfsb200_ctx* ctx, uint64_t* unused_mask /*[N]*/, int32_t* n_in_fov /*[N]*/);
/* per-particle status bits of the last update: 1 = capacity overflow (of the update, or of a predict / append since the
 * update before), 2 = Murty branch taken, 4 = a partition beyond the assignment-sum tables, 8 = births dropped by the
 * last predict / append (cleared by the next update), 16 = a Murty-compatibility record did not fit its buffer,
 * 32 = a birth candidate did not fit the particle's candidate list (cleared by the next update) */
int rfsb200_get_flags(rfsb200_ctx* ctx, int32_t* flags /*[N]*/);

/* ---- utilities --------------------------------------------------------------------
 * rfs::MatPerm::calc (src/MatrixPermanent.cpp:41-113) for a batch of n x n matrices, on the
 * device, fp64.  Non-square input cannot be expressed here; n must be in [1, 24]. */
int rfsb200_permanent(rfsb200_ctx* ctx, const double* A /*[batch][n][n]*/, int32_t n,
                      int32_t batch, double* out /*[batch]*/);

/* ---- measurement aid ----------------------------------------------------------------
 * rfsb200_profile_begin arms a ring of CUDA event pairs: each of the next max_updates calls of
 * rfsb200_update() records one pair around the fused update kernel alone, on the ctx stream.
 * rfsb200_profile_read synchronises, returns the device time of each recorded launch in
 * microseconds (n = number recorded, at most cap are written) and disarms the ring.  Used by
 * bench.py for the live roofline figure; costs nothing when not armed. */
int rfsb200_profile_begin(rfsb200_ctx* ctx, int32_t max_updates);
/* Per-phase times of the last update that ran with RFSB200_UPDATE_STAGE_TIMES.  The update is ONE fused kernel, so a
 * phase has no wall time of its own: the kernel's wall time (device global timer, first CTA in to last CTA out) is
 * split into set-up (measurement tables), the particle loop and the weight-sum epilogue, and the particle loop is
 * attributed to the phases by their share of the warp cycles spent in them (SM clock read by every warp at the
 * phase boundaries of every particle). */
typedef struct rfsb200_stage_times {
  double kernel_us;          /* first CTA in -> last CTA out                                                         */
  double setup_us;           /* ... -> last CTA has built its measurement tables                                      */
  double particles_us;       /* ... -> last warp is out of particles                                                  */
  double epilogue_us;        /* ... -> end: [sum w, sum w^2], cross-GPU sum, normalisation                            */
  double share_load;         /* shares of the warp cycles inside the particle loop (sum = 1):  queue + bulk loads     */
  double share_map_update_kf;/*   GM-PHD corrector incl. the EKF updates          (TimingInfo::mapUpdate_kf)          */
  double share_weighting;    /*   normalisers, SC-PHD weight, posterior weights + multi-feature importanceWeighting
                                  (TimingInfo::particleWeighting)                                                    */
  double share_merge;        /*   GaussianMixture::merge                          (TimingInfo::mapMerge)              */
  double share_prune;        /*   prune + sort + store                            (TimingInfo::mapPrune)              */
  double warp_cycles;        /* total warp cycles inside the particle loop                                            */
  int32_t warps_per_cta;     /* launch shape of the stage-timing kernel                                               */
  int32_t reserved_i;
  double reserved[4];
} rfsb200_stage_times;
int rfsb200_get_stage_times(rfsb200_ctx* ctx, rfsb200_stage_times* out);
int rfsb200_profile_read(rfsb200_ctx* ctx, float* kernel_us /*[cap]*/, int32_t cap, int32_t* n);

/* Pinned host memory helpers (so H2D/D2H copies can overlap and run at PCIe speed, and so that rfsb200_update_host
 * can let the kernels read / write the buffers directly).  The library remembers the ranges it hands out together with
 * their device aliases: release them with rfsb200_host_free, not with cudaFreeHost. */
int rfsb200_host_alloc(void** ptr, uint64_t bytes);
int rfsb200_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* RFSB200_H */
