/*
 * phd_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * fp64 CPU restatement of the reference's PHD measurement-update path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under rfs-slam_b200/ may include, link or call this.
 *
 * The same extern "C" signature is implemented twice:
 *   phd_oracle_update  (oracle/phd_oracle.cpp)   our restatement of the reference algorithm
 *   phd_ref_update     (oracle/ref_harness.cpp)  the reference's OWN sources compiled against
 *                                                compat shims -> oracle/_ref/libphd_ref.so
 * so the restatement can be pinned against the real thing on identical inputs.
 */
#ifndef PHD_ORACLE_H
#define PHD_ORACLE_H

#include <stdint.h>
#include "../include/rfsb200.h"   /* only for the POD descriptors (model / filter cfg) */

#ifdef __cplusplus
extern "C" {
#endif

/* stage: how far RBPHDFilter::update() is run before the state is dumped */
#define PHD_STAGE_UPDATE_MAP 1  /* after updateMap (include/RBPHDFilter.hpp:475-478)            */
#define PHD_STAGE_WEIGHTING  2  /* after importanceWeighting (:488-491; no-op for SC-PHD)        */
#define PHD_STAGE_MERGE      3  /* after GaussianMixture::merge (:501-505), holes removed        */
#define PHD_STAGE_FULL       4  /* after prune (:513-516); weights NOT normalised                */

/* sort_mode: tie-break of GaussianMixture::sortByWeight (include/GaussianMixture.hpp:523-534) */
#define PHD_SORT_STD    0  /* std::sort with the weight-only comparator, as the reference (Q9)   */
#define PHD_SORT_STABLE 1  /* weight descending, then current list position: what the GPU does   */

typedef struct phd_io {
  /* in */
  int32_t N;
  const int32_t* count_in;   /* [N] */
  const double* mean_in;     /* packed [sum][2] */
  const double* cov_in;      /* packed [sum][3] (xx,xy,yy) */
  const double* w_in;        /* packed [sum] */
  const double* pose;        /* [N][3] */
  const double* pose_cov;    /* mode 0: NULL, 1: [6], 2: [N][6] upper triangle of 3x3 */
  int32_t pose_cov_mode;
  const double* weight_in;   /* [N] */
  const double* Z;           /* [nZ][2] */
  int32_t nZ;
  const rfsb200_model_desc* model;
  const rfsb200_filter_cfg* cfg;
  int32_t sort_mode;
  int32_t stage;
  int32_t n_threads;         /* OpenMP threads (<=0: default) */
  /* out */
  int64_t cap_total;         /* capacity (in Gaussians) of the packed output arrays */
  int32_t* count_out;        /* [N] */
  double* mean_out;          /* packed [..][2] */
  double* cov_out;           /* packed [..][3] */
  double* w_out;             /* packed */
  double* wprev_out;         /* packed (weight_prev; may be NULL) */
  double* weight_out;        /* [N] unnormalised particle weights */
  uint64_t* unused_mask;     /* [N] bit z = measurement z unused (may be NULL) */
  int32_t* n_in_fov;         /* [N] (may be NULL) */
  int32_t* flags;            /* [N] bit 1 (value 2): Murty branch taken (may be NULL) */
  double elapsed_s;          /* wall time of the update loop proper */
} phd_io;

int phd_oracle_update(phd_io* io);

/* Victoria Park plugin set (model_id RFSB200_MODEL_VICTORIAPARK): same struct with 3-D arrays —
 * mean [..][3], cov [..][6] (xx,xy,xz,yy,yz,zz), Z [nZ][3]; pose_cov is ignored (the model builds
 * a zero-covariance pose, src/MeasurementModel_VictoriaPark.cpp:112-114).
 *   phd_oracle_update_vp (oracle/phd_oracle_vp.cpp)   restatement
 *   phd_ref_update_vp    (oracle/ref_harness_vp.cpp)  the reference's own sources */
int phd_oracle_update_vp(phd_io* io);
/* MeasurementModel_VictoriaPark::probabilityOfDetection for one landmark (lcov6 = upper triangle) */
double phd_oracle_vp_pd(const rfsb200_model_desc* md, const double* pose, const double* lx,
                        const double* lcov6, int* close_out);

/* rfs::MatPerm::calc restatement (src/MatrixPermanent.cpp:41-113); A row-major n x n */
double phd_oracle_permanent(const double* A, int n);

/* number of assignments visited by the PermutationLexicographic enumeration for (nM, nZ)
 * (src/PermutationLexicographic.cpp:38-96), a KAT for the enumerators */
int64_t phd_oracle_lexi_count(int nM, int nZ);

/* rfsMeasurementLikelihood on a given likelihood table L [nE][nZ] (row-major), evalPd[nE],
 * clutter[nZ]; returns the product over partitions BEFORE the division by the clutter
 * integral (include/RBPHDFilter.hpp:866-994).  flags bit 1 (value 2) set if Murty was used. */
double phd_oracle_partition_likelihood(const double* L, int nE, int nZ, const double* evalPd,
                                       const double* clutter, int32_t* flags);

/* CostMatrixGeneral::partition() + getPartitionSize() for p < nP (src/CostMatrix.cpp:92-173) */
int phd_oracle_partition(const double* L, int nR, int nC, int* nRows, int* nCols, int* isZero);

/* k-best assignment sum exactly as include/RBPHDFilter.hpp:920-959 on one partition
 * Cp [nR][nC] of likelihoods (0 = no edge). */
double phd_oracle_murty_sum(const double* Lp, int nR, int nC, const double* rowPd,
                            const double* colClutter);

/* addBirthGaussians() in its candidate-list form (include/RBPHDFilter.hpp:1000-1080), all particles.
 * The model decides the dimensions: RngBrg ld = 2 / nc = 3 / md = 2, Victoria Park ld = 3 / nc = 6 / md = 3.
 *   phd_oracle_birth_candidates (oracle/phd_oracle_births.cpp)   restatement
 *   phd_ref_birth_candidates    (oracle/ref_harness*.cpp)        the reference's own addBirthGaussians() */
typedef struct phd_birth_io {
  const rfsb200_model_desc* model;
  int32_t N, nZ;
  int32_t cand_cap;            /* slots per particle in the cand_* arrays                                  */
  int32_t add_cap;             /* slots per particle in the add_* arrays                                   */
  int32_t resample_occurred;   /* resampleOccured_: lists are looked up through parent[] (:1005-1011)      */
  uint32_t count_thr;          /* birthGaussianMeasurementCountThreshold_                                  */
  uint32_t check_thr;          /* birthGaussianMeasurementCheckThreshold_                                  */
  uint32_t cur_count_thr;      /* birthGaussianCurrentMeasurementCountThreshold_                           */
  double support_dist;         /* birthGaussianMeasurementSupportDist_                                     */
  const double* pose;          /* [N][3]                                                                   */
  const double* pose_cov;      /* [N][6] upper triangle or NULL (RngBrg only)                              */
  const double* Z;             /* [nZ][md]                                                                 */
  const int32_t* parent;       /* [N] getParentId() of the particle in slot i                              */
  uint64_t* unused;            /* [N] in: unused_measurements_ as bit masks; out: 0                        */
  const int32_t* nfov;         /* [N] nLandmarksInFOV_                                                     */
  int32_t* cand_n;             /* [N] in / out: length of birthGaussians_[i]                               */
  double* cand_mean;           /* [N][cand_cap][ld] in / out                                               */
  double* cand_cov;            /* [N][cand_cap][nc] in / out, upper triangle                               */
  int32_t* cand_support;       /* [N][cand_cap] nSupportingMeasurements                                    */
  int32_t* cand_checks;        /* [N][cand_cap] nChecks                                                    */
  int32_t* add_n;              /* [N] out: Gaussians that became real, in the order of the addGaussian calls */
  double* add_mean;            /* [N][add_cap][ld] out                                                     */
  double* add_cov;             /* [N][add_cap][nc] out                                                     */
} phd_birth_io;
int phd_oracle_birth_candidates(phd_birth_io* io);

#ifdef __cplusplus
}
#endif
#endif
