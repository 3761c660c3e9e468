/*
 * phd_oracle_vp.cpp — TEST INFRASTRUCTURE, NOT PRODUCT (see phd_oracle.h).
 *
 * fp64 CPU restatement of the per-particle PHD measurement update for the Victoria Park plugin
 * set of kykleung/RFS-SLAM (BASELINE config 5): 3-D landmarks (x, y, diameter), 3-D measurements
 * (range, bearing, diameter), probability of detection evaluated against the raw lidar scan.
 * Plain C++; every function cites the reference lines it follows (paths relative to
 * /root/reference).  The filter logic (updateMap, importanceWeighting, merge, prune) is the same
 * template code as for the 2-D model (include/RBPHDFilter.hpp, include/GaussianMixture.hpp); what
 * differs are the plugin functions and the dimension.
 *
 * Parity pin: checked against oracle/_ref/libphd_ref.so (phd_ref_update_vp / phd_ref_vp_pd =
 * the reference's own MeasurementModel_VictoriaPark.cpp etc. compiled unmodified), see
 * tests/test_oracle_golden.py and tests/golden/phd_vp_*.npz.
 *
 * One deliberate definition where the reference has undefined behaviour: a lidar-scan index
 * >= scan_n (the reference indexes a 361-entry vector with values up to 719,
 * src/MeasurementModel_VictoriaPark.cpp:249-255) counts no lidar point, which is what the reference
 * binary does in practice (the heap bytes behind the vector are neither > minrange nor == 0).
 */
#include "phd_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double PI = acos(-1); /* include/RandomVec.hpp:55 */

struct G3 { /* include/GaussianMixture.hpp:60-64 + Landmark3d */
  double x[3];
  double P[9]; /* row-major 3x3 */
  double w, wprev;
  bool alive;
};

struct CtxVP {
  const rfsb200_model_desc* md;
  const rfsb200_filter_cfg* fc;
  const double* Z; /* [nZ][3] */
  int nZ;
  int sort_mode;
};

/* ---- 3x3 helpers: cofactor inverse / determinant (Eigen's fixed-size closed forms) ---- */
inline double det3(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
inline void inv3(const double* A, double* Ai) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  Ai[0] = c00 * id;
  Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id;
  Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id;
  Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id;
  Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id;
  Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
inline void mul3(const double* A, const double* B, double* C) { /* C = A B */
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = A[r * 3] * B[c];
      for (int k = 1; k < 3; k++) s += A[r * 3 + k] * B[k * 3 + c];
      C[r * 3 + c] = s;
    }
}
inline void mul3t(const double* A, const double* B, double* C) { /* C = A B^T */
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = A[r * 3] * B[c * 3];
      for (int k = 1; k < 3; k++) s += A[r * 3 + k] * B[c * 3 + k];
      C[r * 3 + c] = s;
    }
}
/* (e^T Ai) e — include/RandomVec.hpp:387-407 */
inline double quad3(const double* Ai, const double* e) {
  double r[3];
  for (int c = 0; c < 3; c++) r[c] = e[0] * Ai[c] + e[1] * Ai[3 + c] + e[2] * Ai[6 + c];
  return r[0] * e[0] + r[1] * e[1] + r[2] * e[2];
}
/* include/RandomVec.hpp:415-451, nDim = 3 */
inline double gauss_pdf3(const double* Sinv, double det, const double* e, double* md2_out) {
  double factor = sqrt(pow(2 * PI, 3) * det);
  double md2 = quad3(Sinv, e);
  double l = exp(-0.5 * md2) / factor;
  if (l != l) l = 0;
  if (md2_out) *md2_out = md2;
  return l;
}

/* src/MeasurementModel_VictoriaPark.cpp:104-151 on top of src/MeasurementModel_RngBrg.cpp:70-115 with
 * the pose rotated by -pi/2 and a ZERO pose covariance (the transformed pose is built from the
 * mean only, :112-114).  Always "valid": the 2-D model's range test is ignored (:127,:139).
 * P = 3x3 landmark covariance; zexp[3], S[9], H[9] (H may be NULL). */
void vp_measure(const rfsb200_model_desc& md, const double* pose, const double* lx, const double* P,
                double* zexp, double* S, double* H) {
  const double th = pose[2] - PI / 2;
  double dx = lx[0] - pose[0], dy = lx[1] - pose[1];
  double range2 = pow(dx, 2) + pow(dy, 2);
  double range = sqrt(range2);
  double bearing = atan2(dy, dx) - th;
  while (bearing > PI) bearing -= 2 * PI;
  while (bearing < -PI) bearing += 2 * PI;
  zexp[0] = range;
  zexp[1] = bearing;
  zexp[2] = lx[2];
  const double Hl[4] = {dx / range, dy / range, -dy / range2, dx / range2};
  const double P2[4] = {P[0], P[1], P[3], P[4]};
  const double HP[4] = {Hl[0] * P2[0] + Hl[1] * P2[2], Hl[0] * P2[1] + Hl[1] * P2[3],
                        Hl[2] * P2[0] + Hl[3] * P2[2], Hl[2] * P2[1] + Hl[3] * P2[3]};
  const double A[4] = {HP[0] * Hl[0] + HP[1] * Hl[1], HP[0] * Hl[2] + HP[1] * Hl[3],
                       HP[2] * Hl[0] + HP[3] * Hl[1], HP[2] * Hl[2] + HP[3] * Hl[3]};
  for (int k = 0; k < 9; k++) S[k] = 0;
  /* the 2-D model's R_ is the top-left block of R (setNoise :66-72) */
  S[0] = A[0] + md.R[0];
  S[1] = A[1] + md.R[1];
  S[3] = A[2] + md.R[3];
  S[4] = A[3] + md.R[4];
  S[8] = P[8] + md.R[8] + pow(range, 2) * md.Slb; /* :131,:143 */
  if (H) {
    for (int k = 0; k < 9; k++) H[k] = 0;
    H[0] = Hl[0]; H[1] = Hl[1]; H[3] = Hl[2]; H[4] = Hl[3];
    H[8] = 1;
  }
}

inline double scan_at(const rfsb200_model_desc& md, int b) {
  /* reference: out-of-bounds read (heap bytes behind the vector: neither a range nor exactly 0) -> no point */
  return (b >= 0 && b < md.scan_n) ? md.scan[b] : std::numeric_limits<double>::quiet_NaN();
}

/* src/MeasurementModel_VictoriaPark.cpp:202-266 */
double vp_pd2(const rfsb200_model_desc& md, const double* pose, const double* lx, const double* P, bool& close) {
  close = false;
  double z[3], S[9];
  vp_measure(md, pose, lx, P, z, S, nullptr);
  const double dist = z[0], angle = z[1];
  if (angle > md.bearing_max || angle < md.bearing_min || dist < md.range_min || dist > md.range_max) return 0;
  const double modified_radius = z[2] / 2;
  const double gamma = atan(modified_radius / z[0]);
  const int maxNumPoints = (int)floor(2 * gamma * 720.0 / (2 * PI));
  const size_t tn = (size_t)md.pd_table_n;
  if (tn > (size_t)maxNumPoints && md.pd_table[maxNumPoints] == 0) return 0; /* int -> size_t compare as in the reference */
  if (tn > (size_t)maxNumPoints && md.pd_table[maxNumPoints] < md.buffer_zone_pd) close = true;
  int minb = (int)ceil((angle - gamma) * 720.0 / (2 * PI));
  int maxb = minb + maxNumPoints;
  while (minb >= 720) minb -= 720;
  while (minb < 0) minb += 720;
  while (maxb >= 720) maxb -= 720;
  while (maxb < 0) maxb += 720;
  int numPoints = 0;
  const double minrange = dist - modified_radius - 6 * 0.03;
  if ((maxb - minb + 720) % 720 > 0) {
    for (int b = minb; b != maxb; b = (b + 1) % 720) {
      const double s = scan_at(md, b);
      if (s > minrange || s == 0) numPoints++;
    }
  }
  if (numPoints >= (int)tn) numPoints = (int)tn - 1;
  if (md.pd_table[numPoints] == 0) close = false;
  return md.pd_table[numPoints];
}

/* src/MeasurementModel_VictoriaPark.cpp:153-199 */
double vp_pd(const rfsb200_model_desc& md, const double* pose, const double* lx, const double* P, bool& close) {
  double z[3], S[9];
  vp_measure(md, pose, lx, P, z, S, nullptr);
  /* :164: atan2(bearing, range) + theta — as written in the reference */
  const double angle = atan2(z[1], z[0]) + pose[2];
  const double perp[2] = {-sin(angle), cos(angle)};
  double sd = (perp[0] * P[0] + perp[1] * P[3]) * perp[0] + (perp[0] * P[1] + perp[1] * P[4]) * perp[1];
  sd = 3 * sqrt(sd);
  sd = std::max(sd, 0.2);
  double pmin = std::numeric_limits<double>::infinity(), pmax = -std::numeric_limits<double>::infinity();
  auto take = [&](double v) { pmin = std::min(pmin, v); pmax = std::max(pmax, v); };
  double l2[3] = {lx[0], lx[1], lx[2]};
  for (int i = 1; (i - 1) * (2 * lx[2]) < sd; i++) {
    l2[0] = lx[0] + i * 2 * lx[2] * perp[0];
    l2[1] = lx[1] + i * 2 * lx[2] * perp[1];
    take(vp_pd2(md, pose, l2, P, close));
    l2[0] = lx[0] - i * 2 * lx[2] * perp[0];
    l2[1] = lx[1] - i * 2 * lx[2] * perp[1];
    take(vp_pd2(md, pose, l2, P, close));
    if (i > 4096) break; /* the reference loops forever for a non-positive diameter */
  }
  take(vp_pd2(md, pose, lx, P, close));
  if (pmin == 0 && pmax > 0) close = true;
  return pmax;
}

/* include/KalmanFilter_VictoriaPark.hpp:56-74: wrap first, then the thresholds */
bool vp_innovation(const rfsb200_model_desc& md, const double* zexp, const double* zact, double* innov) {
  for (int k = 0; k < 3; k++) innov[k] = zact[k] - zexp[k];
  while (innov[1] > PI) innov[1] -= 2 * PI;
  while (innov[1] < -PI) innov[1] += 2 * PI;
  if (md.innov_thr_range > 0 && fabs(innov[0]) > md.innov_thr_range) return false;
  if (md.innov_thr_bearing > 0 && fabs(innov[1]) > md.innov_thr_bearing) return false;
  return true;
}

struct LmkNew3 {
  double x[3];
  double P[9];
};

/* include/KalmanFilter.hpp:261-342 */
void kf_correct_batch3(const CtxVP& c, const double* pose, const G3& lm, std::vector<LmkNew3>& lmNew,
                       std::vector<double>& lik, std::vector<double>& md2v) {
  const int nZ = c.nZ;
  double zexp[3], S[9], H[9];
  vp_measure(*c.md, pose, lm.x, lm.P, zexp, S, H); /* always valid */
  double Sinv[9];
  inv3(S, Sinv);
  double PHt[9], K[9], KH[9], IKH[9], Pu[9], Ps[9];
  mul3t(lm.P, H, PHt);
  mul3(PHt, Sinv, K);
  mul3(K, H, KH);
  for (int k = 0; k < 9; k++) IKH[k] = ((k % 4 == 0) ? 1.0 : 0.0) - KH[k];
  mul3(IKH, lm.P, Pu);
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) Ps[r * 3 + cc] = (Pu[r * 3 + cc] + Pu[cc * 3 + r]) / 2;
  const double detS = det3(S);
  for (int i = 0; i < nZ; i++) {
    const double* zact = c.Z + 3 * i;
    double innov[3];
    if (vp_innovation(*c.md, zexp, zact, innov)) {
      for (int r = 0; r < 3; r++)
        lmNew[i].x[r] = lm.x[r] + (K[r * 3] * innov[0] + K[r * 3 + 1] * innov[1] + K[r * 3 + 2] * innov[2]);
      memcpy(lmNew[i].P, Ps, sizeof(Ps));
      /* Q3: the likelihood uses the UNWRAPPED difference z - zexp (KalmanFilter.hpp:319) */
      double e[3] = {zact[0] - zexp[0], zact[1] - zexp[1], zact[2] - zexp[2]};
      double md2;
      double zl = gauss_pdf3(Sinv, detS, e, &md2);
      if (zl != zl) zl = 0;
      lik[i] = zl;
      md2v[i] = md2;
    } else {
      lik[i] = 0;
      md2v[i] = 0;
    }
  }
}

/* include/GaussianMixture.hpp:523-534 */
void sort_by_weight3(std::vector<G3>& g, int sort_mode) {
  if (sort_mode == PHD_SORT_STD)
    std::sort(g.begin(), g.end(), [](G3 a, G3 b) { return a.w > b.w; });
  else
    std::stable_sort(g.begin(), g.end(), [](const G3& a, const G3& b) { return a.w > b.w; });
}

/* include/RBPHDFilter.hpp:543-725 */
void update_map3(const CtxVP& c, const double* pose, std::vector<G3>& gm, double& pweight, uint64_t& unused_mask,
                 int32_t& n_in_fov) {
  const int nZ = c.nZ;
  const rfsb200_filter_cfg& fc = *c.fc;
  const unsigned nM = gm.size();
  unused_mask = 0;
  n_in_fov = 0;
  if (nM == 0) { /* :559-564 (Q10) */
    for (int z = 0; z < nZ; z++) unused_mask |= (1ull << z);
    return;
  }
  std::vector<double> Pd(nM);
  std::vector<int> closeLim(nM);
  double w_km_sum = std::numeric_limits<double>::denorm_min();
  double likelihoodProd = 1;
  if (fc.use_cluster_process)
    for (unsigned m = 0; m < nM; m++) w_km_sum += gm[m].w;
  std::vector<double> W((size_t)nM * nZ);
  std::vector<char> Mok((size_t)nM * nZ);
  std::vector<LmkNew3> Mtab((size_t)nM * nZ);
  const double thr2 = fc.new_gaussian_create_innov_md_threshold * fc.new_gaussian_create_innov_md_threshold;
  std::vector<double> lik(nZ), md2(nZ);
  std::vector<LmkNew3> lmNew(nZ);
  for (unsigned m = 0; m < nM; m++) { /* :597-641 */
    bool close;
    Pd[m] = vp_pd(*c.md, pose, gm[m].x, gm[m].P, close);
    if (close) {
      closeLim[m] = 1;
      Pd[m] = 1; /* Q2 */
    } else
      closeLim[m] = 0;
    const double Pd_times_w_km = Pd[m] * gm[m].w;
    if (Pd[m] != 0) {
      n_in_fov++;
      kf_correct_batch3(c, pose, gm[m], lmNew, lik, md2);
      for (int z = 0; z < nZ; z++) {
        if (lik[z] == 0 || md2[z] > thr2) {
          Mok[m * nZ + z] = 0;
          W[m * nZ + z] = 0;
        } else {
          Mok[m * nZ + z] = 1;
          Mtab[m * nZ + z] = lmNew[z];
          W[m * nZ + z] = Pd_times_w_km * lik[z];
        }
      }
    } else {
      for (int z = 0; z < nZ; z++) {
        Mok[m * nZ + z] = 0;
        W[m * nZ + z] = 0;
      }
    }
  }
  for (int z = 0; z < nZ; z++) { /* :644-659 */
    double sum = c.md->clutter_intensity;
    for (unsigned m = 0; m < nM; m++) sum += W[m * nZ + z];
    if (fc.use_cluster_process) likelihoodProd *= sum;
    for (unsigned m = 0; m < nM; m++) W[m * nZ + z] = W[m * nZ + z] / sum;
  }
  if (fc.use_cluster_process) pweight = exp(w_km_sum) * likelihoodProd * pweight; /* :661-668 (Q4) */
  for (unsigned m = 0; m < nM; m++) /* :675-683 */
    for (int z = 0; z < nZ; z++)
      if (Mok[m * nZ + z] && W[m * nZ + z] > 0) {
        G3 g;
        memcpy(g.x, Mtab[m * nZ + z].x, sizeof(g.x));
        memcpy(g.P, Mtab[m * nZ + z].P, sizeof(g.P));
        g.w = W[m * nZ + z];
        g.wprev = 0;
        g.alive = true;
        gm.push_back(g);
      }
  for (unsigned m = 0; m < nM; m++) { /* :686-706 */
    double w_km = gm[m].w;
    double w_k = (1 - Pd[m]) * w_km;
    if (closeLim[m] == 1 && w_km > fc.birth_gaussian_weight) {
      double weight_sum_m = 0;
      for (int z = 0; z < nZ; z++) weight_sum_m += W[m * nZ + z];
      double delta_w = Pd[m] * w_km - weight_sum_m;
      if (delta_w > 0) {
        w_k += delta_w;
        if (w_k > 1) w_k = 1;
      }
    }
    gm[m].wprev = gm[m].w;
    gm[m].w = w_k;
  }
  for (int z = 0; z < nZ; z++) { /* :709-720 */
    bool used = false;
    for (unsigned m = 0; m < nM; m++)
      if (W[m * nZ + z] != 0) {
        used = true;
        break;
      }
    if (!used) unused_mask |= (1ull << z);
  }
}

/* include/RBPHDFilter.hpp:821-997 (the table; the partition logic is shared with the 2-D oracle) */
double rfs_measurement_likelihood3(const CtxVP& c, const double* pose, const std::vector<G3>& gm,
                                   const std::vector<unsigned>& evalIdx, const std::vector<double>& evalPd,
                                   int32_t* flags) {
  const int nM = evalIdx.size();
  const int nZ = c.nZ;
  const double thr = c.fc->meas_likelihood_md_threshold * c.fc->meas_likelihood_md_threshold;
  std::vector<double> L((size_t)nM * nZ);
  const double zeroP[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int m = 0; m < nM; m++) {
    double zexp[3], S[9], Sinv[9];
    vp_measure(*c.md, pose, gm[evalIdx[m]].x, zeroP, zexp, S, nullptr); /* :850-852 covariance zeroed */
    inv3(S, Sinv);
    const double detS = det3(S);
    const double Pd = evalPd[m];
    for (int n = 0; n < nZ; n++) {
      double e[3] = {c.Z[3 * n] - zexp[0], c.Z[3 * n + 1] - zexp[1], c.Z[3 * n + 2] - zexp[2]};
      double md2;
      L[m * nZ + n] = gauss_pdf3(Sinv, detS, e, &md2) * Pd;
      if (md2 > thr) L[m * nZ + n] = 0;
    }
  }
  std::vector<double> clutter(nZ, c.md->clutter_intensity);
  double l = phd_oracle_partition_likelihood(L.data(), nM, nZ, evalPd.data(), clutter.data(), flags);
  return l / c.md->clutter_integral;
}

/* include/RBPHDFilter.hpp:728-819 */
void importance_weighting3(const CtxVP& c, const double* pose, std::vector<G3>& gm, double& pweight, int32_t* flags) {
  const rfsb200_filter_cfg& fc = *c.fc;
  const unsigned nM = gm.size();
  int nEvalPoints = (unsigned)fc.eval_point_count > nM ? (int)nM : fc.eval_point_count; /* Q14 */
  std::vector<unsigned> evalIdx;
  std::vector<double> evalPd;
  if (nEvalPoints == 0) {
    pweight = std::numeric_limits<double>::denorm_min();
    return;
  }
  sort_by_weight3(gm, c.sort_mode);
  for (unsigned m = 0; m < nM; m++) {
    if (gm[m].w < fc.eval_point_gaussian_weight) break;
    bool close;
    double Pd = vp_pd(*c.md, pose, gm[m].x, gm[m].P, close);
    if (Pd > 0) {
      evalIdx.push_back(m);
      evalPd.push_back(Pd);
    }
    if (nEvalPoints != -1 && (int)evalIdx.size() >= nEvalPoints) break;
  }
  nEvalPoints = evalIdx.size();
  double sumBefore = 0, sumAfter = 0;
  for (unsigned m = 0; m < nM; m++) {
    sumBefore += gm[m].wprev;
    sumAfter += gm[m].w;
  }
  double prodBefore = 1, prodAfter = 1;
  std::vector<double> Pinv(9 * nM), fac(nM);
  for (unsigned m = 0; m < nM; m++) {
    inv3(gm[m].P, &Pinv[9 * m]);
    fac[m] = sqrt(pow(2 * PI, 3) * det3(gm[m].P));
  }
  for (int e = 0; e < nEvalPoints; e++) {
    const G3& ge = gm[evalIdx[e]];
    double vb = std::numeric_limits<double>::denorm_min();
    double va = std::numeric_limits<double>::denorm_min();
    for (unsigned m = 0; m < nM; m++) {
      double d[3] = {ge.x[0] - gm[m].x[0], ge.x[1] - gm[m].x[1], ge.x[2] - gm[m].x[2]};
      double md2 = quad3(&Pinv[9 * m], d);
      double lk = exp(-0.5 * md2) / fac[m];
      if (lk != lk) lk = 0;
      vb += gm[m].wprev * lk;
      va += gm[m].w * lk;
    }
    prodBefore *= vb;
    prodAfter *= va;
  }
  double ml = rfs_measurement_likelihood3(c, pose, gm, evalIdx, evalPd, flags);
  double overall = ml * prodBefore / prodAfter * exp(sumAfter - sumBefore);
  pweight = overall * pweight;
}

/* include/GaussianMixture.hpp:419-475 */
bool merge_pair3(std::vector<G3>& g, unsigned i1, unsigned i2, double t, double f) {
  if (!g[i1].alive || !g[i2].alive) return false;
  const double w1 = g[i1].w, w2 = g[i2].w;
  const double t2 = t * t;
  double Pi[9], d[3];
  inv3(g[i1].P, Pi);
  for (int k = 0; k < 3; k++) d[k] = g[i2].x[k] - g[i1].x[k];
  if (quad3(Pi, d) > t2) {
    inv3(g[i2].P, Pi);
    for (int k = 0; k < 3; k++) d[k] = g[i1].x[k] - g[i2].x[k];
    if (quad3(Pi, d) > t2) return false;
  }
  const double wm = w1 + w2;
  if (wm == 0) return false;
  double xm[3], e1[3], e2[3], Sm[9];
  for (int k = 0; k < 3; k++) {
    xm[k] = (g[i1].x[k] * w1 + g[i2].x[k] * w2) / wm;
    e1[k] = xm[k] - g[i1].x[k];
    e2[k] = xm[k] - g[i2].x[k];
  }
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) {
      const double a = w1 * (g[i1].P[r * 3 + cc] + f * e1[r] * e1[cc]);
      const double b = w2 * (g[i2].P[r * 3 + cc] + f * e2[r] * e2[cc]);
      Sm[r * 3 + cc] = (a + b) / wm;
    }
  memcpy(g[i1].x, xm, sizeof(xm));
  memcpy(g[i1].P, Sm, sizeof(Sm));
  g[i1].w = wm;
  g[i1].wprev = 0;
  g[i2].alive = false;
  g[i2].w = 0;
  g[i2].wprev = 0;
  return true;
}

/* include/GaussianMixture.hpp:394-416 */
void merge_all3(std::vector<G3>& g, double t, double f) {
  const unsigned n = g.size();
  for (unsigned i = 0; i < n; i++) {
    if (!g[i].alive) continue;
    for (unsigned j = i + 1; j < n; j++) merge_pair3(g, i, j, t, f);
  }
}

/* include/GaussianMixture.hpp:477-521 */
void prune3(std::vector<G3>& g, double t, int sort_mode) {
  if (g.size() < 1) return;
  sort_by_weight3(g, sort_mode);
  unsigned min_idx = 0, max_idx = g.size() - 1;
  unsigned idx = (max_idx + min_idx) / 2;
  unsigned idx_old = idx + 1;
  double w = g[idx].w;
  while (idx != idx_old) {
    if (w <= t)
      max_idx = idx;
    else if (w > t)
      min_idx = idx;
    idx_old = idx;
    idx = (max_idx + min_idx) / 2;
    w = g[idx].w;
  }
  while (w >= t) {
    idx++;
    if (idx >= g.size()) break;
    w = g[idx].w;
  }
  g.resize(idx);
}

}  // namespace

/* same contract as phd_oracle_update with 3-D arrays: mean [..][3], cov [..][6] (xx,xy,xz,yy,yz,zz), Z [nZ][3];
 * the pose covariance is ignored by this model (Q1: VP builds a zero-covariance pose) */
extern "C" int phd_oracle_update_vp(phd_io* io) {
  if (!io || !io->model || !io->cfg || io->N < 0 || io->nZ < 0 || io->nZ > 64) return -1;
  if (io->model->model_id != RFSB200_MODEL_VICTORIAPARK) return -5;
  const rfsb200_model_desc& md = *io->model;
  if (md.pd_table_n < 1 || md.pd_table_n > 16 || md.scan_n < 0 || (md.scan_n > 0 && !md.scan)) return -1;
  const int N = io->N;
  std::vector<int64_t> off(N + 1, 0);
  for (int i = 0; i < N; i++) off[i + 1] = off[i] + io->count_in[i];
  std::vector<std::vector<G3>> maps(N);
  for (int i = 0; i < N; i++) {
    maps[i].resize(io->count_in[i]);
    for (int m = 0; m < io->count_in[i]; m++) {
      const int64_t k = off[i] + m;
      G3& g = maps[i][m];
      for (int d = 0; d < 3; d++) g.x[d] = io->mean_in[3 * k + d];
      const double* c = io->cov_in + 6 * k;
      g.P[0] = c[0]; g.P[1] = c[1]; g.P[2] = c[2];
      g.P[3] = c[1]; g.P[4] = c[3]; g.P[5] = c[4];
      g.P[6] = c[2]; g.P[7] = c[4]; g.P[8] = c[5];
      g.w = io->w_in[k];
      g.wprev = 0;
      g.alive = true;
    }
  }
  std::vector<double> pw(io->weight_in, io->weight_in + N);
  std::vector<uint64_t> unused(N, 0);
  std::vector<int32_t> nfov(N, 0), flags(N, 0);
  CtxVP c{io->model, io->cfg, io->Z, io->nZ, io->sort_mode};
#ifdef _OPENMP
  if (io->n_threads > 0) omp_set_num_threads(io->n_threads);
#endif
  auto t0 = std::chrono::steady_clock::now();
  if (io->nZ > 0) { /* include/RBPHDFilter.hpp:451-452 (Q11) */
#pragma omp parallel
    {
#pragma omp for
      for (int i = 0; i < N; i++) update_map3(c, io->pose + 3 * (size_t)i, maps[i], pw[i], unused[i], nfov[i]);
      if (!io->cfg->use_cluster_process && io->stage >= PHD_STAGE_WEIGHTING) {
#pragma omp for
        for (int i = 0; i < N; i++) importance_weighting3(c, io->pose + 3 * (size_t)i, maps[i], pw[i], &flags[i]);
      }
      if (io->stage >= PHD_STAGE_MERGE) {
#pragma omp for
        for (int i = 0; i < N; i++) merge_all3(maps[i], io->cfg->merging_threshold, io->cfg->merging_cov_inflation_factor);
      }
      if (io->stage >= PHD_STAGE_FULL) {
#pragma omp for
        for (int i = 0; i < N; i++) prune3(maps[i], io->cfg->pruning_threshold, io->sort_mode);
      }
    }
  }
  io->elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int64_t k = 0;
  for (int i = 0; i < N; i++) {
    int cnt = 0;
    for (const G3& g : maps[i]) {
      if (!g.alive) continue;
      if (k >= io->cap_total) return -4;
      for (int d = 0; d < 3; d++) io->mean_out[3 * k + d] = g.x[d];
      double* cv = io->cov_out + 6 * k;
      cv[0] = g.P[0]; cv[1] = g.P[1]; cv[2] = g.P[2]; cv[3] = g.P[4]; cv[4] = g.P[5]; cv[5] = g.P[8];
      io->w_out[k] = g.w;
      if (io->wprev_out) io->wprev_out[k] = g.wprev;
      k++;
      cnt++;
    }
    io->count_out[i] = cnt;
    io->weight_out[i] = pw[i];
    if (io->unused_mask) io->unused_mask[i] = unused[i];
    if (io->n_in_fov) io->n_in_fov[i] = nfov[i];
    if (io->flags) io->flags[i] = flags[i];
  }
  return 0;
}

/* MeasurementModel_VictoriaPark::probabilityOfDetection restated (probe; pinned against phd_ref_vp_pd) */
extern "C" double phd_oracle_vp_pd(const rfsb200_model_desc* md, const double* pose, const double* lx,
                                   const double* lcov6, int* close_out) {
  const double P[9] = {lcov6[0], lcov6[1], lcov6[2], lcov6[1], lcov6[3], lcov6[4], lcov6[2], lcov6[4], lcov6[5]};
  bool close = false;
  const double pd = vp_pd(*md, pose, lx, P, close);
  if (close_out) *close_out = close ? 1 : 0;
  return pd;
}
