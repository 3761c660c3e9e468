/*
 * phd_oracle_births.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * fp64 CPU restatement of RBPHDFilter::addBirthGaussians() in its candidate-list form
 * (reference include/RBPHDFilter.hpp:1000-1080) for the two plugin sets of the device build:
 *   MeasurementModel_RngBrg + KalmanFilter_RngBrg            (2-D landmarks, 2-D measurements)
 *   MeasurementModel_VictoriaPark + KalmanFilter_VictoriaPark (3-D landmarks, 3-D measurements)
 * Pinned against the reference's own addBirthGaussians() (phd_ref_birth_candidates in
 * oracle/ref_harness*.cpp) by tests/golden/make_golden_births.py -> tests/golden/births_*.npz.
 */
#include <cmath>
#include <cstring>
#include <vector>

#include "phd_oracle.h"

namespace {

const double PI = 3.14159265358979323846;

template <int D>
struct Cand { /* BirthGaussianCandidate (include/RBPHDFilter.hpp:258-262): a landmark + two counters */
  double x[D];
  double P[D * D];
  unsigned nSupport, nChecks;
};

/* closed-form inverses, as Eigen evaluates fixed-size 2x2 / 3x3 (cofactors times 1 / det) */
inline void inv(const double (&A)[4], double (&Ai)[4]) {
  const double id = 1.0 / (A[0] * A[3] - A[1] * A[2]);
  Ai[0] = A[3] * id; Ai[1] = -A[1] * id; Ai[2] = -A[2] * id; Ai[3] = A[0] * id;
}
inline void inv(const double (&A)[9], double (&Ai)[9]) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
template <int D>
inline void mul(const double* A, const double* B, double* C, bool bt) { /* C = A B or A B^T */
  for (int r = 0; r < D; r++)
    for (int c = 0; c < D; c++) {
      double s = 0;
      for (int k = 0; k < D; k++) s += A[r * D + k] * (bt ? B[c * D + k] : B[k * D + c]);
      C[r * D + c] = s;
    }
}

/* the 2-D range-bearing geometry both models share (src/MeasurementModel_RngBrg.cpp:70-115); th = heading used */
struct RB {
  double range, bearing, Hl[4];
  double A[4]; /* Hl P2 Hl^T */
};
inline RB range_bearing(const double* pose, double th, const double* lx, const double* P2) {
  RB g;
  const double dx = lx[0] - pose[0], dy = lx[1] - pose[1];
  const double range2 = pow(dx, 2) + pow(dy, 2);
  g.range = sqrt(range2);
  g.bearing = atan2(dy, dx) - th;
  while (g.bearing > PI) g.bearing -= 2 * PI;
  while (g.bearing < -PI) g.bearing += 2 * PI;
  g.Hl[0] = dx / g.range; g.Hl[1] = dy / g.range; g.Hl[2] = -dy / range2; g.Hl[3] = dx / range2;
  double HP[4];
  mul<2>(g.Hl, P2, HP, false);
  mul<2>(HP, g.Hl, g.A, true);
  return g;
}

/* measure(): expected measurement, its covariance S and the Jacobian wrt the landmark; returns the model's
 * validity flag (RngBrg: range inside the sensing limits; Victoria Park: always true, :104-151) */
bool measure(const rfsb200_model_desc& md, const double* pose, const double* Sx, const Cand<2>& c, double* zexp,
             double (&S)[4], double (&H)[4]) {
  const RB g = range_bearing(pose, pose[2], c.x, c.P);
  zexp[0] = g.range;
  zexp[1] = g.bearing;
  const double dx = c.x[0] - pose[0], dy = c.x[1] - pose[1], r2 = g.range * g.range;
  const double Hr[6] = {-dx / g.range, -dy / g.range, 0, dy / r2, -dx / r2, -1};
  double B[4] = {0, 0, 0, 0};
  if (Sx) {
    double HS[6];
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Hr[i * 3 + k] * Sx[k * 3 + j];
        HS[i * 3 + j] = s;
      }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += HS[i * 3 + k] * Hr[j * 3 + k];
        B[i * 2 + j] = s;
      }
  }
  for (int k = 0; k < 4; k++) { S[k] = g.A[k] + B[k] + md.R[k]; H[k] = g.Hl[k]; }
  return !(g.range > md.range_max || g.range < md.range_min);
}
bool measure(const rfsb200_model_desc& md, const double* pose, const double*, const Cand<3>& c, double* zexp,
             double (&S)[9], double (&H)[9]) {
  const double P2[4] = {c.P[0], c.P[1], c.P[3], c.P[4]};
  const RB g = range_bearing(pose, pose[2] - PI / 2, c.x, P2);
  zexp[0] = g.range;
  zexp[1] = g.bearing;
  zexp[2] = c.x[2];
  for (int k = 0; k < 9; k++) { S[k] = 0; H[k] = 0; }
  S[0] = g.A[0] + md.R[0]; S[1] = g.A[1] + md.R[1]; S[3] = g.A[2] + md.R[3]; S[4] = g.A[3] + md.R[4];
  S[8] = c.P[8] + md.R[8] + pow(g.range, 2) * md.Slb;
  H[0] = g.Hl[0]; H[1] = g.Hl[1]; H[3] = g.Hl[2]; H[4] = g.Hl[3]; H[8] = 1;
  return true;
}

/* calculateInnovation: src/KalmanFilter_RngBrg.cpp:52-65 (range test before the wrap) and
 * include/KalmanFilter_VictoriaPark.hpp:56-74 (wrap first) */
bool innovation(const rfsb200_model_desc& md, const double* zexp, const double* z, double (&v)[2]) {
  v[0] = z[0] - zexp[0];
  v[1] = z[1] - zexp[1];
  if (md.innov_thr_range > 0 && fabs(v[0]) > md.innov_thr_range) return false;
  while (v[1] > PI) v[1] -= 2 * PI;
  while (v[1] < -PI) v[1] += 2 * PI;
  if (md.innov_thr_bearing > 0 && fabs(v[1]) > md.innov_thr_bearing) return false;
  return true;
}
bool innovation(const rfsb200_model_desc& md, const double* zexp, const double* z, double (&v)[3]) {
  for (int k = 0; k < 3; k++) v[k] = z[k] - zexp[k];
  while (v[1] > PI) v[1] -= 2 * PI;
  while (v[1] < -PI) v[1] += 2 * PI;
  if (md.innov_thr_range > 0 && fabs(v[0]) > md.innov_thr_range) return false;
  if (md.innov_thr_bearing > 0 && fabs(v[1]) > md.innov_thr_bearing) return false;
  return true;
}

/* inverseMeasure: src/MeasurementModel_RngBrg.cpp:117-136, src/MeasurementModel_VictoriaPark.cpp:75-102 */
void inverse_measure(const rfsb200_model_desc& md, const double* pose, const double* z, Cand<2>& c) {
  const double a = pose[2] + z[1];
  c.x[0] = pose[0] + z[0] * cos(a);
  c.x[1] = pose[1] + z[0] * sin(a);
  const double Hi[4] = {cos(a), -z[0] * sin(a), sin(a), z[0] * cos(a)};
  const double R2[4] = {md.R[0], md.R[1], md.R[2], md.R[3]};
  double T[4];
  mul<2>(Hi, R2, T, false);
  mul<2>(T, Hi, c.P, true);
}
void inverse_measure(const rfsb200_model_desc& md, const double* pose, const double* z, Cand<3>& c) {
  const double a = pose[2] - PI / 2 + z[1];
  c.x[0] = pose[0] + z[0] * cos(a);
  c.x[1] = pose[1] + z[0] * sin(a);
  c.x[2] = z[2];
  const double Hi[4] = {cos(a), -z[0] * sin(a), sin(a), z[0] * cos(a)};
  const double R2[4] = {md.R[0], md.R[1], md.R[3], md.R[4]};
  double T[4], C2[4];
  mul<2>(Hi, R2, T, false);
  mul<2>(T, Hi, C2, true);
  for (int k = 0; k < 9; k++) c.P[k] = 0;
  c.P[0] = C2[0]; c.P[1] = C2[1]; c.P[3] = C2[2]; c.P[4] = C2[3];
  c.P[8] = md.R[8];
}

template <int D>
void run(phd_birth_io* io) {
  constexpr int NC = D * (D + 1) / 2;
  const rfsb200_model_desc& md = *io->model;
  const int N = io->N;
  std::vector<std::vector<Cand<D> > > lists(N);
  for (int i = 0; i < N; i++) {
    lists[i].resize(io->cand_n[i]);
    for (int k = 0; k < io->cand_n[i]; k++) {
      Cand<D>& c = lists[i][k];
      const size_t s = (size_t)i * io->cand_cap + k;
      for (int d = 0; d < D; d++) c.x[d] = io->cand_mean[s * D + d];
      for (int r = 0, q = 0; r < D; r++)
        for (int cc = r; cc < D; cc++, q++) c.P[r * D + cc] = c.P[cc * D + r] = io->cand_cov[s * NC + q];
      c.nSupport = (unsigned)io->cand_support[s];
      c.nChecks = (unsigned)io->cand_checks[s];
    }
  }
  const double sup2 = io->support_dist * io->support_dist;
  for (int i = 0; i < N; i++) { /* ascending i, in place: a copy sees what its parent slot holds at that moment */
    if (io->resample_occurred) {
      const int par = io->parent[i];
      if (par != i) {
        io->unused[i] = io->unused[par];
        lists[i] = lists[par];
      }
    }
    std::vector<Cand<D> >& cand = lists[i];
    const double* pose = io->pose + 3 * i;
    double Sx[9];
    const double* Sxp = NULL;
    if (D == 2 && io->pose_cov) {
      const double* c6 = io->pose_cov + 6 * i;
      const double full[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
      memcpy(Sx, full, sizeof(full));
      Sxp = Sx;
    }
    int nAdd = 0;
    auto add_real = [&](const Cand<D>& c) {
      if (nAdd < io->add_cap) {
        const size_t s = (size_t)i * io->add_cap + nAdd;
        for (int d = 0; d < D; d++) io->add_mean[s * D + d] = c.x[d];
        for (int r = 0, q = 0; r < D; r++)
          for (int cc = r; cc < D; cc++, q++) io->add_cov[s * NC + q] = c.P[r * D + cc];
      }
      nAdd++;
    };
    const unsigned nfov = (unsigned)io->nfov[i];
    for (int zi = io->nZ - 1; zi >= 0; zi--) { /* pop_back on an ascending list: descending index (:1013-1018) */
      if (!((io->unused[i] >> zi) & 1ull)) continue;
      const double* z = io->Z + (size_t)D * zi;
      bool isNew = true;
      for (size_t k = 0; k < cand.size(); k++) {
        Cand<D>& c = cand[k];
        double zexp[D], S[D * D], H[D * D], Sinv[D * D];
        const bool valid = measure(md, pose, Sxp, c, zexp, S, H); /* :1027-1029: the flag is not looked at here */
        inv(S, Sinv);
        double e[D];
        for (int d = 0; d < D; d++) e[d] = z[d] - zexp[d]; /* RandomVec::mahalanobisDist2: plain difference */
        double d2 = 0;
        for (int r = 0; r < D; r++) {
          double t = 0;
          for (int q = 0; q < D; q++) t += e[q] * Sinv[q * D + r];
          d2 += t * e[r];
        }
        if (d2 <= sup2) {
          /* kfs_[0].correct(x, z, *it, *it) (include/KalmanFilter.hpp:211-258); its result is ignored */
          double v[D];
          if (valid && innovation(md, zexp, z, v)) {
            double PHt[D * D], K[D * D], KH[D * D], IKH[D * D], Pu[D * D];
            mul<D>(c.P, H, PHt, true);
            mul<D>(PHt, Sinv, K, false);
            mul<D>(K, H, KH, false);
            for (int q = 0; q < D * D; q++) IKH[q] = ((q % (D + 1) == 0) ? 1.0 : 0.0) - KH[q];
            mul<D>(IKH, c.P, Pu, false);
            double xn[D];
            for (int r = 0; r < D; r++) {
              double t = 0;
              for (int q = 0; q < D; q++) t += K[r * D + q] * v[q];
              xn[r] = c.x[r] + t;
            }
            for (int r = 0; r < D; r++) {
              c.x[r] = xn[r];
              for (int q = 0; q < D; q++) c.P[r * D + q] = (Pu[r * D + q] + Pu[q * D + r]) / 2;
            }
          }
          c.nSupport++;
          isNew = false;
          break;
        }
      }
      if (isNew) {
        Cand<D> c;
        c.nSupport = 1;
        c.nChecks = 0;
        inverse_measure(md, pose, z, c);
        if (io->count_thr == 1 || nfov <= io->cur_count_thr) add_real(c);
        else cand.push_back(c);
      }
    }
    io->unused[i] = 0;
    /* :1056-1075.  Erasing the LAST candidate leaves the inner loop with it == end(); the for statement then
     * increments end(), which on libstdc++'s circular list is begin(): the remaining candidates get another pass. */
    size_t it = 0;
    while (it < cand.size()) {
      cand[it].nChecks++;
      bool wrapped = false;
      while (cand[it].nSupport >= io->count_thr || cand[it].nChecks > io->check_thr || nfov <= io->cur_count_thr) {
        if (cand[it].nSupport >= io->count_thr) add_real(cand[it]);
        else if (nfov <= io->cur_count_thr) add_real(cand[it]);
        cand.erase(cand.begin() + it);
        if (it < cand.size()) cand[it].nChecks++;
        else { wrapped = true; break; }
      }
      if (wrapped) it = 0;
      else it++;
    }
    io->add_n[i] = nAdd;
  }
  for (int i = 0; i < N; i++) {
    const int n = (int)lists[i].size();
    io->cand_n[i] = n;
    for (int k = 0; k < n && k < io->cand_cap; k++) {
      const Cand<D>& c = lists[i][k];
      const size_t s = (size_t)i * io->cand_cap + k;
      for (int d = 0; d < D; d++) io->cand_mean[s * D + d] = c.x[d];
      for (int r = 0, q = 0; r < D; r++)
        for (int cc = r; cc < D; cc++, q++) io->cand_cov[s * NC + q] = c.P[r * D + cc];
      io->cand_support[s] = (int32_t)c.nSupport;
      io->cand_checks[s] = (int32_t)c.nChecks;
    }
  }
}

}  // namespace

extern "C" int phd_oracle_birth_candidates(phd_birth_io* io) {
  if (!io || !io->model || io->N <= 0 || io->nZ < 0 || io->nZ > 64) return -1;
  if (io->model->model_id == RFSB200_MODEL_VICTORIAPARK) run<3>(io);
  else run<2>(io);
  return 0;
}
