// birth_kernels.cuh — RBPHDFilter::addBirthGaussians() in its candidate-list form on the device, sm_100a.
//
// Reference: include/RBPHDFilter.hpp:1000-1080 (used when birthGaussianMeasurementCountThreshold_ != 1, e.g.
// cfg/rbphdslam_VictoriaPark_artificialClutter.xml:71-77).  Per particle, strictly in sequence: every measurement the
// last update left unused (consumed from the back of unused_measurements_, i.e. descending index) either supports the
// first candidate within birthGaussianMeasurementSupportDist_ (Mahalanobis distance of the expected measurement;
// the candidate takes a Kalman correction) or becomes a new candidate at inverseMeasure(pose, z); afterwards every
// candidate is checked once (nChecks++) and leaves the list when it has enough support (it becomes a real Gaussian),
// is too old, or the particle has few landmarks in view (real Gaussian too).  The list logic is serial by
// construction, so ONE THREAD owns a particle; all arithmetic is fp64 (as on the host) and only the Gaussians that
// become real are rounded to the map's element type.  A few microseconds for thousands of particles: the kernel is
// bound by the latency of one thread's chain, not by any throughput limit — what it buys is that the masks, the
// candidate lists and the maps never leave the device between two updates.
//
// Plugin arithmetic restated here in fp64: MeasurementModel_RngBrg::measure / inverseMeasure
// (src/MeasurementModel_RngBrg.cpp:70-136), MeasurementModel_VictoriaPark::measure / inverseMeasure
// (src/MeasurementModel_VictoriaPark.cpp:75-151), KalmanFilter::correct (include/KalmanFilter.hpp:211-258) with
// calculateInnovation of KalmanFilter_RngBrg (src/KalmanFilter_RngBrg.cpp:52-65) and KalmanFilter_VictoriaPark
// (include/KalmanFilter_VictoriaPark.hpp:56-74), RandomVec::mahalanobisDist2 (include/RandomVec.hpp:387-394).
#pragma once

#include "phd_kernels.cuh"

namespace rfsb200 {

// one candidate record: mean[D] | covariance, upper triangle [D (D + 1) / 2] | nSupportingMeasurements | nChecks
__host__ __device__ constexpr int cand_rec(int D) { return D + D * (D + 1) / 2 + 2; }

struct BirthCandParams {
  int N, cap, cand_cap, nZ, pass, pcov_mode;
  const int* parent;                 // [N] or NULL (no resampling since the last call)
  const int* level;                  // [N] with parent: the launch (pass) in which the particle is processed
  const double* pose64;              // [N][3] pose of the last update
  const void* pcov;                  // T[8] per particle (mode 2) or shared (mode 1): 2-D model only
  unsigned long long* unused;        // [N]
  const int* nfov;                   // [N]
  const double* cand_in;             // [N][cand_cap][REC]
  const int* cand_n_in;              // [N]
  double* cand_out;
  int* cand_n_out;
  void* gm;                          // T planes of the committed maps
  int* cnt;
  int* flags;
  double R[9], Slb, range_min, range_max, thr_r, thr_b;
  double support_d2, birth_w;
  unsigned count_thr, check_thr, cur_thr;
  double Z[MAX_Z * 3];
};

template <int D>
struct BCand {
  double x[D];
  double P[D * D];   // full, symmetric
};

namespace bc {
constexpr double PI_D = 3.14159265358979323846;

template <int D>
__device__ __forceinline__ void mm(const double* A, const double* B, double* C, bool bt) {
#pragma unroll
  for (int r = 0; r < D; r++)
#pragma unroll
    for (int c = 0; c < D; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < D; k++) s += A[r * D + k] * (bt ? B[c * D + k] : B[k * D + c]);
      C[r * D + c] = s;
    }
}
__device__ __forceinline__ void inv(const double (&A)[4], double (&Ai)[4]) {
  const double id = 1.0 / (A[0] * A[3] - A[1] * A[2]);
  Ai[0] = A[3] * id; Ai[1] = -A[1] * id; Ai[2] = -A[2] * id; Ai[3] = A[0] * id;
}
__device__ __forceinline__ void inv(const double (&A)[9], double (&Ai)[9]) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// range / bearing of a point seen from (pose, heading th) with the Jacobian wrt the point and Hl P2 Hl^T
struct RB {
  double range, bearing, Hl[4], A[4], dx, dy;
};
__device__ __forceinline__ RB range_bearing(const double* pose, double th, const double* lx, const double* P2) {
  RB g;
  g.dx = lx[0] - pose[0];
  g.dy = lx[1] - pose[1];
  const double range2 = g.dx * g.dx + g.dy * g.dy;
  g.range = sqrt(range2);
  g.bearing = atan2(g.dy, g.dx) - th;
  while (g.bearing > PI_D) g.bearing -= 2 * PI_D;
  while (g.bearing < -PI_D) g.bearing += 2 * PI_D;
  g.Hl[0] = g.dx / g.range; g.Hl[1] = g.dy / g.range; g.Hl[2] = -g.dy / range2; g.Hl[3] = g.dx / range2;
  double HP[4];
  mm<2>(g.Hl, P2, HP, false);
  mm<2>(HP, g.Hl, g.A, true);
  return g;
}

// measure(): expected measurement, innovation covariance, Jacobian; returns the model's validity flag
__device__ __forceinline__ bool measure(const BirthCandParams& p, const double* pose, const double* Sx, const BCand<2>& c,
                                        double* zexp, double (&S)[4], double (&H)[4]) {
  const RB g = range_bearing(pose, pose[2], c.x, c.P);
  zexp[0] = g.range;
  zexp[1] = g.bearing;
  double B[4] = {0, 0, 0, 0};
  if (Sx) {   // Hr Sx Hr^T, Hr = d(range, bearing) / d(pose)
    const double r2 = g.range * g.range;
    const double Hr[6] = {-g.dx / g.range, -g.dy / g.range, 0, g.dy / r2, -g.dx / r2, -1};
    double HS[6];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) s += Hr[i * 3 + k] * Sx[k * 3 + j];
        HS[i * 3 + j] = s;
      }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) s += HS[i * 3 + k] * Hr[j * 3 + k];
        B[i * 2 + j] = s;
      }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) { S[k] = g.A[k] + B[k] + p.R[k]; H[k] = g.Hl[k]; }
  return !(g.range > p.range_max || g.range < p.range_min);
}
__device__ __forceinline__ bool measure(const BirthCandParams& p, const double* pose, const double*, const BCand<3>& c,
                                        double* zexp, double (&S)[9], double (&H)[9]) {
  const double P2[4] = {c.P[0], c.P[1], c.P[3], c.P[4]};
  const RB g = range_bearing(pose, pose[2] - PI_D / 2, c.x, P2);   // the sensor looks along theta - pi / 2; zero pose covariance
  zexp[0] = g.range;
  zexp[1] = g.bearing;
  zexp[2] = c.x[2];
#pragma unroll
  for (int k = 0; k < 9; k++) { S[k] = 0; H[k] = 0; }
  S[0] = g.A[0] + p.R[0]; S[1] = g.A[1] + p.R[1]; S[3] = g.A[2] + p.R[3]; S[4] = g.A[3] + p.R[4];
  S[8] = c.P[8] + p.R[8] + g.range * g.range * p.Slb;
  H[0] = g.Hl[0]; H[1] = g.Hl[1]; H[3] = g.Hl[2]; H[4] = g.Hl[3]; H[8] = 1;
  return true;
}

__device__ __forceinline__ bool innovation(const BirthCandParams& p, const double* zexp, const double* z, double (&v)[2]) {
  v[0] = z[0] - zexp[0];
  v[1] = z[1] - zexp[1];
  if (p.thr_r > 0 && fabs(v[0]) > p.thr_r) return false;   // range test before the wrap (KalmanFilter_RngBrg)
  while (v[1] > PI_D) v[1] -= 2 * PI_D;
  while (v[1] < -PI_D) v[1] += 2 * PI_D;
  if (p.thr_b > 0 && fabs(v[1]) > p.thr_b) return false;
  return true;
}
__device__ __forceinline__ bool innovation(const BirthCandParams& p, const double* zexp, const double* z, double (&v)[3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) v[k] = z[k] - zexp[k];
  while (v[1] > PI_D) v[1] -= 2 * PI_D;   // wrap first (KalmanFilter_VictoriaPark)
  while (v[1] < -PI_D) v[1] += 2 * PI_D;
  if (p.thr_r > 0 && fabs(v[0]) > p.thr_r) return false;
  if (p.thr_b > 0 && fabs(v[1]) > p.thr_b) return false;
  return true;
}

__device__ __forceinline__ void inverse_measure(const BirthCandParams& p, const double* pose, const double* z, BCand<2>& c) {
  double sn, cs;
  sincos(pose[2] + z[1], &sn, &cs);
  c.x[0] = pose[0] + z[0] * cs;
  c.x[1] = pose[1] + z[0] * sn;
  const double Hi[4] = {cs, -z[0] * sn, sn, z[0] * cs};
  const double R2[4] = {p.R[0], p.R[1], p.R[2], p.R[3]};
  double T[4];
  mm<2>(Hi, R2, T, false);
  mm<2>(T, Hi, c.P, true);
}
__device__ __forceinline__ void inverse_measure(const BirthCandParams& p, const double* pose, const double* z, BCand<3>& c) {
  double sn, cs;
  sincos(pose[2] - PI_D / 2 + z[1], &sn, &cs);
  c.x[0] = pose[0] + z[0] * cs;
  c.x[1] = pose[1] + z[0] * sn;
  c.x[2] = z[2];
  const double Hi[4] = {cs, -z[0] * sn, sn, z[0] * cs};
  const double R2[4] = {p.R[0], p.R[1], p.R[3], p.R[4]};
  double T[4], C2[4];
  mm<2>(Hi, R2, T, false);
  mm<2>(T, Hi, C2, true);
#pragma unroll
  for (int k = 0; k < 9; k++) c.P[k] = 0;
  c.P[0] = C2[0]; c.P[1] = C2[1]; c.P[3] = C2[2]; c.P[4] = C2[3];
  c.P[8] = p.R[8];
}

template <int D>
__device__ __forceinline__ void load_cand(const double* rec, BCand<D>& c) {
#pragma unroll
  for (int d = 0; d < D; d++) c.x[d] = rec[d];
  int q = D;
#pragma unroll
  for (int r = 0; r < D; r++)
#pragma unroll
    for (int cc = r; cc < D; cc++, q++) c.P[r * D + cc] = c.P[cc * D + r] = rec[q];
}
template <int D>
__device__ __forceinline__ void store_cand(double* rec, const BCand<D>& c) {
#pragma unroll
  for (int d = 0; d < D; d++) rec[d] = c.x[d];
  int q = D;
#pragma unroll
  for (int r = 0; r < D; r++)
#pragma unroll
    for (int cc = r; cc < D; cc++, q++) rec[q] = c.P[r * D + cc];
}
}  // namespace bc

// After a resampling the reference's loop (ascending i, in place) lets particle i take over the list its parent slot
// holds AT THAT MOMENT: a parent slot above i is still untouched (its list as it was: the input buffer), a parent slot
// below i has had its own turn already (the list AFTER it, and no unused measurements are left to copy, so only the
// check loop acts).  Parent ids are not slot numbers after the first resampling, so the lower parent may itself have
// taken its list from a still lower slot: level[i] = 0 if parent >= i, else level[parent] + 1 (computed by the host),
// one launch per level, a particle of level l reads the result its parent wrote in an earlier launch.
template <typename T, int D>
__global__ void birth_candidates_kernel(const BirthCandParams p) {
  constexpr int NC = D * (D + 1) / 2;
  constexpr int REC = cand_rec(D);
  constexpr int NPL = D + NC + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  int par = p.parent ? p.parent[i] : i;
  if (par < 0 || par >= p.N) par = i;
  if ((p.parent ? p.level[i] : 0) != p.pass) return;
  const bool after = par < i;   // the parent slot has had its turn: take its result
  const double* src = (after ? p.cand_out : p.cand_in) + (size_t)par * p.cand_cap * REC;
  int n = (after ? p.cand_n_out : p.cand_n_in)[par];
  n = n < 0 ? 0 : (n > p.cand_cap ? p.cand_cap : n);
  double* dst = p.cand_out + (size_t)i * p.cand_cap * REC;
  for (int k = 0; k < n * REC; k++) dst[k] = src[k];

  unsigned long long mask = after ? 0ull : p.unused[i];
  if (p.nZ < 64) mask &= (1ull << p.nZ) - 1ull;
  const double pose[3] = {p.pose64[3 * i], p.pose64[3 * i + 1], p.pose64[3 * i + 2]};
  double Sx[9];
  const double* Sxp = nullptr;
  if (D == 2 && p.pcov_mode != 0) {
    const T* c6 = reinterpret_cast<const T*>(p.pcov) + (p.pcov_mode == 2 ? (size_t)8 * i : 0);
    Sx[0] = (double)c6[0]; Sx[1] = (double)c6[1]; Sx[2] = (double)c6[2];
    Sx[3] = (double)c6[1]; Sx[4] = (double)c6[3]; Sx[5] = (double)c6[4];
    Sx[6] = (double)c6[2]; Sx[7] = (double)c6[4]; Sx[8] = (double)c6[5];
    Sxp = Sx;
  }
  const unsigned nfov = (unsigned)p.nfov[i];
  T* g = reinterpret_cast<T*>(p.gm) + (size_t)i * NPL * p.cap;
  int cnt = p.cnt[i];
  cnt = cnt < 0 ? 0 : (cnt > p.cap ? p.cap : cnt);
  int flags = 0;
  auto add_real = [&](const BCand<D>& c) {   // GaussianMixture::addGaussian: behind the existing ones
    if (cnt < p.cap) {
      int pl = 0;
#pragma unroll
      for (int d = 0; d < D; d++, pl++) g[(size_t)pl * p.cap + cnt] = (T)c.x[d];
#pragma unroll
      for (int r = 0; r < D; r++)
#pragma unroll
        for (int cc = r; cc < D; cc++, pl++) g[(size_t)pl * p.cap + cnt] = (T)c.P[r * D + cc];
      g[(size_t)pl * p.cap + cnt] = (T)p.birth_w;
      cnt++;
    } else {
      flags |= FLAG_OVERFLOW | FLAG_BIRTH_OVERFLOW;
    }
  };

  for (int zi = p.nZ - 1; zi >= 0; zi--) {
    if (!((mask >> zi) & 1ull)) continue;
    double z[D];
#pragma unroll
    for (int d = 0; d < D; d++) z[d] = p.Z[D * zi + d];
    bool isNew = true;
    for (int k = 0; k < n; k++) {
      double* rec = dst + (size_t)k * REC;
      BCand<D> c;
      bc::load_cand<D>(rec, c);
      double zexp[D], S[D * D], H[D * D], Sinv[D * D];
      const bool valid = bc::measure(p, pose, Sxp, c, zexp, S, H);
      bc::inv(S, Sinv);
      double e[D];
#pragma unroll
      for (int d = 0; d < D; d++) e[d] = z[d] - zexp[d];   // plain difference, no wrap (RandomVec::mahalanobisDist2)
      double d2 = 0;
#pragma unroll
      for (int r = 0; r < D; r++) {
        double t = 0;
#pragma unroll
        for (int q = 0; q < D; q++) t += e[q] * Sinv[q * D + r];
        d2 += t * e[r];
      }
      if (d2 <= p.support_d2) {
        double v[D];
        if (valid && bc::innovation(p, zexp, z, v)) {   // KalmanFilter::correct(x, z, *it, *it); a refused update changes nothing
          double PHt[D * D], K[D * D], KH[D * D], Pu[D * D];
          bc::mm<D>(c.P, H, PHt, true);
          bc::mm<D>(PHt, Sinv, K, false);
          bc::mm<D>(K, H, KH, false);
#pragma unroll
          for (int q = 0; q < D * D; q++) KH[q] = ((q % (D + 1) == 0) ? 1.0 : 0.0) - KH[q];
          bc::mm<D>(KH, c.P, Pu, false);
          BCand<D> u;
#pragma unroll
          for (int r = 0; r < D; r++) {
            double t = 0;
#pragma unroll
            for (int q = 0; q < D; q++) t += K[r * D + q] * v[q];
            u.x[r] = c.x[r] + t;
#pragma unroll
            for (int q = 0; q < D; q++) u.P[r * D + q] = (Pu[r * D + q] + Pu[q * D + r]) / 2;
          }
          bc::store_cand<D>(rec, u);
        }
        rec[REC - 2] += 1.0;   // nSupportingMeasurements++
        isNew = false;
        break;
      }
    }
    if (isNew) {
      BCand<D> c;
      bc::inverse_measure(p, pose, z, c);
      if (p.count_thr == 1u || nfov <= p.cur_thr) {
        add_real(c);
      } else if (n < p.cand_cap) {
        double* rec = dst + (size_t)n * REC;
        bc::store_cand<D>(rec, c);
        rec[REC - 2] = 1.0;
        rec[REC - 1] = 0.0;
        n++;
      } else {
        flags |= FLAG_CAND_OVERFLOW;
      }
    }
  }
  p.unused[i] = 0ull;

  // the check pass (:1056-1075).  Erasing the LAST element of the list ends the reference's inner loop with the
  // iterator at end(); its for statement then increments end(), which on libstdc++'s circular list is begin(): the
  // remaining candidates are visited again (nChecks++ each).  Reproduced as built.
  const double count_thr = (double)p.count_thr, check_thr = (double)p.check_thr;
  int it = 0;
  while (it < n) {
    dst[(size_t)it * REC + REC - 1] += 1.0;
    bool wrapped = false;
    while (dst[(size_t)it * REC + REC - 2] >= count_thr || dst[(size_t)it * REC + REC - 1] > check_thr || nfov <= p.cur_thr) {
      if (dst[(size_t)it * REC + REC - 2] >= count_thr || nfov <= p.cur_thr) {
        BCand<D> c;
        bc::load_cand<D>(dst + (size_t)it * REC, c);
        add_real(c);
      }
      for (int k = (it + 1) * REC; k < n * REC; k++) dst[k - REC] = dst[k];   // erase(it)
      n--;
      if (it < n) dst[(size_t)it * REC + REC - 1] += 1.0;
      else { wrapped = true; break; }
    }
    if (wrapped) it = 0;
    else it++;
  }
  p.cand_n_out[i] = n;
  p.cnt[i] = cnt;
  if (flags) p.flags[i] |= flags;
}

}  // namespace rfsb200
