// phd_vp_kernels.cuh — the fused per-particle PHD measurement update for the Victoria Park plugin
// set (rfs::MeasurementModel_VictoriaPark + KalmanFilter_VictoriaPark, BASELINE config 5), sm_100a.
//
// Same organisation as phd_kernels.cuh (one warp per particle, persistent CTAs, one launch per
// update) with 3-D landmarks (x, y, diameter), 3-D measurements (range, bearing, diameter) and the
// probability of detection evaluated against the raw lidar scan:
//   planes in HBM  x, y, d, Pxx, Pxy, Pxd, Pyy, Pyd, Pdd, w            (10 per particle, [N][10][cap])
//   S0  TMA bulk loads of the 10 planes
//   S1  probabilityOfDetection (src/MeasurementModel_VictoriaPark.cpp:153-266) in fp64 from the stored
//       state — it is a chain of floor / ceil / table look-ups whose outcome must not depend on the
//       storage precision — then the EKF corrector (include/KalmanFilter.hpp:261-342 with
//       measure() of src/MeasurementModel_VictoriaPark.cpp:104-151 and calculateInnovation() of
//       include/KalmanFilter_VictoriaPark.hpp:56-74) in T
//   S2-S4, S5, S7, S8 as in the 2-D kernel (the filter template code is the same,
//       include/RBPHDFilter.hpp:543-997); S5's partition logic and S8 are the shared device functions
//   S6  greedy GaussianMixture::merge in the reference's order with a conservative distance
//       pre-filter (|e|^2 > t^2 max(tr P_i, tr P_j) cannot pass either Mahalanobis test)
#pragma once
#include "phd_kernels.cuh"

namespace rfsb200 {

constexpr int VP_NPL = 10;   // planes of the state
constexpr int VP_WP = 9;     // weight plane
constexpr int VP_SCAN_MAX = 720;
constexpr int VP_PD_MAX = 16;
constexpr int VP_EP = 12;    // scalars per eval point in the L-table stage

template <typename T>
struct VPParams {
  KParams<T> k;   // shapes, state pointers, filter configuration, reductions; of the model fields only
                  // kappa / log_kappa / log_clutter_integral / thr_r / thr_b are used
  T R00, R01, R10, R11, R22, Slb;
  double rmin, rmax, bmin, bmax, buf_pd;
  double pd_table[VP_PD_MAX];
  int pd_n, scan_n;
  const double* scan;   // device, [scan_n]
};

// ---- probabilityOfDetection, fp64 ---------------------------------------------------------------
struct VPGeom {
  double rmin, rmax, bmin, bmax, buf_pd;
  const double* pd;     // [pd_n]  (shared memory)
  const double* scan;   // [scan_n] (shared memory)
  int pd_n, scan_n;
};

constexpr double VP_PI = 3.14159265358979323846;

// probabilityOfDetection2 (src/MeasurementModel_VictoriaPark.cpp:202-266); th = pose theta - pi/2
__device__ __forceinline__ double vp_pd2(const VPGeom& g, double px, double py, double th, double lx, double ly,
                                         double ld, bool& close) {
  close = false;
  const double dx = lx - px, dy = ly - py;
  const double dist = sqrt(dx * dx + dy * dy);
  double angle = atan2(dy, dx) - th;
  while (angle > VP_PI) angle -= 2 * VP_PI;
  while (angle < -VP_PI) angle += 2 * VP_PI;
  if (angle > g.bmax || angle < g.bmin || dist < g.rmin || dist > g.rmax) return 0.0;
  const double mr = ld / 2;
  const double gamma = atan(mr / dist);
  const int maxNumPoints = (int)floor(2 * gamma * 720.0 / (2 * VP_PI));
  const bool in_tab = maxNumPoints >= 0 && g.pd_n > maxNumPoints;   // size_t > int compare of the reference
  if (in_tab && g.pd[maxNumPoints] == 0.0) return 0.0;
  if (in_tab && g.pd[maxNumPoints] < g.buf_pd) close = true;
  int minb = (int)ceil((angle - gamma) * 720.0 / (2 * VP_PI));
  int maxb = minb + maxNumPoints;
  while (minb >= 720) minb -= 720;
  while (minb < 0) minb += 720;
  while (maxb >= 720) maxb -= 720;
  while (maxb < 0) maxb += 720;
  int numPoints = 0;
  const double minrange = dist - mr - 6 * 0.03;
  if ((maxb - minb + 720) % 720 > 0) {
    for (int b = minb; b != maxb; b = (b + 1) % 720) {
      // beyond the scan the reference reads past the end of its vector (361 entries, indices up to 719):
      // whatever lies there is neither a plausible range nor exactly 0, so such a beam counts no point
      const double s = b < g.scan_n ? g.scan[b] : __longlong_as_double(0x7ff8000000000000LL);
      if (s > minrange || s == 0.0) numPoints++;
    }
  }
  if (numPoints >= g.pd_n) numPoints = g.pd_n - 1;
  if (g.pd[numPoints] == 0.0) close = false;
  return g.pd[numPoints];
}

// probabilityOfDetection (:153-199).  pth = pose theta; (pxx, pxy, pyy) = position block of the covariance
__device__ __forceinline__ double vp_pd(const VPGeom& g, double px, double py, double pth, double lx, double ly,
                                        double ld, double pxx, double pxy, double pyy, bool& close) {
  const double th = pth - VP_PI / 2;
  const double dx = lx - px, dy = ly - py;
  const double r0 = sqrt(dx * dx + dy * dy);
  double b0 = atan2(dy, dx) - th;
  while (b0 > VP_PI) b0 -= 2 * VP_PI;
  while (b0 < -VP_PI) b0 += 2 * VP_PI;
  const double angle = atan2(b0, r0) + pth;   // :164, as written in the reference
  double sn, cs;
  sincos(angle, &sn, &cs);
  const double ex = -sn, ey = cs;
  double sd = (ex * pxx + ey * pxy) * ex + (ex * pxy + ey * pyy) * ey;
  sd = 3 * sqrt(sd);
  sd = (sd < 0.2) ? 0.2 : sd;   // std::max(sd, 0.2)
  close = false;
  // every evaluation point lies within sd + 2 d of the mean: a component that far outside the
  // range limits has P_D = 0 at all of them
  if (ld > 0.0 && sd < 1e300) {
    const double reach = (sd + 2 * ld) * 1.000001 + 1e-9;
    if (r0 > g.rmax + reach || r0 < g.rmin - reach) return 0.0;
  }
  double pmin = 1e300, pmax = -1e300;
  for (int i = 1; (i - 1) * (2 * ld) < sd; i++) {
    const double s = i * 2 * ld;
    double v = vp_pd2(g, px, py, th, lx + s * ex, ly + s * ey, ld, close);
    pmin = v < pmin ? v : pmin;
    pmax = v > pmax ? v : pmax;
    v = vp_pd2(g, px, py, th, lx - s * ex, ly - s * ey, ld, close);
    pmin = v < pmin ? v : pmin;
    pmax = v > pmax ? v : pmax;
    if (i > 4096) break;   // non-positive diameter: the reference never terminates
  }
  const double v = vp_pd2(g, px, py, th, lx, ly, ld, close);
  pmin = v < pmin ? v : pmin;
  pmax = v > pmax ? v : pmax;
  if (pmin == 0.0 && pmax > 0.0) close = true;
  return pmax;
}

// ---- probabilityOfDetection, fp32 with error bounds (fp32 kernels) -----------------------------------------------
// The detection probability is a chain of discrete decisions (range / bearing limits, floor / ceil onto the half-degree
// beams, "is this return behind the trunk") whose outcome must be the reference's, and a table look-up.  The fp32 kernels
// evaluate the geometry in fp32 together with a bound on the rounding error of every compared quantity; a decision that
// lies inside its bound is "too close to call" and the whole component is re-evaluated in fp64 (vp_pd above) — a few
// components in a thousand.  The values returned are entries of the fp64 table either way, so a decisive fp32
// evaluation returns the same bits as the fp64 one.
struct VPFast {
  float rmin, rmax, bmin, bmax;
  const float* scan;   // [scan_n] float copy of the scan (shared memory); exact zeros stay exact
};

// probabilityOfDetection2 at the point pose + (dx, dy); e_pos = bound on the error of dx and dy.  false: undecided.
// The value is an entry of the P_D table or the literal 0: returned as `code` (table index, VP_PD_ZERO = literal 0) so
// that a lane can hand its result to another one in a single word (vp_pd_warp).
constexpr int VP_PD_ZERO = 63;
__device__ __forceinline__ bool vp_pd2_fast_code(const VPGeom& g, const VPFast& f, float dx, float dy, float th, float ld,
                                                 float e_pos, int& code, bool& close) {
  constexpr float PI_F = 3.14159265358979323846f;
  close = false;
  const float d2 = dx * dx + dy * dy;
  const float dist = sqrtf(d2);
  if (!(dist > 1e-3f)) return false;
  const float e_dist = 1.5f * e_pos + 3e-7f * dist;
  float angle = M<float>::atan2_(dy, dx) - th;
  const float e_ang = 1.5f * e_pos / dist + 1.2e-6f;   // atan2 (1.1e-7 + input error), th and the subtraction (|angle| < 8)
  if (angle > PI_F) angle -= 2 * PI_F;
  if (angle < -PI_F) angle += 2 * PI_F;
  if (!(fabsf(angle) < PI_F - e_ang)) return false;   // at the wrap (or |th| beyond one turn)
  if (fabsf(angle - f.bmax) < e_ang || fabsf(angle - f.bmin) < e_ang || fabsf(dist - f.rmin) < e_dist || fabsf(dist - f.rmax) < e_dist)
    return false;
  if (angle > f.bmax || angle < f.bmin || dist < f.rmin || dist > f.rmax) { code = VP_PD_ZERO; return true; }
  const float mr = 0.5f * ld;
  const float gamma = M<float>::atan2_(mr, dist);   // = atan(mr / dist), both positive
  const float e_gam = mr * e_dist / d2 + 3e-7f;
  const float v = gamma * 229.18311805232929f;       // 2 gamma 720 / (2 pi)
  const float e_v = 229.2f * e_gam + 3e-7f * v + 1e-6f;
  const float fl = floorf(v);
  if (v - fl < e_v || fl + 1.0f - v < e_v) return false;
  const int maxNumPoints = (int)fl;
  const bool in_tab = maxNumPoints >= 0 && g.pd_n > maxNumPoints;
  if (in_tab && g.pd[maxNumPoints] == 0.0) { code = VP_PD_ZERO; return true; }
  if (in_tab && g.pd[maxNumPoints] < g.buf_pd) close = true;
  const float u = (angle - gamma) * 114.59155902616465f;   // 720 / (2 pi)
  const float e_u = 114.6f * (e_ang + e_gam) + 3e-7f * fabsf(u) + 1e-5f;
  const float cu = ceilf(u);
  if (cu - u < e_u || u - (cu - 1.0f) < e_u) return false;
  int minb = (int)cu;
  int maxb = minb + maxNumPoints;
  while (minb >= 720) minb -= 720;
  while (minb < 0) minb += 720;
  while (maxb >= 720) maxb -= 720;
  while (maxb < 0) maxb += 720;
  int numPoints = 0;
  const float minrange = dist - mr - 0.18f;
  const float e_s = e_dist + 2e-6f;
  {
    // beams minb, minb + 1, ... (mod 720) up to maxb - 1: four per trip, their loads in flight together, no branches
    // (a return too close to minrange makes the whole evaluation undecided wherever it sits, so it is only a flag)
    const int len = (maxb - minb + 720) % 720;
    bool undecided = false;
    for (int i0 = 0; i0 < len; i0 += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int i = i0 + k;
        int b = minb + i;
        b = b >= 720 ? b - 720 : b;
        const bool on = i < len && b < g.scan_n;   // (see vp_pd2: a beam beyond the scan counts no point)
        const float s = f.scan[on ? b : 0];
        const bool none = s == 0.0f;
        undecided = undecided || (on && !none && fabsf(s - minrange) < e_s + 3e-7f * s);
        numPoints += (on && (none || s > minrange)) ? 1 : 0;
      }
    }
    if (undecided) return false;
  }
  if (numPoints >= g.pd_n) numPoints = g.pd_n - 1;
  if (g.pd[numPoints] == 0.0) close = false;
  code = numPoints;
  return true;
}

__device__ __forceinline__ bool vp_pd2_fast(const VPGeom& g, const VPFast& f, float dx, float dy, float th, float ld,
                                            float e_pos, double& val, bool& close) {
  int code;
  if (!vp_pd2_fast_code(g, f, dx, dy, th, ld, e_pos, code, close)) return false;
  val = code == VP_PD_ZERO ? 0.0 : g.pd[code];
  return true;
}

// probabilityOfDetection in fp32; false: something was too close to call (the caller runs vp_pd in fp64)
__device__ __forceinline__ bool vp_pd_fast(const VPGeom& g, const VPFast& f, float px, float py, float pth, float lx, float ly,
                                           float ld, float pxx, float pxy, float pyy, double& pd, bool& close) {
  constexpr float PI_F = 3.14159265358979323846f;
  const float th = pth - 0.5f * PI_F;
  const float dx = lx - px, dy = ly - py;             // fp32 inputs: rounded once (half an ulp of the result)
  const float e0 = 4e-8f * (fabsf(dx) + fabsf(dy)) + 1e-7f;
  const float r0 = sqrtf(dx * dx + dy * dy);
  float b0 = M<float>::atan2_(dy, dx) - th;
  if (b0 > PI_F) b0 -= 2 * PI_F;
  if (b0 < -PI_F) b0 += 2 * PI_F;
  if (!(fabsf(b0) < PI_F - 2e-6f)) return false;      // the reference's wrap decides the evaluation direction here
  const float angle = M<float>::atan2_(b0, r0) + pth;
  float sn, cs;
  __sincosf(angle, &sn, &cs);
  const float ex = -sn, ey = cs;                      // direction error ~2e-6 (atan2 inputs, __sincosf)
  float sd = (ex * pxx + ey * pxy) * ex + (ex * pxy + ey * pyy) * ey;
  if (!(sd >= 0.0f) || !(ld > 0.0f)) return false;
  sd = 3.0f * sqrtf(sd);
  sd = sd < 0.2f ? 0.2f : sd;
  if (!(sd < 1e30f)) return false;
  close = false;
  {
    const float reach = (sd + 2.0f * ld) * 1.001f + 1e-3f;   // wider than the fp64 bound: skipping less is always right
    if (r0 > f.rmax + reach || r0 < f.rmin - reach) { pd = 0.0; return true; }
  }
  double pmin = 1e300, pmax = -1e300;
  const float step = 2.0f * ld;
  const float e_sd = 2e-5f * sd + 1e-6f;
  for (int i = 1; i <= 4096; i++) {
    const float lim = (float)(i - 1) * step;
    if (fabsf(lim - sd) < e_sd) return false;        // how many evaluation points there are is a decision too
    if (!(lim < sd)) break;
    const float s = (float)i * step;
    const float e_pos = e0 + 4e-6f * s + 2e-6f;       // direction error x offset, roundings of the sums
    double v;
    if (!vp_pd2_fast(g, f, dx + s * ex, dy + s * ey, th, ld, e_pos, v, close)) return false;
    pmin = v < pmin ? v : pmin;
    pmax = v > pmax ? v : pmax;
    if (!vp_pd2_fast(g, f, dx - s * ex, dy - s * ey, th, ld, e_pos, v, close)) return false;
    pmin = v < pmin ? v : pmin;
    pmax = v > pmax ? v : pmax;
  }
  double v;
  if (!vp_pd2_fast(g, f, dx, dy, th, ld, e0, v, close)) return false;
  pmin = v < pmin ? v : pmin;
  pmax = v > pmax ? v : pmax;
  if (pmin == 0.0 && pmax > 0.0) close = true;
  pd = pmax;
  return true;
}

template <typename T>
__device__ __forceinline__ double vp_pd_any(const VPGeom& g, const VPFast& f, T px, T py, T pth, T lx, T ly, T ld, T pxx, T pxy,
                                            T pyy, bool& close) {
  if constexpr (sizeof(T) == 4) {
    double pd;
    if (vp_pd_fast(g, f, px, py, pth, lx, ly, ld, pxx, pxy, pyy, pd, close)) return pd;
  }
  return vp_pd(g, (double)px, (double)py, (double)pth, (double)lx, (double)ly, (double)ld, (double)pxx, (double)pxy,
               (double)pyy, close);
}

// probabilityOfDetection for the (up to 32) landmarks a warp holds, one per lane (have = this lane has one).  Same
// decisions and the same values as vp_pd_any per lane, organised for the warp: a landmark is evaluated at 2 K + 1 points
// across its 3-sigma extent (vp_pd_fast), K differs from landmark to landmark and most landmarks of a chunk are out of
// reach altogether — evaluated lane by lane, a handful of lanes would walk through their points while the others wait.
// Instead every lane first runs the cheap part for its own landmark (reach test, direction, K), the evaluation points
// of all landmarks of the chunk are numbered through (warp scan), lane i evaluates point i, i + 32, ... whoever it
// belongs to (parameters fetched from the owner by shuffle), and the owners collect the coded results of their points
// by shuffle again.  A point too close to call sends its landmark to the fp64 evaluation, as before.
template <typename T>
__device__ __forceinline__ double vp_pd_warp(const VPGeom& g, const VPFast& f, T px_, T py_, T pth_, bool have, T lx_, T ly_, T ld_,
                                             T pxx_, T pxy_, T pyy_, bool& close, int lane) {
  close = false;
  if constexpr (sizeof(T) != 4) {
    if (!have) return 0.0;
    return vp_pd(g, (double)px_, (double)py_, (double)pth_, (double)lx_, (double)ly_, (double)ld_, (double)pxx_, (double)pxy_,
                 (double)pyy_, close);
  } else {
    constexpr float PI_F = 3.14159265358979323846f;
    const float px = px_, py = py_, pth = pth_, lx = lx_, ly = ly_, ld = ld_, pxx = pxx_, pxy = pxy_, pyy = pyy_;
    const float th = pth - 0.5f * PI_F;
    // ---- own landmark: the part of vp_pd_fast before its evaluation loop ----
    enum { NONE = 0, DECIDED = 1, FP64 = 2, EVAL = 3 };
    int state = have ? EVAL : NONE;
    const float dx = lx - px, dy = ly - py;
    const float e0 = 4e-8f * (fabsf(dx) + fabsf(dy)) + 1e-7f;
    float ex = 0.0f, ey = 0.0f;
    int K = 0;
    if (have) {
      const float r0 = sqrtf(dx * dx + dy * dy);
      float b0 = M<float>::atan2_(dy, dx) - th;
      if (b0 > PI_F) b0 -= 2 * PI_F;
      if (b0 < -PI_F) b0 += 2 * PI_F;
      if (!(fabsf(b0) < PI_F - 2e-6f)) state = FP64;
      const float angle = M<float>::atan2_(b0, r0) + pth;
      float sn, cs;
      __sincosf(angle, &sn, &cs);
      ex = -sn; ey = cs;
      float sd = (ex * pxx + ey * pxy) * ex + (ex * pxy + ey * pyy) * ey;
      if (!(sd >= 0.0f) || !(ld > 0.0f)) state = FP64;
      sd = 3.0f * sqrtf(sd);
      sd = sd < 0.2f ? 0.2f : sd;
      if (!(sd < 1e30f)) state = FP64;
      if (state == EVAL) {
        const float reach = (sd + 2.0f * ld) * 1.001f + 1e-3f;
        if (r0 > f.rmax + reach || r0 < f.rmin - reach) state = DECIDED;   // P_D = 0, not close
      }
      if (state == EVAL) {
        const float step = 2.0f * ld;
        const float e_sd = 2e-5f * sd + 1e-6f;
        for (int i = 1; i <= 64; i++) {
          const float lim = (float)(i - 1) * step;
          if (fabsf(lim - sd) < e_sd) { state = FP64; break; }   // how many evaluation points there are is a decision too
          if (!(lim < sd)) break;
          K = i;
          if (i == 64) state = FP64;                             // a very thin trunk under a very wide covariance
        }
      }
    }
    // ---- the evaluation points of the chunk, numbered through ----
    const int cnt = state == EVAL ? 2 * K + 1 : 0;
    const int incl = warp_incl_scan(cnt, lane);
    const int excl = incl - cnt;
    const int total = __shfl_sync(FULL, incl, 31);
    double pmin = 1e300, pmax = -1e300;
    bool bad = false, close_c = false;
    for (int it0 = 0; it0 < total; it0 += 32) {
      const int item = it0 + lane;
      int owner = 0;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {   // smallest lane whose inclusive count exceeds the item
        const int v = __shfl_sync(FULL, incl, owner + sh - 1);
        if (item >= v) owner += sh;
      }
      owner &= 31;
      const int t = item - __shfl_sync(FULL, excl, owner);
      const float odx = __shfl_sync(FULL, dx, owner), ody = __shfl_sync(FULL, dy, owner);
      const float oex = __shfl_sync(FULL, ex, owner), oey = __shfl_sync(FULL, ey, owner);
      const float old = __shfl_sync(FULL, ld, owner), oe0 = __shfl_sync(FULL, e0, owner);
      const int oK = __shfl_sync(FULL, K, owner);
      int res = 1;
      if (item < total) {
        int code = VP_PD_ZERO;
        bool cl = false;
        // t == 2 K: the landmark itself (offset 0: the sums below return odx, ody unchanged), evaluated last by the
        // reference, so its "close" flag is the one that stays; the others: +-(i * 2 ld) along the perpendicular
        const bool centre = t == 2 * oK;
        const float s = centre ? 0.0f : (float)(t / 2 + 1) * (2.0f * old);
        const float e_pos = centre ? oe0 : oe0 + 4e-6f * s + 2e-6f;   // direction error x offset, roundings of the sums
        const float sg = (t & 1) ? -s : s;
        const bool ok = vp_pd2_fast_code(g, f, odx + sg * oex, ody + sg * oey, th, old, e_pos, code, cl);
        res = centre ? 256 : 0;
        res |= (ok ? 1 : 0) | (cl ? 2 : 0) | (code << 2);
      }
      // owners collect: the points of a landmark sit in consecutive lanes
      int first = excl > it0 ? excl - it0 : 0;
      int last = incl < it0 + 32 ? incl - it0 : 32;
      const int mine = last > first ? last - first : 0;
      const int most = 32 - __clz((int)__reduce_or_sync(FULL, mine ? 1u << (mine - 1) : 0u));   // max over the lanes
      for (int j = 0; j < most; j++) {
        const int r = __shfl_sync(FULL, res, (first + j) & 31);
        if (j < mine) {
          if (!(r & 1)) bad = true;
          const int code = (r >> 2) & 63;
          const double v = code == VP_PD_ZERO ? 0.0 : g.pd[code];
          pmin = v < pmin ? v : pmin;
          pmax = v > pmax ? v : pmax;
          if (r & 256) close_c = (r & 2) != 0;
        }
      }
    }
    if (state == EVAL && !bad) {
      close = close_c;
      if (pmin == 0.0 && pmax > 0.0) close = true;
      return pmax;
    }
    if (state == FP64 || (state == EVAL && bad))
      return vp_pd(g, (double)px, (double)py, (double)pth, (double)lx, (double)ly, (double)ld, (double)pxx, (double)pxy, (double)pyy,
                   close);
    return 0.0;   // out of reach, or no landmark in this lane
  }
}

// ---- 3x3 symmetric helpers (upper triangle a00 a01 a02 a11 a12 a22) ---------------------------------
template <typename T>
struct Sym3 {
  T a00, a01, a02, a11, a12, a22;
};
template <typename T>
__device__ __forceinline__ T sym3_inv(const Sym3<T>& a, Sym3<T>& o) {   // returns det
  const T c00 = a.a11 * a.a22 - a.a12 * a.a12;
  const T c01 = a.a02 * a.a12 - a.a01 * a.a22;
  const T c02 = a.a01 * a.a12 - a.a02 * a.a11;
  const T det = a.a00 * c00 + a.a01 * c01 + a.a02 * c02;
  const T id = T(1) / det;
  o.a00 = c00 * id;
  o.a01 = c01 * id;
  o.a02 = c02 * id;
  o.a11 = (a.a00 * a.a22 - a.a02 * a.a02) * id;
  o.a12 = (a.a01 * a.a02 - a.a00 * a.a12) * id;
  o.a22 = (a.a00 * a.a11 - a.a01 * a.a01) * id;
  return det;
}
template <typename T>
__device__ __forceinline__ T sym3_quad(const Sym3<T>& a, T x, T y, T z) {
  return (a.a00 * x + a.a01 * y + a.a02 * z) * x + (a.a01 * x + a.a11 * y + a.a12 * z) * y +
         (a.a02 * x + a.a12 * y + a.a22 * z) * z;
}
template <typename T>
__device__ __forceinline__ bool sym3_pd(const Sym3<T>& a) {
  return (a.a00 > T(0)) && (a.a00 * a.a11 - a.a01 * a.a01 > T(0)) &&
         (a.a00 * (a.a11 * a.a22 - a.a12 * a.a12) - a.a01 * (a.a01 * a.a22 - a.a12 * a.a02) +
              a.a02 * (a.a01 * a.a12 - a.a11 * a.a02) > T(0));
}
template <typename T>
__device__ __forceinline__ void load_sym3(Sym3<T>& s, const T* cur, int W, int i) {
  s.a00 = cur[3 * W + i]; s.a01 = cur[4 * W + i]; s.a02 = cur[5 * W + i];
  s.a11 = cur[6 * W + i]; s.a12 = cur[7 * W + i]; s.a22 = cur[8 * W + i];
}

// ---- shared memory layout ------------------------------------------------------------------------------
//  per CTA : lidar scan double[720] | P_D table double[16] | Z T[3 * MAX_Z]
//  per warp: planes T[NPL][W] (NPL = 10, +1 weight_prev in multi-feature mode) | aux u32[W] | colsum T[MAX_Z] |
//            evalIdx int[MAX_EVAL] | mbarrier | work region (vp_region_bytes).  The eval points' P_D, which the
//            intensity and the L-table stages both need, lives in colsum[] (dead after S3).
template <typename T>
__host__ __device__ inline int vp_cta_bytes() {
  return (int)((VP_SCAN_MAX * 8 + VP_PD_MAX * 8 + 3 * MAX_Z * sizeof(T) + (sizeof(T) == 4 ? VP_SCAN_MAX * 4 : 0) + 127) & ~127);
}
template <typename T>
__host__ __device__ inline int vp_mf_fixed_bytes(int n_eval, int zcap) {
  const int ltab = (n_eval * zcap + 3) & ~3;
  return (int)(8 * (MAX_EVAL + MAX_COMP + 2 * (1 << DP_MAXB)) + 4 * MAX_COMP + sizeof(T) * (n_eval * VP_EP + ltab) + 15) & ~15;
}
// work region of a warp, used by stages that follow each other:
//   sort of S5 / prune   keys u64[W] | order u16[W] | one temporary plane T[W]    (merge: reach T[W] aliases the keys,
//                                                                                  firstCand aliases order)
//   intensity of S5      7 planes T[W]
//   L-table stage of S5  rowmask / components / DP tables / eval block / L table
// (the fp32 kernels accept any work capacity W that is a multiple of 32; the bitonic sorts pad to a power of two, so
//  the key array holds vp_key_cap(W) entries)
__host__ __device__ inline int vp_key_cap(int W) {
  int p = 32;
  while (p < W) p <<= 1;
  return p;
}
template <typename T>
__host__ __device__ inline int vp_region_bytes(int W, int mf, int n_eval, int zcap) {
  int r = vp_key_cap(W) * 8 + ((W * 2 + 15) & ~15) + W * (int)sizeof(T);
  if (mf) {
    const int a = vp_mf_fixed_bytes<T>(n_eval, zcap), i7 = 7 * W * (int)sizeof(T);
    r = a > r ? a : r;
    r = i7 > r ? i7 : r;
  }
  return (r + 15) & ~15;
}
template <typename T>
__host__ __device__ inline int vp_warp_bytes(int W, int mf, int n_eval, int zcap) {
  const int planes = mf ? VP_NPL + 1 : VP_NPL;
  const int b = planes * W * (int)sizeof(T) + W * 4 + MAX_Z * (int)sizeof(T) + MAX_EVAL * 4 + 16 +
                vp_region_bytes<T>(W, mf, n_eval, zcap);
  return (b + 127) & ~127;
}

template <typename T>
struct VPRow {
  T x, y, d, w;
  Sym3<T> P, I;
  T reach2;
};
template <typename T>
__device__ __forceinline__ void vp_row_refresh(VPRow<T>& r, T tt) {
  sym3_inv(r.P, r.I);
  r.reach2 = sym3_pd(r.P) ? tt * (r.P.a00 + r.P.a11 + r.P.a22) : M<T>::inf();
}

template <typename T, bool MF>
__global__ void __launch_bounds__(512, 1)
phd_update_vp_kernel(const __grid_constant__ VPParams<T> vp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const KParams<T>& p = vp.k;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int W = p.W;
  const int nZ = p.nZ;
  constexpr int NPL = MF ? VP_NPL + 1 : VP_NPL;
  constexpr int WPREV = VP_NPL;   // weight_prev plane (MF only)

  double* scan_s = reinterpret_cast<double*>(smem_raw);
  double* pd_s = scan_s + VP_SCAN_MAX;
  T* zs = reinterpret_cast<T*>(pd_s + VP_PD_MAX);   // [3 * MAX_Z]: (zr, zb, zd) per measurement
  float* scan_f = reinterpret_cast<float*>(zs + 3 * MAX_Z);   // [720] fp32 kernels only: the scan for vp_pd_fast
  unsigned char* wb = smem_raw + vp_cta_bytes<T>() + (size_t)warp * p.warp_bytes;
  T* cur = reinterpret_cast<T*>(wb);
  unsigned* aux = reinterpret_cast<unsigned*>(cur + NPL * W);
  T* colsum = reinterpret_cast<T*>(aux + W);
  int* evalIdx = reinterpret_cast<int*>(colsum + MAX_Z);
  uint64_t* bar = reinterpret_cast<uint64_t*>(evalIdx + MAX_EVAL);
  unsigned char* mfs = reinterpret_cast<unsigned char*>(bar + 2);       // the work region (vp_region_bytes)
  unsigned long long* k64 = reinterpret_cast<unsigned long long*>(mfs);
  unsigned short* order = reinterpret_cast<unsigned short*>(k64 + vp_key_cap(W));
  T* sort_tmp = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(order) + ((W * 2 + 15) & ~15));
  T* rad2 = sort_tmp;                                                   // merge only (the sort of S5 is over)

  for (int k = threadIdx.x; k < VP_SCAN_MAX; k += blockDim.x) {
    const double v = k < vp.scan_n ? vp.scan[k] : 0.0;
    scan_s[k] = v;
    if constexpr (sizeof(T) == 4) scan_f[k] = (float)v;
  }
  for (int k = threadIdx.x; k < VP_PD_MAX; k += blockDim.x) pd_s[k] = k < vp.pd_n ? vp.pd_table[k] : 0.0;
  for (int k = threadIdx.x; k < 3 * nZ; k += blockDim.x) {
    const T v = (T)p.Zval[k];
    zs[k] = v;
    if (blockIdx.x == 0) p.Zdev_w[k] = v;
  }
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  // deferred cross-GPU sums of the previous step (KParams::comm_pending, as in phd_update_kernel): thread r picks up rank
  // r's pair from this GPU's own mailbox
  __shared__ double prev_sum[8];
  if ((int)threadIdx.x < p.comm_pending) {
    const CommSlot* mine = reinterpret_cast<const CommSlot*>(p.comm_peer[p.comm_rank]) + (int)(p.comm_prev_epoch & 1ull) * 8;
    const unsigned long long t0 = globaltimer_ns();
    double a1 = 0, a2 = 0;
    while (!comm_recv(mine + threadIdx.x, p.comm_prev_epoch, a1, a2)) {
      if (globaltimer_ns() - t0 > p.comm_timeout_ns) {   // a peer never finished the previous step
        *p.comm_error = 1;
        a1 = __longlong_as_double(0x7ff8000000000000LL);
        break;
      }
    }
    prev_sum[threadIdx.x] = a1;
  }
  __syncthreads();
  VPGeom geom;
  geom.rmin = vp.rmin; geom.rmax = vp.rmax; geom.bmin = vp.bmin; geom.bmax = vp.bmax; geom.buf_pd = vp.buf_pd;
  geom.pd = pd_s; geom.scan = scan_s; geom.pd_n = vp.pd_n; geom.scan_n = vp.scan_n > VP_SCAN_MAX ? VP_SCAN_MAX : vp.scan_n;
  VPFast fast;
  fast.rmin = (float)vp.rmin; fast.rmax = (float)vp.rmax; fast.bmin = (float)vp.bmin; fast.bmax = (float)vp.bmax;
  fast.scan = scan_f;

  uint32_t phase = 0;
  unsigned long long tot_in = 0, tot_out = 0;
  int max_out = 0, n_over = 0, n_murty = 0;
  unsigned mstat[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T norm3 = T(1) / M<T>::sqrt_(M<T>::TWO_PI * M<T>::TWO_PI * M<T>::TWO_PI);

  while (true) {
    int pi = 0;
    if (lane == 0) pi = (int)atomicAdd(p.work_counter, 1u);
    pi = __shfl_sync(FULL, pi, 0);
    if (pi >= p.N) break;

    double w_prev_particle = p.w_in[pi];
    if (p.comm_pending_scale) {   // ParticleFilter::normalizeWeights of the previous step, applied on the way in
      double total = 0;
      for (int r = 0; r < p.comm_pending; r++) total += prev_sum[r];
      w_prev_particle = w_prev_particle / total;
    }
    int nM = p.cnt_in[pi];
    nM = nM < 0 ? 0 : (nM > p.cap ? p.cap : nM);
    int flags = (p.flags[pi] & (FLAG_BIRTH_OVERFLOW | FLAG_CAND_OVERFLOW)) ? FLAG_OVERFLOW : 0;
    if (nM > W) { nM = W; flags |= FLAG_OVERFLOW; }

    // ---------------- S0: TMA bulk loads ---------------------------------------------------------
    if (nM > 0 && lane == 0) {
      const uint32_t bytes = (uint32_t)(((nM + 3) & ~3) * sizeof(T));
      fence_proxy_async();
      mbar_expect_tx(bar, VP_NPL * bytes);
      const T* src = p.gm_in + (size_t)pi * VP_NPL * p.cap;
#pragma unroll
      for (int k = 0; k < VP_NPL; k++) tma_load_1d(cur + k * W, src + (size_t)k * p.cap, bytes, bar);
    }
    const T px = p.pose[4 * pi], py = p.pose[4 * pi + 1], pth = p.pose[4 * pi + 2];
    const T pths = pth - T(M<T>::PI / 2);   // sensor frame (pose theta - pi/2, :110-112)
    if (nM > 0) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }

    // ---------------- S1: detection probability + corrector -------------------------------------------
    int nS = 0;
    double wsum_d = 0;
    int nfov = 0;
    bool over = false;
    for (int base = 0; base < nM; base += 32) {
      const int m = base + lane;
      unsigned long long mask = 0;
      T x = 0, y = 0, d = 0, Pdw = 0;
      Sym3<T> P{};
      T zr_hat = 0, zb_hat = 0, i00 = 0, i01 = 0, i11 = 0, i22 = 0, norm = 0;
      T ph00 = 0, ph01 = 0, ph10 = 0, ph11 = 0, ph20 = 0, ph21 = 0;   // (P H^T)[r][0..1]; column 2 is P[r][2]
      if (m < nM) {
        x = cur[m]; y = cur[W + m]; d = cur[2 * W + m];
        load_sym3(P, cur, W, m);
      }
      bool close = false;
      T Pd = (T)vp_pd_warp<T>(geom, fast, px, py, pth, m < nM, x, y, d, P.a00, P.a01, P.a11, close, lane);
      if (m < nM) {
        const T w = cur[VP_WP * W + m];
        wsum_d += (double)w;
        if (close) Pd = T(1);   // Q2 (include/RBPHDFilter.hpp:604-606)
        if (Pd != T(0)) nfov++;
        if (MF) cur[WPREV * W + m] = w;
        const bool fix = close && (w > p.birth_w);
        cur[VP_WP * W + m] = fix ? w : (T(1) - Pd) * w;
        aux[m] = fix ? 1u : 0u;
        if (Pd != T(0)) {
          // measure(): always valid for this model (the 2-D model's range test is ignored, :127,:139)
          const T dx = x - px, dy = y - py;
          const T r2 = dx * dx + dy * dy;
          const T r = M<T>::sqrt_(r2);
          const T invr = T(1) / r;
          const T c = dx * invr, s = dy * invr;
          const T h10 = -s * invr, h11 = c * invr;
          zr_hat = r;
          zb_hat = wrap_pi<T>(M<T>::atan2_(dy, dx) - pths);
          ph00 = P.a00 * c + P.a01 * s;   ph01 = P.a00 * h10 + P.a01 * h11;
          ph10 = P.a01 * c + P.a11 * s;   ph11 = P.a01 * h10 + P.a11 * h11;
          ph20 = P.a02 * c + P.a12 * s;   ph21 = P.a02 * h10 + P.a12 * h11;
          const T s00 = c * ph00 + s * ph10 + vp.R00;
          const T s01 = c * ph01 + s * ph11 + vp.R01;
          const T s10 = h10 * ph00 + h11 * ph10 + vp.R10;
          const T s11 = h10 * ph01 + h11 * ph11 + vp.R11;
          const T s22 = P.a22 + vp.R22 + r2 * vp.Slb;
          const T det2 = s00 * s11 - s01 * s10;
          const T invdet = T(1) / det2;
          i00 = s11 * invdet; i01 = -s01 * invdet; i11 = s00 * invdet;   // R01 == R10 in every use of the model
          i22 = T(1) / s22;
          norm = norm3 / M<T>::sqrt_(det2 * s22);
          Pdw = Pd * w;
          for (int z = 0; z < nZ; z++) {
            const T nr = zs[3 * z] - zr_hat;
            const T nb_raw = zs[3 * z + 1] - zb_hat;
            const T nb = wrap_pi<T>(nb_raw);
            if (p.thr_r > T(0) && M<T>::abs_(nr) > p.thr_r) continue;
            if (p.thr_b > T(0) && M<T>::abs_(nb) > p.thr_b) continue;
            const T nd = zs[3 * z + 2] - d;
            // Q3: likelihood and gate use the UNWRAPPED difference
            const T md2 = (nr * i00 + nb_raw * i01) * nr + (nr * i01 + nb_raw * i11) * nb_raw + nd * nd * i22;
            if (md2 > p.gate2) continue;
            const T lik = M<T>::exp_(T(-0.5) * md2) * norm;
            if (!(lik == lik) || lik == T(0)) continue;
            if (!(Pdw * lik > T(0))) continue;
            mask |= (1ull << z);
          }
        }
      }
      const int cnt = __popcll(mask);
      const int incl = warp_incl_scan(cnt, lane);
      int off = nM + nS + incl - cnt;
      nS += __shfl_sync(FULL, incl, 31);
      if (cnt) {
        // K = P H^T S^-1 ; P+ = sym((I - K H) P) = sym(P - K (P H^T)^T)   (include/KalmanFilter.hpp:297-302)
        const T k00 = ph00 * i00 + ph01 * i01, k01 = ph00 * i01 + ph01 * i11, k02 = P.a02 * i22;
        const T k10 = ph10 * i00 + ph11 * i01, k11 = ph10 * i01 + ph11 * i11, k12 = P.a12 * i22;
        const T k20 = ph20 * i00 + ph21 * i01, k21 = ph20 * i01 + ph21 * i11, k22 = P.a22 * i22;
        // M[r][c] = sum_k K[r][k] * PH[c][k]
        const T m00 = k00 * ph00 + k01 * ph01 + k02 * P.a02;
        const T m01 = k00 * ph10 + k01 * ph11 + k02 * P.a12;
        const T m02 = k00 * ph20 + k01 * ph21 + k02 * P.a22;
        const T m10 = k10 * ph00 + k11 * ph01 + k12 * P.a02;
        const T m11 = k10 * ph10 + k11 * ph11 + k12 * P.a12;
        const T m12 = k10 * ph20 + k11 * ph21 + k12 * P.a22;
        const T m20 = k20 * ph00 + k21 * ph01 + k22 * P.a02;
        const T m21 = k20 * ph10 + k21 * ph11 + k22 * P.a12;
        const T m22 = k20 * ph20 + k21 * ph21 + k22 * P.a22;
        const T n00 = P.a00 - m00, n11 = P.a11 - m11, n22 = P.a22 - m22;
        const T n01 = ((P.a01 - m01) + (P.a01 - m10)) * T(0.5);
        const T n02 = ((P.a02 - m02) + (P.a02 - m20)) * T(0.5);
        const T n12 = ((P.a12 - m12) + (P.a12 - m21)) * T(0.5);
        while (mask) {
          const int z = __ffsll((long long)mask) - 1;
          mask &= mask - 1;
          if (off >= W) { over = true; break; }
          const T nr = zs[3 * z] - zr_hat;
          const T nb_raw = zs[3 * z + 1] - zb_hat;
          const T nb = wrap_pi<T>(nb_raw);
          const T nd = zs[3 * z + 2] - d;
          const T md2 = (nr * i00 + nb_raw * i01) * nr + (nr * i01 + nb_raw * i11) * nb_raw + nd * nd * i22;
          const T lik = M<T>::exp_(T(-0.5) * md2) * norm;
          cur[off] = x + (k00 * nr + k01 * nb + k02 * nd);
          cur[W + off] = y + (k10 * nr + k11 * nb + k12 * nd);
          cur[2 * W + off] = d + (k20 * nr + k21 * nb + k22 * nd);
          cur[3 * W + off] = n00; cur[4 * W + off] = n01; cur[5 * W + off] = n02;
          cur[6 * W + off] = n11; cur[7 * W + off] = n12; cur[8 * W + off] = n22;
          cur[VP_WP * W + off] = Pdw * lik;
          if (MF) cur[WPREV * W + off] = T(0);
          aux[off] = ((unsigned)m << 8) | (unsigned)z;
          off++;
        }
      }
      __syncwarp();
    }
    if (__any_sync(FULL, over)) flags |= FLAG_OVERFLOW;
    if (nM + nS > W) nS = W - nM;
    const int n = nM + nS;
    nfov = warp_sum(nfov);
    wsum_d = warp_sum(wsum_d);
    __syncwarp();

    double weight_new = w_prev_particle;
    unsigned long long unused_mask = 0;
    T* wpl = cur + VP_WP * W;
    if (nM == 0) {
      unused_mask = (nZ >= 64) ? ~0ull : ((1ull << nZ) - 1ull);   // :559-564 (Q10)
    } else {
      // ---------------- S2: per-measurement normalisers (:644-659) ---------------------------------
      double ll = 0;
      for (int zb0 = 0; zb0 < nZ; zb0 += 32) {
        const int z = zb0 + lane;
        bool used = false;
        if (z < nZ) {
          T sum = p.kappa;
          for (int s = nM; s < n; s++) {
            if ((int)(aux[s] & 0xffu) == z) { sum += wpl[s]; used = true; }
          }
          colsum[z] = sum;
          ll += log((double)sum);
        }
        const unsigned b = __ballot_sync(FULL, (z < nZ) && !used);
        unused_mask |= ((unsigned long long)b) << zb0;
      }
      ll = warp_sum(ll);
      __syncwarp();
      if (p.use_sc) weight_new = exp(wsum_d + ll) * w_prev_particle;   // :661-668 (Q4)
      // ---------------- S3: posterior weights of the new Gaussians ----------------------------------
      for (int s = nM + lane; s < n; s += 32) wpl[s] = wpl[s] / colsum[aux[s] & 0xffu];
      __syncwarp();
      // ---------------- S4: sensing-limit heuristic (:692-703, Q2) ----------------------------------
      for (int mb = 0; mb < nM; mb += 32) {
        const int m = mb + lane;
        const unsigned am = (m < nM) ? aux[m] : 0u;
        if (__any_sync(FULL, am != 0u)) {
          if (am) {
            const T w_km = wpl[m];   // still the pre-update weight
            T rowsum = T(0);
            for (int s = nM; s < n; s++)
              if ((int)(aux[s] >> 8) == m) rowsum += wpl[s];
            const T delta = w_km - rowsum;   // Pd[m] == 1 here
            T w_k = T(0);
            if (delta > T(0)) {
              w_k += delta;
              if (w_k > T(1)) w_k = T(1);
            }
            wpl[m] = w_k;
          }
        }
      }
      __syncwarp();
    }

    // ---------------- S5: multi-feature importance weighting (:728-819) ------------------------------
    if constexpr (MF) {
      const int nEvalCfg = p.n_eval < n ? p.n_eval : n;
      if (nEvalCfg == 0) {
        weight_new = 4.9406564584124654e-324;   // denorm_min (:742-745, Q10)
      } else {
        T* ia = reinterpret_cast<T*>(mfs);   // [7][W]; dead before the L-table stage takes the region over
        {  // sortByWeight (:746): weight descending, ties by position
          const int P2 = next_pow2(n);
          if constexpr (sizeof(T) == 4) {
            for (int k = lane; k < P2; k += 32)
              k64[k] = (k < n) ? (((unsigned long long)__float_as_uint((float)wpl[k]) << 32) |
                                  (unsigned long long)(0xffffffffu - (unsigned)k))
                               : 0ull;
            __syncwarp();
            warp_sort_desc64(k64, P2, lane);
            for (int k = lane; k < n; k += 32) order[k] = (unsigned short)(0xffffffffu - (unsigned)(k64[k] & 0xffffffffull));
          } else {
            T* kw = reinterpret_cast<T*>(k64);
            for (int k = lane; k < P2; k += 32) {
              kw[k] = (k < n) ? wpl[k] : -M<T>::inf();
              aux[k] = (unsigned)k;
            }
            __syncwarp();
            warp_bitonic(kw, aux, P2, lane);
            for (int k = lane; k < n; k += 32) order[k] = (unsigned short)aux[k];
          }
          __syncwarp();
          if (sizeof(T) == 4 && n <= 256) {
            // a lane's (at most 8) source positions and its values of one plane in registers: gather, barrier, store —
            // no temporary plane, the permutation is read once for all planes
            int pk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) pk[j] = (lane + 32 * j < n) ? (int)order[lane + 32 * j] : 0;
            for (int pl = 0; pl < NPL; pl++) {
              T v[8];
#pragma unroll
              for (int j = 0; j < 8; j++) v[j] = cur[pl * W + pk[j]];
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; j++)
                if (lane + 32 * j < n) cur[pl * W + lane + 32 * j] = v[j];
              __syncwarp();
            }
          } else {
            T* tmp = sort_tmp;
            for (int pl = 0; pl < NPL; pl++) {
              for (int k = lane; k < n; k += 32) tmp[k] = cur[pl * W + order[k]];
              __syncwarp();
              for (int k = lane; k < n; k += 32) cur[pl * W + k] = tmp[k];
              __syncwarp();
            }
          }
        }
        // eval points (:747-762): sorted order, w >= min weight, model P_D > 0, the first nEvalCfg
        unsigned long long* rowmask = reinterpret_cast<unsigned long long*>(mfs);   // [MAX_EVAL]
        unsigned long long* compC = rowmask + MAX_EVAL;                             // [MAX_COMP]
        double* f0 = reinterpret_cast<double*>(compC + MAX_COMP);                   // [1<<DP_MAXB]
        double* f1 = f0 + (1 << DP_MAXB);
        unsigned* compR = reinterpret_cast<unsigned*>(f1 + (1 << DP_MAXB));         // [MAX_COMP]
        T* ep = reinterpret_cast<T*>(compR + MAX_COMP);   // [n_eval_cap][VP_EP - 1] eval-point block
        T* evalPd = colsum;                               // [MAX_EVAL] (colsum is dead after S3; MAX_EVAL <= MAX_Z)
        T* L = ep + p.n_eval_cap * VP_EP;                 // [nE][nZ]
        int nE = 0;
        for (int base = 0; base < n && nE < nEvalCfg; base += 32) {
          const int m = base + lane;
          bool elig = false, heavy = false;
          T pdm = 0;
          if (m < n) heavy = !(wpl[m] < p.eval_min_w);
          {
            const int mm = heavy ? m : 0;
            bool close;
            pdm = (T)vp_pd_warp<T>(geom, fast, px, py, pth, heavy, cur[mm], cur[W + mm], cur[2 * W + mm], cur[3 * W + mm],
                                   cur[4 * W + mm], cur[6 * W + mm], close, lane);
            elig = heavy && pdm > T(0);
          }
          const unsigned be = __ballot_sync(FULL, elig);
          const int rank = nE + __popc(be & ((1u << lane) - 1u));
          if (elig && rank < nEvalCfg) { evalIdx[rank] = m; evalPd[rank] = pdm; }   // raw model value (Q12)
          nE += __popc(be);
          if (!__all_sync(FULL, heavy)) break;
        }
        if (nE > nEvalCfg) nE = nEvalCfg;
        __syncwarp();
        double sw_prev = 0, sw_now = 0;   // :765-773
        for (int m = lane; m < n; m += 32) { sw_prev += (double)cur[WPREV * W + m]; sw_now += (double)wpl[m]; }
        sw_prev = warp_sum(sw_prev);
        sw_now = warp_sum(sw_now);
        // intensity at the eval points before / after the update (:776-800), log domain
        for (int m = lane; m < n; m += 32) {
          Sym3<T> Pm, Im;
          load_sym3(Pm, cur, W, m);
          const T det = sym3_inv(Pm, Im);
          ia[m] = Im.a00; ia[W + m] = Im.a01; ia[2 * W + m] = Im.a02;
          ia[3 * W + m] = Im.a11; ia[4 * W + m] = Im.a12; ia[5 * W + m] = Im.a22;
          ia[6 * W + m] = M<T>::log_(M<T>::sqrt_(M<T>::TWO_PI * M<T>::TWO_PI * M<T>::TWO_PI * det));
        }
        __syncwarp();
        double lp_before = 0, lp_after = 0;
        {
          const T CUT = sizeof(T) == 4 ? T(25) : T(45);
          const T LOG_DENORM_MIN = T(-744.4400719213812);
          const int groups = (nE <= 16) ? 2 : 1;
          const int e = groups == 2 ? (lane & 15) : lane;
          const int h = groups == 2 ? (lane >> 4) : 0;
          T mb = -M<T>::inf(), sb = 0, ma = -M<T>::inf(), sa = 0;
          if (e < nE) {
            const int ei = evalIdx[e];
            const T xe = cur[ei], ye = cur[W + ei], de = cur[2 * W + ei];
            const int half = (n + groups - 1) / groups;
            const int m0 = h * half, m1 = (m0 + half < n) ? m0 + half : n;
            for (int m = m0; m < m1; m++) {
              const T dx = xe - cur[m], dy = ye - cur[W + m], dd = de - cur[2 * W + m];
              const T j01 = ia[W + m], j02 = ia[2 * W + m], j12 = ia[4 * W + m];
              const T md2 = (ia[m] * dx + j01 * dy + j02 * dd) * dx + (j01 * dx + ia[3 * W + m] * dy + j12 * dd) * dy +
                            (j02 * dx + j12 * dy + ia[5 * W + m] * dd) * dd;
              const T t = T(-0.5) * md2 - ia[6 * W + m];
              const T wp = cur[WPREV * W + m], wn = wpl[m];
              if (wp > T(0)) {
                if (t > mb) { sb = sb * M<T>::exp_(mb - t) + wp; mb = t; }
                else if (t > mb - CUT) sb += wp * M<T>::exp_(t - mb);
              }
              if (wn > T(0)) {
                if (t > ma) { sa = sa * M<T>::exp_(ma - t) + wn; ma = t; }
                else if (t > ma - CUT) sa += wn * M<T>::exp_(t - ma);
              }
            }
          }
          if (groups == 2) {
            T mo = __shfl_xor_sync(FULL, mb, 16), so = __shfl_xor_sync(FULL, sb, 16);
            T mx = mo > mb ? mo : mb;
            if (mx > -M<T>::inf()) sb = sb * M<T>::exp_(mb - mx) + so * M<T>::exp_(mo - mx);
            mb = mx;
            mo = __shfl_xor_sync(FULL, ma, 16); so = __shfl_xor_sync(FULL, sa, 16);
            mx = mo > ma ? mo : ma;
            if (mx > -M<T>::inf()) sa = sa * M<T>::exp_(ma - mx) + so * M<T>::exp_(mo - mx);
            ma = mx;
          }
          T lvb = (mb > -M<T>::inf() && sb > T(0)) ? mb + M<T>::log_(sb) : LOG_DENORM_MIN;
          T lva = (ma > -M<T>::inf() && sa > T(0)) ? ma + M<T>::log_(sa) : LOG_DENORM_MIN;
          if (lvb < LOG_DENORM_MIN) lvb = LOG_DENORM_MIN;
          if (lva < LOG_DENORM_MIN) lva = LOG_DENORM_MIN;
          const bool mine = (e < nE) && (h == 0);
          lp_before = warp_sum(mine ? (double)lvb : 0.0);
          lp_after = warp_sum(mine ? (double)lva : 0.0);
        }
        __syncwarp();
        // rfsMeasurementLikelihood (:821-997): L table with the landmark covariance zeroed (:850-852), so
        // S_e = blockdiag(R[0:2,0:2], R22 + r^2 Slb)
        if (lane < nE) {
          const int ei = evalIdx[lane];
          const T dx = cur[ei] - px, dy = cur[W + ei] - py;
          const T r2 = dx * dx + dy * dy;
          const T s22 = vp.R22 + r2 * vp.Slb;
          const T det2 = vp.R00 * vp.R11 - vp.R01 * vp.R10;
          const T invdet = T(1) / det2;
          T* q = ep + lane * (VP_EP - 1);
          q[0] = M<T>::sqrt_(r2);
          q[1] = wrap_pi<T>(M<T>::atan2_(dy, dx) - pths);
          q[2] = cur[2 * W + ei];
          q[3] = vp.R11 * invdet; q[4] = -vp.R01 * invdet; q[5] = vp.R00 * invdet;
          q[6] = T(1) / s22;
          q[7] = evalPd[lane] * norm3 / M<T>::sqrt_(det2 * s22);
        }
        __syncwarp();
        for (int k = lane; k < nE * nZ; k += 32) {
          const int e = k / nZ, z = k - e * nZ;
          const T* q = ep + e * (VP_EP - 1);
          const T nr = zs[3 * z] - q[0], nb = zs[3 * z + 1] - q[1], nd = zs[3 * z + 2] - q[2];
          const T md2 = (nr * q[3] + nb * q[4]) * nr + (nr * q[4] + nb * q[5]) * nb + nd * nd * q[6];
          T l = M<T>::exp_(T(-0.5) * md2) * q[7];
          if (!(l == l)) l = T(0);
          if (md2 > p.wl_gate2) l = T(0);
          L[k] = l;
        }
        __syncwarp();
        double* gdp = p.dp_scratch ? p.dp_scratch + ((size_t)(blockIdx.x * (blockDim.x >> 5) + warp) << (p.dp_gmaxb + 1)) : nullptr;
        MurtyOut mo;
        mo.buf = p.murty_buf; mo.count = p.murty_count; mo.cap_words = p.murty_cap_words; mo.pi = pi;
        double logL = mf_partition_loglik<T>(L, evalPd, nE, nZ, rowmask, compC, f0, f1, compR, p.sum_method,
                                            p.log_kappa, flags, lane, gdp, p.dp_gmaxb, p.dp_onchip, mo);
        logL -= p.log_clutter_integral;
        weight_new = exp(logL + (lp_before - lp_after) + (sw_now - sw_prev)) * w_prev_particle;   // :808-812
        __syncwarp();
      }
    }

    // ---------------- S6: merge (include/GaussianMixture.hpp:394-475), the reference's order ------------
    if (n > 1) {
      const T t2 = p.merge_t2, f = p.merge_f;
      const T tt = t2 * T(1.001);
      for (int j = lane; j < n; j += 32) {
        Sym3<T> Pj;
        load_sym3(Pj, cur, W, j);
        rad2[j] = sym3_pd(Pj) ? tt * (Pj.a00 + Pj.a11 + Pj.a22) : M<T>::inf();
      }
      __syncwarp();
      // Pre-pass, one lane per row: the first later component within reach of the row (distance test only).
      // When row i gets its turn it is still in its original state and so is every live j > i (rows are only
      // modified while they are the absorbing row), so a row without such a partner cannot merge with
      // anything and is skipped; the others start their exact scan at that partner.
      unsigned short* firstCand = order;   // [W]: free between the sort of S5 and the prune
      T rmax2 = T(0);
      for (int j = lane; j < n; j += 32) rmax2 = M<T>::max_(rmax2, rad2[j]);
      rmax2 = warp_max(rmax2);
      if (rmax2 < M<T>::inf()) {
        // every partner of a row lies within sqrt(rmax2) of it, in x in particular: sort the components by x (keys
        // in the sort scratch, free during the merge) and look only at the neighbours inside that window
        const int P2 = next_pow2(n);
        for (int k = lane; k < P2; k += 32) {
          unsigned long long key = 0ull;   // pads sort to the end (descending)
          if (k < n) {
            unsigned u = __float_as_uint((float)cur[k]);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);        // order-preserving map of the float to unsigned
            key = ((unsigned long long)u << 32) | (unsigned long long)(k + 1);
          }
          k64[k] = key;
        }
        __syncwarp();
        warp_sort_desc64(k64, P2, lane);
        // the keys order by x rounded to fp32 (and the window bound is evaluated in T): widen the window a little
        const T rwin = M<T>::sqrt_(rmax2) * T(1.001) + T(1e-3);
        for (int sb = 0; sb < n; sb += 32) {
          const int sp = sb + lane;
          if (sp < n) {
            const int i = (int)(k64[sp] & 0xffffffffull) - 1;
            const T xi = cur[i], yi = cur[W + i], di = cur[2 * W + i], ri = rad2[i];
            int best = 0xffff;
            for (int dir = -1; dir <= 1; dir += 2) {
              for (int t = sp + dir; t >= 0 && t < n; t += dir) {
                const int j = (int)(k64[t] & 0xffffffffull) - 1;
                const T ex = cur[j] - xi;
                if (M<T>::abs_(ex) > rwin) break;
                const T ey = cur[W + j] - yi, ed = cur[2 * W + j] - di;
                if (j > i && j < best && !(ex * ex + ey * ey + ed * ed > M<T>::max_(ri, rad2[j]))) best = j;
              }
            }
            firstCand[i] = (unsigned short)best;
          }
        }
      } else {
        for (int ib = 0; ib < n; ib += 32) {   // some covariance is not PD (infinite reach): every pair
          const int i = ib + lane;
          unsigned short jm = 0xffffu;
          if (i < n - 1) {
            const T xi = cur[i], yi = cur[W + i], di = cur[2 * W + i], ri = rad2[i];
            for (int j = i + 1; j < n; j++) {
              const T ex = cur[j] - xi, ey = cur[W + j] - yi, ed = cur[2 * W + j] - di;
              if (!(ex * ex + ey * ey + ed * ed > M<T>::max_(ri, rad2[j]))) { jm = (unsigned short)j; break; }
            }
          }
          if (i < n) firstCand[i] = jm;
        }
      }
      __syncwarp();
      for (int i = 0; i < n - 1; i++) {
        const int j0 = firstCand[i];
        if (j0 == 0xffff) continue;    // nothing within reach
        if (wpl[i] < T(0)) continue;   // hole
        VPRow<T> r;
        r.x = cur[i]; r.y = cur[W + i]; r.d = cur[2 * W + i]; r.w = wpl[i];
        load_sym3(r.P, cur, W, i);
        bool have_inv = false, changed = false;
        r.reach2 = rad2[i];
        int jstart = j0;
        while (jstart < n) {
          int found = -1;
          for (int base = jstart; base < n; base += 32) {
            const int j = base + lane;
            bool cand = false;
            T ex = 0, ey = 0, ed = 0;
            if (j < n && wpl[j] >= T(0)) {
              ex = cur[j] - r.x; ey = cur[W + j] - r.y; ed = cur[2 * W + j] - r.d;
              const T e2 = ex * ex + ey * ey + ed * ed;
              cand = !(e2 > M<T>::max_(r.reach2, rad2[j]));
            }
            bool pass = false;
            if (__any_sync(FULL, cand)) {
              if (!have_inv) { sym3_inv(r.P, r.I); have_inv = true; }   // warp-uniform
              if (cand) {
                pass = !(sym3_quad(r.I, ex, ey, ed) > t2);
                if (!pass) {
                  Sym3<T> Pj, Ij;
                  load_sym3(Pj, cur, W, j);
                  sym3_inv(Pj, Ij);
                  pass = !(sym3_quad(Ij, ex, ey, ed) > t2);
                }
              }
            }
            const unsigned bp = __ballot_sync(FULL, pass);
            if (bp) { found = base + __ffs(bp) - 1; break; }
          }
          if (found < 0) break;
          jstart = found + 1;
          // absorb `found` into the row (every lane computes the same values)
          const T w1 = r.w, w2 = wpl[found];
          const T wm = w1 + w2;
          if (wm == T(0)) continue;   // :447-449 the reference moves on to the next j
          const T x2 = cur[found], y2 = cur[W + found], d2 = cur[2 * W + found];
          Sym3<T> Q;
          load_sym3(Q, cur, W, found);
          const T iw = T(1) / wm;
          const T xm = (r.x * w1 + x2 * w2) * iw, ym = (r.y * w1 + y2 * w2) * iw, dm = (r.d * w1 + d2 * w2) * iw;
          const T ax = xm - r.x, ay = ym - r.y, ad = dm - r.d, bx = xm - x2, by = ym - y2, bd = dm - d2;
          Sym3<T> S;
          S.a00 = (w1 * (r.P.a00 + f * ax * ax) + w2 * (Q.a00 + f * bx * bx)) * iw;
          S.a01 = (w1 * (r.P.a01 + f * ax * ay) + w2 * (Q.a01 + f * bx * by)) * iw;
          S.a02 = (w1 * (r.P.a02 + f * ax * ad) + w2 * (Q.a02 + f * bx * bd)) * iw;
          S.a11 = (w1 * (r.P.a11 + f * ay * ay) + w2 * (Q.a11 + f * by * by)) * iw;
          S.a12 = (w1 * (r.P.a12 + f * ay * ad) + w2 * (Q.a12 + f * by * bd)) * iw;
          S.a22 = (w1 * (r.P.a22 + f * ad * ad) + w2 * (Q.a22 + f * bd * bd)) * iw;
          r.x = xm; r.y = ym; r.d = dm; r.w = wm; r.P = S;
          vp_row_refresh(r, tt);
          have_inv = true;
          changed = true;
          __syncwarp();
          if (lane == 0) wpl[found] = T(-1);   // hole
          __syncwarp();
        }
        if (changed) {
          __syncwarp();
          if (lane == 0) {
            cur[i] = r.x; cur[W + i] = r.y; cur[2 * W + i] = r.d;
            cur[3 * W + i] = r.P.a00; cur[4 * W + i] = r.P.a01; cur[5 * W + i] = r.P.a02;
            cur[6 * W + i] = r.P.a11; cur[7 * W + i] = r.P.a12; cur[8 * W + i] = r.P.a22;
            wpl[i] = r.w;
            rad2[i] = r.reach2;
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();

    // ---------------- S7: prune (include/GaussianMixture.hpp:477-521) + store ---------------------------
    int n_out = 0;
    {
      if constexpr (sizeof(T) == 4) {
        for (int base = 0; base < n; base += 32) {
          const int k = base + lane;
          bool keep = false;
          T w = 0;
          if (k < n) { w = wpl[k]; keep = (w >= p.prune_t) && (w >= T(0)); }
          __syncwarp();   // rad2 aliases the keys: every lane has read its weights of this chunk before keys are written
          const unsigned b = __ballot_sync(FULL, keep);
          if (keep) {
            const int pos = n_out + __popc(b & ((1u << lane) - 1u));
            k64[pos] = ((unsigned long long)__float_as_uint((float)w) << 32) | (unsigned long long)(0xffffffffu - (unsigned)k);
          }
          n_out += __popc(b);
        }
        const int P2 = next_pow2(n_out);
        for (int k = n_out + lane; k < P2; k += 32) k64[k] = 0ull;
        __syncwarp();
        if (n_out > 1) warp_sort_desc64(k64, P2, lane);
        for (int k = lane; k < n_out; k += 32) order[k] = (unsigned short)(0xffffffffu - (unsigned)(k64[k] & 0xffffffffull));
      } else {
        T* kw = reinterpret_cast<T*>(k64);
        for (int base = 0; base < n; base += 32) {
          const int k = base + lane;
          bool keep = false;
          T w = 0;
          if (k < n) { w = wpl[k]; keep = (w >= p.prune_t) && (w >= T(0)); }
          const unsigned b = __ballot_sync(FULL, keep);
          if (keep) {
            const int pos = n_out + __popc(b & ((1u << lane) - 1u));
            kw[pos] = w;
            aux[pos] = (unsigned)k;
          }
          n_out += __popc(b);
        }
        const int P2 = next_pow2(n_out);
        for (int k = n_out + lane; k < P2; k += 32) { kw[k] = -M<T>::inf(); aux[k] = 0xffffffffu; }
        __syncwarp();
        if (n_out > 1) warp_bitonic(kw, aux, P2, lane);
        for (int k = lane; k < n_out; k += 32) order[k] = (unsigned short)aux[k];
      }
      __syncwarp();
      if (n_out > p.cap) { n_out = p.cap; flags |= FLAG_OVERFLOW; }
      T* dst = p.gm_out + (size_t)pi * VP_NPL * p.cap;
      for (int k = lane; k < n_out; k += 32) {
        const unsigned src = order[k];
#pragma unroll
        for (int pl = 0; pl < VP_NPL; pl++) dst[(size_t)pl * p.cap + k] = cur[pl * W + src];
      }
    }
    if (lane == 0) {
      p.cnt_out[pi] = n_out;
      p.w_out[pi] = weight_new;
      p.unused[pi] = unused_mask;
      p.nfov[pi] = nfov;
      p.flags[pi] = flags;
      store_host_results(p, pi, weight_new, unused_mask, nfov);
    }
    tot_in += (unsigned long long)nM;
    tot_out += (unsigned long long)n_out;
    max_out = n_out > max_out ? n_out : max_out;
    if (flags & (FLAG_OVERFLOW | FLAG_DP_OVERFLOW)) n_over++;
    if (flags & FLAG_MURTY) n_murty++;
    __syncwarp();
  }
  step_epilogue<T>(p, lane, warp, tot_in, tot_out, max_out, n_over, n_murty, 0, mstat);
}

// ------------------------------------------------------------------------------------------------
// Map part of RBPHDFilter::predict() for the Victoria Park model: births at
// MeasurementModel_VictoriaPark::inverseMeasure (src/...VictoriaPark.cpp:75-102: the 2-D inverse model
// with the pose rotated by -pi/2, diameter copied, covariance blockdiag(Hinv R2 Hinv^T, R22)), then P += Q.
template <typename T>
struct VPPredictParams {
  T* gm; int* cnt; unsigned long long* unused; int* flags;
  const T* pose; const T* Z;
  int N, cap, nZ, add_births, add_q;
  T R00, R01, R10, R11, R22, birth_w;
  T q[6];
};

template <typename T>
__global__ void predict_maps_vp_kernel(const VPPredictParams<T> p) {
  const int lane = threadIdx.x & 31;
  const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pi >= p.N) return;
  T* g = p.gm + (size_t)pi * VP_NPL * p.cap;
  int n = p.cnt[pi];
  n = n < 0 ? 0 : (n > p.cap ? p.cap : n);
  if (p.add_births) {
    unsigned long long mask = p.unused[pi];
    if (p.nZ < 64) mask &= (1ull << p.nZ) - 1ull;
    const int nb = __popcll(mask);
    const T px = p.pose[4 * pi], py = p.pose[4 * pi + 1], pth = p.pose[4 * pi + 2] - T(M<T>::PI / 2);
    bool over = false;
    for (int k = lane; k < nb; k += 32) {
      unsigned long long m = mask;   // k-th highest set bit (the reference pops from the back)
      for (int s = 0; s < k; s++) m &= ~(1ull << (63 - __clzll((long long)m)));
      const int z = 63 - __clzll((long long)m);
      const T r = p.Z[3 * z], b = p.Z[3 * z + 1], d = p.Z[3 * z + 2];
      T sn, cs;
      if constexpr (sizeof(T) == 4) sincosf(pth + b, &sn, &cs); else sincos(pth + b, &sn, &cs);
      const T h00 = cs, h01 = -r * sn, h10 = sn, h11 = r * cs;
      const T a00 = h00 * p.R00 + h01 * p.R10, a01 = h00 * p.R01 + h01 * p.R11;
      const T a10 = h10 * p.R00 + h11 * p.R10, a11 = h10 * p.R01 + h11 * p.R11;
      const int idx = n + k;
      if (idx < p.cap) {
        g[idx] = px + r * cs;
        g[p.cap + idx] = py + r * sn;
        g[2 * p.cap + idx] = d;
        g[3 * p.cap + idx] = a00 * h00 + a01 * h01;
        g[4 * p.cap + idx] = a00 * h10 + a01 * h11;
        g[5 * p.cap + idx] = T(0);
        g[6 * p.cap + idx] = a10 * h10 + a11 * h11;
        g[7 * p.cap + idx] = T(0);
        g[8 * p.cap + idx] = p.R22;
        g[9 * p.cap + idx] = p.birth_w;
      } else {
        over = true;
      }
    }
    over = __any_sync(FULL, over);
    n = (n + nb > p.cap) ? p.cap : n + nb;
    if (lane == 0) {
      p.cnt[pi] = n;
      p.unused[pi] = 0ull;
      if (over) p.flags[pi] |= FLAG_OVERFLOW | FLAG_BIRTH_OVERFLOW;
    }
    __syncwarp();
  }
  if (p.add_q) {
    for (int j = lane; j < n; j += 32) {
#pragma unroll
      for (int k = 0; k < 6; k++) g[(3 + k) * p.cap + j] += p.q[k];
    }
  }
}

}  // namespace rfsb200
