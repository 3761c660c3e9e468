/*
 * phd_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT (see phd_oracle.h).
 *
 * fp64 CPU restatement of the per-particle PHD measurement update of kykleung/RFS-SLAM.
 * Plain C++ (no Eigen / Boost); every function cites the reference lines it follows
 * (paths relative to /root/reference).  Quirks Q1..Q14 of SURVEY.md §8a are reproduced.
 *
 * Parity pin: this file is checked (tests/test_oracle_vs_ref.py, tests/golden/) against
 * oracle/_ref/libphd_ref.so = the reference's own headers/TUs compiled unmodified against
 * oracle/compat shims.  The reference's only known-answer test on this path
 * (test/MatrixPermanentTest.hpp:66-85) is reproduced in tests/test_oracle_kat.py.
 */
#include "phd_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <queue>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double PI = acos(-1); /* include/RandomVec.hpp:55 */

struct Gaussian {   /* include/GaussianMixture.hpp:60-64 + Landmark2d (mean, cov) */
  double x[2];
  double P[4];      /* row-major 2x2 */
  double w;
  double wprev;
  bool alive;       /* false = NULL hole left by merge */
};

struct Ctx {
  const rfsb200_model_desc* md;
  const rfsb200_filter_cfg* fc;
  const double* Z;
  int nZ;
  int sort_mode;
};

/* ---- 2x2 helpers (Eigen fixed-size closed forms: inverse = adj * (1/det)) ---- */
inline double det2(const double* A) { return A[0] * A[3] - A[1] * A[2]; }
inline void inv2(const double* A, double* Ai) {
  double invdet = 1.0 / det2(A);
  Ai[0] = A[3] * invdet;
  Ai[1] = -A[1] * invdet;
  Ai[2] = -A[2] * invdet;
  Ai[3] = A[0] * invdet;
}
/* e^T * Ai * e evaluated as (e^T Ai) e — include/RandomVec.hpp:387-407 */
inline double quad2(const double* Ai, const double* e) {
  double r0 = e[0] * Ai[0] + e[1] * Ai[2];
  double r1 = e[0] * Ai[1] + e[1] * Ai[3];
  return r0 * e[0] + r1 * e[1];
}
/* include/RandomVec.hpp:415-451 */
inline double gauss_pdf(const double* Sinv, double det, const double* e, double* md2_out) {
  double factor = sqrt(pow(2 * PI, 2) * det);
  double md2 = quad2(Sinv, e);
  double l = exp(-0.5 * md2) / factor;
  if (l != l) l = 0;
  if (md2_out) *md2_out = md2;
  return l;
}

/* src/MeasurementModel_RngBrg.cpp:70-115.  Sx = 3x3 pose covariance (row-major) */
bool rngbrg_measure(const rfsb200_model_desc& md, const double* pose, const double* Sx,
                    const double* lx, const double* P, double* zexp, double* S, double* H) {
  double dx = lx[0] - pose[0], dy = lx[1] - pose[1];
  double range2 = pow(dx, 2) + pow(dy, 2);
  double range = sqrt(range2);
  double bearing = atan2(dy, dx) - pose[2];
  while (bearing > PI) bearing -= 2 * PI;
  while (bearing < -PI) bearing += 2 * PI;
  zexp[0] = range;
  zexp[1] = bearing;
  double Hl[4] = {dx / range, dy / range, -dy / range2, dx / range2};
  double Hr[6] = {-dx / range, -dy / range, 0, dy / range2, -dx / range2, -1};
  /* cov = Hl P Hl^T + Hr Sx Hr^T + R */
  double HP[4] = {Hl[0] * P[0] + Hl[1] * P[2], Hl[0] * P[1] + Hl[1] * P[3],
                  Hl[2] * P[0] + Hl[3] * P[2], Hl[2] * P[1] + Hl[3] * P[3]};
  double A[4] = {HP[0] * Hl[0] + HP[1] * Hl[1], HP[0] * Hl[2] + HP[1] * Hl[3],
                 HP[2] * Hl[0] + HP[3] * Hl[1], HP[2] * Hl[2] + HP[3] * Hl[3]};
  double HS[6];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Hr[i * 3 + k] * Sx[k * 3 + j];
      HS[i * 3 + j] = s;
    }
  double B[4];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += HS[i * 3 + k] * Hr[j * 3 + k];
      B[i * 2 + j] = s;
    }
  S[0] = A[0] + B[0] + md.R[0];
  S[1] = A[1] + B[1] + md.R[1];
  S[2] = A[2] + B[2] + md.R[2];
  S[3] = A[3] + B[3] + md.R[3];
  if (H) memcpy(H, Hl, sizeof(Hl));
  if (range > md.range_max || range < md.range_min) return false;
  return true;
}

/* src/MeasurementModel_RngBrg.cpp:138-167 */
double rngbrg_pd(const rfsb200_model_desc& md, const double* pose, const double* lx,
                 bool& close) {
  close = false;
  double range = sqrt(pow(lx[0] - pose[0], 2) + pow(lx[1] - pose[1], 2));
  double Pd;
  if (range <= md.range_max && range >= md.range_min) {
    Pd = md.Pd;
    if (range >= (md.range_max - md.range_buffer) || range <= (md.range_min + md.range_buffer))
      close = true;
  } else {
    Pd = 0;
    if (range <= (md.range_max + md.range_buffer) && range >= (md.range_min - md.range_buffer))
      close = true;
  }
  return Pd;
}

/* src/KalmanFilter_RngBrg.cpp:52-65 */
bool rngbrg_innovation(const rfsb200_model_desc& md, const double* zexp, const double* zact,
                       double* innov) {
  innov[0] = zact[0] - zexp[0];
  innov[1] = zact[1] - zexp[1];
  if (md.innov_thr_range > 0 && fabs(innov[0]) > md.innov_thr_range) return false;
  while (innov[1] > PI) innov[1] -= 2 * PI;
  while (innov[1] < -PI) innov[1] += 2 * PI;
  if (md.innov_thr_bearing > 0 && fabs(innov[1]) > md.innov_thr_bearing) return false;
  return true;
}

struct LmkNew {
  double x[2];
  double P[4];
};

/* include/KalmanFilter.hpp:261-342 (batch correct: one landmark x all measurements) */
void kf_correct_batch(const Ctx& c, const double* pose, const double* Sx, const Gaussian& lm,
                      std::vector<LmkNew>& lmNew, std::vector<double>& lik,
                      std::vector<double>& md2v) {
  const int nZ = c.nZ;
  double zexp[2], S[4], H[4];
  if (!rngbrg_measure(*c.md, pose, Sx, lm.x, lm.P, zexp, S, H)) {
    for (int i = 0; i < nZ; i++) {
      lik[i] = 0;
      md2v[i] = 0;
    }
    return;
  }
  double Sinv[4];
  inv2(S, Sinv);
  const double* P = lm.P;
  /* K = P H^T S^-1 */
  double PHt[4] = {P[0] * H[0] + P[1] * H[1], P[0] * H[2] + P[1] * H[3],
                   P[2] * H[0] + P[3] * H[1], P[2] * H[2] + P[3] * H[3]};
  double K[4] = {PHt[0] * Sinv[0] + PHt[1] * Sinv[2], PHt[0] * Sinv[1] + PHt[1] * Sinv[3],
                 PHt[2] * Sinv[0] + PHt[3] * Sinv[2], PHt[2] * Sinv[1] + PHt[3] * Sinv[3]};
  /* P_updated = (I - K H) P, then symmetrised */
  double IKH[4] = {1 - (K[0] * H[0] + K[1] * H[2]), 0 - (K[0] * H[1] + K[1] * H[3]),
                   0 - (K[2] * H[0] + K[3] * H[2]), 1 - (K[2] * H[1] + K[3] * H[3])};
  double Pu[4] = {IKH[0] * P[0] + IKH[1] * P[2], IKH[0] * P[1] + IKH[1] * P[3],
                  IKH[2] * P[0] + IKH[3] * P[2], IKH[2] * P[1] + IKH[3] * P[3]};
  double Ps[4] = {(Pu[0] + Pu[0]) / 2, (Pu[1] + Pu[2]) / 2, (Pu[2] + Pu[1]) / 2,
                  (Pu[3] + Pu[3]) / 2};
  double detS = det2(S);
  for (int i = 0; i < nZ; i++) {
    const double* zact = c.Z + 2 * i;
    double innov[2];
    if (rngbrg_innovation(*c.md, zexp, zact, innov)) {
      lmNew[i].x[0] = lm.x[0] + (K[0] * innov[0] + K[1] * innov[1]);
      lmNew[i].x[1] = lm.x[1] + (K[2] * innov[0] + K[3] * innov[1]);
      memcpy(lmNew[i].P, Ps, sizeof(Ps));
      /* Q3: the likelihood uses the UNWRAPPED difference z - zexp (KalmanFilter.hpp:319) */
      double e[2] = {zact[0] - zexp[0], zact[1] - zexp[1]};
      double md2;
      double zl = gauss_pdf(Sinv, detS, e, &md2);
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
void sort_by_weight(std::vector<Gaussian>& g, int sort_mode) {
  if (sort_mode == PHD_SORT_STD) {
    std::sort(g.begin(), g.end(), [](Gaussian a, Gaussian b) { return a.w > b.w; });
  } else {
    std::stable_sort(g.begin(), g.end(),
                     [](const Gaussian& a, const Gaussian& b) { return a.w > b.w; });
  }
}

/* include/RBPHDFilter.hpp:543-725 */
void update_map(const Ctx& c, const double* pose, const double* Sx, std::vector<Gaussian>& gm,
                double& pweight, uint64_t& unused_mask, int32_t& n_in_fov) {
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
  if (fc.use_cluster_process) {
    for (unsigned m = 0; m < nM; m++) w_km_sum += gm[m].w;
  }
  std::vector<double> W((size_t)nM * nZ);
  std::vector<char> Mok((size_t)nM * nZ);
  std::vector<LmkNew> Mtab((size_t)nM * nZ);
  const double thr2 =
      fc.new_gaussian_create_innov_md_threshold * fc.new_gaussian_create_innov_md_threshold;
  std::vector<double> lik(nZ), md2(nZ);
  std::vector<LmkNew> lmNew(nZ);

  for (unsigned m = 0; m < nM; m++) { /* :597-641 */
    bool close;
    Pd[m] = rngbrg_pd(*c.md, pose, gm[m].x, close);
    if (close) {
      closeLim[m] = 1;
      Pd[m] = 1; /* Q2 */
    } else
      closeLim[m] = 0;
    double w_km = gm[m].w;
    double Pd_times_w_km = Pd[m] * w_km;
    if (Pd[m] != 0) {
      n_in_fov++;
      kf_correct_batch(c, pose, Sx, gm[m], lmNew, lik, md2);
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
  if (fc.use_cluster_process) { /* :661-668 (Q4) */
    pweight = exp(w_km_sum) * likelihoodProd * pweight;
  }
  for (unsigned m = 0; m < nM; m++) /* :675-683 */
    for (int z = 0; z < nZ; z++)
      if (Mok[m * nZ + z] && W[m * nZ + z] > 0) {
        Gaussian g;
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
    gm[m].wprev = gm[m].w; /* GaussianMixture::setWeight :339-344 */
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

/* ---- src/PermutationLexicographic.cpp:38-96 ------------------------------------------ */
struct PermLexi {
  unsigned nM, nZ, oSize, nP;
  bool last;
  std::vector<unsigned> o;
  PermLexi(unsigned nM_, unsigned nZ_, bool includeClutter) : nM(nM_), nZ(nZ_), nP(0), last(false) {
    if (nM != nZ) includeClutter = true;
    oSize = includeClutter ? nM + nZ : nM;
    o.resize(oSize);
    for (unsigned i = 0; i < oSize; i++) o[i] = (i < nZ) ? i : nZ;
  }
  unsigned next(unsigned* perm) {
    if (last) {
      for (unsigned i = 0; i < oSize; i++) perm[i] = 0;
      return 0;
    }
    for (unsigned i = 0; i < oSize; i++) perm[i] = o[i];
    if (oSize > 0) {
      unsigned u = nM, v = oSize - 1;
      while (u < v) {
        std::swap(o[u], o[v]);
        u++;
        v--;
      }
    }
    last = !std::next_permutation(o.begin(), o.end());
    nP++;
    return nP;
  }
};

/* ---- include/HungarianMethod.hpp:90-587 (maximize = true path; C modified then restored) */
bool hungarian_run(double** C, int n, int* soln, double* cost) {
  std::vector<double> lx(n), ly(n), slack(n);
  std::vector<int> xy(n), yx(n), p(2 * n);
  std::vector<char> S(n), T(n), NS(n), x_q(n), y_q(n);
  std::queue<int> q;
  int x, x_t, y, root = 0;
  bool pickFreeVertex = true, updateLabel;
  for (int i = 0; i < n; i++) {
    xy[i] = -1;
    S[i] = 0;
    yx[i] = -1;
    T[i] = 0;
  }
  double offset = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      if (C[i][j] < offset) offset = C[i][j];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) C[i][j] -= offset;
  /* Step 1 :189-221 */
  for (int xx = 0; xx < n; xx++) {
    lx[xx] = 0.0;
    ly[xx] = 0.0;
    for (int yy = 0; yy < n; yy++) {
      if (C[xx][yy] >= lx[xx]) {
        lx[xx] = C[xx][yy];
        xy[xx] = yy;
      }
    }
    int yy = xy[xx];
    x_t = yx[yy];
    if (yx[yy] != -1) {
      if (C[xx][yy] > C[x_t][yy]) {
        xy[x_t] = -1;
        yx[yy] = xx;
      } else {
        xy[xx] = -1;
      }
    } else {
      yx[yy] = xx;
    }
  }
  while (true) {
    if (pickFreeVertex) { /* Step 2 :243-326 */
      for (x = 0; x < n; x++) S[x] = 0;
      for (y = 0; y < n; y++) {
        T[y] = 0;
        NS[y] = 0;
      }
      for (x = 0; x < n; x++)
        if (xy[x] == -1) break;
      if (x == n) {
        if (offset != 0)
          for (x = 0; x < n; x++)
            for (y = 0; y < n; y++) C[x][y] = C[x][y] + offset;
        *cost = 0;
        for (x = 0; x < n; x++) {
          soln[x] = xy[x];
          *cost += C[x][xy[x]];
        }
        return true;
      }
      root = x;
      S[x] = 1;
      for (y = 0; y < n; y++) {
        slack[y] = lx[x] + ly[y] - C[x][y];
        if (fabs(slack[y]) < 1e-14) {
          slack[y] = 0;
          NS[y] = 1;
        }
      }
    }
    updateLabel = true; /* Step 3 :328-399 */
    for (y = 0; y < n; y++)
      if (NS[y] != T[y]) {
        updateLabel = false;
        break;
      }
    if (updateLabel) {
      double a = std::numeric_limits<double>::max();
      for (y = 0; y < n; y++)
        if (!T[y]) a = fmin(a, slack[y]);
      for (x = 0; x < n; x++)
        if (S[x]) lx[x] -= a;
      for (y = 0; y < n; y++)
        if (T[y]) ly[y] += a;
      for (y = 0; y < n; y++) {
        if (!T[y]) slack[y] -= a;
        if (slack[y] == 0) NS[y] = 1;
      }
    }
    for (y = 0; y < n; y++) /* Step 4 :401-585 */
      if (NS[y] && !T[y]) break;
    if (y >= n) return false; /* the reference would read yx[n] here: out of bounds (UB) */
    x_t = yx[y];
    if (x_t == -1) {
      bool found = false;
      int target = y + n;
      std::queue<int> q_empty;
      std::swap(q, q_empty);
      q.push(root);
      for (x = 0; x < n; x++) {
        x_q[x] = 0;
        y_q[x] = 0;
      }
      x_q[root] = 1;
      for (x = 0; x < 2 * n; x++) p[x] = -1;
      int t = q.front();
      while (!q.empty()) {
        t = q.front();
        if (t == target) {
          while (t != root) {
            if (t >= n) {
              x_t = p[t];
              xy[x_t] = t - n;
              yx[t - n] = x_t;
            }
            t = p[t];
          }
          found = true;
          break;
        }
        q.pop();
        if (t < n) {
          for (y = 0; y < n; y++)
            if (fabs(lx[t] + ly[y] - C[t][y]) < 1e-12 && !y_q[y] && xy[t] != y) {
              y_q[y] = 1;
              p[y + n] = t;
              q.push(y + n);
            }
        } else {
          t -= n;
          for (x = 0; x < n; x++)
            if (fabs(lx[x] + ly[t] - C[x][t]) < 1e-12 && S[x] && !x_q[x] && yx[t] == x) {
              x_q[x] = 1;
              p[x] = t + n;
              q.push(x);
            }
        }
      }
      if (!found) return false;
      pickFreeVertex = true;
    } else {
      S[x_t] = 1;
      T[y] = 1;
      for (int y_t = 0; y_t < n; y_t++)
        if (fabs(lx[x_t] + ly[y_t] - C[x_t][y_t]) < 1e-14) NS[y_t] = 1;
      for (int yy = 0; yy < n; yy++) {
        double s = lx[x_t] + ly[yy] - C[x_t][yy];
        if (s < slack[yy]) slack[yy] = s;
      }
      pickFreeVertex = false;
    }
  }
}

/* ---- include/MurtyAlgorithm.hpp + src/MurtyAlgorithm.cpp:106-336 ----------------------- */
struct MNode {
  int id;
  std::vector<int> a;
  double s;
  MNode* parent;
  bool has_assignment;
  std::vector<std::unique_ptr<MNode>> children;
  MNode(int id_) : id(id_), s(0), parent(nullptr), has_assignment(false) {}
};
struct MNodeCmp {
  bool operator()(MNode* a, MNode* b) const { return a->s < b->s; }
};
struct Murty {
  int k;
  int n;
  double** C;
  std::vector<std::vector<double>> Ct_store;
  std::vector<double*> Ct;
  std::unique_ptr<MNode> root;
  std::priority_queue<MNode*, std::vector<MNode*>, MNodeCmp> pq;
  double big;
  int rnR, rnC;
  Murty(double** C_, int n_, double bigNum = 10000)
      : k(0), n(n_), C(C_), root(new MNode(0)), big(bigNum), rnR(n_), rnC(n_) {
    Ct_store.assign(n, std::vector<double>(n));
    Ct.resize(n);
    for (int i = 0; i < n; i++) Ct[i] = Ct_store[i].data();
  }
  void setRealAssignmentBlock(int nR, int nC) {
    rnC = nC;
    rnR = nR;
    if (rnC > n) rnC = n;
    if (rnR > n) rnR = n;
  }
  /* returns rank or -1; score in *score */
  int findNextBest(double* score) {
    if (k == 0) {
      std::vector<int> a(n);
      double s;
      hungarian_run(C, n, a.data(), &s);
      root->a = a;
      root->s = s;
      root->has_assignment = true;
      k++;
      pq.push(root.get());
      *score = s;
      return k;
    }
    if (pq.empty()) {
      *score = 0;
      return -1;
    }
    MNode* parent = pq.top();
    int parent_partition = parent->id;
    pq.pop();
    const std::vector<int>& a_parent = parent->a;
    int partitionMax = rnR;
    if (rnR == n) partitionMax = n - 1;
    for (int nn = parent_partition; nn < partitionMax; nn++) {
      MNode* p = new MNode(nn);
      p->parent = parent;
      parent->children.emplace_back(p);
      std::vector<int> a(n);
      std::vector<char> freeCol(n, 1);
      double fixedScore = 0;
      for (int i = 0; i < parent_partition; i++) {
        a[i] = a_parent[i];
        freeCol[a[i]] = 0;
        fixedScore += C[i][a[i]];
      }
      for (int i = parent_partition; i < nn; i++) {
        a[i] = a_parent[i];
        freeCol[a[i]] = 0;
        fixedScore += C[i][a[i]];
      }
      int nFree = n - nn;
      std::vector<int> rowRemap(nFree), rowRemapR(n, 0), colRemap(nFree), colRemapR(n, 0);
      int nFreeCols = 0;
      for (int i = 0; i < nFree; i++) {
        rowRemap[i] = nn + i;
        rowRemapR[nn + i] = i;
      }
      for (int j = 0; j < n; j++)
        if (freeCol[j]) {
          colRemap[nFreeCols] = j;
          colRemapR[j] = nFreeCols;
          nFreeCols++;
        }
      for (int i = 0; i < nFree; i++)
        for (int j = 0; j < nFree; j++) Ct[i][j] = C[rowRemap[i]][colRemap[j]];
      MNode* current = p;
      MNode* next;
      do { /* negative constraints :243-262 (indices as in the reference, quirks included) */
        int currentPart = current->id;
        next = current->parent;
        const std::vector<int>& na = next->a;
        int di = rowRemapR[currentPart];
        int dj = colRemapR[na[currentPart]];
        Ct[di][dj] = -big;
        if (dj >= rnC) {
          for (int yy = 0; yy < nFree; yy++)
            if (yy >= rnC) Ct[di][yy] = -big;
        }
        current = next;
      } while (current != root.get() && current->id >= p->id);
      bool possible = false;
      int constraintRow = rowRemapR[p->id];
      for (int j = 0; j < nFree; j++)
        if (Ct[constraintRow][j] != -big) {
          possible = true;
          break;
        }
      if (possible) {
        std::vector<int> aTmp(nFree);
        double s = 0;
        hungarian_run(Ct.data(), nFree, aTmp.data(), &s);
        double s2 = 0;
        for (int i = 0; i < nFree; i++) {
          int ia = rowRemap[i];
          int ja = colRemap[aTmp[i]];
          a[ia] = ja;
          s2 += C[ia][a[ia]];
        }
        s2 += fixedScore;
        p->a = a;
        p->s = s2;
        p->has_assignment = true;
        pq.push(p);
      }
    }
    if (pq.empty()) {
      *score = 0;
      return -1;
    }
    MNode* hi = pq.top();
    *score = hi->s;
    k++;
    return k;
  }
};

/* include/RBPHDFilter.hpp:904-959: log table + extended matrix + Murty k<=200 */
double murty_sum(const double* Lp, int nR, int nC, const double* rowPd, const double* colClutter) {
  const double BIG_NEG_NUM = -1000;
  int n = nR + nC;
  std::vector<std::vector<double>> store(n, std::vector<double>(n, 0.0));
  std::vector<double*> Cp(n);
  for (int i = 0; i < n; i++) Cp[i] = store[i].data();
  for (int r = 0; r < nR; r++)
    for (int c = 0; c < nC; c++) {
      double v = Lp[r * nC + c];
      if (v == 0)
        v = BIG_NEG_NUM;
      else {
        v = log(v);
        if (v < BIG_NEG_NUM) v = BIG_NEG_NUM;
      }
      Cp[r][c] = v;
    }
  for (int r = 0; r < nR; r++)
    for (int c = nC; c < n; c++) Cp[r][c] = (r == c - nC) ? log(1 - rowPd[r]) : BIG_NEG_NUM;
  for (int r = nR; r < n; r++)
    for (int c = 0; c < nC; c++) Cp[r][c] = (r - nR == c) ? log(colClutter[c]) : BIG_NEG_NUM;
  for (int r = nR; r < n; r++)
    for (int c = nC; c < n; c++) Cp[r][c] = 0;
  Murty murty(Cp.data(), n);
  murty.setRealAssignmentBlock(nR, nC);
  double pl = 0, score = 0;
  for (int k = 0; k < 200; k++) {
    int rank = murty.findNextBest(&score);
    if (rank == -1 || score < BIG_NEG_NUM) break;
    pl += exp(score);
  }
  return pl;
}

/* include/RBPHDFilter.hpp:961-988 on one partition; Lp [nR][nC] likelihoods */
double lexi_sum(const double* Lp, int nR, int nC, const double* rowPd, const double* colClutter) {
  const double BIG_NEG_NUM = -1000;
  std::vector<double> Cp((size_t)nR * nC);
  for (int r = 0; r < nR; r++)
    for (int c = 0; c < nC; c++) {
      double v = Lp[r * nC + c];
      if (v == 0)
        v = BIG_NEG_NUM;
      else {
        v = log(v);
        if (v < BIG_NEG_NUM) v = BIG_NEG_NUM;
      }
      Cp[r * nC + c] = v;
    }
  double pl = 0;
  std::vector<unsigned> o(nR + nC + 1);
  PermLexi lexi(nR, nC, true);
  unsigned nPerm = lexi.next(o.data());
  while (nPerm != 0) {
    double ll = 0;
    for (int a = 0; a < nR; a++) {
      if ((int)o[a] < nC)
        ll += Cp[a * nC + o[a]];
      else
        ll += log(1 - rowPd[a]);
    }
    for (int a = nR; a < nR + nC; a++)
      if ((int)o[a] < nC) ll += log(colClutter[o[a]]);
    pl += exp(ll);
    nPerm = lexi.next(o.data());
  }
  return pl;
}

/*
 * src/CostMatrix.cpp:92-157 (partition) + include/RBPHDFilter.hpp:866-994 (loop over p < nP),
 * including Q6: components keep their original labels although nP = ncc - nMerged.
 * Labels: boost::connected_components numbers components in order of their lowest vertex
 * (vertices = rows 0..nR-1 then columns nR..nR+nC-1).
 */
struct Partitioning {
  std::vector<std::vector<unsigned>> ci, cj;
  int combinedZero, nP;
};

Partitioning partition_table(const double* L, int nE, int nZ) {
  Partitioning P;
  int nV = nE + nZ;
  std::vector<int> label(nV, -1);
  int ncc = 0;
  for (int v = 0; v < nV; v++) {
    if (label[v] != -1) continue;
    /* flood fill */
    std::vector<int> stack(1, v);
    label[v] = ncc;
    while (!stack.empty()) {
      int u = stack.back();
      stack.pop_back();
      if (u < nE) {
        for (int z = 0; z < nZ; z++)
          if (L[u * nZ + z] != 0 && label[nE + z] == -1) {
            label[nE + z] = ncc;
            stack.push_back(nE + z);
          }
      } else {
        int z = u - nE;
        for (int e = 0; e < nE; e++)
          if (L[e * nZ + z] != 0 && label[e] == -1) {
            label[e] = ncc;
            stack.push_back(e);
          }
      }
    }
    ncc++;
  }
  P.ci.assign(ncc, std::vector<unsigned>());
  P.cj.assign(ncc, std::vector<unsigned>());
  for (int i = 0; i < nE; i++) P.ci[label[i]].push_back(i);
  for (int j = nE; j < nV; j++) P.cj[label[j]].push_back(j - nE);
  int combinedZero = -1, nMerged = 0;
  for (int n = 0; n < ncc; n++) {
    if (P.ci[n].size() == 0 || P.cj[n].size() == 0) {
      if (combinedZero == -1)
        combinedZero = n;
      else if (P.ci[n].size() != 0) {
        P.ci[combinedZero].push_back(P.ci[n][0]);
        nMerged++;
      } else {
        P.cj[combinedZero].push_back(P.cj[n][0]);
        nMerged++;
      }
    }
  }
  P.combinedZero = combinedZero;
  P.nP = ncc - nMerged;
  return P;
}

double partition_likelihood(const double* L, int nE, int nZ, const double* evalPd,
                            const double* clutter, int32_t* flags) {
  int nV = nE + nZ;
  if (nV == 0) return 1;
  Partitioning P = partition_table(L, nE, nZ);
  std::vector<std::vector<unsigned>>& ci = P.ci;
  std::vector<std::vector<unsigned>>& cj = P.cj;
  const int combinedZero = P.combinedZero;
  int nP = P.nP;
  double l = 1;
  for (int p = 0; p < nP; p++) {
    int nRows = ci[p].size(), nCols = cj[p].size();
    bool isZero = (p == combinedZero);
    bool useMurty = true;
    if (nRows + nCols <= 8 || isZero) useMurty = false;
    double pl = 0;
    if (isZero) { /* :891-900 (Q5) */
      pl = 1;
      for (int r = 0; r < nRows; r++) pl *= evalPd[ci[p][r]];
      for (int cc = 0; cc < nCols; cc++) pl *= clutter[cj[p][cc]];
    } else {
      std::vector<double> Lp((size_t)nRows * nCols), rp(nRows), cp(nCols);
      for (int r = 0; r < nRows; r++) {
        rp[r] = evalPd[ci[p][r]];
        for (int cc = 0; cc < nCols; cc++) Lp[r * nCols + cc] = L[ci[p][r] * nZ + cj[p][cc]];
      }
      for (int cc = 0; cc < nCols; cc++) cp[cc] = clutter[cj[p][cc]];
      if (useMurty) {
        if (flags) *flags |= 2;
        pl = murty_sum(Lp.data(), nRows, nCols, rp.data(), cp.data());
      } else {
        pl = lexi_sum(Lp.data(), nRows, nCols, rp.data(), cp.data());
      }
    }
    l *= pl;
  }
  return l;
}

/* include/RBPHDFilter.hpp:821-997 */
double rfs_measurement_likelihood(const Ctx& c, const double* pose, const double* Sx,
                                  const std::vector<Gaussian>& gm,
                                  const std::vector<unsigned>& evalIdx,
                                  const std::vector<double>& evalPd, int32_t* flags) {
  const int nM = evalIdx.size();
  const int nZ = c.nZ;
  const double thr = c.fc->meas_likelihood_md_threshold * c.fc->meas_likelihood_md_threshold;
  std::vector<double> L((size_t)nM * nZ);
  const double zeroP[4] = {0, 0, 0, 0};
  for (int m = 0; m < nM; m++) {
    double zexp[2], S[4];
    rngbrg_measure(*c.md, pose, Sx, gm[evalIdx[m]].x, zeroP, zexp, S, nullptr);
    double Sinv[4];
    inv2(S, Sinv);
    double detS = det2(S);
    double Pd = evalPd[m];
    for (int n = 0; n < nZ; n++) {
      double e[2] = {c.Z[2 * n] - zexp[0], c.Z[2 * n + 1] - zexp[1]};
      double md2;
      L[m * nZ + n] = gauss_pdf(Sinv, detS, e, &md2) * Pd;
      if (md2 > thr) L[m * nZ + n] = 0;
    }
  }
  std::vector<double> clutter(nZ, c.md->clutter_intensity);
  double l = partition_likelihood(L.data(), nM, nZ, evalPd.data(), clutter.data(), flags);
  return l / c.md->clutter_integral;
}

/* include/RBPHDFilter.hpp:728-819 */
void importance_weighting(const Ctx& c, const double* pose, const double* Sx,
                          std::vector<Gaussian>& gm, double& pweight, int32_t* flags) {
  const rfsb200_filter_cfg& fc = *c.fc;
  const unsigned nM = gm.size();
  /* :737 int vs unsigned compare (Q14) */
  int nEvalPoints = (unsigned)fc.eval_point_count > nM ? (int)nM : fc.eval_point_count;
  std::vector<unsigned> evalIdx;
  std::vector<double> evalPd;
  if (nEvalPoints == 0) {
    pweight = std::numeric_limits<double>::denorm_min();
    return;
  }
  sort_by_weight(gm, c.sort_mode);
  for (unsigned m = 0; m < nM; m++) {
    if (gm[m].w < fc.eval_point_gaussian_weight) break;
    bool close;
    double Pd = rngbrg_pd(*c.md, pose, gm[m].x, close);
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
  /* per-component inverse / pdf factor (cached lazily in the reference; same values) */
  std::vector<double> Pinv(4 * nM), fac(nM);
  for (unsigned m = 0; m < nM; m++) {
    inv2(gm[m].P, &Pinv[4 * m]);
    fac[m] = sqrt(pow(2 * PI, 2) * det2(gm[m].P));
  }
  for (int e = 0; e < nEvalPoints; e++) {
    const Gaussian& ge = gm[evalIdx[e]];
    double vb = std::numeric_limits<double>::denorm_min();
    double va = std::numeric_limits<double>::denorm_min();
    for (unsigned m = 0; m < nM; m++) {
      double d[2] = {ge.x[0] - gm[m].x[0], ge.x[1] - gm[m].x[1]};
      double md2 = quad2(&Pinv[4 * m], d);
      double lk = exp(-0.5 * md2) / fac[m];
      if (lk != lk) lk = 0;
      vb += gm[m].wprev * lk;
      va += gm[m].w * lk;
    }
    prodBefore *= vb;
    prodAfter *= va;
  }
  double ml = rfs_measurement_likelihood(c, pose, Sx, gm, evalIdx, evalPd, flags);
  double overall = ml * prodBefore / prodAfter * exp(sumAfter - sumBefore);
  pweight = overall * pweight;
}

/* include/GaussianMixture.hpp:419-475 */
bool merge_pair(std::vector<Gaussian>& g, unsigned i1, unsigned i2, double t, double f) {
  if (!g[i1].alive || !g[i2].alive) return false;
  double w1 = g[i1].w, w2 = g[i2].w;
  double t2 = t * t;
  double Pi[4], d[2];
  inv2(g[i1].P, Pi);
  d[0] = g[i2].x[0] - g[i1].x[0];
  d[1] = g[i2].x[1] - g[i1].x[1];
  double d1 = quad2(Pi, d);
  if (d1 > t2) {
    inv2(g[i2].P, Pi);
    d[0] = g[i1].x[0] - g[i2].x[0];
    d[1] = g[i1].x[1] - g[i2].x[1];
    double d2 = quad2(Pi, d);
    if (d2 > t2) return false;
  }
  double wm = w1 + w2;
  if (wm == 0) return false;
  double xm[2] = {(g[i1].x[0] * w1 + g[i2].x[0] * w2) / wm, (g[i1].x[1] * w1 + g[i2].x[1] * w2) / wm};
  double e1[2] = {xm[0] - g[i1].x[0], xm[1] - g[i1].x[1]};
  double e2[2] = {xm[0] - g[i2].x[0], xm[1] - g[i2].x[1]};
  double Sm[4];
  for (int r = 0; r < 2; r++)
    for (int cidx = 0; cidx < 2; cidx++) {
      double a = w1 * (g[i1].P[r * 2 + cidx] + f * e1[r] * e1[cidx]);
      double b = w2 * (g[i2].P[r * 2 + cidx] + f * e2[r] * e2[cidx]);
      Sm[r * 2 + cidx] = (a + b) / wm;
    }
  g[i1].x[0] = xm[0];
  g[i1].x[1] = xm[1];
  memcpy(g[i1].P, Sm, sizeof(Sm));
  g[i1].w = wm;
  g[i1].wprev = 0;
  g[i2].alive = false; /* removeGaussian :310-322 */
  g[i2].w = 0;
  g[i2].wprev = 0;
  return true;
}

/* include/GaussianMixture.hpp:394-416 */
void merge_all(std::vector<Gaussian>& g, double t, double f) {
  unsigned n = g.size();
  for (unsigned i = 0; i < n; i++) {
    if (!g[i].alive) continue;
    for (unsigned j = i + 1; j < n; j++) merge_pair(g, i, j, t, f);
  }
}

/* include/GaussianMixture.hpp:477-521: keeps w >= t, drops holes; output weight-descending */
void prune(std::vector<Gaussian>& g, double t, int sort_mode) {
  if (g.size() < 1) return;
  sort_by_weight(g, sort_mode);
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

void expand_pose_cov(const phd_io* io, int i, double* Sx) {
  for (int k = 0; k < 9; k++) Sx[k] = 0;
  const double* s = nullptr;
  if (io->pose_cov_mode == 1) s = io->pose_cov;
  if (io->pose_cov_mode == 2) s = io->pose_cov + 6 * (size_t)i;
  if (!s) return;
  Sx[0] = s[0]; Sx[1] = s[1]; Sx[2] = s[2];
  Sx[3] = s[1]; Sx[4] = s[3]; Sx[5] = s[4];
  Sx[6] = s[2]; Sx[7] = s[4]; Sx[8] = s[5];
}

}  // namespace

extern "C" int phd_oracle_update(phd_io* io) {
  if (!io || !io->model || !io->cfg || io->N < 0 || io->nZ < 0 || io->nZ > 64) return -1;
  if (io->model->model_id != RFSB200_MODEL_RNGBRG) return -5;
  const int N = io->N;
  std::vector<int64_t> off(N + 1, 0);
  for (int i = 0; i < N; i++) off[i + 1] = off[i] + io->count_in[i];
  std::vector<std::vector<Gaussian>> maps(N);
  for (int i = 0; i < N; i++) {
    maps[i].resize(io->count_in[i]);
    for (int m = 0; m < io->count_in[i]; m++) {
      int64_t k = off[i] + m;
      Gaussian& g = maps[i][m];
      g.x[0] = io->mean_in[2 * k];
      g.x[1] = io->mean_in[2 * k + 1];
      g.P[0] = io->cov_in[3 * k];
      g.P[1] = g.P[2] = io->cov_in[3 * k + 1];
      g.P[3] = io->cov_in[3 * k + 2];
      g.w = io->w_in[k];
      g.wprev = 0;
      g.alive = true;
    }
  }
  std::vector<double> pw(io->weight_in, io->weight_in + N);
  std::vector<uint64_t> unused(N, 0);
  std::vector<int32_t> nfov(N, 0), flags(N, 0);
  Ctx c{io->model, io->cfg, io->Z, io->nZ, io->sort_mode};
#ifdef _OPENMP
  if (io->n_threads > 0) omp_set_num_threads(io->n_threads);
#endif
  auto t0 = std::chrono::steady_clock::now();
  if (io->nZ > 0) { /* include/RBPHDFilter.hpp:451-452 (Q11) */
#pragma omp parallel
    {
#pragma omp for
      for (int i = 0; i < N; i++) {
        double Sx[9];
        expand_pose_cov(io, i, Sx);
        update_map(c, io->pose + 3 * (size_t)i, Sx, maps[i], pw[i], unused[i], nfov[i]);
      }
      if (!io->cfg->use_cluster_process && io->stage >= PHD_STAGE_WEIGHTING) {
#pragma omp for
        for (int i = 0; i < N; i++) {
          double Sx[9];
          expand_pose_cov(io, i, Sx);
          importance_weighting(c, io->pose + 3 * (size_t)i, Sx, maps[i], pw[i], &flags[i]);
        }
      }
      if (io->stage >= PHD_STAGE_MERGE) {
#pragma omp for
        for (int i = 0; i < N; i++)
          merge_all(maps[i], io->cfg->merging_threshold, io->cfg->merging_cov_inflation_factor);
      }
      if (io->stage >= PHD_STAGE_FULL) {
#pragma omp for
        for (int i = 0; i < N; i++) prune(maps[i], io->cfg->pruning_threshold, io->sort_mode);
      }
    }
  }
  io->elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int64_t k = 0;
  for (int i = 0; i < N; i++) {
    int cnt = 0;
    for (const Gaussian& g : maps[i]) {
      if (!g.alive) continue;
      if (k >= io->cap_total) return -4;
      io->mean_out[2 * k] = g.x[0];
      io->mean_out[2 * k + 1] = g.x[1];
      io->cov_out[3 * k] = g.P[0];
      io->cov_out[3 * k + 1] = g.P[1];
      io->cov_out[3 * k + 2] = g.P[3];
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

/* src/MatrixPermanent.cpp:41-113 */
extern "C" double phd_oracle_permanent(const double* A, int n) {
  std::vector<double> x(n);
  std::vector<int> g(n, 0);
  double p = 0, s = -1;
  for (int i = 0; i < n; i++) {
    double rs = 0;
    for (int j = 0; j < n; j++) rs += A[i * n + j];
    x[i] = A[i * n + n - 1] - 0.5 * rs;
  }
  p = s;
  for (int i = 0; i < n; i++) p *= x[i];
  for (int k = 2; k <= pow(2, n - 1); k++) {
    int j = 0;
    if (k % 2 == 0)
      j = 0;
    else {
      j = 1;
      while (g[j - 1] == 0) j++;
    }
    s *= -1;
    double z = 1 - 2 * g[j];
    g[j] = !g[j];
    double xp = 1;
    for (int i = 0; i < n; i++) {
      x[i] += z * A[i * n + j];
      xp *= x[i];
    }
    p += s * xp;
  }
  double r = 2 * p;
  if (n % 2 != 0) r *= -1;
  return r;
}

extern "C" int64_t phd_oracle_lexi_count(int nM, int nZ) {
  PermLexi pl(nM, nZ, true);
  std::vector<unsigned> o(nM + nZ + 1);
  int64_t cnt = 0;
  while (pl.next(o.data()) != 0) cnt++;
  return cnt;
}

extern "C" double phd_oracle_partition_likelihood(const double* L, int nE, int nZ,
                                                  const double* evalPd, const double* clutter,
                                                  int32_t* flags) {
  return partition_likelihood(L, nE, nZ, evalPd, clutter, flags);
}

extern "C" int phd_oracle_partition(const double* L, int nR, int nC, int* nRows, int* nCols, int* isZero) {
  Partitioning P = partition_table(L, nR, nC);
  for (int p = 0; p < P.nP; p++) {
    nRows[p] = (int)P.ci[p].size();
    nCols[p] = (int)P.cj[p].size();
    isZero[p] = (p == P.combinedZero) ? 1 : 0;
  }
  return P.nP;
}

extern "C" double phd_oracle_murty_sum(const double* Lp, int nR, int nC, const double* rowPd,
                                       const double* colClutter) {
  return murty_sum(Lp, nR, nC, rowPd, colClutter);
}
