// phd_kernels.cuh — the fused per-particle PHD measurement update for sm_100a.
//
// One launch processes every particle of the shard: ONE WARP PER PARTICLE, persistent CTAs (one CTA per SM of up to 20
// warps — the fp32 single-cluster kernel — chosen at configure time); a warp's first particle is static, the rest are
// handed out by a global atomic queue, drawn one particle ahead with the next particle's planes prefetched into L2.
// Per particle (reference include/RBPHDFilter.hpp):
//   S0  TMA bulk loads (cp.async.bulk + mbarrier) of the particle's 6 SoA planes HBM -> smem
//   S1  GM-PHD corrector, updateMap :597-641 — EKF innovation / likelihood between every Gaussian
//       and every measurement in registers; gated survivors appended (m-major, z-minor)
//   S2  per-measurement normalisers kappa + sum_m W (:644-659) and SC-PHD weight (:661-668)
//   S3  posterior weights of the new Gaussians, S4 missed-detection weights + sensing-limit
//       heuristic (:686-706), unused-measurement mask (:709-720)
//   S5  multi-feature importance weighting (:728-819, rfsMeasurementLikelihood :821-997)
//   S6  greedy GaussianMixture::merge (include/GaussianMixture.hpp:394-475), exact order
//   S7  prune (:477-521): keep w >= t, weight-descending; gather from shared memory, coalesced stores to HBM
//   S8  deterministic [sum w, sum w^2] reduction by the last CTA (ParticleFilter.hpp:352-363,406-411)
//
// No tensor cores: the 2x2 / 2x3 EKF blocks are register math; the kernel is instruction-issue / latency bound
// (7.6 k warp instructions per particle against 5.8 kB of HBM traffic).
#pragma once
#include "common.cuh"

namespace rfsb200 {

constexpr int MAX_WARPS_PER_CTA = 20;
constexpr int MAX_Z = 64;
constexpr int MAX_EVAL = 32;
constexpr int DP_MAXB = 7;     // assignment-sum DP in shared memory: the smaller side of a partition has <= 7 members
constexpr int DP_GMAXB = 15;   // ... in the per-warp global workspace (KParams::dp_scratch): <= 15 members
constexpr int MAX_COMP = 96;   // connected components of the (eval point, measurement) graph
constexpr int MAX_PAIRS = 144; // merge: passing pairs (+ cluster links) kept per particle

// flag bits written per particle
constexpr int FLAG_OVERFLOW = 1;
constexpr int FLAG_BIRTH_OVERFLOW = 8;   // set by predict / append when births did not fit gm_capacity; the next update turns
                                         // it into FLAG_OVERFLOW of its own result (and counts it), so the drop is not lost
constexpr int FLAG_CAND_OVERFLOW = 32;  // rfsb200_birth_candidates: a new candidate did not fit the particle's candidate list (sticky like bit 8)
constexpr int FLAG_MURTY = 2;       // a partition with nR + nC > 8 (reference would use Murty-200)
constexpr int FLAG_DP_OVERFLOW = 4; // partition too large for the on-chip DP
constexpr int FLAG_MURTY_DROPPED = 16;   // Murty compatibility: the record of a partition did not fit the record buffer

// Murty compatibility (rfsb200_filter_cfg::murty_compat, quirk Q7): for a partition with nR + nC > 8 the reference adds up
// Murty's 200 best assignments only (include/RBPHDFilter.hpp:904-959).  The device always adds up ALL assignments; with
// the switch on it also writes the partition out — eval points' P_D, the likelihood block, the log of its exact sum —
// and the host replaces the exact sum by the truncated one (murty_compat.hpp) before the weights are summed.
struct MurtyOut {
  unsigned long long* buf;   // records: [pi | nR << 32] [nC] [log exact sum] [P_D x nR] [L x nR x nC], 8-byte words
  unsigned int* count;       // [0] words used, [1] records written
  unsigned int cap_words;
  int pi;
};

struct CommSlot { double s1, s2; unsigned long long epoch; unsigned long long pad; };   // 32 B; mailbox = [4][8]: banks 0 / 1 the
                                                                                         // weight sums of even / odd update epochs, 2 / 3 rfsb200_comm_barrier
constexpr int COMM_BANKS = 4;
// A slot travels as four 8-byte words, each = 32 bits of payload | (epoch tag << 32): every word is written atomically
// and carries its own arrival flag, so the sender needs no fence between data and flag and the receiver simply polls
// until all four tags match (the "LL" protocol of collective libraries: one NVLink traversal of latency).
__device__ __forceinline__ void comm_send(CommSlot* dst, double s1, double s2, unsigned long long epoch) {
  const unsigned long long a = (unsigned long long)__double_as_longlong(s1), b = (unsigned long long)__double_as_longlong(s2);
  const unsigned long long tag = (epoch & 0xffffffffull) << 32;
  unsigned long long* q = reinterpret_cast<unsigned long long*>(dst);
  st_relaxed_sys_u64(q + 0, (a & 0xffffffffull) | tag);
  st_relaxed_sys_u64(q + 1, (a >> 32) | tag);
  st_relaxed_sys_u64(q + 2, (b & 0xffffffffull) | tag);
  st_relaxed_sys_u64(q + 3, (b >> 32) | tag);
}
__device__ __forceinline__ bool comm_recv(const CommSlot* src, unsigned long long epoch, double& s1, double& s2) {
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(src);
  const unsigned long long w0 = ld_relaxed_sys_u64(q + 0), w1 = ld_relaxed_sys_u64(q + 1), w2 = ld_relaxed_sys_u64(q + 2),
                           w3 = ld_relaxed_sys_u64(q + 3);
  const unsigned long long tag = epoch & 0xffffffffull;
  if ((w0 >> 32) != tag || (w1 >> 32) != tag || (w2 >> 32) != tag || (w3 >> 32) != tag) return false;
  s1 = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  s2 = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
  return true;
}

template <typename T>
struct KParams {
  // model (MeasurementModel_RngBrg / KalmanFilter_RngBrg configs)
  T R00, R01, R11;
  T Pd, kappa;
  T rmin, rmax, rbuf;
  T thr_r, thr_b;
  // filter config
  T birth_w, gate2, eval_min_w, wl_gate2, merge_t2, merge_f, prune_t;
  int n_eval, use_sc, sum_method, merge_algo;
  double log_clutter_integral;
  double log_kappa;
  // shapes
  int N, cap, W, nZ, pose_cov_mode;
  int warp_bytes;  // shared memory per warp
  int mf_bytes;    // multi-feature scratch inside it
  int n_eval_cap, zcap;   // sizes the multi-feature scratch was laid out for
  // state in / out
  const T* gm_in;
  const int* cnt_in;
  const double* w_in;
  const T* pose;      // [N][4]
  const T* pose_cov;  // [8] or [N][8]
  const T* Z;         // [nZ][2] device copy of the batch (written by CTA 0 of the update kernel, read by predict_maps)
  // the measurement batch travels BY VALUE, fp64 as the caller holds it: no copy precedes the launch
  double Zval[MAX_Z * 3];
  T* Zdev_w;          // where CTA 0 leaves the T copy
  // host-facing step (rfsb200_update_host, pinned caller buffers): poses / particle weights / pose covariances are read
  // straight from the caller's memory over PCIe by the update kernel (host_in_convert), which leaves the device copies
  // every stage reads (NULL pose_h: the device arrays are current)
  const double* pose_h;     // [N][3]
  const double* weight_h;   // [N] or NULL = keep
  const double* pcov_h;     // [N][6] (pose_cov_mode 2) or NULL
  double cov6[6];           // pose_cov_mode 1 with pose_h: the shared covariance by value
  T* pose_w;                // [N][4]
  double* pose64_w;         // [N][3]
  T* pcov_w;                // [N][8] / [8]
  double* w_front;          // [N] particle weights of the state the step starts from
  unsigned long long* hin_ready;   // [grid] slice flags of the host-facing step (see host_in_convert)
  unsigned long long* done_host;   // pinned flag the last CTA sets to done_value when every result is in host memory
  unsigned long long done_value;   // number of this host-facing launch (> 0)
  T* gm_out;
  int* cnt_out;
  double* w_out;
  unsigned long long* unused;
  int* nfov;
  int* flags;
  // reductions
  double* sums;                   // [2]
  unsigned long long* totals;     // [0]=gm_in total [1]=gm_out total
  int* istats;                    // [0]=max out [1]=n_overflow [2]=n_murty [3]=merge recomputations
  unsigned int* mstats;           // [8] merge statistics (fallback reasons, pairs, clusters)
  unsigned int* ticket;
  unsigned int* work_counter;     // dynamic particle queue, re-armed by the last CTA
  // fused cross-GPU sum of [sum w, sum w^2] over peer memory (NVLink / NVSwitch), see S8
  int comm_rank, comm_world, fused_normalize;
  unsigned long long comm_epoch;
  unsigned long long* murty_buf;    // Murty compatibility (see MurtyOut); NULL = off
  unsigned int* murty_count;
  unsigned int murty_cap_words;
  void* comm_peer[8];               // mailbox of every rank (own included), mapped into this process
  int* comm_error;                  // set to 1 if a peer did not arrive in time
  unsigned long long comm_timeout_ns;   // how long the last CTA waits for the peers (RFSB200_COMM_TIMEOUT_MS, default 2 s)
  // deferred consumption of the cross-GPU sums (RFSB200_UPDATE_DEFER_NORMALIZE): a launch with comm_defer sends its pair
  // to the peers and ends without waiting (weights stay unnormalised); the NEXT launch (comm_pending) finds the pairs of
  // epoch comm_prev_epoch in its mailbox during set-up — they have been there for a whole step — and divides the incoming
  // particle weights by their total while it loads them: the same division on the same operands as the eager
  // normalisation, so the results are bit-identical, but no rank ever waits for the slowest one inside a step.
  int comm_defer, comm_pending;   // comm_pending = number of ranks whose pair is to be picked up (0: nothing is open)
  int comm_pending_scale;         // the open normalisation is that of the weights this launch reads (else: of a buffer it overwrites)
  unsigned long long comm_prev_epoch;
  unsigned long long* stats_out;  // [13] totals/istats/mstats of the finished step, published by the last CTA
  // multi-feature weighting: global workspace of the assignment-sum DP for partitions beyond the on-chip tables,
  // 2 x (1 << dp_gmaxb) doubles per warp of the grid (NULL: none); dp_onchip = largest smaller side summed on chip
  double* dp_scratch;
  int dp_gmaxb, dp_onchip;
  // host-facing step (rfsb200_update_host with pinned, device-accessible caller buffers): the results are ALSO stored
  // straight into the caller's host memory (posted writes over PCIe), so no device-to-host copy follows the kernel.
  // NULL = not used.
  double* w_host;                   // [N] particle weights as the launch leaves them (normalised on the fused path)
  unsigned long long* unused_host;  // [N]
  int* nfov_host;                   // [N]
  unsigned long long* stats_host;   // [15] sums[2] (as bits) followed by stats_out[13]
  // stage-timing build of the kernels only (RFSB200_UPDATE_STAGE_TIMES): [0..11] SM cycles of all warps per stage
  // (see StageClock), [12] first CTA start, [13] last CTA past its set-up, [14] last warp out of particles, [15] last
  // CTA done — %globaltimer ns
  unsigned long long* prof;
};


// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T wrap_pi(T a) {
  // same result as the reference's while-loops (src/MeasurementModel_RngBrg.cpp:92-93,
  // src/KalmanFilter_RngBrg.cpp:58-61); the loops only run for |a| > 3*pi
  if (a > M<T>::PI) a -= M<T>::TWO_PI;
  if (a < -M<T>::PI) a += M<T>::TWO_PI;
  if (!(M<T>::abs_(a) <= M<T>::PI)) {
    while (a > M<T>::PI) a -= M<T>::TWO_PI;
    while (a < -M<T>::PI) a += M<T>::TWO_PI;
  }
  return a;
}

// ordering used by every sort: weight descending, then position ascending (stable)
template <typename T>
__device__ __forceinline__ bool before(T wa, unsigned ia, T wb, unsigned ib) {
  return (wa > wb) || (wa == wb && ia < ib);
}

// Bitonic sort of P (power of two) (key, idx) pairs in shared memory by one warp.
template <typename T>
__device__ __forceinline__ void warp_bitonic(T* kw, unsigned* ki, int P, int lane) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int q = i + j;
        bool up = ((i & k) == 0);
        T wa = kw[i], wb = kw[q];
        unsigned ia = ki[i], ib = ki[q];
        bool a_first = before(wa, ia, wb, ib);
        if (a_first != up) {
          kw[i] = wb; kw[q] = wa;
          ki[i] = ib; ki[q] = ia;
        }
      }
      __syncwarp();
    }
  }
}

// Bitonic sort, DESCENDING, of P (power of two) packed u64 keys in shared memory by one warp.
__device__ __forceinline__ void warp_bitonic_desc64(unsigned long long* key, int P, int lane) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int q = i + j;
        const bool up = ((i & k) == 0);
        const unsigned long long a = key[i], b = key[q];
        if ((a > b) != up) { key[i] = b; key[q] = a; }
      }
      __syncwarp();
    }
  }
}

// The same sort with the keys in REGISTERS: 32 * KPL keys, lane l holds elements l * KPL .. l * KPL + KPL - 1, so the
// compare-exchanges at distance j < KPL stay inside a lane (no shared-memory traffic, no warp barrier) and those at
// distance j >= KPL are one shuffle per key with lane l ^ (j / KPL).  Same network, same result as
// warp_bitonic_desc64; about half its instructions.  Not inlined: one copy per KPL serves every call site.
template <int KPL>
__device__ __noinline__ void warp_sort_desc64_reg(unsigned long long* key, int lane) {
  constexpr int P = 32 * KPL;
  unsigned long long v[KPL];
#pragma unroll
  for (int r = 0; r < KPL; r++) v[r] = key[lane * KPL + r];
#pragma unroll
  for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < KPL) {
#pragma unroll
        for (int r = 0; r < KPL; r++) {
          const int q = r ^ j;
          if (q > r) {
            const bool up = (((lane * KPL + r) & k) == 0);
            const unsigned long long a = v[r], b = v[q];
            const bool sw = (a > b) != up;
            v[r] = sw ? b : a;
            v[q] = sw ? a : b;
          }
        }
      } else {
        const int lj = j / KPL;
        const bool lower = (lane & lj) == 0;
#pragma unroll
        for (int r = 0; r < KPL; r++) {
          const unsigned long long o = __shfl_xor_sync(FULL, v[r], lj);
          const bool up = (((lane * KPL + r) & k) == 0);   // (the two ends of a pair agree on bit k: j < k)
          const bool keep_max = (lower == up);
          const bool mine_bigger = v[r] > o;
          v[r] = (mine_bigger == keep_max) ? v[r] : o;
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < KPL; r++) key[lane * KPL + r] = v[r];
  __syncwarp();
}

// descending sort of P (power of two) packed keys (entries beyond the live count padded with 0 by the caller): the
// register network for 128 and 256 keys, the shared-memory one for the small and the very large sets.  Used by the 2-D
// kernels (C3 shape with multi-feature weighting: 323 -> 280 us); the Victoria Park kernels keep the shared-memory
// network, the register one measured 1-5 % slower there (126 live registers around the call).
__device__ __forceinline__ void warp_sort_desc64(unsigned long long* key, int P, int lane) {
  if (P == 256) warp_sort_desc64_reg<8>(key, lane);
  else if (P == 128) warp_sort_desc64_reg<4>(key, lane);
  else warp_bitonic_desc64(key, P, lane);
}

__device__ __forceinline__ int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// Stage timing (the reference's TimingInfo, include/RBPHDFilter.hpp:152-167: mapUpdate_kf, particleWeighting, mapMerge,
// mapPrune): lane 0 of every warp reads the SM clock at the stage boundaries of every particle and accumulates the
// differences; the totals of all warps go to KParams::prof at the end of the launch.  Compiled in only in the PROF
// instantiations of the kernels; the product kernels carry none of it.
constexpr int STAGE_LOAD = 0, STAGE_CORRECT = 1, STAGE_WEIGHT = 2, STAGE_MFWEIGHT = 3, STAGE_MERGE = 4, STAGE_PRUNE = 5,
              STAGE_M1 = 6, STAGE_M2 = 7, STAGE_M3 = 8, STAGE_M4 = 9, N_STAGES = 10;   // M1..M4: inside the merge (STAGE_MERGE = the rest of it)
template <bool PROF>
struct StageClock {
  long long t;
  unsigned long long acc[N_STAGES];
  __device__ __forceinline__ void start() {
    if constexpr (PROF) {
      t = clock64();
#pragma unroll
      for (int k = 0; k < N_STAGES; k++) acc[k] = 0ull;
    }
  }
  __device__ __forceinline__ void mark(int k) {
    if constexpr (PROF) {
      const long long n = clock64();
      acc[k] += (unsigned long long)(n - t);
      t = n;
    }
  }
  __device__ __forceinline__ void flush(unsigned long long* prof, int lane) {
    if constexpr (PROF) {
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N_STAGES; k++) atomicAdd(&prof[k], acc[k]);
        atomicMax(&prof[14], globaltimer_ns());
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// S6: GaussianMixture::merge.  cur = 7 planes of W; holes are marked by weight < 0.
// Exact reference semantics: rows i ascending; for each live i, j ascending from i+1; test with
// the CURRENT state of i; on success i absorbs j (moment matching with inflation f) and j dies.
template <typename T>
struct MergeRow {
  T x, y, pxx, pxy, pyy, w, i00, i01, i11;
};

template <typename T>
__device__ __forceinline__ void inv_sym2(T a, T b, T c, T& i00, T& i01, T& i11) {
  T invdet = T(1) / (a * c - b * b);
  i00 = c * invdet;
  i01 = -b * invdet;
  i11 = a * invdet;
}

template <typename T>
__device__ __forceinline__ bool merge_test(const MergeRow<T>& r, const T* cur, int W, int j, T t2) {
  T dx = cur[j] - r.x, dy = cur[W + j] - r.y;
  T d1 = (dx * r.i00 + dy * r.i01) * dx + (dx * r.i01 + dy * r.i11) * dy;
  if (!(d1 > t2)) return true;
  T j00, j01, j11;
  inv_sym2(cur[2 * W + j], cur[3 * W + j], cur[4 * W + j], j00, j01, j11);
  T d2 = (dx * j00 + dy * j01) * dx + (dx * j01 + dy * j11) * dy;
  return !(d2 > t2);
}

template <typename T>
__device__ __forceinline__ void load_row(MergeRow<T>& r, const T* cur, int W, int i) {
  r.x = cur[i]; r.y = cur[W + i];
  r.pxx = cur[2 * W + i]; r.pxy = cur[3 * W + i]; r.pyy = cur[4 * W + i];
  r.w = cur[5 * W + i];
  inv_sym2(r.pxx, r.pxy, r.pyy, r.i00, r.i01, r.i11);
}

// absorb j into row r (all lanes compute the same values; lane 0 writes). returns false if w_m == 0.
template <typename T>
__device__ __forceinline__ bool merge_absorb(MergeRow<T>& r, T* cur, int W, int i, int j, T f, int lane) {
  T w1 = r.w, w2 = cur[5 * W + j];
  T wm = w1 + w2;
  if (wm == T(0)) return false;
  T x2 = cur[j], y2 = cur[W + j];
  T q00 = cur[2 * W + j], q01 = cur[3 * W + j], q11 = cur[4 * W + j];
  const T iw = T(1) / wm;
  T xm = (r.x * w1 + x2 * w2) * iw, ym = (r.y * w1 + y2 * w2) * iw;
  T ax = xm - r.x, ay = ym - r.y, bx = xm - x2, by = ym - y2;
  T s00 = (w1 * (r.pxx + f * ax * ax) + w2 * (q00 + f * bx * bx)) * iw;
  T s01 = (w1 * (r.pxy + f * ax * ay) + w2 * (q01 + f * bx * by)) * iw;
  T s11 = (w1 * (r.pyy + f * ay * ay) + w2 * (q11 + f * by * by)) * iw;
  r.x = xm; r.y = ym; r.pxx = s00; r.pxy = s01; r.pyy = s11; r.w = wm;
  inv_sym2(s00, s01, s11, r.i00, r.i01, r.i11);
  __syncwarp();
  if (lane == 0) {
    cur[i] = xm; cur[W + i] = ym;
    cur[2 * W + i] = s00; cur[3 * W + i] = s01; cur[4 * W + i] = s11;
    cur[5 * W + i] = wm; cur[6 * W + i] = T(0);
    cur[5 * W + j] = T(-1);  // hole
    cur[6 * W + j] = T(0);
  }
  __syncwarp();
  return true;
}

// scan j in [jstart, n) for the first live j passing the test against row r (rad2 = optional
// conservative pre-filter plane: reject when |d|^2 > max(rad2_i, rad2_j)); returns -1 if none
template <typename T, bool PREFILTER>
__device__ __forceinline__ int merge_scan(const MergeRow<T>& r, const T* cur, const T* rad2, T rad2_i,
                                          int W, int jstart, int n, T t2, int lane) {
  for (int base = jstart; base < n; base += 32) {
    int j = base + lane;
    bool pass = false;
    if (j < n) {
      T wj = cur[5 * W + j];
      if (wj >= T(0)) {
        bool cand = true;
        if (PREFILTER) {
          T dx = cur[j] - r.x, dy = cur[W + j] - r.y;
          T rr = M<T>::max_(rad2_i, rad2[j]);
          cand = !(dx * dx + dy * dy > rr);
        }
        if (cand) pass = merge_test(r, cur, W, j, t2);
      }
    }
    unsigned b = __ballot_sync(FULL, pass);
    if (b) return base + __ffs(b) - 1;
  }
  return -1;
}

// conservative squared reach of a component for the merge test: d_mahalanobis >= |d|^2 / lambda_max
// and lambda_max <= trace for a PSD matrix; non-PD covariances get an infinite reach.
template <typename T>
__device__ __forceinline__ T merge_reach2(T pxx, T pxy, T pyy, T t2) {
  bool pd = (pxx > T(0)) && (pyy > T(0)) && (pxx * pyy - pxy * pxy > T(0));
  return pd ? t2 * (pxx + pyy) * T(1.001) : M<T>::inf();
}

template <typename T>
__device__ void merge_bruteforce(T* cur, int W, int n, T t2, T f, int lane) {
  for (int i = 0; i < n; i++) {
    if (cur[5 * W + i] < T(0)) continue;
    MergeRow<T> r;
    load_row(r, cur, W, i);
    int jstart = i + 1;
    while (true) {
      int j = merge_scan<T, false>(r, cur, nullptr, T(0), W, jstart, n, t2, lane);
      if (j < 0) break;
      merge_absorb(r, cur, W, i, j, f, lane);
      jstart = j + 1;
    }
  }
}

// Clustered merge: same results as merge_bruteforce, or MERGE_FALLBACK with nothing modified.
//
//  The greedy merge only ever changes components that pass the Mahalanobis test with something,
//  and a component can only pass the test with a partner closer than max(reach_a, reach_b), where
//  reach^2 = t^2 * trace(P) bounds t^2 * lambda_max(P).  So:
//   M1  counting sort of the components on <= 256 cells over x whose width is >= the largest
//       reach: all partners of a component lie in its own and the two adjacent cells.
//   M2  every component (in cell order) tests the components behind it up to the end of the next
//       cell: distance pre-filter, then the exact test on the original parameters -> list of
//       passing pairs.
//   M3  connected components ("clusters") of the passing-pair graph by label propagation.
//   M4  ONE LANE PER CLUSTER runs the reference's sequential loop restricted to its cluster: rows
//       ascending, partners ascending, exact test against the current row (kept in registers).
//       Once a row has absorbed something it is re-tested against every neighbour (cell lookup):
//       an unowned neighbour that passes is claimed (CAS) and absorbed, exactly as the reference
//       would; a neighbour owned by ANOTHER cluster means the two clusters interact and the
//       lane-parallel schedule could differ from the sequential one.
//       Nothing is written to the mixture during M4: merged rows go to a log and deaths to a
//       bitmap, so on a conflict the two clusters are simply linked and M3/M4 run again.
//   M5  commit the log.
//  Clusters are independent in the reference's order as long as no row of one ever comes within
//  reach of a component of another, which is what the conflict test checks, conservatively.
constexpr int MERGE_OK = 0;
constexpr int MERGE_FALLBACK = 1;   // nothing modified: run the exhaustive merge instead
constexpr int MAX_CLUSTERS = 64;
constexpr int MAX_ROWLOG = 48;
constexpr int MEMBERS_CAP = 2 * MAX_PAIRS;   // members of all clusters together (every member is an end of a passing pair)
constexpr int MAX_MERGE_ROUNDS = 4;
constexpr int MERGE_MSTART = (MAX_CLUSTERS + 2 + 3) & ~3;   // entries of memberStart[]
constexpr unsigned NO_OWNER = 0xffffffffu;

template <typename T>
struct XY { T x, y; };

template <typename T>
struct MergeScratch {
  unsigned* label;           // [W] cluster label (smallest member index) / NO_OWNER   (aliases aux; M2: candidate list)
  unsigned short* order;     // [W] cell order -> component index
  unsigned short* cellStart; // [264]
  unsigned* counters;        // [4]: 0 = number of pairs, 1 = logged rows
  // --- region that is dead outside the merge (the prune sort keys alias all of it) ---
  XY<T>* sxy;                // [W + 4] M1/M2 only: (x, y) in cell order, four sentinels behind   } same storage
  unsigned short* members;   // [MEMBERS_CAP] members of cluster c at memberStart[c] .. memberStart[c + 1]   } (M3+ only)
  unsigned short* memberStart; // [MAX_CLUSTERS + 2]                    }
  T* rowLog;                 // [MAX_ROWLOG][6]                         }
  unsigned short* rowIdx;    // [MAX_ROWLOG]                            }
  unsigned short* headPre;   // [32] cluster heads in front of each word of headBits   }
  unsigned* pairs;           // [MAX_PAIRS]   (M1: the scatter cursors, 128 words)
  unsigned* memberCount;     // [MAX_CLUSTERS]
  unsigned* deadBits;        // [W / 32] bit per component
  unsigned* headBits;        // [W / 32] cluster heads
  unsigned* seenBits;        // [W / 32] first-visit marks of the pair ends
  int nwords;                // W / 32
  T* keys;                   // [W] prune sort keys (fp64) / packed u64 keys (fp32)
};

template <typename T>
__host__ __device__ inline int merge_region_a_bytes(int W) {
  const int rl = (MAX_ROWLOG * 6 * (int)sizeof(T) + 7) & ~7;
  const int a = rl + MEMBERS_CAP * 2 + MERGE_MSTART * 2 + MAX_ROWLOG * 2 + 32 * 2;
  const int b = 2 * (W + 4) * (int)sizeof(T);   // sxy[W] + four sentinels
  return ((a > b ? a : b) + 15) & ~15;
}
template <typename T>
__host__ __device__ inline int merge_only_bytes(int W) {
  const int a = merge_region_a_bytes<T>(W) + MAX_PAIRS * 4 + MAX_CLUSTERS * 4 + 3 * ((W + 31) >> 5) * 4;
  const int b = W * 8;   // prune sort keys: T[W] (fp64) or packed u64[W] (fp32)
  return ((a > b ? a : b) + 15) & ~15;
}
// bytes in front of the merge-only region: order u16[W] | cellStart u16[264] | counters u32[4]
__host__ __device__ inline int merge_head_bytes(int W) { return (2 * W + 264 * 2 + 16 + 15) & ~15; }
// label[] is the caller's aux array and is not counted here
template <typename T>
__host__ __device__ inline int merge_scratch_bytes(int W) {
  return merge_head_bytes(W) + merge_only_bytes<T>(W);
}

template <typename T>
__device__ __forceinline__ MergeScratch<T> carve_merge_scratch(unsigned char* base, unsigned* aux, int W) {
  MergeScratch<T> m;
  m.label = aux;
  m.counters = reinterpret_cast<unsigned*>(base);
  m.order = reinterpret_cast<unsigned short*>(m.counters + 4);
  m.cellStart = m.order + W;
  unsigned char* r = base + merge_head_bytes(W);
  m.keys = reinterpret_cast<T*>(r);
  m.sxy = reinterpret_cast<XY<T>*>(r);
  m.rowLog = reinterpret_cast<T*>(r);
  m.members = reinterpret_cast<unsigned short*>(r + ((MAX_ROWLOG * 6 * (int)sizeof(T) + 7) & ~7));
  m.memberStart = m.members + MEMBERS_CAP;
  m.rowIdx = m.memberStart + MERGE_MSTART;
  m.headPre = m.rowIdx + MAX_ROWLOG;
  m.pairs = reinterpret_cast<unsigned*>(r + merge_region_a_bytes<T>(W));
  m.memberCount = m.pairs + MAX_PAIRS;
  m.nwords = (W + 31) >> 5;
  m.deadBits = m.memberCount + MAX_CLUSTERS;
  m.headBits = m.deadBits + m.nwords;
  m.seenBits = m.headBits + m.nwords;
  return m;
}

// exact test of the pair (a, b) on the parameters stored in cur (include/GaussianMixture.hpp:435-442,
// the row is a)
template <typename T>
__device__ __forceinline__ bool merge_test_pair(const T* cur, int W, int a, int b, T t2) {
  const T dx = cur[b] - cur[a], dy = cur[W + b] - cur[W + a];
  T i00, i01, i11;
  inv_sym2(cur[2 * W + a], cur[3 * W + a], cur[4 * W + a], i00, i01, i11);
  const T d1 = (dx * i00 + dy * i01) * dx + (dx * i01 + dy * i11) * dy;
  if (!(d1 > t2)) return true;
  inv_sym2(cur[2 * W + b], cur[3 * W + b], cur[4 * W + b], i00, i01, i11);
  const T d2 = (dx * i00 + dy * i01) * dx + (dx * i01 + dy * i11) * dy;
  return !(d2 > t2);
}

// lane-local absorb: row r (registers) takes component j (read-only); false if the weights sum to 0
template <typename T>
__device__ __forceinline__ bool lane_absorb(MergeRow<T>& r, const T* cur, int W, int j, T f) {
  const T w1 = r.w, w2 = cur[5 * W + j];
  const T wm = w1 + w2;
  if (wm == T(0)) return false;
  const T x2 = cur[j], y2 = cur[W + j];
  const T q00 = cur[2 * W + j], q01 = cur[3 * W + j], q11 = cur[4 * W + j];
  const T iw = T(1) / wm;
  const T xm = (r.x * w1 + x2 * w2) * iw, ym = (r.y * w1 + y2 * w2) * iw;
  const T ax = xm - r.x, ay = ym - r.y, bx = xm - x2, by = ym - y2;
  const T s00 = (w1 * (r.pxx + f * ax * ax) + w2 * (q00 + f * bx * bx)) * iw;
  const T s01 = (w1 * (r.pxy + f * ax * ay) + w2 * (q01 + f * bx * by)) * iw;
  const T s11 = (w1 * (r.pyy + f * ay * ay) + w2 * (q11 + f * by * by)) * iw;
  r.x = xm; r.y = ym; r.pxx = s00; r.pxy = s01; r.pyy = s11; r.w = wm;
  inv_sym2(s00, s01, s11, r.i00, r.i01, r.i11);
  return true;
}

// 16-byte fill of a word array (n4 = number of uint4 entries) by the warp
struct alignas(16) Word4 { unsigned a, b, c, d; };
struct alignas(16) Quad { float v[4]; };   // four consecutive fp32 entries of a plane in one 16-byte shared-memory load
__device__ __forceinline__ void warp_fill4(void* dst, unsigned v, int n4, int lane) {
  Word4* d = reinterpret_cast<Word4*>(dst);
  const Word4 q{v, v, v, v};
  for (int k = lane; k < n4; k += 32) d[k] = q;
}

template <typename T, bool PROF>
__device__ __forceinline__ int merge_clustered(T* cur, const MergeScratch<T>& ms, int W, int n, T t2, T f, bool has_wprev,
                               int lane, unsigned (&mstat)[8], T xmin, T xmax, T trmax, bool bad, StageClock<PROF>& clk) {
  // ---- M1: counting sort on cells.  xmin / xmax / trmax (largest trace(P)) / bad (some covariance
  //      is not PD) over the n components were gathered by the corrector (warp-uniform values). ----
  if (bad) { mstat[0]++; return MERGE_FALLBACK; }   // a non-PD covariance has no finite reach
  const T tt = t2 * T(1.001);                // reach^2 of a component = tt * trace(P)
  const T rmax2 = trmax * tt;
  if (!(rmax2 < M<T>::inf()) || !(xmax - xmin < M<T>::inf())) { mstat[0]++; return MERGE_FALLBACK; }
  const unsigned lt = (1u << lane) - 1u;
  for (int c = lane; c < 132; c += 32) reinterpret_cast<unsigned*>(ms.cellStart)[c] = 0u;
  const T span = xmax - xmin;
  const T rmax = M<T>::sqrt_(rmax2) * T(1.001);
  T cw = span * T(1.0 / 255.5);
  cw = cw > rmax ? cw : rmax;
  const T invw = cw > T(0) ? T(1) / cw : T(0);
  auto cell_of = [&](T x) -> int {   // monotone in x; NaN -> 0
    return (int)M<T>::min_(M<T>::max_((x - xmin) * invw, T(0)), T(255));
  };
  __syncwarp();
  for (int j = lane; j < n; j += 32) {   // histogram at cell+1; 16-bit counters packed in words
    const int c1 = cell_of(cur[j]) + 1;
    atomicAdd(reinterpret_cast<unsigned*>(ms.cellStart) + (c1 >> 1), (c1 & 1) ? 0x10000u : 1u);
  }
  __syncwarp();
  {  // inclusive prefix over entries 1..256: lane owns 8 consecutive entries
    int loc[8];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { loc[k] = ms.cellStart[1 + lane * 8 + k]; sum += loc[k]; }
    const int incl = warp_incl_scan(sum, lane);
    int run = incl - sum;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) { run += loc[k]; ms.cellStart[1 + lane * 8 + k] = (unsigned short)run; }
    __syncwarp();
    if (lane == 0) { ms.cellStart[257] = (unsigned short)n; ms.cellStart[258] = (unsigned short)n; }
  }
  // scatter: per-cell cursors (cell c -> next free position), 16-bit counters packed in words
  unsigned* cursor = ms.pairs;   // the pair list is not in use yet (MAX_PAIRS >= 128 words)
  for (int c = lane; c < 128; c += 32) cursor[c] = reinterpret_cast<const unsigned*>(ms.cellStart)[c];
  __syncwarp();
  for (int j = lane; j < n; j += 32) {
    XY<T> q;
    q.x = cur[j]; q.y = cur[W + j];
    const int c = cell_of(q.x);
    const unsigned old = atomicAdd(cursor + (c >> 1), (c & 1) ? 0x10000u : 1u);
    const unsigned pos = (c & 1) ? (old >> 16) : (old & 0xffffu);
    ms.order[pos] = (unsigned short)j;
    ms.sxy[pos] = q;
  }
  if (lane < 4) { XY<T> far; far.x = M<T>::inf(); far.y = M<T>::inf(); ms.sxy[n + lane] = far; }   // sentinels (see M2a)
  __syncwarp();
  // (the order inside a cell is schedule dependent; nothing below depends on it: M2 visits every
  //  pair of one cell or of adjacent cells exactly once, M4 takes minima over index)
  clk.mark(STAGE_M1);
  // ---- M2: candidate pairs ------------------------------------------------------------------------
  //  a) by distance: position s against the positions behind it up to the end of the next cell, four at a time,
  //     branch-free; the hits (|d|^2 <= largest reach^2) are compacted into a list in label[] (not in use yet).
  //     Positions past the end of the next cell need no test of their own: they are at least one cell width
  //     (>= the largest reach) away in x and cannot hit; four sentinels at infinity follow the last position.
  //  b) lane-parallel over that list: the two components' own reaches, then the exact test on the original
  //     parameters -> list of passing pairs (collected in the storage of sxy[], which is dead by then, and copied
  //     to pairs[] at the end: the positions of the list live in pairs[] until then)
  unsigned* candHits = ms.label;   // [W] per position with a hit: bit b = position s + 1 + b is within the largest reach
  unsigned short* candPos = reinterpret_cast<unsigned short*>(ms.pairs);   // [2 * MAX_PAIRS] its position s (the cursors are dead)
  constexpr int CAND_CAP = 2 * MAX_PAIRS;
  int ncand = 0;
  bool far_window = false;
  for (int sb = 0; sb < n; sb += 32) {
    const int s = sb + lane;
    const bool live = s < n;
    XY<T> me;
    me.x = M<T>::inf(); me.y = M<T>::inf();   // (inf - inf = NaN: a dead lane never hits)
    if (live) me = ms.sxy[s];
    const int end = live ? (int)ms.cellStart[cell_of(me.x) + 2] : 0;
    unsigned hits = 0;
    int sh = 0;
    for (int t0 = s + 1; __any_sync(FULL, t0 < end); t0 += 4, sh += 4) {
      const int tb = t0 < end ? t0 : n;   // a lane that is through reads the sentinels
      unsigned hm = 0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const XY<T> q = ms.sxy[tb + u];
        const T dx = q.x - me.x, dy = q.y - me.y;
        hm |= (dx * dx + dy * dy <= rmax2) ? (1u << u) : 0u;
      }
      if (sh < 32) hits |= hm << sh;
      else far_window |= hm != 0u;   // more than 32 positions inside one window: not handled here
    }
    const unsigned bh = __ballot_sync(FULL, hits != 0u);
    if (bh) {
      const int pos = ncand + __popc(bh & lt);
      if (hits != 0u && pos < W && pos < CAND_CAP) { candHits[pos] = hits; candPos[pos] = (unsigned short)s; }
      ncand += __popc(bh);
    }
  }
  __syncwarp();
  if (ncand > W || ncand > CAND_CAP || __any_sync(FULL, far_window)) { mstat[1]++; return MERGE_FALLBACK; }
  int npass = 0;
  unsigned* passing = reinterpret_cast<unsigned*>(ms.sxy);   // [MAX_PAIRS]
  for (int base = 0; base < ncand; base += 32) {
    const int q = base + lane;
    unsigned hits = 0;
    int sp = 0;
    if (q < ncand) { hits = candHits[q]; sp = candPos[q]; }
    const int j = ms.order[sp];
    const T ax = cur[j], ay = cur[W + j];
    const T rj = tt * (cur[2 * W + j] + cur[4 * W + j]);
    while (__any_sync(FULL, hits != 0u)) {
      unsigned key = 0;
      bool pass = false;
      if (hits) {
        const int tp = sp + __ffs(hits);
        hits &= hits - 1;
        const int k = ms.order[tp];
        const T dx = cur[k] - ax, dy = cur[W + k] - ay;
        const T d2 = dx * dx + dy * dy;
        const T rk = tt * (cur[2 * W + k] + cur[4 * W + k]);
        if (!(d2 > M<T>::max_(rj, rk))) {
          const int lo = j < k ? j : k, hi = j < k ? k : j;
          key = ((unsigned)lo << 16) | (unsigned)hi;
          pass = merge_test_pair(cur, W, lo, hi, t2);
        }
      }
      const unsigned bp = __ballot_sync(FULL, pass);
      if (bp) {
        const int pos = npass + __popc(bp & lt);
        if (pass && pos < MAX_PAIRS) passing[pos] = key;
        npass += __popc(bp);
      }
    }
  }
  __syncwarp();
  for (int k = lane; k < npass && k < MAX_PAIRS; k += 32) ms.pairs[k] = passing[k];
  clk.mark(STAGE_M2);
  if (npass == 0) return MERGE_OK;
  if (npass > MAX_PAIRS) { mstat[1]++; return MERGE_FALLBACK; }
  mstat[5] += (unsigned)npass;
  // the candidate list is dead: label[] takes its array over
  warp_fill4(ms.label, NO_OWNER, W >> 2, lane);
  if (lane < 4) ms.counters[lane] = lane == 0 ? (unsigned)npass : 0u;
  for (int c = lane; c < MAX_CLUSTERS; c += 32) ms.memberCount[c] = 0u;
  if (lane < ms.nwords) { ms.deadBits[lane] = 0u; ms.headBits[lane] = 0u; ms.seenBits[lane] = 0u; }
  __syncwarp();

  for (int round = 0; round < MAX_MERGE_ROUNDS; round++) {
    const int np = (int)ms.counters[0];
    // ---- M3: clusters by label propagation (converges to the smallest index of each cluster) ------
    for (int k = lane; k < np; k += 32) {
      const unsigned key = ms.pairs[k];
      ms.label[key >> 16] = key >> 16;
      ms.label[key & 0xffffu] = key & 0xffffu;
    }
    __syncwarp();
    while (true) {
      bool changed = false;
      for (int k = lane; k < np; k += 32) {
        const unsigned key = ms.pairs[k];
        const unsigned a = key >> 16, b = key & 0xffffu;
        const unsigned la = ms.label[a], lb = ms.label[b];
        if (la != lb) {
          const unsigned m = la < lb ? la : lb;
          atomicMin(&ms.label[a], m);
          atomicMin(&ms.label[b], m);
          changed = true;
        }
      }
      __syncwarp();
      if (!__any_sync(FULL, changed)) break;
    }
    // ---- cluster slots and member lists, through the ends of the pairs (every member is one): the first visit
    //      of a component is detected on a bitmap.  Slot of a cluster = rank of its head among the heads. --------
    for (int k = lane; k < np; k += 32) {
      const unsigned key = ms.pairs[k];
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const unsigned e = side ? (key & 0xffffu) : (key >> 16);
        if (ms.label[e] == e) atomicOr(&ms.headBits[e >> 5], 1u << (e & 31));
      }
    }
    __syncwarp();
    int nclusters;
    {
      const int cnt = lane < ms.nwords ? __popc(ms.headBits[lane]) : 0;
      const int incl = warp_incl_scan(cnt, lane);
      nclusters = __shfl_sync(FULL, incl, 31);
      ms.headPre[lane] = (unsigned short)(incl - cnt);   // heads in front of word `lane`
    }
    if (nclusters > MAX_CLUSTERS) { mstat[2]++; return MERGE_FALLBACK; }
    __syncwarp();
    auto slot_of = [&](unsigned head) -> int {
      return (int)ms.headPre[head >> 5] + __popc(ms.headBits[head >> 5] & ((1u << (head & 31)) - 1u));
    };
    for (int k = lane; k < np; k += 32) {   // member counts
      const unsigned key = ms.pairs[k];
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const unsigned e = side ? (key & 0xffffu) : (key >> 16);
        const unsigned bit = 1u << (e & 31);
        if (!(atomicOr(&ms.seenBits[e >> 5], bit) & bit)) atomicAdd(&ms.memberCount[slot_of(ms.label[e])], 1u);
      }
    }
    __syncwarp();
    {  // exclusive scan of the counts -> memberStart; the counts become the fill cursors
      const int c0 = 2 * lane < nclusters ? (int)ms.memberCount[2 * lane] : 0;
      const int c1 = 2 * lane + 1 < nclusters ? (int)ms.memberCount[2 * lane + 1] : 0;
      const int incl = warp_incl_scan(c0 + c1, lane);
      const int ex = incl - (c0 + c1);
      ms.memberStart[2 * lane] = (unsigned short)ex;
      ms.memberStart[2 * lane + 1] = (unsigned short)(ex + c0);
      if (lane == 31) ms.memberStart[64] = (unsigned short)incl;
      ms.memberCount[2 * lane] = (unsigned)ex;
      ms.memberCount[2 * lane + 1] = (unsigned)(ex + c0);
    }
    __syncwarp();
    for (int k = lane; k < np; k += 32) {   // fill (a set bit is cleared by its first visitor)
      const unsigned key = ms.pairs[k];
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const unsigned e = side ? (key & 0xffffu) : (key >> 16);
        const unsigned bit = 1u << (e & 31);
        if (atomicAnd(&ms.seenBits[e >> 5], ~bit) & bit) {
          const unsigned pos = atomicAdd(&ms.memberCount[slot_of(ms.label[e])], 1u);
          ms.members[pos] = (unsigned short)e;
        }
      }
    }
    __syncwarp();
    clk.mark(STAGE_M3);
    // ---- M4: one lane per cluster; read-only on the mixture -----------------------------------------
    bool conflict = false, logfull = false;
    auto is_dead = [&](int k) -> bool { return (ms.deadBits[k >> 5] >> (k & 31)) & 1u; };
    for (int cb = 0; cb < nclusters; cb += 32) {
      const int sl = cb + lane;
      if (sl < nclusters) {
        unsigned short* mem = ms.members + ms.memberStart[sl];
        const int nm = (int)ms.memberStart[sl + 1] - (int)ms.memberStart[sl];
        for (int u = 1; u < nm; u++) {   // members ascending
          const unsigned short v = mem[u];
          int q = u - 1;
          while (q >= 0 && mem[q] > v) { mem[q + 1] = mem[q]; q--; }
          mem[q + 1] = v;
        }
        const unsigned myLabel = mem[0];
        for (int a = 0; a < nm && !conflict; a++) {
          const int i = mem[a];
          if (is_dead(i)) continue;
          MergeRow<T> r;
          load_row(r, cur, W, i);
          bool changed = false;
          int jlast = i;
          while (true) {
            int best = 0x7fffffff;
            for (int b = a + 1; b < nm; b++) {       // members, ascending: the first hit is the smallest
              const int k = mem[b];
              if (k > jlast && !is_dead(k) && merge_test(r, cur, W, k, t2)) { best = k; break; }
            }
            if (changed) {
              // neighbours of the changed row: every partner lies within sqrt(max(reach_row, rmax2))
              const T ri2 = tt * (r.pxx + r.pyy);
              const bool pd = (r.pxx > T(0)) && (r.pyy > T(0)) && (r.pxx * r.pyy - r.pxy * r.pxy > T(0));
              if (!pd) { logfull = true; break; }
              const T R2 = M<T>::max_(ri2, rmax2);
              const T R = M<T>::sqrt_(R2) * T(1.001);
              // a partner k has |x_k - x_row| <= R and cell_of is monotone: its cell lies in [cell(x - R), cell(x + R)]
              const int clo = cell_of(r.x - R), chi = cell_of(r.x + R);
              const int t0 = ms.cellStart[clo], t1 = ms.cellStart[chi + 1];
              for (int t = t0; t < t1; t++) {
                const int k = ms.order[t];
                if (k <= jlast || k >= best) continue;
                const T dx = cur[k] - r.x, dy = cur[W + k] - r.y;
                const T d2 = dx * dx + dy * dy;
                if (d2 > R2) continue;
                if (d2 > M<T>::max_(ri2, tt * (cur[2 * W + k] + cur[4 * W + k]))) continue;
                const unsigned owner = ms.label[k];
                if (owner == myLabel) continue;            // a member (tested above) or already absorbed
                if (owner != NO_OWNER) {                   // another cluster: link the two and redo M3/M4
                  conflict = true;
                  const unsigned q = atomicAdd(&ms.counters[0], 1u);
                  const unsigned lo = owner < myLabel ? owner : myLabel, hi = owner < myLabel ? myLabel : owner;
                  if (q < (unsigned)MAX_PAIRS) ms.pairs[q] = (lo << 16) | hi;
                  else logfull = true;
                  break;
                }
                if (merge_test(r, cur, W, k, t2)) best = k;
              }
              if (conflict) break;
            }
            if (best == 0x7fffffff) break;
            if (ms.label[best] != myLabel) {
              const unsigned prev = atomicCAS(&ms.label[best], NO_OWNER, myLabel);
              if (prev != NO_OWNER) {                      // claimed by another cluster meanwhile
                conflict = true;
                const unsigned q = atomicAdd(&ms.counters[0], 1u);
                const unsigned lo = prev < myLabel ? prev : myLabel, hi = prev < myLabel ? myLabel : prev;
                if (q < (unsigned)MAX_PAIRS) ms.pairs[q] = (lo << 16) | hi;
                else logfull = true;
                break;
              }
            }
            jlast = best;
            if (lane_absorb(r, cur, W, best, f)) {   // w == 0: the reference moves on to j+1
              changed = true;
              atomicOr(&ms.deadBits[best >> 5], 1u << (best & 31));
            }
          }
          if (changed && !conflict) {
            const unsigned q = atomicAdd(&ms.counters[1], 1u);
            if (q < (unsigned)MAX_ROWLOG) {
              ms.rowIdx[q] = (unsigned short)i;
              T* e = ms.rowLog + q * 6;
              e[0] = r.x; e[1] = r.y; e[2] = r.pxx; e[3] = r.pxy; e[4] = r.pyy; e[5] = r.w;
            } else {
              logfull = true;
            }
          }
        }
      }
      __syncwarp();
    }
    clk.mark(STAGE_M4);
    if (__any_sync(FULL, logfull)) { mstat[3]++; return MERGE_FALLBACK; }
    if (!__any_sync(FULL, conflict)) {
      // ---- M5: commit --------------------------------------------------------------------------------
      mstat[6] += (unsigned)nclusters;
      const int nrows = (int)ms.counters[1];
      for (int q = lane; q < nrows; q += 32) {
        const int i = ms.rowIdx[q];
        const T* e = ms.rowLog + q * 6;
        cur[i] = e[0]; cur[W + i] = e[1];
        cur[2 * W + i] = e[2]; cur[3 * W + i] = e[3]; cur[4 * W + i] = e[4];
        cur[5 * W + i] = e[5];
        if (has_wprev) cur[6 * W + i] = T(0);
      }
      {  // holes: lane = word of the death bitmap
        unsigned d = lane < ms.nwords ? ms.deadBits[lane] : 0u;
        while (d) {
          const int j = lane * 32 + __ffs(d) - 1;
          d &= d - 1;
          cur[5 * W + j] = T(-1);
        }
      }
      __syncwarp();
      return MERGE_OK;
    }
    // interacting clusters: their heads were linked in the pair list; reset and go round again
    mstat[4]++;
    warp_fill4(ms.label, NO_OWNER, W >> 2, lane);
    for (int c = lane; c < MAX_CLUSTERS; c += 32) ms.memberCount[c] = 0u;
    if (lane < ms.nwords) { ms.deadBits[lane] = 0u; ms.headBits[lane] = 0u; ms.seenBits[lane] = 0u; }
    if (lane == 0) ms.counters[1] = 0u;
    __syncwarp();
    if ((int)ms.counters[0] > MAX_PAIRS) { mstat[1]++; return MERGE_FALLBACK; }
  }
  mstat[2]++;
  return MERGE_FALLBACK;   // still interacting after MAX_MERGE_ROUNDS
}

// ------------------------------------------------------------------------------------------------
// S5 helpers: assignment sum of one partition by DP over subsets of its smaller side (fp64).
//  rows: eval points (miss factor 1-Pd), cols: measurements (clutter factor kappa), L = Pd*pdf.
//  Equals the reference's exhaustive enumeration (include/RBPHDFilter.hpp:961-988).
template <typename T>
__device__ double partition_dp(const T* L, int nZ, unsigned rmask, unsigned long long cmask,
                               const T* evalPd, double kappa, double* f0, double* f1, int lane) {
  int nR = __popc(rmask), nC = __popcll(cmask);
  bool rowsSmall = nR <= nC;
  int b = rowsSmall ? nR : nC;
  int S = 1 << b;
  // index lists
  int small[DP_GMAXB];
  {
    int k = 0;
    if (rowsSmall) { unsigned m = rmask; while (m) { int r = __ffs(m) - 1; m &= m - 1; if (k < DP_GMAXB) small[k] = r; k++; } }
    else { unsigned long long m = cmask; while (m) { int c = __ffsll((long long)m) - 1; m &= m - 1; if (k < DP_GMAXB) small[k] = c; k++; } }
  }
  for (int s = lane; s < S; s += 32) f0[s] = (s == 0) ? 1.0 : 0.0;
  __syncwarp();
  double* fa = f0;
  double* fb = f1;
  if (rowsSmall) {
    unsigned long long m = cmask;
    while (m) {  // columns one at a time: clutter, or assigned to a row in the mask
      int c = __ffsll((long long)m) - 1; m &= m - 1;
      for (int s = lane; s < S; s += 32) {
        double v = fa[s] * kappa;
        for (int k = 0; k < b; k++) if (s & (1 << k)) v += fa[s ^ (1 << k)] * (double)L[small[k] * nZ + c];
        fb[s] = v;
      }
      __syncwarp();
      double* t = fa; fa = fb; fb = t;
    }
    double tot = 0;
    for (int s = lane; s < S; s += 32) {
      double v = fa[s];
      for (int k = 0; k < b; k++) if (!(s & (1 << k))) v *= (1.0 - (double)evalPd[small[k]]);
      tot += v;
    }
    tot = warp_sum(tot);
    __syncwarp();
    return tot;
  } else {
    unsigned m = rmask;
    while (m) {  // rows one at a time: missed, or assigned to a free column in the mask
      int r = __ffs(m) - 1; m &= m - 1;
      double miss = 1.0 - (double)evalPd[r];
      for (int s = lane; s < S; s += 32) {
        double v = fa[s] * miss;
        for (int k = 0; k < b; k++) if (s & (1 << k)) v += fa[s ^ (1 << k)] * (double)L[r * nZ + small[k]];
        fb[s] = v;
      }
      __syncwarp();
      double* t = fa; fa = fb; fb = t;
    }
    double tot = 0;
    for (int s = lane; s < S; s += 32) {
      double v = fa[s];
      for (int k = 0; k < b; k++) if (!(s & (1 << k))) v *= kappa;
      tot += v;
    }
    tot = warp_sum(tot);
    __syncwarp();
    return tot;
  }
}

// Nijenhuis-Wilf / Gray-code permanent (src/MatrixPermanent.cpp:41-113) of an n x n fp64 matrix in
// shared or global memory, evaluated by one warp: the 2^(n-1) subsets are split across lanes.
__device__ inline double warp_permanent(const double* A, int n, int lane) {
  if (n == 1) return A[0];
  // Glynn-style formula equivalent to the reference's NW walk:
  // perm = 2 * (-1)^n... evaluated as sum over delta in {+-1}^(n-1) (last column sign fixed)
  // x_i(S) = A(i,n-1) - 1/2 sum_j A(i,j) + sum_{j in S} A(i,j);  perm = (-1)^(n-1) * 2 * sum_S (-1)^|S| prod_i x_i(S)
  unsigned long long total = 1ull << (n - 1);
  double acc = 0;
  for (unsigned long long s = lane; s < total; s += 32) {
    double prod = 1;
    for (int i = 0; i < n; i++) {
      double rs = 0, xs = 0;
      for (int j = 0; j < n; j++) {
        double a = A[i * n + j];
        rs += a;
        if (j < n - 1 && ((s >> j) & 1)) xs += a;
      }
      prod *= (A[i * n + n - 1] - 0.5 * rs + xs);
    }
    acc += (__popcll(s) & 1) ? -prod : prod;
  }
  acc = warp_sum(acc);
  double r = 2 * acc;
  if ((n - 1) & 1) r = -r;
  return r;
}

// ------------------------------------------------------------------------------------------------
// S5 helper shared by the 2-D and the Victoria Park kernels: rfsMeasurementLikelihood's partition
// logic (include/RBPHDFilter.hpp:866-994, src/CostMatrix.cpp:92-227) on a likelihood table
// L [nE][nZ] (0 = no edge) held in shared memory.  Returns sum over the visited partitions of
// log(partition likelihood) — the caller subtracts log(clutter integral).  Whole warp; scratch:
// rowmask [MAX_EVAL], compC [MAX_COMP], f0 / f1 [1 << DP_MAXB], compR [MAX_COMP].  Partitions whose smaller side has
// more than dp_onchip members (<= DP_MAXB) are summed in this warp's block of the global workspace gdp
// (2 x (1 << gmaxb) doubles; NULL: such partitions are skipped and the particle is flagged FLAG_DP_OVERFLOW).
template <typename T>
__device__ __forceinline__ double mf_partition_loglik(const T* L, const T* evalPd, int nE, int nZ,
                                                      unsigned long long* rowmask, unsigned long long* compC,
                                                      double* f0, double* f1, unsigned* compR, int sum_method,
                                                      double log_kappa, int& flags, int lane, double* gdp, int gmaxb,
                                                      int dp_onchip, const MurtyOut& mo) {
  double logL = 0;
  if (lane < nE) {
    unsigned long long rm = 0;
    for (int z = 0; z < nZ; z++) if (L[lane * nZ + z] != T(0)) rm |= (1ull << z);
    rowmask[lane] = rm;
  }
  __syncwarp();
  // connected components of the (eval point, measurement) graph, numbered by their lowest vertex
  // (rows first, then columns) like boost::connected_components (src/CostMatrix.cpp:98-109):
  // label propagation, lane = row; a label converges to the smallest vertex of its component.
  int ncc = 0;
  {
    unsigned* lc = reinterpret_cast<unsigned*>(f0);   // [MAX_Z] column labels (f0 is free until the DP)
    for (int z = lane; z < nZ; z += 32) lc[z] = (unsigned)(nE + z);
    const unsigned long long rm = (lane < nE) ? rowmask[lane] : 0ull;
    unsigned lr = (lane < nE) ? (unsigned)lane : 0xffffffffu;
    __syncwarp();
    while (true) {
      bool ch = false;
      unsigned long long m = rm;
      unsigned best = lr;
      while (m) { const int z = __ffsll((long long)m) - 1; m &= m - 1; const unsigned l = lc[z]; best = l < best ? l : best; }
      if (best < lr) { lr = best; ch = true; }
      __syncwarp();
      m = rm;
      while (m) { const int z = __ffsll((long long)m) - 1; m &= m - 1; if (lc[z] > lr) { atomicMin(&lc[z], lr); ch = true; } }
      __syncwarp();
      if (!__any_sync(FULL, ch)) break;
    }
    const unsigned bR = __ballot_sync(FULL, lane < nE && lr == (unsigned)lane);          // row-rooted components
    const unsigned bC0 = __ballot_sync(FULL, lane < nZ && lc[lane < nZ ? lane : 0] == (unsigned)(nE + lane));
    const unsigned bC1 = __ballot_sync(FULL, lane + 32 < nZ && lc[lane + 32 < nZ ? lane + 32 : 0] == (unsigned)(nE + lane + 32));
    const int nRowRoots = __popc(bR);
    ncc = nRowRoots + __popc(bC0) + __popc(bC1);
    // masks of the row-rooted components: lane c takes the c-th root
    {
      const int root = (lane < nRowRoots) ? (int)__fns(bR, 0, lane + 1) : -1;
      unsigned cr = 0;
      for (int e = 0; e < nE; e++) {
        const unsigned le = __shfl_sync(FULL, lr, e);
        if ((int)le == root) cr |= 1u << e;
      }
      unsigned long long cc = 0;
      for (int z = 0; z < nZ; z++) if ((int)lc[z] == root) cc |= 1ull << z;
      if (lane < nRowRoots) { compR[lane] = cr; compC[lane] = cc; }
    }
    // isolated columns are components of their own, after all the row-rooted ones
    if ((bC0 >> lane) & 1u) {
      const int idx = nRowRoots + __popc(bC0 & ((1u << lane) - 1u));
      compR[idx] = 0u; compC[idx] = 1ull << lane;
    }
    if ((bC1 >> lane) & 1u) {
      const int idx = nRowRoots + __popc(bC0) + __popc(bC1 & ((1u << lane) - 1u));
      compR[idx] = 0u; compC[idx] = 1ull << (lane + 32);
    }
  }
  __syncwarp();
  // CostMatrixGeneral::partition (:111-144): one-sided components fold into the first one
  int combinedZero = -1, nMerged = 0;
  unsigned zeroR = 0;
  unsigned long long zeroC = 0;
  {
    int nOne = 0;
    for (int cb = 0; cb < ncc; cb += 32) {
      const int c = cb + lane;
      unsigned cr = 0;
      unsigned long long cc = 0;
      bool one = false;
      if (c < ncc) { cr = compR[c]; cc = compC[c]; one = (cr == 0u) || (cc == 0ull); }
      const unsigned bo = __ballot_sync(FULL, one);
      if (bo && combinedZero < 0) combinedZero = cb + __ffs(bo) - 1;
      nOne += __popc(bo);
      zeroR |= __reduce_or_sync(FULL, one ? cr : 0u);
      zeroC |= (unsigned long long)__reduce_or_sync(FULL, one ? (unsigned)(cc & 0xffffffffull) : 0u) |
               ((unsigned long long)__reduce_or_sync(FULL, one ? (unsigned)(cc >> 32) : 0u) << 32);
    }
    nMerged = nOne > 0 ? nOne - 1 : 0;
  }
  const int nP = ncc - nMerged;
  // Partition likelihoods (Q6: original labels, only p < nP are visited).  Lane-parallel: the
  // zero partition and partitions with a single row or a single column have closed forms;
  // the rest go through the warp-wide subset DP (or the matrix-permanent identity).
  {
    const double kap = exp(log_kappa);
    double mylog = 0;
    for (int pb = 0; pb < nP; pb += 32) {
      const int pp = pb + lane;
      bool hard = false;
      if (pp < nP) {
        if (pp == combinedZero) {   // :891-900 (Q5: Pd, not 1-Pd)
          unsigned m = zeroR;
          while (m) { const int r = __ffs(m) - 1; m &= m - 1; mylog += log((double)evalPd[r]); }
          mylog += (double)__popcll(zeroC) * log_kappa;
        } else {
          const unsigned cr = compR[pp];
          const unsigned long long cc = compC[pp];
          const int nR = __popc(cr), nC = __popcll(cc);
          if (nR + nC > 8) flags |= FLAG_MURTY;
          if (sum_method == 0 && nR == 1) {
            // one eval point: missed (all measurements clutter) or detected by one of them
            const int r = __ffs(cr) - 1;
            double sumL = 0;
            unsigned long long m = cc;
            while (m) { const int c = __ffsll((long long)m) - 1; m &= m - 1; sumL += (double)L[r * nZ + c]; }
            double kpow = 1;   // kappa^(nC-1)
            for (int k = 1; k < nC; k++) kpow *= kap;
            const double miss = 1.0 - (double)evalPd[r];
            mylog += log(nC == 0 ? miss : miss * kpow * kap + kpow * sumL);   // nC == 0: a one-sided component revisited (Q6)
          } else if (sum_method == 0 && nC == 1) {
            // one measurement: clutter (all eval points missed) or it detects one of them
            const int c = __ffsll((long long)cc) - 1;
            double allmiss = 1;
            unsigned m = cr;
            while (m) { const int r = __ffs(m) - 1; m &= m - 1; allmiss *= 1.0 - (double)evalPd[r]; }
            double tot = kap * allmiss;
            m = cr;
            while (m) {
              const int r = __ffs(m) - 1; m &= m - 1;
              double others = 1;
              unsigned m2 = cr & ~(1u << r);
              while (m2) { const int r2 = __ffs(m2) - 1; m2 &= m2 - 1; others *= 1.0 - (double)evalPd[r2]; }
              tot += (double)L[r * nZ + c] * others;
            }
            mylog += log(tot);
          } else {
            hard = true;
          }
        }
      }
      unsigned bh = __ballot_sync(FULL, hard);
      while (bh) {   // warp-wide paths, one partition at a time
        const int hp = pb + __ffs(bh) - 1;
        bh &= bh - 1;
        const unsigned cr = compR[hp];
        const unsigned long long cc = compC[hp];
        const int nR = __popc(cr), nC = __popcll(cc);
        const int bsmall = nR < nC ? nR : nC;
        const bool offchip = bsmall > dp_onchip;
        if (offchip && (gdp == nullptr || bsmall > gmaxb)) { flags |= FLAG_DP_OVERFLOW; continue; }
        double* const d0 = offchip ? gdp : f0;
        double* const d1 = offchip ? gdp + (1 << gmaxb) : f1;
        __syncwarp();
        double plog;
        if (sum_method == 1 && nR + nC <= 11) {
          // MatPerm path: sum = (prod kappa) * perm([[L/kappa, diag(1-Pd)],[ones]]) / nC!
          const int nn = nR + nC;
          double* Am = f0;   // nn*nn <= 121 doubles: f0 alone holds 128 (f0 and f1 are not adjacent in every layout)
          int ridx[12], cidx[12];
          { int k = 0; unsigned m = cr; while (m) { ridx[k++] = __ffs(m) - 1; m &= m - 1; }
            k = 0; unsigned long long mc = cc; while (mc) { cidx[k++] = __ffsll((long long)mc) - 1; mc &= mc - 1; } }
          for (int k = lane; k < nn * nn; k += 32) {
            const int r = k / nn, c = k - r * nn;
            double v;
            if (r < nR) {
              if (c < nC) v = (double)L[ridx[r] * nZ + cidx[c]] / kap;
              else v = (c - nC == r) ? 1.0 - (double)evalPd[ridx[r]] : 0.0;
            } else v = 1.0;
            Am[k] = v;
          }
          __syncwarp();
          const double perm = warp_permanent(Am, nn, lane);
          // The Gray-code / Ryser-type formula (the reference's MatPerm::calc) adds 2^(nn-1) signed products, each
          // bounded by the product of the row sums: with L / kappa spanning many decades the cancellation eats the
          // result.  A-posteriori bound on the relative error; if it is not negligible the partition is summed by
          // the subset DP instead (all terms non-negative, no cancellation).
          double rs = 0;
          if (lane < nn) for (int c = 0; c < nn; c++) rs += fabs(Am[lane * nn + c]);
          double bound = 1;
          for (int r = 0; r < nn; r++) bound *= __shfl_sync(FULL, rs, r);
          const double rel_err = ldexp(1.2e-16, nn - 1) * bound / fabs(perm);
          __syncwarp();
          if (rel_err < 1e-11) {
            double fact = 1; for (int k = 2; k <= nC; k++) fact *= k;
            plog = log(perm / fact) + (double)nC * log_kappa;
          } else {
            plog = log(partition_dp<T>(L, nZ, cr, cc, evalPd, kap, d0, d1, lane));
          }
          __syncwarp();
        } else {
          plog = log(partition_dp<T>(L, nZ, cr, cc, evalPd, kap, d0, d1, lane));
        }
        logL += plog;
        if (mo.buf != nullptr && nR + nC > 8) {   // Murty compatibility: the partition goes to the host
          const unsigned words = 3u + (unsigned)nR + (unsigned)(nR * nC);
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(&mo.count[0], words);
          base = __shfl_sync(FULL, base, 0);
          if (base + words <= mo.cap_words) {
            unsigned long long* rec = mo.buf + base;
            if (lane == 0) {
              rec[0] = (unsigned long long)(unsigned)mo.pi | ((unsigned long long)(unsigned)nR << 32);
              rec[1] = (unsigned long long)(unsigned)nC;
              rec[2] = (unsigned long long)__double_as_longlong(plog);
            }
            for (int k = lane; k < nR; k += 32) {
              const int r = (int)__fns(cr, 0, k + 1);
              rec[3 + k] = (unsigned long long)__double_as_longlong((double)evalPd[r]);
            }
            for (int k = lane; k < nR * nC; k += 32) {
              const int ri = k / nC, ci = k - ri * nC;
              const int r = (int)__fns(cr, 0, ri + 1);
              const unsigned lo = (unsigned)(cc & 0xffffffffull), hi = (unsigned)(cc >> 32);
              const int nlo = __popc(lo);
              const int c = ci < nlo ? (int)__fns(lo, 0, ci + 1) : 32 + (int)__fns(hi, 0, ci - nlo + 1);
              rec[3 + nR + k] = (unsigned long long)__double_as_longlong((double)L[r * nZ + c]);
            }
            __threadfence();
            if (lane == 0) atomicAdd(&mo.count[1], 1u);
          } else {
            flags |= FLAG_MURTY_DROPPED;
          }
        }
      }
    }
    logL += warp_sum(mylog);
    flags = __reduce_or_sync(FULL, (unsigned)flags);
  }
  return logL;
}

// per-particle results into the caller's host buffers (see KParams::w_host); the weight only where no
// normalisation follows in this launch (the fused path stores the normalised weights in the epilogue)
template <typename T>
__device__ __forceinline__ void store_host_results(const KParams<T>& p, int pi, double weight, unsigned long long unused_mask,
                                                   int nfov) {
  if (p.unused_host) p.unused_host[pi] = unused_mask;
  if (p.nfov_host) p.nfov_host[pi] = nfov;
  if (p.w_host && !p.fused_normalize) p.w_host[pi] = weight;
}

// ------------------------------------------------------------------------------------------------
// End of a step, shared by the 2-D and the Victoria Park kernels: per-warp statistics, then S8 —
// deterministic [sum w, sum w^2] by the last CTA (+ the fused cross-GPU sum and normalisation).
template <typename T>
__device__ __forceinline__ void step_epilogue(const KParams<T>& p, int lane, int warp, unsigned long long tot_in,
                                              unsigned long long tot_out, int max_out, int n_over, int n_murty,
                                              int n_fallback, const unsigned (&mstat)[8]) {
  if (lane == 0) {
    if (tot_in) atomicAdd(&p.totals[0], tot_in);
    if (tot_out) atomicAdd(&p.totals[1], tot_out);
    atomicMax(&p.istats[0], max_out);
    if (n_over) atomicAdd(&p.istats[1], n_over);
    if (n_murty) atomicAdd(&p.istats[2], n_murty);
    if (n_fallback) atomicAdd(&p.istats[3], n_fallback);
#pragma unroll
    for (int k = 0; k < 7; k++)
      if (mstat[k]) atomicAdd(&p.mstats[k], mstat[k]);
  }

  // ---------------- S8: deterministic [sum w, sum w^2] by the last CTA ----------------------------
  __shared__ bool is_last;
  __shared__ double red[2][MAX_WARPS_PER_CTA];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(p.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed summation order (bit-reproducible): thread t owns particles t, t + nt, ...  This runs on ONE CTA at the
    // end of every step and is pure L2 latency, so a shard of up to EPI_REG particles per thread is loaded in one
    // round trip and stays in registers for the normalisation below; larger shards stream, four loads in flight.
    constexpr int EPI_REG = 16;
    const bool in_regs = p.N <= EPI_REG * (int)blockDim.x;
    double wreg[EPI_REG];
    double s1, s2;
    if (in_regs) {
      const int nt = blockDim.x;
#pragma unroll
      for (int k = 0; k < EPI_REG; k++) {
        const int i = threadIdx.x + k * nt;
        wreg[k] = i < p.N ? __ldcg(p.w_out + i) : 0.0;
      }
      double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
#pragma unroll
      for (int k = 0; k < EPI_REG; k += 2) {
        a0 += wreg[k]; a1 += wreg[k + 1];
        b0 += wreg[k] * wreg[k]; b1 += wreg[k + 1] * wreg[k + 1];
      }
      s1 = a0 + a1;
      s2 = b0 + b1;
    } else {
      const int nt = blockDim.x;
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
      int i = threadIdx.x;
      for (; i + 3 * nt < p.N; i += 4 * nt) {
        const double w0 = __ldcg(p.w_out + i), w1 = __ldcg(p.w_out + i + nt), w2 = __ldcg(p.w_out + i + 2 * nt),
                     w3 = __ldcg(p.w_out + i + 3 * nt);
        a0 += w0; a1 += w1; a2 += w2; a3 += w3;
        b0 += w0 * w0; b1 += w1 * w1; b2 += w2 * w2; b3 += w3 * w3;
      }
      for (; i < p.N; i += nt) {
        const double w = __ldcg(p.w_out + i);
        a0 += w;
        b0 += w * w;
      }
      s1 = (a0 + a1) + (a2 + a3);
      s2 = (b0 + b1) + (b2 + b3);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    __shared__ double xs[2][8];
    __shared__ int xok;
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); k++) { a += red[0][k]; b += red[1][k]; }
      red[0][0] = a;
      red[1][0] = b;
      xok = 1;
    }
    if (p.comm_world > 1) {
      // ---- fused all-reduce: every rank writes its pair into slot [parity][rank] of EVERY rank's
      // mailbox (remote stores over NVLink), then waits until its own mailbox holds this epoch from
      // all ranks and adds them in rank order — the same bits on every GPU, no extra launch.
      // Slots are double-buffered on the epoch's parity: a rank can be at most one step ahead.
      // Thread r talks to rank r, so the remote stores and the waits of the (up to 8) peers overlap.
      __syncthreads();
      const unsigned long long e = p.comm_epoch;
      const int par = (int)(e & 1ull);
      if ((int)threadIdx.x < p.comm_world) {
        const int r = threadIdx.x;
        CommSlot* dst = reinterpret_cast<CommSlot*>(p.comm_peer[r]) + par * 8 + p.comm_rank;
        comm_send(dst, red[0][0], red[1][0], e);
        if (!p.comm_defer) {
          const CommSlot* mine = reinterpret_cast<const CommSlot*>(p.comm_peer[p.comm_rank]) + par * 8;
          const unsigned long long t0 = globaltimer_ns();
          double a1 = 0, a2 = 0;
          bool ok = true;
          while (!comm_recv(mine + r, e, a1, a2)) {
            if (globaltimer_ns() - t0 > p.comm_timeout_ns) { ok = false; break; }   // a peer never launched
          }
          if (ok) {
            xs[0][r] = a1;
            xs[1][r] = a2;
          } else {
            atomicAnd(&xok, 0);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0 && !p.comm_defer) {   // (deferred: the pair stays the local one; the next launch adds them up)
        double ta = 0, tb = 0;
        for (int r = 0; r < p.comm_world; r++) { ta += xs[0][r]; tb += xs[1][r]; }
        if (!xok) { *p.comm_error = 1; ta = __longlong_as_double(0x7ff8000000000000LL); tb = ta; }
        red[0][0] = ta;
        red[1][0] = tb;
      }
    }
    if (p.fused_normalize && in_regs) {   // ParticleFilter::normalizeWeights, in the same launch
      __syncthreads();
      const double total = red[0][0];
      const int nt = blockDim.x;
#pragma unroll
      for (int k = 0; k < EPI_REG; k++) {
        const int i = threadIdx.x + k * nt;
        if (i < p.N) {
          const double w = wreg[k] / total;
          p.w_out[i] = w;
          if (p.w_host) p.w_host[i] = w;
        }
      }
    } else if (p.fused_normalize) {
      __syncthreads();
      const double total = red[0][0];
      const int nt = blockDim.x;
      int i = threadIdx.x;
      for (; i + 3 * nt < p.N; i += 4 * nt) {
        const double w0 = __ldcg(p.w_out + i) / total, w1 = __ldcg(p.w_out + i + nt) / total,
                     w2 = __ldcg(p.w_out + i + 2 * nt) / total, w3 = __ldcg(p.w_out + i + 3 * nt) / total;
        p.w_out[i] = w0; p.w_out[i + nt] = w1; p.w_out[i + 2 * nt] = w2; p.w_out[i + 3 * nt] = w3;
        if (p.w_host) { p.w_host[i] = w0; p.w_host[i + nt] = w1; p.w_host[i + 2 * nt] = w2; p.w_host[i + 3 * nt] = w3; }
      }
      for (; i < p.N; i += nt) {
        const double w = __ldcg(p.w_out + i) / total;
        p.w_out[i] = w;
        if (p.w_host) p.w_host[i] = w;
      }
    }
    // (after the normalisation, which every thread of the CTA is waiting for)
    if (threadIdx.x == 0) {
      const double a = red[0][0], b = red[1][0];
      p.sums[0] = a;
      p.sums[1] = b;
      // publish the step statistics and re-arm the accumulators / queue for the next launch
      p.stats_out[0] = __ldcg(&p.totals[0]);
      p.stats_out[1] = __ldcg(&p.totals[1]);
      p.stats_out[2] = (unsigned long long)(unsigned)__ldcg(&p.istats[0]);
      p.stats_out[3] = (unsigned long long)(unsigned)__ldcg(&p.istats[1]);
      p.stats_out[4] = (unsigned long long)(unsigned)__ldcg(&p.istats[2]);
      p.stats_out[5] = (unsigned long long)(unsigned)__ldcg(&p.istats[3]);
      for (int k = 0; k < 7; k++) { p.stats_out[6 + k] = (unsigned long long)__ldcg(&p.mstats[k]); p.mstats[k] = 0u; }
      if (p.stats_host) {
        p.stats_host[0] = (unsigned long long)__double_as_longlong(a);
        p.stats_host[1] = (unsigned long long)__double_as_longlong(b);
        for (int k = 0; k < 13; k++) p.stats_host[2 + k] = p.stats_out[k];
      }
      p.totals[0] = 0; p.totals[1] = 0;
      p.istats[0] = 0; p.istats[1] = 0; p.istats[2] = 0; p.istats[3] = 0;
      *p.ticket = 0;
      *p.work_counter = 0;
    }
    if (p.done_host) {
      // host-facing step: every result of the launch is in the caller's memory (the other CTAs' stores are ordered
      // before their tickets, this CTA's before the barrier): tell the host, which polls this word instead of
      // waiting for the stream
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(p.done_host) = p.done_value;
      }
    }
  }
}

// rfsb200_comm_barrier: the ranks of a connected group meet on the stream (one warp, thread r talks to rank r): a
// flag in banks 2 / 3 of every rank's mailbox, written over NVLink peer memory; a rank leaves when it has seen all of them.
struct CommPeers { void* p[8]; };
__global__ void comm_barrier_kernel(const CommPeers peers, int rank, int world, unsigned long long epoch, int* comm_error,
                                    unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  if (r < world) {
    const int bank = 2 + (int)(epoch & 1ull);
    CommSlot* dst = reinterpret_cast<CommSlot*>(peers.p[r]) + bank * 8 + rank;
    comm_send(dst, 0.0, 0.0, epoch);
    const CommSlot* mine = reinterpret_cast<const CommSlot*>(peers.p[rank]) + bank * 8;
    const unsigned long long t0 = globaltimer_ns();
    double a1, a2;
    while (!comm_recv(mine + r, epoch, a1, a2)) {
      if (globaltimer_ns() - t0 > timeout_ns) { *comm_error = 1; break; }
    }
  }
}

// rfsb200_comm_resolve: the normalisation a deferred step (KParams::comm_defer) left open, for a consumer other than
// the next update (the weights are read, resampled, exported ...): one CTA picks up the pairs of that epoch from the
// mailbox, adds them in rank order and divides the weights by the total — the eager epilogue's arithmetic.
__global__ void comm_resolve_kernel(const CommPeers peers, int rank, int world, unsigned long long epoch, double* w, int N,
                                    double* sums, int* comm_error, unsigned long long timeout_ns) {
  __shared__ double xs[2][8];
  if ((int)threadIdx.x < world) {
    const CommSlot* mine = reinterpret_cast<const CommSlot*>(peers.p[rank]) + (int)(epoch & 1ull) * 8;
    const unsigned long long t0 = globaltimer_ns();
    double a1 = 0, a2 = 0;
    while (!comm_recv(mine + threadIdx.x, epoch, a1, a2)) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        *comm_error = 1;
        a1 = a2 = __longlong_as_double(0x7ff8000000000000LL);
        break;
      }
    }
    xs[0][threadIdx.x] = a1;
    xs[1][threadIdx.x] = a2;
  }
  __syncthreads();
  double ta = 0, tb = 0;
  for (int r = 0; r < world; r++) { ta += xs[0][r]; tb += xs[1][r]; }
  for (int i = threadIdx.x; i < N; i += blockDim.x) w[i] = w[i] / ta;
  if (threadIdx.x == 0) { sums[0] = ta; sums[1] = tb; }
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (bytes); host and device agree through these helpers
//  per warp: planes [NPL][W] (NPL = 6, or 7 with weight_prev in multi-feature mode) | work region | aux u32[W] |
//            colsum T[MAX_Z] | evalIdx int[MAX_EVAL] | mbarrier
//  per CTA : the measurement batch + the two window tables of the corrector
// The work region is the merge scratch.  In multi-feature mode the stages of S5 run before the merge and
// reuse the same bytes, one after the other (each stage's data is dead when the next one starts):
//   sort       perm = order[], keys = keys[] (+ a temporary plane behind the keys in the fp64 build)
//   intensity  6 planes T[W] (inverse covariance, log normaliser, two bound coefficients)
//   L table    rowmask | compC | f1 | (f0) | compR | eval-point block | L    (f0 lives in aux[] when W >= 256)
// so the region is max(merge scratch, 6 planes, L-table stage) — with the defaults (W = 256, 15 eval points,
// <= 32 measurements, fp32) 6 kB, 14.5 kB per warp in total: 15 resident warps per SM instead of 8 with a
// separate multi-feature scratch.
template <typename T>
__host__ __device__ inline int mf_stage3_bytes(int W, int n_eval, int zcap) {
  const int ltab = (n_eval * zcap + 3) & ~3;
  const int f0 = (W * 4 >= 8 * (1 << DP_MAXB)) ? 0 : 8 * (1 << DP_MAXB);
  return (int)(8 * (MAX_EVAL + MAX_COMP + (1 << DP_MAXB)) + f0 + 4 * MAX_COMP + sizeof(T) * (MAX_EVAL * 8 + ltab) + 15) & ~15;
}
template <typename T>
__host__ __device__ inline int mf_region_bytes(int W, int n_eval, int zcap) {
  int r = merge_scratch_bytes<T>(W);
  const int a = 6 * W * (int)sizeof(T), b = mf_stage3_bytes<T>(W, n_eval, zcap);
  r = a > r ? a : r;
  r = b > r ? b : r;
  return (r + 15) & ~15;
}
// region_bytes: merge_scratch_bytes (single-cluster) or mf_region_bytes (multi-feature)
template <typename T>
__host__ __device__ inline int warp_bytes_for(int W, int multi_feature, int region_bytes) {
  const int planes = multi_feature ? 7 : 6;   // MF carries weight_prev as a 7th plane
  // colsum[] aliases the merge's cell table (dead outside the merge); evalIdx[] exists in multi-feature mode only
  int b = planes * W * (int)sizeof(T) + region_bytes + W * 4 + (multi_feature ? MAX_EVAL * 4 : 0) + 16;
  return (b + 15) & ~15;
}
constexpr int NBINS = 256;   // bins of the range / bearing window tables
// Z block: (zr,zb) pairs T[2*MAX_Z] | range table u64[NBINS+1] | bearing table u64[NBINS+1] | bin params T[4] | bins u8[2*MAX_Z]
template <typename T>
__host__ __device__ inline int z_bytes() {
  return (int)((2 * MAX_Z * sizeof(T) + 2 * (NBINS + 1) * 8 + 4 * sizeof(T) + 2 * MAX_Z + 127) & ~127);
}

// Host-facing step (rfsb200_update_host with pinned caller buffers): poses, particle weights and pose covariances are read
// from the CALLER's memory by the update kernel itself.  The particles are cut into one slice per CTA; every CTA converts
// its slice when it starts — the reads over PCIe are issued before the measurement tables are built and consumed after,
// so their latency is hidden — and publishes it through hin_ready[slice] = number of the launch.  The particle queue is
// global, so a warp may draw a particle of another CTA's slice: if that flag is not up yet the warp gives the owner 20 us
// (CTAs do not start at the same instant) and then converts the slice itself (a grid that is not fully resident; the CPU
// interpreter of tests/simt, which runs the CTAs one after the other).  Conversions are idempotent — every converter
// writes the same bits — so no claim protocol is needed and nobody can wait forever.
template <typename T>
__device__ __forceinline__ void host_in_store(const KParams<T>& p, int i, double x, double y, double th) {
  p.pose64_w[3 * (size_t)i] = x; p.pose64_w[3 * (size_t)i + 1] = y; p.pose64_w[3 * (size_t)i + 2] = th;
  T* q = p.pose_w + 4 * (size_t)i;
  q[0] = (T)x; q[1] = (T)y; q[2] = (T)th; q[3] = T(0);
}
// particles [lo + first, hi) of slice s (tid strides by nthreads); skip_pose_below: particles below this index already
// have pose and weight (the CTA's early reads) and only need their covariance
template <typename T>
__device__ __forceinline__ void host_in_convert(const KParams<T>& p, int s, int slice, int tid, int nthreads, int skip_pose_below) {
  const int lo = s * slice, hi = (lo + slice < p.N) ? lo + slice : p.N;
  for (int i = lo + tid; i < hi; i += nthreads) {
    if (i >= skip_pose_below) {
      host_in_store(p, i, p.pose_h[3 * (size_t)i], p.pose_h[3 * (size_t)i + 1], p.pose_h[3 * (size_t)i + 2]);
      if (p.weight_h) p.w_front[i] = p.weight_h[i];
    }
    if (p.pcov_h) {
      T* c = p.pcov_w + 8 * (size_t)i;
#pragma unroll
      for (int k = 0; k < 6; k++) c[k] = (T)p.pcov_h[6 * (size_t)i + k];
      c[6] = T(0); c[7] = T(0);
    }
  }
  if (s == 0 && tid == 0 && p.pose_cov_mode == 1) {
#pragma unroll
    for (int k = 0; k < 6; k++) p.pcov_w[k] = (T)p.cov6[k];
    p.pcov_w[6] = T(0); p.pcov_w[7] = T(0);
  }
}
// whole warp, slice s not published when the warp looked: convert it (again) and publish
template <typename T>
__device__ __forceinline__ void host_in_self_service(const KParams<T>& p, int s, int slice, int lane) {
  {  // the owner is usually a few microseconds from publishing (CTAs do not start at the same instant): give it that long
    int up = 0;
    if (lane == 0) {
      const unsigned long long t0 = globaltimer_ns();
      while (!(up = (ld_acquire_gpu_u64(p.hin_ready + s) == p.done_value)) && globaltimer_ns() - t0 < 20000ull) {
      }
    }
    if (__shfl_sync(FULL, up, 0)) return;
  }
  host_in_convert(p, s, slice, lane, 32, 0);
  __threadfence();
  __syncwarp();
  if (lane == 0) st_release_gpu_u64(p.hin_ready + s, p.done_value);
  __syncwarp();
}

// threads per CTA the kernel is compiled for: the fp32 single-cluster kernel fits 96 registers (20 warps per SM),
// the multi-feature and the fp64 kernels need 128 (16 warps)
template <typename T, bool MF, bool PROF = false>
constexpr int update_max_threads() { return (sizeof(T) == 4 && !MF && !PROF) ? 32 * MAX_WARPS_PER_CTA : 512; }

// WT: the work capacity W as a compile-time constant (0: p.W at run time).  With W known, every plane of a warp's
// block is addressed as one register plus an immediate offset; the fp32 kernels are instantiated for W = 256.
template <typename T, bool MF, int WT, bool PROF = false>
__global__ void __launch_bounds__((update_max_threads<T, MF, PROF>()), 1)
phd_update_kernel(const __grid_constant__ KParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int W = WT ? WT : p.W;
  const int nZ = p.nZ;
  constexpr int NPL = MF ? 7 : 6;   // MF = multi-feature weighting (p.use_sc == 0)
  if constexpr (PROF) {
    if (threadIdx.x == 0) atomicMin(&p.prof[12], globaltimer_ns());
  }

  T* zs = reinterpret_cast<T*>(smem_raw);          // [2*MAX_Z]: zr[z] at 2z, zb[z] at 2z+1
  unsigned long long* tabR = reinterpret_cast<unsigned long long*>(zs + 2 * MAX_Z);   // [NBINS+1]
  unsigned long long* tabB = tabR + (NBINS + 1);                                      // [NBINS+1]
  T* binp = reinterpret_cast<T*>(tabB + (NBINS + 1));   // r0, inv_r, b0, inv_b
  unsigned char* zbin = reinterpret_cast<unsigned char*>(binp + 4);                   // [2*MAX_Z]
  unsigned char* wb = smem_raw + z_bytes<T>() + (size_t)warp * p.warp_bytes;
  T* bufA = reinterpret_cast<T*>(wb);
  unsigned char* after = reinterpret_cast<unsigned char*>(bufA + NPL * W);
  unsigned* aux = reinterpret_cast<unsigned*>(after + (MF ? p.mf_bytes : merge_scratch_bytes<T>(W)));  // [W]
  const MergeScratch<T> ms = carve_merge_scratch<T>(after, aux, W);
  T* colsum = reinterpret_cast<T*>(ms.cellStart);             // [MAX_Z] S2 / S3 only (528 bytes; the merge's table is dead)
  int* evalIdx = reinterpret_cast<int*>(aux + W);             // [MAX_EVAL] multi-feature mode only
  uint64_t* bar = reinterpret_cast<uint64_t*>(evalIdx + (MF ? MAX_EVAL : 0));
  unsigned char* mfs = after;   // multi-feature stages reuse the work region (see mf_region_bytes)

  // ---- host-facing step: this CTA's slice of the caller's pinned inputs (see host_in_convert): the reads are issued
  //      here, the values are stored and published behind the table set-up below --------------------------------------
  const int hin_slice = (p.N + (int)gridDim.x - 1) / (int)gridDim.x;
  const int hin_lo = (int)blockIdx.x * hin_slice;
  int hin_early = 0;   // particles of the slice whose pose and weight are read here: flat, fully coalesced 8-byte loads
  double hin_v0 = 0, hin_v1 = 0, hin_w = 0;
  if (p.pose_h) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    int cnt = p.N - hin_lo;
    cnt = cnt < 0 ? 0 : (cnt > hin_slice ? hin_slice : cnt);
    hin_early = cnt < (2 * nt) / 3 ? cnt : (2 * nt) / 3;
    const double* src = p.pose_h + 3 * (size_t)hin_lo;
    if (tid < 3 * hin_early) hin_v0 = src[tid];
    if (tid + nt < 3 * hin_early) hin_v1 = src[tid + nt];
    if (p.weight_h && tid < hin_early) hin_w = p.weight_h[hin_lo + tid];
  }
  // ---- deferred cross-GPU sums of the previous step (KParams::comm_pending): thread r picks up rank r's pair from this
  //      GPU's own mailbox (it arrived while the previous launch was still running or long before this one started)
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
  // ---- the measurement batch and the corrector's window tables (once per CTA) ------------------
  // tabR[b] = set of measurements whose range bin is < b, tabB likewise on the bearing: the
  // measurements with range in [lo,hi] are a subset of tabR[bin(hi)+1] & ~tabR[bin(lo)].
  for (int k = threadIdx.x; k < 2 * nZ; k += blockDim.x) {
    const T v = (T)p.Zval[k];
    zs[k] = v;
    if (blockIdx.x == 0) p.Zdev_w[k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    T lo = M<T>::inf(), hi = -M<T>::inf();
    for (int z = 0; z < nZ; z++) {
      const T v = zs[2 * z + threadIdx.x];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    const T span = hi - lo;
    binp[2 * threadIdx.x] = lo;
    binp[2 * threadIdx.x + 1] = (span > T(0) && span < M<T>::inf()) ? T(NBINS - 0.001) / span : T(0);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * nZ; k += blockDim.x) {
    const int comp = k & 1;
    const T v = (zs[k] - binp[2 * comp]) * binp[2 * comp + 1];
    int b = (v >= T(NBINS - 1)) ? NBINS - 1 : (int)v;
    zbin[k] = (unsigned char)(b < 0 ? 0 : b);
  }
  __syncthreads();
  for (int b = threadIdx.x; b <= NBINS; b += blockDim.x) { tabR[b] = 0ull; tabB[b] = 0ull; }
  __syncthreads();
  // exact bin sets at [bin+1], then an exclusive prefix-OR: tab[b] = measurements with bin < b
  for (int k = threadIdx.x; k < 2 * nZ; k += blockDim.x) {
    unsigned long long* tab = (k & 1) ? tabB : tabR;
    atomicOr(&tab[(int)zbin[k] + 1], 1ull << (k >> 1));
  }
  __syncthreads();
  for (int tsel = warp; tsel < 2; tsel += (int)(blockDim.x >> 5)) {   // one warp per table (a 1-warp CTA does both)
    unsigned long long* tab = tsel ? tabB : tabR;
    unsigned long long loc[8];
    unsigned long long acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { acc |= tab[1 + lane * 8 + k]; loc[k] = acc; }
    unsigned long long run = acc;   // inclusive OR-scan of the lane totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(FULL, run, o);
      if (lane >= o) run |= u;
    }
    unsigned long long before_me = __shfl_up_sync(FULL, run, 1);
    if (lane == 0) before_me = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) tab[1 + lane * 8 + k] = loc[k] | before_me;
  }
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (p.pose_h) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    double* stage = reinterpret_cast<double*>(smem_raw + z_bytes<T>());   // the warps' blocks are not in use yet
    if (tid < 3 * hin_early) stage[tid] = hin_v0;
    if (tid + nt < 3 * hin_early) stage[tid + nt] = hin_v1;
    __syncthreads();
    if (tid < hin_early) {
      host_in_store(p, hin_lo + tid, stage[3 * tid], stage[3 * tid + 1], stage[3 * tid + 2]);
      if (p.weight_h) p.w_front[hin_lo + tid] = hin_w;
    }
    host_in_convert(p, (int)blockIdx.x, hin_slice, tid, nt, hin_lo + hin_early);   // the rest of the slice, covariances
    __threadfence();
  }
  __syncthreads();
  if (p.pose_h && threadIdx.x == 0) st_release_gpu_u64(p.hin_ready + blockIdx.x, p.done_value);
  if constexpr (PROF) {
    if (threadIdx.x == 0) atomicMax(&p.prof[13], globaltimer_ns());
  }
  const T binR0 = binp[0], binRi = binp[1], binB0 = binp[2], binBi = binp[3];
  auto bin_of = [](T v, T v0, T inv) -> int {   // monotone in v; NaN -> 0
    return (int)M<T>::min_(M<T>::max_((v - v0) * inv, T(0)), T(NBINS - 1));
  };
  const T r_hi_in = p.rmax - p.rbuf, r_lo_in = p.rmin + p.rbuf, r_hi_out = p.rmax + p.rbuf, r_lo_out = p.rmin - p.rbuf;

  StageClock<PROF> clk;
  clk.start();
  uint32_t phase = 0;
  unsigned long long tot_in = 0, tot_out = 0;
  int max_out = 0, n_over = 0, n_murty = 0, n_fallback = 0;
  unsigned mstat[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // merge statistics: fallbacks by reason, pairs, clusters

  // Particle queue.  The particles are cut into one slice of S per CTA (the slices of the host-facing step, above).  A
  // warp's FIRST particle is static — particle `warp` of its CTA's slice: no atomic, and in a host-facing step no
  // dependency on another CTA's conversion while the grid is still starting up.  Everything else (the particles
  // warps .. S-1 of every slice) goes through one global queue, one atomic per particle (lane 0): the NEXT particle is
  // drawn while the current one is being merged, and its planes are prefetched into L2 before the current one is stored,
  // so neither the atomic nor the HBM latency of the bulk loads sits on the warp's critical path.
  const int q_first = (int)(blockDim.x >> 5);                          // static particles per slice
  const int q_per_slice = hin_slice > q_first ? hin_slice - q_first : 0;   // queued particles of a full slice
  auto queue_particle = [&](int q) -> int {   // queue ticket -> particle index (>= N: the queue is empty)
    if (q_per_slice == 0) return p.N;
    const int sl = q / q_per_slice;
    const int i = sl * hin_slice + q_first + (q - sl * q_per_slice);
    return (sl < (int)gridDim.x && i < p.N) ? i : p.N;
  };
  int pi = (int)blockIdx.x * hin_slice + warp;
  if (warp >= hin_slice || pi >= p.N) {   // no static particle for this warp (a small shard): straight to the queue
    int q = 0;
    if (lane == 0) q = (int)atomicAdd(p.work_counter, 1u);
    pi = queue_particle(__shfl_sync(FULL, q, 0));
  }
  int hin_ok = 0;   // lane 0: the slice of the particle in hand was published when it was looked up
  if (p.pose_h && lane == 0 && pi < p.N)   // (this CTA's own slice is on the device: the barrier above)
    hin_ok = (pi / hin_slice == (int)blockIdx.x) || ld_acquire_gpu_u64(p.hin_ready + pi / hin_slice) == p.done_value;
  while (pi < p.N) {
    int pi_next = 0;   // lane 0 only until the end of the iteration
    if (p.pose_h) {   // (host-facing step) the inputs of this particle are on the device?  (flag read one particle ahead)
      if (!__shfl_sync(FULL, hin_ok, 0)) host_in_self_service(p, pi / hin_slice, hin_slice, lane);
    }
    double w_prev_particle = p.w_in[pi];
    if (p.comm_pending_scale) {   // ParticleFilter::normalizeWeights of the previous step, applied on the way in (rank order: same bits everywhere)
      double total = 0;
      for (int r = 0; r < p.comm_pending; r++) total += prev_sum[r];
      w_prev_particle = w_prev_particle / total;
    }

    T* cur = bufA;
    int nM = p.cnt_in[pi];
    nM = nM < 0 ? 0 : (nM > p.cap ? p.cap : nM);
    int flags = (p.flags[pi] & (FLAG_BIRTH_OVERFLOW | FLAG_CAND_OVERFLOW)) ? FLAG_OVERFLOW : 0;
    if (nM > W) { nM = W; flags |= FLAG_OVERFLOW; }

    // ---------------- S0: TMA bulk loads -------------------------------------------------
    if (nM > 0) {
      if (lane == 0) {
        uint32_t bytes = (uint32_t)(((nM + 3) & ~3) * sizeof(T));
        fence_proxy_async();
        mbar_expect_tx(bar, 6 * bytes);
        const T* src = p.gm_in + (size_t)pi * 6 * p.cap;
#pragma unroll
        for (int k = 0; k < 6; k++) tma_load_1d(cur + k * W, src + (size_t)k * p.cap, bytes, bar);
      }
    }
    const T px = p.pose[4 * pi], py = p.pose[4 * pi + 1], pth = p.pose[4 * pi + 2];
    T c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
    if (p.pose_cov_mode == 1 && p.pose_h) {   // host-facing step: the shared covariance came by value (its device copy is slice 0's)
      c00 = (T)p.cov6[0]; c01 = (T)p.cov6[1]; c02 = (T)p.cov6[2]; c11 = (T)p.cov6[3]; c12 = (T)p.cov6[4]; c22 = (T)p.cov6[5];
    } else if (p.pose_cov_mode) {
      const T* pc = p.pose_cov + (p.pose_cov_mode == 2 ? (size_t)pi * 8 : 0);
      c00 = pc[0]; c01 = pc[1]; c02 = pc[2]; c11 = pc[3]; c12 = pc[4]; c22 = pc[5];
    }
    if (nM > 0) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }

    clk.mark(STAGE_LOAD);
    // ---------------- S1: corrector --------------------------------------------------------
    int nS = 0;                 // survivors appended so far
    double wsum_d = 0;          // sum of pre-update weights (SC-PHD)
    int nfov = 0;
    bool over = false;
    // S1a: every component — detection probability, missed-detection weight, and a CHEAP conservative
    // test for "some measurement could pass the gate": S_rr <= tr(P) + tr(Sigma_xy) + R_rr and
    // S_bb <= tr(P)/r^2 + tr(Sigma)(1/r^2 + 1) + R_bb for PSD P and Sigma, so a component whose
    // widened windows hold no measurement is finished here.  The others are queued for S1b.
    unsigned short* candIdx = ms.order;   // [W] (free until the merge)
    T* candW = ms.keys;                    // [W] pre-update weight of the queued components
    // components under the sensing-limit heuristic of S4 (few): their indices; aux[m] = (first child << 8) | children
    unsigned short* fixList = reinterpret_cast<unsigned short*>(ms.keys + W);   // [W] (behind candW, inside the work region)
    int ncand = 0, nfix = 0;
    // statistics the merge needs over all n components (gathered here while the data is in registers)
    T g_xmin = M<T>::inf(), g_xmax = -M<T>::inf(), g_trmax = T(0);
    bool g_bad = false;
    const T pc_xy = c00 + c11, pc_all = c00 + c11 + c22;
    const bool pc_ok = (c00 >= T(0)) && (c11 >= T(0)) && (c22 >= T(0));
    for (int base = 0; base < nM; base += 32) {
      const int m = base + lane;
      bool queue = false, fix = false;
      T w = 0;
      if (m < nM) {
        const T x = cur[m], y = cur[W + m];
        const T pxx = cur[2 * W + m], pxy = cur[3 * W + m], pyy = cur[4 * W + m];
        const bool pd = (pxx > T(0)) && (pyy > T(0)) && (pxx * pyy - pxy * pxy > T(0));
        g_bad |= !pd;
        g_xmin = x < g_xmin ? x : g_xmin;
        g_xmax = x > g_xmax ? x : g_xmax;
        g_trmax = M<T>::max_(g_trmax, pxx + pyy);
        w = cur[5 * W + m];
        wsum_d += (double)w;
        const T dx = x - px, dy = y - py;
        const T r2 = dx * dx + dy * dy;
        const T r = M<T>::sqrt_(r2);
        // probabilityOfDetection (src/MeasurementModel_RngBrg.cpp:138-167) + Q2 override
        T Pd;
        bool close = false;
        const bool inrange = (r <= p.rmax) && (r >= p.rmin);
        if (inrange) {
          Pd = p.Pd;
          close = (r >= r_hi_in) || (r <= r_lo_in);
        } else {
          Pd = T(0);
          close = (r <= r_hi_out) && (r >= r_lo_out);
        }
        if (close) Pd = T(1);
        if (Pd != T(0)) nfov++;
        // missed-detection weight (:686-706); the sensing-limit heuristic is patched in S4,
        // which still finds the pre-update weight of the flagged components in the weight plane
        if (MF) cur[6 * W + m] = w;              // weight_prev (only importanceWeighting reads it)
        fix = close && (w > p.birth_w);
        cur[5 * W + m] = fix ? w : (T(1) - Pd) * w;
        if (Pd != T(0) && inrange) {              // measure() returns false outside [rmin,rmax]
          queue = true;
          if (pd && pc_ok) {
            const T zb_hat = wrap_pi<T>(M<T>::atan2_(dy, dx) - pth);
            const T ir2 = T(1) / r2;
            const T tr = pxx + pyy;
            const T srr = (tr + pc_xy + p.R00) * T(1.001);
            const T sbb = (tr * ir2 + pc_all * (ir2 + T(1)) + p.R11) * T(1.001);
            const T dr = M<T>::sqrt_(p.gate2 * srr);
            const T db = M<T>::sqrt_(p.gate2 * sbb);
            const unsigned long long mr = tabR[bin_of(r + dr, binR0, binRi) + 1] & ~tabR[bin_of(r - dr, binR0, binRi)];
            const unsigned long long mb = tabB[bin_of(zb_hat + db, binB0, binBi) + 1] & ~tabB[bin_of(zb_hat - db, binB0, binBi)];
            queue = (mr & mb) != 0ull;
          }
        }
      }
      const unsigned bq = __ballot_sync(FULL, queue);
      if (queue) {
        const int pos = ncand + __popc(bq & ((1u << lane) - 1u));
        candIdx[pos] = (unsigned short)m;
        candW[pos] = w;
      }
      ncand += __popc(bq);
      const unsigned bf = __ballot_sync(FULL, fix);
      if (bf) {
        if (fix) {
          fixList[nfix + __popc(bf & ((1u << lane) - 1u))] = (unsigned short)m;
          aux[m] = 0u;   // no children (S1b fills it in for the queued ones)
        }
        nfix += __popc(bf);
      }
    }
    __syncwarp();
    // S1b: the queued components (ascending m) — EKF innovation, exact gate, posterior Gaussians
    for (int base = 0; base < ncand; base += 32) {
      const int q = base + lane;
      unsigned long long cand = 0;
      int m = 0;
      T x = 0, y = 0, pxx = 0, pxy = 0, pyy = 0;
      T zr_hat = 0, zb_hat = 0, i00 = 0, i01 = 0, i11 = 0, norm = 0, Pdw = 0;
      T hp00 = 0, hp01 = 0, hp10 = 0, hp11 = 0;
      bool fixq = false;
      if (q < ncand) {
        m = candIdx[q];
        const T w = candW[q];
        x = cur[m]; y = cur[W + m];
        pxx = cur[2 * W + m]; pxy = cur[3 * W + m]; pyy = cur[4 * W + m];
        const T dx = x - px, dy = y - py;
        const T r2 = dx * dx + dy * dy;
        const T r = M<T>::sqrt_(r2);
        const bool close = (r >= r_hi_in) || (r <= r_lo_in);   // in range here
        const T Pd = close ? T(1) : p.Pd;
        fixq = close && (w > p.birth_w);
        {
          const T invr = T(1) / r;
          const T c = dx * invr, s = dy * invr;
          const T h10 = -s * invr, h11 = c * invr;   // -dy/r^2, dx/r^2
          zr_hat = r;
          zb_hat = wrap_pi<T>(M<T>::atan2_(dy, dx) - pth);
          hp00 = c * pxx + s * pxy;  hp01 = c * pxy + s * pyy;
          hp10 = h10 * pxx + h11 * pxy;  hp11 = h10 * pxy + h11 * pyy;
          T s00 = hp00 * c + hp01 * s + p.R00;
          T s01 = hp00 * h10 + hp01 * h11 + p.R01;
          T s11 = hp10 * h10 + hp11 * h11 + p.R11;
          if (p.pose_cov_mode) {   // Hx Sigma_x Hx^T, Hx = [[-c,-s,0],[s/r,-c/r,-1]]  (Q1)
            const T g0 = -c, g1 = -s, k0 = s * invr, k1 = -c * invr, k2 = T(-1);
            const T a0 = g0 * c00 + g1 * c01, a1 = g0 * c01 + g1 * c11, a2 = g0 * c02 + g1 * c12;
            const T b0 = k0 * c00 + k1 * c01 + k2 * c02, b1 = k0 * c01 + k1 * c11 + k2 * c12,
                    b2 = k0 * c02 + k1 * c12 + k2 * c22;
            s00 += a0 * g0 + a1 * g1;
            s01 += a0 * k0 + a1 * k1 + a2 * k2;
            s11 += b0 * k0 + b1 * k1 + b2 * k2;
          }
          const T det = s00 * s11 - s01 * s01;
          const T invdet = T(1) / det;
          i00 = s11 * invdet; i01 = -s01 * invdet; i11 = s00 * invdet;
          norm = T(1) / M<T>::sqrt_(M<T>::TWO_PI * M<T>::TWO_PI * det);
          Pdw = Pd * w;
          // candidate measurements: md2 >= nu_r^2/S_rr and md2 >= nu_b^2/S_bb for a PD S, so only the
          // measurements inside both windows can pass the gate (bounds slightly widened against
          // rounding); the windows are looked up in the per-CTA tables.
          if (det > T(0) && s00 > T(0) && s11 > T(0)) {
            const T dr = M<T>::sqrt_(p.gate2 * s00) * T(1.0002);
            const T db = M<T>::sqrt_(p.gate2 * s11) * T(1.0002);
            const unsigned long long mr = tabR[bin_of(zr_hat + dr, binR0, binRi) + 1] & ~tabR[bin_of(zr_hat - dr, binR0, binRi)];
            const unsigned long long mb = tabB[bin_of(zb_hat + db, binB0, binBi) + 1] & ~tabB[bin_of(zb_hat - db, binB0, binBi)];
            cand = mr & mb;
          } else {
            cand = (nZ >= 64) ? ~0ull : ((1ull << nZ) - 1ull);   // degenerate S: test everything
          }
        }
      }
      __syncwarp();
      // full gate on the candidates, ascending z
      unsigned long long mask = 0;
      while (cand) {
        const int z = __ffsll((long long)cand) - 1;
        cand &= cand - 1;
        const T nr = zs[2 * z] - zr_hat;
        if (p.thr_r > T(0) && M<T>::abs_(nr) > p.thr_r) continue;
        const T nb_raw = zs[2 * z + 1] - zb_hat;
        const T nb = wrap_pi<T>(nb_raw);
        if (p.thr_b > T(0) && M<T>::abs_(nb) > p.thr_b) continue;
        // Q3: likelihood and gate use the UNWRAPPED difference
        const T md2 = (nr * i00 + nb_raw * i01) * nr + (nr * i01 + nb_raw * i11) * nb_raw;
        if (md2 > p.gate2) continue;
        const T lik = M<T>::exp_(T(-0.5) * md2) * norm;
        if (!(lik == lik) || lik == T(0)) continue;
        if (!(Pdw * lik > T(0))) continue;
        mask |= (1ull << z);
      }
      const int cnt = __popcll(mask);
      const int incl = warp_incl_scan(cnt, lane);
      int off = nM + nS + incl - cnt;
      nS += __shfl_sync(FULL, incl, 31);
      if (cnt && fixq) {   // children of a component under the sensing-limit heuristic: [off, off + stored)
        const int room = W - off;
        const int stored = room < 0 ? 0 : (cnt < room ? cnt : room);
        aux[m] = ((unsigned)off << 8) | (unsigned)stored;
      }
      if (cnt) {
        // K = P H^T S^-1 ; P+ = sym((I-KH)P)  (include/KalmanFilter.hpp:297-302)
        const T k00 = hp00 * i00 + hp10 * i01, k01 = hp00 * i01 + hp10 * i11;
        const T k10 = hp01 * i00 + hp11 * i01, k11 = hp01 * i01 + hp11 * i11;
        const T n00 = pxx - (k00 * hp00 + k01 * hp10);
        const T n01a = pxy - (k00 * hp01 + k01 * hp11);
        const T n01b = pxy - (k10 * hp00 + k11 * hp10);
        const T n11 = pyy - (k10 * hp01 + k11 * hp11);
        const T n01 = (n01a + n01b) * T(0.5);
        g_bad |= !((n00 > T(0)) && (n11 > T(0)) && (n00 * n11 - n01 * n01 > T(0)));
        g_trmax = M<T>::max_(g_trmax, n00 + n11);
        while (mask) {
          const int z = __ffsll((long long)mask) - 1;
          mask &= mask - 1;
          if (off >= W) { over = true; break; }
          const T nr = zs[2 * z] - zr_hat;
          const T nb_raw = zs[2 * z + 1] - zb_hat;
          const T nb = wrap_pi<T>(nb_raw);
          const T md2 = (nr * i00 + nb_raw * i01) * nr + (nr * i01 + nb_raw * i11) * nb_raw;
          const T lik = M<T>::exp_(T(-0.5) * md2) * norm;
          const T xn = x + (k00 * nr + k01 * nb);
          g_xmin = xn < g_xmin ? xn : g_xmin;
          g_xmax = xn > g_xmax ? xn : g_xmax;
          cur[off] = xn;
          cur[W + off] = y + (k10 * nr + k11 * nb);
          cur[2 * W + off] = n00; cur[3 * W + off] = n01; cur[4 * W + off] = n11;
          cur[5 * W + off] = Pdw * lik;   // un-normalised; divided by the column sum in S3
          if (MF) cur[6 * W + off] = T(0);
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
    g_xmin = warp_min(g_xmin);
    g_xmax = warp_max(g_xmax);
    g_trmax = warp_max(g_trmax);
    g_bad = __any_sync(FULL, g_bad);
    __syncwarp();

    clk.mark(STAGE_CORRECT);
    double weight_new = w_prev_particle;
    unsigned long long unused_mask = 0;
    if (nM == 0) {
      // :559-564 all measurements unused; SC weight untouched (Q10)
      unused_mask = (nZ >= 64) ? ~0ull : ((1ull << nZ) - 1ull);
    } else {
      // ---------------- S2: per-measurement normalisers ------------------------------------
      double ll = 0;
      for (int zb0 = 0; zb0 < nZ; zb0 += 32) {
        const int z = zb0 + lane;
        bool used = false;
        if (z < nZ) {
          T sum = p.kappa;
#pragma unroll 4
          for (int s = nM; s < n; s++) {
            const unsigned u = aux[s];
            if ((int)(u & 0xffu) == z) { sum += cur[5 * W + s]; used = true; }
          }
          colsum[z] = sum;
          ll += (double)M<T>::log_(sum);   // fp32 build: MUFU log (|error| ~1e-7 per term, the weight tolerance is 2e-3)
        }
        const unsigned b = __ballot_sync(FULL, (z < nZ) && !used);
        unused_mask |= ((unsigned long long)b) << zb0;
      }
      ll = warp_sum(ll);
      __syncwarp();
      if (p.use_sc) {  // :661-668 (Q4): exp(sum w) * prod_z(kappa + sum_m W) * w_prev
        weight_new = exp(wsum_d + ll) * w_prev_particle;
      }
      // ---------------- S3: posterior weights of the new Gaussians --------------------------
      for (int s = nM + lane; s < n; s += 32) {
        const unsigned u = aux[s];
        cur[5 * W + s] = cur[5 * W + s] / colsum[u & 0xffu];
      }
      __syncwarp();
      // ---------------- S4: sensing-limit heuristic (:692-703, Q2) --------------------------
      for (int q = lane; q < nfix; q += 32) {   // the flagged components (S1a); their children are contiguous (S1b)
        const int m = fixList[q];
        const unsigned ch = aux[m];
        const int s0 = (int)(ch >> 8), s1 = s0 + (int)(ch & 0xffu);
        const T w_km = cur[5 * W + m];   // still the pre-update weight (see S1)
        T rowsum = T(0);
        for (int s = s0; s < s1; s++) rowsum += cur[5 * W + s];
        const T delta = w_km - rowsum;   // Pd[m] == 1 here
        T w_k = (T(1) - T(1)) * w_km;
        if (delta > T(0)) {
          w_k += delta;
          if (w_k > T(1)) w_k = T(1);
        }
        cur[5 * W + m] = w_k;
      }
      __syncwarp();
    }

    clk.mark(STAGE_WEIGHT);
    // ---------------- S5: multi-feature importance weighting ----------------------------------
    if constexpr (MF) {
      int nEvalCfg = p.n_eval < n ? p.n_eval : n;
      if (nEvalCfg == 0) {
        weight_new = 4.9406564584124654e-324;  // denorm_min (:742-745, Q10)
      } else {
        // sortByWeight (:746): weight descending, ties by position.  Keys are sorted in the merge
        // scratch, then every plane is permuted through one temporary plane.
        {
          unsigned short* perm = ms.order;   // [W] sorted position -> current position
          if constexpr (sizeof(T) == 4) {
            unsigned long long* k64 = reinterpret_cast<unsigned long long*>(ms.keys);
            const int P = next_pow2(n);
            for (int k = lane; k < P; k += 32) {
              // weights are >= 0 here: the bit pattern orders like the value; low word = ~position
              k64[k] = (k < n) ? (((unsigned long long)__float_as_uint((float)cur[5 * W + k]) << 32) |
                                  (unsigned long long)(0xffffffffu - (unsigned)k))
                               : 0ull;
            }
            __syncwarp();
            warp_sort_desc64(k64, P, lane);
            for (int k = lane; k < n; k += 32) perm[k] = (unsigned short)(0xffffffffu - (unsigned)(k64[k] & 0xffffffffull));
          } else {
            T* kw = ms.keys;
            const int P = next_pow2(n);
            for (int k = lane; k < P; k += 32) {
              kw[k] = (k < n) ? cur[5 * W + k] : -M<T>::inf();
              aux[k] = (unsigned)k;
            }
            __syncwarp();
            warp_bitonic(kw, aux, P, lane);
            for (int k = lane; k < n; k += 32) perm[k] = (unsigned short)aux[k];
          }
          __syncwarp();
          T* tmp = reinterpret_cast<T*>(aux);   // [W] words; T = double uses mf scratch instead
          if constexpr (sizeof(T) == 8) tmp = reinterpret_cast<T*>(ms.keys) + W;   // behind the keys, inside the merge scratch
          if constexpr (WT != 0 && WT <= 256 && sizeof(T) == 4) {
            // compile-time capacity: a lane's (at most 8) source positions and the values of one plane live in
            // registers — gather, barrier, store; no temporary plane, the permutation is read once for all planes
            constexpr int KPL = WT / 32;
            int pk[KPL];
#pragma unroll
            for (int j = 0; j < KPL; j++) pk[j] = (lane + 32 * j < n) ? (int)perm[lane + 32 * j] : 0;
            for (int pl = 0; pl < 7; pl++) {
              T v[KPL];
#pragma unroll
              for (int j = 0; j < KPL; j++) v[j] = cur[pl * W + pk[j]];
              __syncwarp();
#pragma unroll
              for (int j = 0; j < KPL; j++)
                if (lane + 32 * j < n) cur[pl * W + lane + 32 * j] = v[j];
              __syncwarp();
            }
          } else {
            for (int pl = 0; pl < 7; pl++) {
              for (int k = lane; k < n; k += 32) tmp[k] = cur[pl * W + perm[k]];
              __syncwarp();
              for (int k = lane; k < n; k += 32) cur[pl * W + k] = tmp[k];
              __syncwarp();
            }
          }
        }
        // eval points (:747-762): sorted order, w >= min weight, raw Pd > 0, first nEvalCfg
        int nE = 0;
        for (int base = 0; base < n && nE < nEvalCfg; base += 32) {
          const int m = base + lane;
          bool elig = false;
          bool heavy = false;
          if (m < n) {
            heavy = !(cur[5 * W + m] < p.eval_min_w);
            if (heavy) {
              const T dx = cur[m] - px, dy = cur[W + m] - py;
              const T r = M<T>::sqrt_(dx * dx + dy * dy);
              elig = (r <= p.rmax) && (r >= p.rmin) && (p.Pd > T(0));
            }
          }
          const unsigned be = __ballot_sync(FULL, elig);
          const int rank = nE + __popc(be & ((1u << lane) - 1u));
          if (elig && rank < nEvalCfg) evalIdx[rank] = m;
          nE += __popc(be);
          if (!__all_sync(FULL, heavy)) break;
        }
        if (nE > nEvalCfg) nE = nEvalCfg;
        __syncwarp();
        // sums of weights (:765-773)
        double sw_prev = 0, sw_now = 0;
        for (int m = lane; m < n; m += 32) { sw_prev += (double)cur[6 * W + m]; sw_now += (double)cur[5 * W + m]; }
        sw_prev = warp_sum(sw_prev);
        sw_now = warp_sum(sw_now);
        // intensity at the eval points before / after the update (:776-800):
        //   v(e) = denorm_min + sum_m w_m N(x_e; x_m, P_m), once with the previous and once with the new weights.
        // Per-component inverse covariance, log normaliser and two bound coefficients are precomputed (6 planes in
        // the work region); then ONE LANE PER EVAL POINT (two lanes, splitting the components, when there are
        // at most 16 eval points) runs an online log-sum-exp over the components, whose data are
        // broadcast reads.
        T* ia = reinterpret_cast<T*>(mfs);   // [6][W] (the sort's perm / keys are dead)
        for (int m = lane; m < n; m += 32) {
          const T a = cur[2 * W + m], b = cur[3 * W + m], c = cur[4 * W + m];
          const T det = a * c - b * b;
          const T invdet = T(1) / det;
          ia[m] = c * invdet; ia[W + m] = -b * invdet; ia[2 * W + m] = a * invdet;
          ia[3 * W + m] = M<T>::log_(M<T>::TWO_PI * M<T>::sqrt_(det));
          // |d|^2 / lambda_max <= md2 <= |d|^2 / lambda_min with lambda_max <= trace and lambda_min >= det / trace for
          // a PD covariance: two coefficients that bound the exponent of the component at distance |d| from above
          // and from below (0.1 % slack against rounding); 0 / inf = no bound
          const bool pd = (a > T(0)) && (c > T(0)) && (det > T(0));
          ia[4 * W + m] = pd ? T(0.4995) / (a + c) : T(0);
          ia[5 * W + m] = pd ? T(0.5005) * (a + c) * invdet : M<T>::inf();
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
          // pass 1 (cheap): a LOWER bound on the largest exponent of each sum, from the lower bound of every
          // component's exponent; pass 2 skips the components whose UPPER bound is more than CUT below it — they
          // could not contribute to either sum — before touching their inverse covariance or an exp
          const int ei = (e < nE) ? evalIdx[e] : 0;
          const T xe = cur[ei], ye = cur[W + ei];
          const int half = (((n + groups - 1) / groups) + 3) & ~3;   // (a multiple of 4: the fp32 loops read 16 bytes at a time)
          const int m0 = h * half, m1 = (m0 + half < n) ? m0 + half : n;
          T lob = -M<T>::inf(), loa = -M<T>::inf();
          if (e < nE) {
            if constexpr (sizeof(T) == 4) {
              // the component data are the same for every lane (eval point): 16-byte broadcast loads, four components each
              for (int m = m0; m < m1; m += 4) {
                const Quad X = *reinterpret_cast<const Quad*>(cur + m), Y = *reinterpret_cast<const Quad*>(cur + W + m);
                const Quad A5 = *reinterpret_cast<const Quad*>(ia + 5 * W + m), A3 = *reinterpret_cast<const Quad*>(ia + 3 * W + m);
                const Quad WP = *reinterpret_cast<const Quad*>(cur + 6 * W + m), WN = *reinterpret_cast<const Quad*>(cur + 5 * W + m);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const T dx = xe - X.v[k], dy = ye - Y.v[k];
                  const T l = -A5.v[k] * (dx * dx + dy * dy) - A3.v[k];
                  const bool in = m + k < m1;
                  if (in && WP.v[k] > T(0) && l > lob) lob = l;
                  if (in && WN.v[k] > T(0) && l > loa) loa = l;
                }
              }
            } else {
              for (int m = m0; m < m1; m++) {
                const T dx = xe - cur[m], dy = ye - cur[W + m];
                const T l = -ia[5 * W + m] * (dx * dx + dy * dy) - ia[3 * W + m];
                if (cur[6 * W + m] > T(0) && l > lob) lob = l;
                if (cur[5 * W + m] > T(0) && l > loa) loa = l;
              }
            }
          }
          if (groups == 2) {
            const T ob = __shfl_xor_sync(FULL, lob, 16), oa = __shfl_xor_sync(FULL, loa, 16);
            lob = ob > lob ? ob : lob;
            loa = oa > loa ? oa : loa;
          }
          const T skip_below = (lob < loa ? lob : loa) - CUT;   // -inf (no skipping) while a sum has no term at all
          auto add_component = [&](int m, T dx, T dy, T lnm) {
            const T i01 = ia[W + m];
            const T md2 = (dx * ia[m] + dy * i01) * dx + (dx * i01 + dy * ia[2 * W + m]) * dy;
            const T t = T(-0.5) * md2 - lnm;
            const T wp = cur[6 * W + m], wn = cur[5 * W + m];
            if (wp > T(0)) {
              if (t > mb) { sb = sb * M<T>::exp_(mb - t) + wp; mb = t; }
              else if (t > mb - CUT) sb += wp * M<T>::exp_(t - mb);
            }
            if (wn > T(0)) {
              if (t > ma) { sa = sa * M<T>::exp_(ma - t) + wn; ma = t; }
              else if (t > ma - CUT) sa += wn * M<T>::exp_(t - ma);
            }
          };
          if (e < nE) {
            if constexpr (sizeof(T) == 4) {
              for (int m = m0; m < m1; m += 4) {
                const Quad X = *reinterpret_cast<const Quad*>(cur + m), Y = *reinterpret_cast<const Quad*>(cur + W + m);
                const Quad A4 = *reinterpret_cast<const Quad*>(ia + 4 * W + m), A3 = *reinterpret_cast<const Quad*>(ia + 3 * W + m);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const T dx = xe - X.v[k], dy = ye - Y.v[k];
                  const T lnm = A3.v[k];
                  if (m + k < m1 && !(-A4.v[k] * (dx * dx + dy * dy) - lnm <= skip_below)) add_component(m + k, dx, dy, lnm);
                }
              }
            } else {
              for (int m = m0; m < m1; m++) {
                const T dx = xe - cur[m], dy = ye - cur[W + m];
                const T lnm = ia[3 * W + m];
                if (-ia[4 * W + m] * (dx * dx + dy * dy) - lnm <= skip_below) continue;
                add_component(m, dx, dy, lnm);
              }
            }
          }
          if (groups == 2) {   // combine the two halves of each eval point (lanes e and e + 16)
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
        // rfsMeasurementLikelihood (:821-997): L table with the landmark covariance zeroed
        // (the intensity planes are dead: the L-table stage takes the region over)
        unsigned long long* rowmask = reinterpret_cast<unsigned long long*>(mfs);   // [MAX_EVAL]
        unsigned long long* compC = rowmask + MAX_EVAL;                             // [MAX_COMP]
        double* f1 = reinterpret_cast<double*>(compC + MAX_COMP);                   // [1<<DP_MAXB]
        const bool f0_in_aux = W * 4 >= 8 * (1 << DP_MAXB);
        double* f0 = f0_in_aux ? reinterpret_cast<double*>(aux) : f1 + (1 << DP_MAXB);   // [1<<DP_MAXB]
        unsigned* compR = reinterpret_cast<unsigned*>((f0_in_aux ? f1 : f0) + (1 << DP_MAXB));   // [MAX_COMP]
        T* ep = reinterpret_cast<T*>(compR + MAX_COMP);   // [MAX_EVAL][8]: zr, zb, i00, i01, i11, Pd*norm, -, Pd
        T* L = ep + MAX_EVAL * 8;                         // [nE][nZ]
        T* evalPd = ep + MAX_EVAL * 7;  // stride-1 array of Pd per eval point (slot 7 of the ep block region)
        if (lane < nE) {
          const int ei = evalIdx[lane];
          const T dx = cur[ei] - px, dy = cur[W + ei] - py;
          const T r2 = dx * dx + dy * dy;
          const T r = M<T>::sqrt_(r2);
          const T invr = T(1) / r;
          const T c = dx * invr, s = dy * invr;
          T s00 = p.R00, s01 = p.R01, s11 = p.R11;
          if (p.pose_cov_mode) {
            const T g0 = -c, g1 = -s, k0 = s * invr, k1 = -c * invr, k2 = T(-1);
            const T a0 = g0 * c00 + g1 * c01, a1 = g0 * c01 + g1 * c11, a2 = g0 * c02 + g1 * c12;
            const T b0 = k0 * c00 + k1 * c01 + k2 * c02, b1 = k0 * c01 + k1 * c11 + k2 * c12,
                    b2 = k0 * c02 + k1 * c12 + k2 * c22;
            s00 += a0 * g0 + a1 * g1;
            s01 += a0 * k0 + a1 * k1 + a2 * k2;
            s11 += b0 * k0 + b1 * k1 + b2 * k2;
          }
          const T det = s00 * s11 - s01 * s01;
          const T invdet = T(1) / det;
          ep[lane * 7 + 0] = r;
          ep[lane * 7 + 1] = wrap_pi<T>(M<T>::atan2_(dy, dx) - pth);
          ep[lane * 7 + 2] = s11 * invdet;
          ep[lane * 7 + 3] = -s01 * invdet;
          ep[lane * 7 + 4] = s00 * invdet;
          ep[lane * 7 + 5] = p.Pd / M<T>::sqrt_(M<T>::TWO_PI * M<T>::TWO_PI * det);
          evalPd[lane] = p.Pd;   // raw model Pd of an in-range landmark (Q12)
        }
        __syncwarp();
        for (int k = lane; k < nE * nZ; k += 32) {
          const int e = k / nZ, z = k - e * nZ;
          const T nr = zs[2 * z] - ep[e * 7], nb = zs[2 * z + 1] - ep[e * 7 + 1];
          const T j00 = ep[e * 7 + 2], j01 = ep[e * 7 + 3], j11 = ep[e * 7 + 4];
          const T md2 = (nr * j00 + nb * j01) * nr + (nr * j01 + nb * j11) * nb;
          T l = M<T>::exp_(T(-0.5) * md2) * ep[e * 7 + 5];
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
        // :808-812
        weight_new = exp(logL + (lp_before - lp_after) + (sw_now - sw_prev)) * w_prev_particle;
        __syncwarp();
      }
    }

    clk.mark(STAGE_MFWEIGHT);
    // ---------------- S6: merge ------------------------------------------------------------------
    if (lane == 0) pi_next = queue_particle((int)atomicAdd(p.work_counter, 1u));   // first needed after the merge
    if (n > 1) {
      int st = MERGE_FALLBACK;
      if (p.merge_algo != 0) st = merge_clustered<T, PROF>(cur, ms, W, n, p.merge_t2, p.merge_f, MF, lane, mstat, g_xmin, g_xmax, g_trmax, g_bad, clk);
      __syncwarp();
      if (st == MERGE_FALLBACK && p.merge_algo != 0) n_fallback++;
      if (st == MERGE_FALLBACK) merge_bruteforce<T>(cur, W, n, p.merge_t2, p.merge_f, lane);
    }
    __syncwarp();

    clk.mark(STAGE_MERGE);
    // ---------------- S7: prune + store ------------------------------------------------------------
    int nM_next = 0;
    if (lane == 0 && pi_next < p.N) {
      nM_next = p.cnt_in[pi_next];   // first needed after the sort
      if (p.pose_h) hin_ok = ld_acquire_gpu_u64(p.hin_ready + pi_next / hin_slice) == p.done_value;
    }
    int n_out = 0;
    {
      T* kw = ms.keys;   // [W] sort keys (the merge-only scratch is dead)
      unsigned short* sorted = ms.order;   // [W] component index by output position
      if constexpr (sizeof(T) == 4) {
        // fp32: weights are positive, so (weight bits, ~index) packs into one u64 whose descending
        // order is "weight descending, then index ascending"
        unsigned long long* k64 = reinterpret_cast<unsigned long long*>(ms.keys);   // [W] fits: see merge_only_bytes
        for (int base = 0; base < n; base += 32) {
          const int k = base + lane;
          bool keep = false;
          T w = 0;
          if (k < n) { w = cur[5 * W + k]; keep = (w >= p.prune_t) && (w >= T(0)); }
          const unsigned b = __ballot_sync(FULL, keep);
          if (keep) {
            const int pos = n_out + __popc(b & ((1u << lane) - 1u));
            k64[pos] = ((unsigned long long)__float_as_uint((float)w) << 32) | (unsigned long long)(0xffffffffu - (unsigned)k);
          }
          n_out += __popc(b);
        }
        __syncwarp();
        if (n_out <= 64) {
          // rank sort: output position = number of larger keys (keys are distinct)
          const unsigned long long k0 = lane < n_out ? k64[lane] : 0ull;
          const unsigned long long k1 = lane + 32 < n_out ? k64[lane + 32] : 0ull;
          int r0 = 0, r1 = 0;
          {
            // on the weights alone first (32-bit compares); equal weights show up as colliding ranks
            const unsigned* kw32 = reinterpret_cast<const unsigned*>(k64);
            const unsigned w0 = (unsigned)(k0 >> 32), w1 = (unsigned)(k1 >> 32);
            if (n_out <= 32) {
              for (int q = 0; q < n_out; q++) r0 += (kw32[2 * q + 1] > w0) ? 1 : 0;
            } else {
              for (int q = 0; q < n_out; q++) {
                const unsigned wq = kw32[2 * q + 1];
                r0 += (wq > w0) ? 1 : 0;
                r1 += (wq > w1) ? 1 : 0;
              }
            }
            const unsigned long long mine = (lane < n_out ? 1ull << r0 : 0ull) | (lane + 32 < n_out ? 1ull << r1 : 0ull);
            const unsigned long long all = (unsigned long long)__reduce_or_sync(FULL, (unsigned)mine) |
                                           ((unsigned long long)__reduce_or_sync(FULL, (unsigned)(mine >> 32)) << 32);
            const unsigned long long want = n_out == 64 ? ~0ull : (1ull << n_out) - 1ull;
            if (all != want) {   // ties: the full (weight, position) keys decide
              r0 = 0; r1 = 0;
              for (int q = 0; q < n_out; q++) {
                const unsigned long long kq = k64[q];
                r0 += (kq > k0) ? 1 : 0;
                r1 += (kq > k1) ? 1 : 0;
              }
            }
          }
          if (lane < n_out) sorted[r0] = (unsigned short)(0xffffffffu - (unsigned)(k0 & 0xffffffffull));
          if (lane + 32 < n_out) sorted[r1] = (unsigned short)(0xffffffffu - (unsigned)(k1 & 0xffffffffull));
        } else {
          const int P = next_pow2(n_out);
          for (int k = n_out + lane; k < P; k += 32) k64[k] = 0ull;   // below every real key
          __syncwarp();
          warp_sort_desc64(k64, P, lane);
          for (int k = lane; k < n_out; k += 32) sorted[k] = (unsigned short)(0xffffffffu - (unsigned)(k64[k] & 0xffffffffull));
        }
      } else {
        for (int base = 0; base < n; base += 32) {
          const int k = base + lane;
          bool keep = false;
          T w = 0;
          if (k < n) { w = cur[5 * W + k]; keep = (w >= p.prune_t) && (w >= T(0)); }
          const unsigned b = __ballot_sync(FULL, keep);
          if (keep) {
            const int pos = n_out + __popc(b & ((1u << lane) - 1u));
            kw[pos] = w;
            aux[pos] = (unsigned)k;
          }
          n_out += __popc(b);
        }
        const int P = next_pow2(n_out);
        for (int k = n_out + lane; k < P; k += 32) { kw[k] = -M<T>::inf(); aux[k] = 0xffffffffu; }
        __syncwarp();
        if (n_out > 1) warp_bitonic(kw, aux, P, lane);
        for (int k = lane; k < n_out; k += 32) sorted[k] = (unsigned short)aux[k];
      }
      __syncwarp();
      if (n_out > p.cap) { n_out = p.cap; flags |= FLAG_OVERFLOW; }
      if (lane == 0 && nM_next > 0) {   // the next particle's planes on their way into L2
        nM_next = nM_next > p.cap ? p.cap : nM_next;
        const uint32_t bytes = (uint32_t)(((nM_next + 3) & ~3) * sizeof(T));
        const T* src = p.gm_in + (size_t)pi_next * 6 * p.cap;
#pragma unroll
        for (int k = 0; k < 6; k++) tma_prefetch_l2(src + (size_t)k * p.cap, bytes);
      }
      // gather from shared memory, 128-byte coalesced stores to the particle's planes in HBM
      T* dst = p.gm_out + (size_t)pi * 6 * p.cap;
      for (int k = lane; k < n_out; k += 32) {
        const unsigned src = sorted[k];
#pragma unroll
        for (int pl = 0; pl < 6; pl++) dst[(size_t)pl * p.cap + k] = cur[pl * W + src];
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
    pi = __shfl_sync(FULL, pi_next, 0);
    clk.mark(STAGE_PRUNE);
  }
  clk.flush(p.prof, lane);
  step_epilogue<T>(p, lane, warp, tot_in, tot_out, max_out, n_over, n_murty, n_fallback, mstat);
  if constexpr (PROF) {
    if (threadIdx.x == 0) atomicMax(&p.prof[15], globaltimer_ns());
  }
}

// ------------------------------------------------------------------------------------------------
// Map part of RBPHDFilter::predict() (include/RBPHDFilter.hpp:415-442), one warp per particle, in place:
// births from the unused measurements of the last update (descending index, as the reference pops
// them from the back of unused_measurements_, :1013-1052) at inverseMeasure(pose, z)
// (src/MeasurementModel_RngBrg.cpp:117-136), then P += Q on everything (include/ProcessModel.hpp:195-219).
template <typename T>
struct PredictParams {
  T* gm; int* cnt; unsigned long long* unused; int* flags;
  const T* pose; const T* Z;
  int N, cap, nZ, add_births, add_q;
  T R00, R01, R10, R11, q00, q01, q11, birth_w;
};

template <typename T>
__global__ void predict_maps_kernel(const PredictParams<T> p) {
  const int lane = threadIdx.x & 31;
  const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pi >= p.N) return;
  T* g = p.gm + (size_t)pi * 6 * p.cap;
  int n = p.cnt[pi];
  n = n < 0 ? 0 : (n > p.cap ? p.cap : n);
  if (p.add_births) {
    unsigned long long mask = p.unused[pi];
    if (p.nZ < 64) mask &= (1ull << p.nZ) - 1ull;
    const int nb = __popcll(mask);
    const T px = p.pose[4 * pi], py = p.pose[4 * pi + 1], pth = p.pose[4 * pi + 2];
    bool over = false;
    for (int k = lane; k < nb; k += 32) {
      unsigned long long m = mask;   // k-th highest set bit
      for (int s = 0; s < k; s++) m &= ~(1ull << (63 - __clzll((long long)m)));
      const int z = 63 - __clzll((long long)m);
      const T r = p.Z[2 * z], b = p.Z[2 * z + 1];
      T sn, cs;
      if constexpr (sizeof(T) == 4) sincosf(pth + b, &sn, &cs); else sincos(pth + b, &sn, &cs);
      // Hinv = [[c, -r s], [s, r c]] ; cov = Hinv R Hinv^T
      const T h00 = cs, h01 = -r * sn, h10 = sn, h11 = r * cs;
      const T a00 = h00 * p.R00 + h01 * p.R10, a01 = h00 * p.R01 + h01 * p.R11;
      const T a10 = h10 * p.R00 + h11 * p.R10, a11 = h10 * p.R01 + h11 * p.R11;
      const int idx = n + k;
      if (idx < p.cap) {
        g[idx] = px + r * cs;
        g[p.cap + idx] = py + r * sn;
        g[2 * p.cap + idx] = a00 * h00 + a01 * h01;
        g[3 * p.cap + idx] = a00 * h10 + a01 * h11;
        g[4 * p.cap + idx] = a10 * h10 + a11 * h11;
        g[5 * p.cap + idx] = p.birth_w;
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
      g[2 * p.cap + j] += p.q00;
      g[3 * p.cap + j] += p.q01;
      g[4 * p.cap + j] += p.q11;
    }
  }
}

// Data movement of ParticleFilter::resample(): new particle i <- map of src[i], aux of asrc[i].
template <typename T>
__global__ void resample_gather_kernel(const T* __restrict__ gm_in, const int* __restrict__ cnt_in,
                                       const unsigned long long* __restrict__ unused_in, const int* __restrict__ nfov_in,
                                       const double* __restrict__ w_in, const int* __restrict__ src,
                                       const int* __restrict__ asrc, T* __restrict__ gm_out, int* __restrict__ cnt_out,
                                       unsigned long long* __restrict__ unused_out, int* __restrict__ nfov_out,
                                       double* __restrict__ w_out, int set_w, double w_value,
                                       const T* __restrict__ pose_in, const T* __restrict__ pcov_in,
                                       T* __restrict__ pose_out, T* __restrict__ pcov_out, int pose_cov_mode,
                                       const double* __restrict__ pose64_in, double* __restrict__ pose64_out,
                                       int N, int cap, int npl) {
  const int lane = threadIdx.x & 31;
  const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pi >= N) return;
  int s = src[pi];
  s = s < 0 ? 0 : (s >= N ? N - 1 : s);
  int a = asrc ? asrc[pi] : s;   // -1: the new particle starts with no unused measurements
  a = a >= N ? N - 1 : a;
  const int n = cnt_in[s];
  const T* gi = gm_in + (size_t)s * npl * cap;
  T* go = gm_out + (size_t)pi * npl * cap;
  for (int pl = 0; pl < npl; pl++)
    for (int j = lane; j < n; j += 32) go[(size_t)pl * cap + j] = gi[(size_t)pl * cap + j];
  if (lane == 0) {
    cnt_out[pi] = n;
    unused_out[pi] = a >= 0 ? unused_in[a] : 0ull;
    nfov_out[pi] = nfov_in[pi];   // nLandmarksInFOV_ stays with the slot in the reference
    w_out[pi] = set_w ? w_value : w_in[s];
  }
  // the copy carries the pose of its source (Particle::copy): the next predict's births use it
  if (lane < 4) pose_out[4 * pi + lane] = pose_in[4 * s + lane];
  if (lane < 3) pose64_out[3 * pi + lane] = pose64_in[3 * s + lane];   // fp64 copy used by rfsb200_propagate
  if (pose_cov_mode == 2 && lane < 8) pcov_out[8 * pi + lane] = pcov_in[8 * s + lane];
}

// ------------------------------------------------------------------------------------------------
// ParticleFilter::propagate(): ProcessModel::sample() for every particle (include/ProcessModel.hpp:125-150).
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so particle i of step s always sees the same numbers.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two independent N(0,1) from two 32-bit words (Box-Muller on uniforms in (0,1))
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, double& n0, double& n1) {
  const double u1 = ((double)a + 0.5) * (1.0 / 4294967296.0), u2 = ((double)b + 0.5) * (1.0 / 4294967296.0);
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  n0 = r * cs;
  n1 = r * sn;
}

struct MotionParams {
  int model_id, use_model_noise, use_input_noise, n_in;
  double LQ[9];      // lower Cholesky factor of Q (row-major)
  double Lu[9];      // lower Cholesky factor of the input covariance (row-major 3x3, zero padded)
  double u[3];
  double dt, h, l, pdx, pdy;
  unsigned long long seed, step;
};

template <typename T>
__global__ void propagate_kernel(double* __restrict__ pose64, T* __restrict__ pose, const MotionParams m, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double x = pose64[3 * i], y = pose64[3 * i + 1], th = pose64[3 * i + 2];
  double z[6] = {0, 0, 0, 0, 0, 0};
  if (m.use_model_noise || m.use_input_noise) {
    unsigned r[4], q[4];
    philox4x32_10((unsigned)i, (unsigned)m.step, (unsigned)(m.step >> 32), 0u, (unsigned)m.seed, (unsigned)(m.seed >> 32), r);
    philox4x32_10((unsigned)i, (unsigned)m.step, (unsigned)(m.step >> 32), 1u, (unsigned)m.seed, (unsigned)(m.seed >> 32), q);
    box_muller(r[0], r[1], z[0], z[1]);
    box_muller(r[2], r[3], z[2], z[3]);
    box_muller(q[0], q[1], z[4], z[5]);
  }
  double u0 = m.u[0], u1 = m.u[1], u2 = m.u[2];
  if (m.use_input_noise) {   // input_k.sample(in): u + L xi (include/RandomVec.hpp:457-473)
    u0 += m.Lu[0] * z[0];
    u1 += m.Lu[3] * z[0] + m.Lu[4] * z[1];
    if (m.n_in == 3) u2 += m.Lu[6] * z[0] + m.Lu[7] * z[1] + m.Lu[8] * z[2];
  }
  if (m.model_id == 1) {
    // MotionModel_Odometry2d::step (src/ProcessModel_Odometry2D.cpp:41-89): p += C(theta)^T dp ; C_k = C(dtheta) C(theta)
    double st, ct, sd, cd;
    sincos(th, &st, &ct);
    sincos(u2, &sd, &cd);
    x += ct * u0 - st * u1;
    y += st * u0 + ct * u1;
    const double c00 = cd * ct - sd * st, c01 = cd * st + sd * ct;   // first row of C(dtheta) C(theta)
    th = atan2(c01, c00);
  } else {
    // MotionModel_Ackerman2d::step (src/ProcessModel_Ackerman2D.cpp:49-77)
    double sr, cr;
    sincos(th, &sr, &cr);
    const double tu = tan(u1);
    const double v = u0 / (1 - tu * m.h / m.l);
    x += m.dt * (v * cr - v / m.l * tu * (m.pdx * sr + m.pdy * cr));
    y += m.dt * (v * sr + v / m.l * tu * (m.pdx * cr - m.pdy * sr));
    th += m.dt * v / m.l * tu;
    const double PI_ = 3.14159265358979323846;
    if (th > PI_) th -= 2 * PI_;
    else if (th < -PI_) th += 2 * PI_;
  }
  if (m.use_model_noise) {   // s_k.setCov(Q_); s_k.sample(): x += L_Q xi
    const double a = z[3], b = z[4], c = z[5];
    x += m.LQ[0] * a;
    y += m.LQ[3] * a + m.LQ[4] * b;
    th += m.LQ[6] * a + m.LQ[7] * b + m.LQ[8] * c;
  }
  pose64[3 * i] = x; pose64[3 * i + 1] = y; pose64[3 * i + 2] = th;
  pose[4 * i] = (T)x; pose[4 * i + 1] = (T)y; pose[4 * i + 2] = (T)th; pose[4 * i + 3] = T(0);
}

// Murty compatibility: w[idx[k]] *= ratio[k] (the truncated sums of the flagged particles), then the shard's
// [sum w, sum w^2] again, in a fixed order (one CTA)
__global__ void murty_patch_kernel(double* w, const int* idx, const double* ratio, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) w[idx[k]] *= ratio[k];
}
__global__ void weight_sums_kernel(const double* w, int N, double* sums) {
  __shared__ double r1[32], r2[32];
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const double v = w[i]; a += v; b += v * v; }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = a; r2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s1 = 0, s2 = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) { s1 += r1[k]; s2 += r2[k]; }
    sums[0] = s1;
    sums[1] = s2;
  }
}

// w_i /= sum  (ParticleFilter::normalizeWeights, include/ParticleFilter.hpp:352-363)
__global__ void normalize_kernel(double* w, const double* sums, int N, double* w_host) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    const double v = w[i] / sums[0];
    w[i] = v;
    if (w_host) w_host[i] = v;   // the caller's pinned buffer (KParams::w_host)
  }
}

// Inputs of a host-facing step of the Victoria Park kernels, read straight from the caller's pinned host memory
// (zero-copy loads over PCIe) and converted to the device layout: poses (fp64 copy kept for export / propagate), pose
// covariance, particle weights.  (The 2-D update kernels do this themselves, host_in_convert; the measurement batch
// travels by value in the update kernel's parameters either way.)
struct HostInParams {
  const double* pose;     // host [N][3]
  const double* weight;   // host [N] or NULL
  const double* pcov;     // host [N][6] (mode 2) or NULL
  int mode, N;
  double cov6[6];         // mode 1
};
template <typename T>
__global__ void host_in_kernel(const HostInParams h, double* __restrict__ pose64, T* __restrict__ pose, T* __restrict__ pcov,
                               double* __restrict__ w_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < h.N) {
    const double x = h.pose[3 * i], y = h.pose[3 * i + 1], th = h.pose[3 * i + 2];
    pose64[3 * i] = x; pose64[3 * i + 1] = y; pose64[3 * i + 2] = th;
    pose[4 * i] = (T)x; pose[4 * i + 1] = (T)y; pose[4 * i + 2] = (T)th; pose[4 * i + 3] = T(0);
    if (h.mode == 2) {
      for (int k = 0; k < 6; k++) pcov[8 * i + k] = (T)h.pcov[6 * i + k];
      pcov[8 * i + 6] = pcov[8 * i + 7] = T(0);
    }
    if (h.weight) w_dev[i] = h.weight[i];
  }
  if (h.mode == 1 && i == 0) {
    for (int k = 0; k < 6; k++) pcov[k] = (T)h.cov6[k];
    pcov[6] = pcov[7] = T(0);
  }
}

}  // namespace rfsb200
