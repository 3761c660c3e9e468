// rfsb200_abi.cu — extern "C" entry points of include/rfsb200.h.
// Host side of the drop-in boundary: owns device state (particle-major SoA in HBM), converts the
// reference's fp64 host representation to the device layout on the GPU, launches the fused update
// kernel (phd_kernels.cuh) and serves state reads.  There is no CPU fallback anywhere in here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rfsb200.h"
#include "murty_compat.hpp"
#include "phd_kernels.cuh"
#include "phd_vp_kernels.cuh"
#include "birth_kernels.cuh"

using namespace rfsb200;

namespace {

thread_local std::string g_last_error;

struct StateBuf {
  void* gm = nullptr;       // T [N][npl][cap]  (npl = 6: 2-D landmarks, 10: 3-D)
  int* cnt = nullptr;       // [N]
  double* weight = nullptr; // [N]
};

}  // namespace

struct rfsb200_ctx {
  rfsb200_dims dims{};
  int N = 0, cap = 0, W = 0, prec = 32;
  int ld = 2;    // landmark / measurement dimension
  int nc = 3;    // unique covariance entries
  int npl = 6;   // planes per particle in HBM: ld means + nc covariance entries + weight
  double* scan_dev = nullptr;   // Victoria Park: lidar scan [720]
  size_t tsize = 4;
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  StateBuf st[2];
  int front = 0;     // committed state
  int last_out = 0;  // buffer written by the last update (== front unless NO_COMMIT)
  void* pose = nullptr;      // T [N][4]
  void* pose_cov = nullptr;  // T [N][8]
  void* pose_alt = nullptr;      // resample targets (swapped with pose / pose_cov)
  void* pose_cov_alt = nullptr;
  int pose_cov_mode = 0;
  void* Zdev = nullptr;      // T [64][2]
  unsigned long long* unused = nullptr;
  int* nfov = nullptr;
  unsigned long long* unused_alt = nullptr;   // resample target (swapped with unused)
  int* nfov_alt = nullptr;
  int* src_dev = nullptr;                     // [2][N] resample sources
  // fused all-reduce over peer memory
  void* comm_mail = nullptr;      // CommSlot[2][8] of this rank
  void* comm_peer[8] = {};        // every rank's mailbox as seen from this process
  int comm_rank = 0, comm_world = 1;
  unsigned long long comm_epoch = 0;
  unsigned long long comm_bar_epoch = 0;      // rfsb200_comm_barrier
  unsigned long long comm_timeout_ns = 2000000000ull;   // RFSB200_COMM_TIMEOUT_MS
  int* comm_error = nullptr;
  // RFSB200_UPDATE_DEFER_NORMALIZE: the weights of st[comm_pending_buf] still wait for the sums of comm_pending_epoch
  bool comm_local = false;                    // peers connected by rfsb200_comm_connect_local (plain pointers, nothing to close)
  bool comm_pending = false;
  unsigned long long comm_pending_epoch = 0;
  int comm_pending_buf = 0;
  bool consume_pending = false;               // the launch being prepared applies them on the way in (KParams::comm_pending)
  bool consume_scale = false;                 //   ... to the weights it reads (else it only picks the pairs up: the open buffer is overwritten)
  int last_nZ = 0;                            // size of the measurement batch still held in Zdev
  int* flags = nullptr;
  double* sums = nullptr;                 // [2]
  unsigned long long* totals = nullptr;   // [2]
  int* istats = nullptr;                  // [4]
  unsigned int* mstats = nullptr;         // [8]
  unsigned int* ticket = nullptr;
  unsigned int* work_counter = nullptr;
  unsigned long long* stats_out = nullptr;  // [8]
  int cfg_mode_mf = -1;                      // launch configuration was computed for this mode
  int cfg_n_eval = -1;                       //   ... and this eval-point count (multi-feature scratch)
  // staging (device): packed fp64 + offsets
  double* stg = nullptr;       // N*cap*6 doubles
  long long* offs = nullptr;   // [N+1]
  double* stg_small = nullptr; // N*16 doubles (poses, covs, weights)
  // pinned host scratch
  unsigned char* hpin = nullptr;
  size_t hpin_bytes = 0;
  bool have_model = false, have_cfg = false, have_maps = false, have_poses = false;
  rfsb200_model_desc model{};
  rfsb200_filter_cfg cfg{};
  int merge_algo = 1;
  double* dp_scratch = nullptr;              // multi-feature: global workspace of the assignment-sum DP (see KParams)
  size_t dp_scratch_bytes = 0;
  int dp_gmaxb = 0;
  int dp_onchip = DP_MAXB;                   // RFSB200_DP_ONCHIP_MAXB (test aid): smaller partitions take the workspace too
  bool zero_copy = true;                     // rfsb200_update_host: pinned caller buffers are read / written by the kernels directly
  double* w_host = nullptr;                  // device views of the caller's result buffers for the NEXT launch (or NULL)
  unsigned long long* unused_host = nullptr;
  int* nfov_host = nullptr;
  unsigned long long* stats_host = nullptr;
  // host-facing step: device views of the caller's INPUT buffers for the next launch (or NULL), and the completion word
  const double* hin_pose = nullptr; const double* hin_weight = nullptr; const double* hin_pcov = nullptr;
  double hin_cov6[6] = {};
  // Murty compatibility (rfsb200_filter_cfg::murty_compat): partitions written out by the kernel, patch lists
  unsigned long long* murty_buf = nullptr;   // [MURTY_WORDS]
  unsigned int* murty_count = nullptr;       // [2]
  int* murty_idx = nullptr;                  // [N]
  double* murty_ratio = nullptr;             // [N]
  int last_murty_terms = 0;                  // diagnostics: assignments enumerated by the last update
  unsigned long long* hin_ready = nullptr;   // [HIN_SLICES] slice flags of the host-facing step (KParams::hin_ready)
  unsigned long long* done_host = nullptr;   // device view of the pinned completion word for the next launch (or NULL)
  unsigned long long done_seq = 0;
  double Zval[MAX_Z * 3] = {};               // the measurement batch of the next launch, fp64
  // candidate-list births (rfsb200_birth_candidates): per-particle lists, double-buffered; allocated on first use
  double* cand[2] = {nullptr, nullptr};      // [N][BIRTH_CAND_CAP][cand_rec(ld)]
  int* cand_n[2] = {nullptr, nullptr};       // [N]
  int cand_front = 0;
  std::vector<cudaEvent_t> prof_ev;   // event pairs around the update kernel (rfsb200_profile_*)
  int prof_cap = 0, prof_n = 0;
  unsigned long long* prof_dev = nullptr;   // [16] stage-timing build (KParams::prof)
  bool prof_valid = false;                  // an update with RFSB200_UPDATE_STAGE_TIMES ran since the last read
  int prof_nwarps = 0;
  int grid = 0;
  int nwarps = 4;   // warps per CTA of the update kernel
  size_t smem_bytes = 0;
  int warp_bytes = 0;
  int mf_bytes = 0;
  std::string err;
};

namespace {

int fail(rfsb200_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  g_last_error = buf;
  return code;
}

#define CU(ctx, call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ctx, RFSB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

// ---- layout conversion kernels (device side of upload / download) ------------------------------
// packed fp64 (mean[k][2], cov[k][3], w[k]) -> particle-major SoA planes of T. One warp per particle.
template <typename T>
__global__ void pack_soa_kernel(const int* __restrict__ cnt, const long long* __restrict__ offs,
                                const double* __restrict__ mean, const double* __restrict__ cov,
                                const double* __restrict__ w, T* __restrict__ gm, int N, int cap, int ld, int nc) {
  const int lane = threadIdx.x & 31;
  const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pi >= N) return;
  const int n = cnt[pi];
  const long long o = offs[pi];
  T* g = gm + (size_t)pi * (ld + nc + 1) * cap;
  for (int k = lane; k < n; k += 32) {
    const long long s = o + k;
    for (int d = 0; d < ld; d++) g[d * cap + k] = (T)mean[ld * s + d];
    for (int d = 0; d < nc; d++) g[(ld + d) * cap + k] = (T)cov[nc * s + d];
    g[(ld + nc) * cap + k] = (T)w[s];
  }
}

template <typename T>
__global__ void unpack_soa_kernel(const int* __restrict__ cnt, const long long* __restrict__ offs,
                                  const T* __restrict__ gm, double* __restrict__ mean,
                                  double* __restrict__ cov, double* __restrict__ w, int N, int cap, int ld, int nc) {
  const int lane = threadIdx.x & 31;
  const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pi >= N) return;
  const int n = cnt[pi];
  const long long o = offs[pi];
  const T* g = gm + (size_t)pi * (ld + nc + 1) * cap;
  for (int k = lane; k < n; k += 32) {
    const long long s = o + k;
    for (int d = 0; d < ld; d++) mean[ld * s + d] = (double)g[d * cap + k];
    for (int d = 0; d < nc; d++) cov[nc * s + d] = (double)g[(ld + d) * cap + k];
    w[s] = (double)g[(ld + nc) * cap + k];
  }
}

// packed fp64 Gaussians appended behind the existing ones of each particle (births decided by the host)
template <typename T>
__global__ void append_soa_kernel(const int* __restrict__ add, const long long* __restrict__ offs,
                                  const double* __restrict__ mean, const double* __restrict__ cov,
                                  const double* __restrict__ w, T* __restrict__ gm, int* __restrict__ cnt,
                                  int* __restrict__ flags, int N, int cap, int ld, int nc) {
  const int lane = threadIdx.x & 31;
  const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pi >= N) return;
  const int na = add[pi];
  if (na <= 0) return;
  const int n0 = cnt[pi];
  const long long o = offs[pi];
  T* g = gm + (size_t)pi * (ld + nc + 1) * cap;
  const int fit = (n0 + na > cap) ? (cap - n0 > 0 ? cap - n0 : 0) : na;
  for (int k = lane; k < fit; k += 32) {
    const long long s = o + k;
    const int j = n0 + k;
    for (int d = 0; d < ld; d++) g[d * cap + j] = (T)mean[ld * s + d];
    for (int d = 0; d < nc; d++) g[(ld + d) * cap + j] = (T)cov[nc * s + d];
    g[(ld + nc) * cap + j] = (T)w[s];
  }
  __syncwarp();
  if (lane == 0) {
    cnt[pi] = n0 + fit;
    if (fit < na) flags[pi] |= FLAG_OVERFLOW | FLAG_BIRTH_OVERFLOW;
  }
}

// ---- particle records for the cross-GPU exchange: [planes npl*cap T][header 128 B] per particle -----------------
struct ParticleHeader {          // 128 bytes
  double pose64[3];
  double weight;
  unsigned long long unused;
  int cnt, nfov;
  unsigned char pose_t[32];      // T[4]
  unsigned char pcov_t[64];      // T[8]
};
static_assert(sizeof(ParticleHeader) % 16 == 0, "records must keep the planes 16-byte aligned");

template <typename T>
__global__ void export_particles_kernel(const int* __restrict__ idx, int n, const T* __restrict__ gm, const int* __restrict__ cnt,
                                        const double* __restrict__ weight, const T* __restrict__ pose, const T* __restrict__ pcov,
                                        int pose_cov_mode, const double* __restrict__ pose64,
                                        const unsigned long long* __restrict__ unused, const int* __restrict__ nfov,
                                        unsigned char* __restrict__ buf, long long rec, int cap, int npl, int N) {
  const int lane = threadIdx.x & 31;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n) return;
  int i = idx[k];
  i = i < 0 ? 0 : (i >= N ? N - 1 : i);
  unsigned char* r = buf + (size_t)k * rec;
  T* planes = reinterpret_cast<T*>(r);
  const T* g = gm + (size_t)i * npl * cap;
  const int c = cnt[i];
  for (int pl = 0; pl < npl; pl++)
    for (int j = lane; j < c; j += 32) planes[(size_t)pl * cap + j] = g[(size_t)pl * cap + j];
  if (lane == 0) {
    ParticleHeader* h = reinterpret_cast<ParticleHeader*>(r + (size_t)npl * cap * sizeof(T));
    for (int d = 0; d < 3; d++) h->pose64[d] = pose64[3 * i + d];
    h->weight = weight[i];
    h->unused = unused[i];
    h->cnt = c;
    h->nfov = nfov[i];
    T* pt = reinterpret_cast<T*>(h->pose_t);
    for (int d = 0; d < 4; d++) pt[d] = pose[4 * i + d];
    T* pc = reinterpret_cast<T*>(h->pcov_t);
    for (int d = 0; d < 8; d++) pc[d] = pose_cov_mode == 2 ? pcov[8 * i + d] : (pose_cov_mode == 1 ? pcov[d] : T(0));
  }
}

template <typename T>
__global__ void import_particles_kernel(const int* __restrict__ slot, int n, T* __restrict__ gm, int* __restrict__ cnt,
                                        double* __restrict__ weight, T* __restrict__ pose, T* __restrict__ pcov,
                                        int pose_cov_mode, double* __restrict__ pose64, unsigned long long* __restrict__ unused,
                                        int* __restrict__ nfov, const unsigned char* __restrict__ buf, long long rec, int cap,
                                        int npl, int N, double w_value) {
  const int lane = threadIdx.x & 31;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n) return;
  const int i = slot[k];
  if (i < 0 || i >= N) return;
  const unsigned char* r = buf + (size_t)k * rec;
  const T* planes = reinterpret_cast<const T*>(r);
  const ParticleHeader* h = reinterpret_cast<const ParticleHeader*>(r + (size_t)npl * cap * sizeof(T));
  T* g = gm + (size_t)i * npl * cap;
  const int c = h->cnt;
  for (int pl = 0; pl < npl; pl++)
    for (int j = lane; j < c; j += 32) g[(size_t)pl * cap + j] = planes[(size_t)pl * cap + j];
  if (lane == 0) {
    for (int d = 0; d < 3; d++) pose64[3 * i + d] = h->pose64[d];
    weight[i] = w_value;
    unused[i] = h->unused;
    cnt[i] = c;
    nfov[i] = h->nfov;
    const T* pt = reinterpret_cast<const T*>(h->pose_t);
    for (int d = 0; d < 4; d++) pose[4 * i + d] = pt[d];
    if (pose_cov_mode == 2) {
      const T* pc = reinterpret_cast<const T*>(h->pcov_t);
      for (int d = 0; d < 8; d++) pcov[8 * i + d] = pc[d];
    }
  }
}

// exclusive scan of counts -> offsets[N+1]; single CTA (N is at most a few 10^5)
__global__ void scan_counts_kernel(const int* __restrict__ cnt, long long* __restrict__ offs, int N) {
  __shared__ long long part[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const int per = (N + nt - 1) / nt;
  const int lo = t * per, hi = min(N, lo + per);
  long long s = 0;
  for (int i = lo; i < hi; i++) s += cnt[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int i = 0; i < nt; i++) { long long v = part[i]; part[i] = run; run += v; }
    offs[N] = run;
  }
  __syncthreads();
  long long run = part[t];
  for (int i = lo; i < hi; i++) { offs[i] = run; run += cnt[i]; }
}

template <typename T>
__global__ void pose_convert_kernel(const double* __restrict__ pose, const double* __restrict__ pcov,
                                    int mode, T* __restrict__ pose_out, T* __restrict__ pcov_out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    pose_out[4 * i] = (T)pose[3 * i];
    pose_out[4 * i + 1] = (T)pose[3 * i + 1];
    pose_out[4 * i + 2] = (T)pose[3 * i + 2];
    pose_out[4 * i + 3] = T(0);
    if (mode == 2) {
      for (int k = 0; k < 6; k++) pcov_out[8 * i + k] = (T)pcov[6 * i + k];
      pcov_out[8 * i + 6] = pcov_out[8 * i + 7] = T(0);
    }
  }
  if (mode == 1 && i == 0) {
    for (int k = 0; k < 6; k++) pcov_out[k] = (T)pcov[k];
    pcov_out[6] = pcov_out[7] = T(0);
  }
}

// rfs::MatPerm::calc for a batch: one warp per matrix
__global__ void permanent_kernel(const double* __restrict__ A, int n, int batch, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= batch) return;
  const double r = warp_permanent(A + (size_t)b * n * n, n, lane);
  if (lane == 0) out[b] = r;
}

int round_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Warps per CTA.  The per-CTA block (measurement batch + window tables / lidar scan) is built once per CTA and every
// CTA ends with a block-wide barrier, so FEWER, LARGER CTAs are better as long as the SMs hold as many warps: measured on
// C3, 16 warps in one CTA per SM run the step in 141 us, 4 CTAs of 4 warps in 172 us.  Choose, in this order: most
// particles in flight, most SMs in use (small shards), then the largest CTA.
struct LaunchShape { int nwarps = 0, grid = 0; size_t smem = 0; };

template <typename K>
int choose_launch_shape(rfsb200_ctx* c, K kernel, size_t cta_bytes, size_t warp_bytes, int max_warps, LaunchShape* out) {
  int best_nw = 0, best_occ = 0, best_sms = -1;
  long long best_res = -1;
  int nw_lo = 1, nw_hi = max_warps;   // (the kernel's launch bounds) large work capacities / the fp64 build may only fit a few warps
  if (const char* e = getenv("RFSB200_WARPS_PER_CTA")) { nw_lo = nw_hi = std::max(1, std::min(max_warps, atoi(e))); }   // tuning aid
  for (int nw = nw_lo; nw <= nw_hi; nw++) {
    const size_t smem = cta_bytes + (size_t)nw * warp_bytes;
    if (smem > 227 * 1024) break;
    CU(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nw * 32, smem));
    if (occ < 1) continue;
    const long long res = std::min<long long>(c->N, (long long)occ * nw * c->sm_count);
    const int sms = std::min(c->sm_count, (c->N + nw - 1) / nw);
    if (res > best_res || (res == best_res && (sms > best_sms || (sms == best_sms && nw > best_nw)))) {
      best_res = res; best_sms = sms; best_nw = nw; best_occ = occ;
    }
  }
  if (best_nw == 0 || best_occ < 1)
    return fail(c, RFSB200_ECAPACITY, "work_capacity %d needs %zu B shared memory per CTA (> 227 KB)", c->W,
                cta_bytes + (size_t)nw_lo * warp_bytes);
  out->nwarps = best_nw;
  out->smem = cta_bytes + (size_t)best_nw * warp_bytes;
  CU(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->smem));
  const int need = (c->N + best_nw - 1) / best_nw;
  out->grid = std::max(1, std::min(std::min(need, best_occ * c->sm_count), 8192));   // (8192: slice flags of the host-facing step)
  return RFSB200_OK;
}
template <typename K>
int choose_warps_per_cta(rfsb200_ctx* c, K kernel, size_t cta_bytes, size_t warp_bytes, int max_warps) {
  LaunchShape sh;
  const int rc = choose_launch_shape(c, kernel, cta_bytes, warp_bytes, max_warps, &sh);
  if (rc) return rc;
  c->nwarps = sh.nwarps;
  c->smem_bytes = sh.smem;
  c->grid = sh.grid;
  return RFSB200_OK;
}

// Multi-feature weighting: the exact assignment sum of a partition is a DP over the subsets of its smaller side, 2^b
// states.  Up to b = 7 the two tables live in shared memory; beyond, in a per-warp block of this workspace, sized for
// b <= min(eval points, z capacity, 15) and capped at 2 GiB (a smaller b then).  Without it (allocation failed, or
// nothing can exceed 7) larger partitions are skipped and flagged as before.
int ensure_dp_scratch(rfsb200_ctx* c, int n_eval) {
  int b = std::min(std::min(n_eval, (int)c->dims.z_capacity), DP_GMAXB);
  const size_t warps = (size_t)c->grid * c->nwarps;
  while (b > DP_MAXB && warps * (size_t(16) << b) > (size_t(2) << 30)) b--;
  if (b <= c->dp_onchip) b = 0;
  const size_t need = b ? warps * (size_t(16) << b) : 0;
  if (need > c->dp_scratch_bytes) {
    if (c->dp_scratch) cudaFree(c->dp_scratch);
    c->dp_scratch = nullptr;
    c->dp_scratch_bytes = 0;
    if (cudaMalloc((void**)&c->dp_scratch, need) != cudaSuccess) {
      cudaGetLastError();
      c->dp_scratch = nullptr;
      c->dp_gmaxb = 0;
      return RFSB200_OK;   // not fatal: the flag tells the caller which particles were affected
    }
    c->dp_scratch_bytes = need;
  }
  c->dp_gmaxb = (b && c->dp_scratch) ? b : 0;
  return RFSB200_OK;
}

template <typename T, bool MF>
int configure_launch_t(rfsb200_ctx* c) {
  const int mf = MF ? 1 : 0;
  const int n_eval = (MF && c->have_cfg) ? std::max(1, c->cfg.eval_point_count) : MAX_EVAL;
  if (c->cfg_mode_mf == mf && (!MF || c->cfg_n_eval == n_eval)) return RFSB200_OK;
  c->cfg_n_eval = n_eval;
  c->mf_bytes = mf ? mf_region_bytes<T>(c->W, n_eval, c->dims.z_capacity) : 0;   // the work region in multi-feature mode
  c->warp_bytes = warp_bytes_for<T>(c->W, mf, mf ? c->mf_bytes : merge_scratch_bytes<T>(c->W));
  int rc;
  if constexpr (sizeof(T) == 4) {
    rc = c->W == 256 ? choose_warps_per_cta(c, phd_update_kernel<T, MF, 256>, (size_t)z_bytes<T>(), (size_t)c->warp_bytes, update_max_threads<T, MF>() / 32)
                     : choose_warps_per_cta(c, phd_update_kernel<T, MF, 0>, (size_t)z_bytes<T>(), (size_t)c->warp_bytes, update_max_threads<T, MF>() / 32);
  } else {
    rc = choose_warps_per_cta(c, phd_update_kernel<T, MF, 0>, (size_t)z_bytes<T>(), (size_t)c->warp_bytes, update_max_threads<T, MF>() / 32);
  }
  if (rc) return rc;
  if (MF) ensure_dp_scratch(c, n_eval);
  c->cfg_mode_mf = mf;
  return RFSB200_OK;
}
template <typename T, bool MF>
int configure_launch_vp_t(rfsb200_ctx* c) {
  const int mf = MF ? 1 : 0;
  const int n_eval = (MF && c->have_cfg) ? std::max(1, c->cfg.eval_point_count) : MAX_EVAL;
  if (c->cfg_mode_mf == mf && (!MF || c->cfg_n_eval == n_eval)) return RFSB200_OK;
  c->cfg_n_eval = n_eval;
  c->mf_bytes = 0;
  c->warp_bytes = vp_warp_bytes<T>(c->W, mf, n_eval, c->dims.z_capacity);
  int rc = choose_warps_per_cta(c, phd_update_vp_kernel<T, MF>, (size_t)vp_cta_bytes<T>(), (size_t)c->warp_bytes, 16);
  if (rc) return rc;
  if (MF) ensure_dp_scratch(c, n_eval);
  c->cfg_mode_mf = mf;
  return RFSB200_OK;
}
template <typename T>
int configure_launch(rfsb200_ctx* c, int mf) {
  if (c->ld == 3) return mf ? configure_launch_vp_t<T, true>(c) : configure_launch_vp_t<T, false>(c);
  return mf ? configure_launch_t<T, true>(c) : configure_launch_t<T, false>(c);
}

constexpr unsigned MURTY_WORDS = 1u << 20;   // 8 MB of partition records per update

// Murty compatibility is in force for this update (multi-feature weighting with rfsb200_filter_cfg::murty_compat)
inline bool murty_active(const rfsb200_ctx* c) { return c->have_cfg && c->cfg.murty_compat != 0 && !c->cfg.use_cluster_process; }

template <typename T>
int launch_update(rfsb200_ctx* c, int nZ, int out_idx, unsigned flags) {
  KParams<T> p{};
  const rfsb200_model_desc& m = c->model;
  const rfsb200_filter_cfg& f = c->cfg;
  p.R00 = (T)m.R[0]; p.R01 = (T)m.R[1]; p.R11 = (T)m.R[3];
  p.Pd = (T)m.Pd; p.kappa = (T)m.clutter_intensity;
  p.rmin = (T)m.range_min; p.rmax = (T)m.range_max; p.rbuf = (T)m.range_buffer;
  p.thr_r = (T)m.innov_thr_range; p.thr_b = (T)m.innov_thr_bearing;
  p.birth_w = (T)f.birth_gaussian_weight;
  p.gate2 = (T)(f.new_gaussian_create_innov_md_threshold * f.new_gaussian_create_innov_md_threshold);
  p.eval_min_w = (T)f.eval_point_gaussian_weight;
  p.wl_gate2 = (T)(f.meas_likelihood_md_threshold * f.meas_likelihood_md_threshold);
  p.merge_t2 = (T)(f.merging_threshold * f.merging_threshold);
  p.merge_f = (T)f.merging_cov_inflation_factor;
  p.prune_t = (T)f.pruning_threshold;
  p.n_eval = f.eval_point_count; p.use_sc = f.use_cluster_process ? 1 : 0;
  p.sum_method = f.assignment_sum_method; p.merge_algo = c->merge_algo;
  p.log_clutter_integral = log(m.clutter_integral);
  p.log_kappa = log(m.clutter_intensity);
  p.N = c->N; p.cap = c->cap; p.W = c->W; p.nZ = nZ; p.pose_cov_mode = c->pose_cov_mode;
  const StateBuf& in = c->st[c->front];
  const StateBuf& out = c->st[out_idx];
  p.gm_in = (const T*)in.gm; p.cnt_in = in.cnt; p.w_in = in.weight;
  p.pose = (const T*)c->pose; p.pose_cov = (const T*)c->pose_cov;
  p.Z = (const T*)c->Zdev;
  p.Zdev_w = (T*)c->Zdev;
  for (int k = 0; k < c->ld * nZ; k++) p.Zval[k] = c->Zval[k];
  p.pose_h = c->hin_pose; p.weight_h = c->hin_weight; p.pcov_h = c->hin_pcov;
  for (int k = 0; k < 6; k++) p.cov6[k] = c->hin_cov6[k];
  p.pose_w = (T*)c->pose; p.pose64_w = c->stg_small; p.pcov_w = (T*)c->pose_cov; p.w_front = (double*)in.weight;
  p.done_host = c->done_host; p.done_value = c->done_seq; p.hin_ready = c->hin_ready;
  p.gm_out = (T*)out.gm; p.cnt_out = out.cnt; p.w_out = out.weight;
  p.unused = c->unused; p.nfov = c->nfov; p.flags = c->flags;
  p.sums = c->sums; p.totals = c->totals; p.istats = c->istats; p.ticket = c->ticket; p.mstats = c->mstats;
  p.work_counter = c->work_counter; p.stats_out = c->stats_out;
  p.comm_rank = c->comm_rank; p.comm_world = 1; p.fused_normalize = 0; p.comm_epoch = 0; p.comm_error = c->comm_error;
  p.comm_timeout_ns = c->comm_timeout_ns;
  p.unused_host = c->unused_host; p.nfov_host = c->nfov_host; p.stats_host = c->stats_host;
  // the weights go to the host from whichever launch leaves them final: this kernel (fused normalisation, or none at
  // all), else normalize_kernel
  const bool normalize_follows = !(flags & (RFSB200_UPDATE_NO_NORMALIZE | RFSB200_UPDATE_FUSED_ALLREDUCE));
  p.w_host = normalize_follows ? nullptr : c->w_host;
  if (murty_active(c)) {
    // the host replaces the sums of the large partitions before the weights are added up (enqueue_update): this launch
    // stops at the unnormalised weights
    if (!c->murty_buf) {
      CU(c, cudaMalloc((void**)&c->murty_buf, (size_t)MURTY_WORDS * 8));
      CU(c, cudaMalloc((void**)&c->murty_count, 8));
      CU(c, cudaMalloc((void**)&c->murty_idx, (size_t)c->N * 4));
      CU(c, cudaMalloc((void**)&c->murty_ratio, (size_t)c->N * 8));
    }
    CU(c, cudaMemsetAsync(c->murty_count, 0, 8, c->stream));
    p.murty_buf = c->murty_buf; p.murty_count = c->murty_count; p.murty_cap_words = MURTY_WORDS;
    p.w_host = nullptr;
    flags = (flags & ~RFSB200_UPDATE_FUSED_ALLREDUCE) | RFSB200_UPDATE_NO_NORMALIZE;
  }
  p.comm_defer = 0; p.comm_pending = 0; p.comm_pending_scale = 0; p.comm_prev_epoch = 0;
  if (flags & RFSB200_UPDATE_FUSED_ALLREDUCE) {
    p.comm_world = c->comm_world;
    p.fused_normalize = 1;
    p.comm_epoch = c->comm_epoch + 1;   // the ctx moves on only when the launch has succeeded (below)
    for (int r = 0; r < 8; r++) p.comm_peer[r] = c->comm_peer[r];
    if ((flags & RFSB200_UPDATE_DEFER_NORMALIZE) && c->comm_world > 1) { p.comm_defer = 1; p.fused_normalize = 0; }
  }
  if (c->consume_pending) {   // the sums of the previous (deferred) step, applied to the weights on the way in
    p.comm_pending = c->comm_world;
    p.comm_pending_scale = c->consume_scale ? 1 : 0;
    p.comm_prev_epoch = c->comm_pending_epoch;
    for (int r = 0; r < 8; r++) p.comm_peer[r] = c->comm_peer[r];
  }
  const int mf = f.use_cluster_process ? 0 : 1;
  {
    int rc = configure_launch<T>(c, mf);
    if (rc) return rc;
    p.warp_bytes = c->warp_bytes;
    p.mf_bytes = c->mf_bytes;
    p.n_eval_cap = c->cfg_n_eval;
    p.zcap = c->dims.z_capacity;
    p.dp_scratch = (mf && c->dp_gmaxb) ? c->dp_scratch : nullptr;
    p.dp_gmaxb = c->dp_gmaxb;
    p.dp_onchip = c->dp_onchip;
  }
  const bool prof = c->prof_n < c->prof_cap;
  if (prof) CU(c, cudaEventRecord(c->prof_ev[2 * c->prof_n], c->stream));
  if (flags & RFSB200_UPDATE_STAGE_TIMES) {
    // the stage-timing build of the kernel (fp32, 2-D model): same results, clocks read at the stage boundaries
    if constexpr (sizeof(T) == 4) {
      if (c->ld != 2) return fail(c, RFSB200_EUNSUPPORTED, "RFSB200_UPDATE_STAGE_TIMES: 2-D model only");
      if (!c->prof_dev) CU(c, cudaMalloc((void**)&c->prof_dev, 16 * 8));
      unsigned long long init[16] = {};
      init[12] = ~0ull;
      // (pageable source: the copy is staged before the call returns)
      CU(c, cudaMemcpyAsync(c->prof_dev, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
      p.prof = c->prof_dev;
      LaunchShape sh;
      int rc;
      if (mf) rc = choose_launch_shape(c, phd_update_kernel<T, true, 0, true>, (size_t)z_bytes<T>(), (size_t)c->warp_bytes, update_max_threads<T, true, true>() / 32, &sh);
      else rc = choose_launch_shape(c, phd_update_kernel<T, false, 0, true>, (size_t)z_bytes<T>(), (size_t)c->warp_bytes, update_max_threads<T, false, true>() / 32, &sh);
      if (rc) return rc;
      if (mf && c->dp_gmaxb && (size_t)sh.grid * sh.nwarps > (size_t)c->grid * c->nwarps) p.dp_scratch = nullptr;   // workspace sized for the product launch
      if (mf) phd_update_kernel<T, true, 0, true><<<sh.grid, sh.nwarps * 32, sh.smem, c->stream>>>(p);
      else phd_update_kernel<T, false, 0, true><<<sh.grid, sh.nwarps * 32, sh.smem, c->stream>>>(p);
      c->prof_valid = true;
      c->prof_nwarps = sh.nwarps;
    } else {
      return fail(c, RFSB200_EUNSUPPORTED, "RFSB200_UPDATE_STAGE_TIMES: fp32 build only");
    }
  } else if (c->ld == 3) {
    VPParams<T> v{};
    v.k = p;
    v.R00 = (T)m.R[0]; v.R01 = (T)m.R[1]; v.R10 = (T)m.R[3]; v.R11 = (T)m.R[4]; v.R22 = (T)m.R[8];
    v.Slb = (T)m.Slb;
    v.rmin = m.range_min; v.rmax = m.range_max; v.bmin = m.bearing_min; v.bmax = m.bearing_max;
    v.buf_pd = m.buffer_zone_pd;
    v.pd_n = m.pd_table_n; v.scan_n = m.scan_n; v.scan = c->scan_dev;
    for (int k = 0; k < VP_PD_MAX; k++) v.pd_table[k] = k < m.pd_table_n ? m.pd_table[k] : 0.0;
    if (mf) phd_update_vp_kernel<T, true><<<c->grid, c->nwarps * 32, c->smem_bytes, c->stream>>>(v);
    else phd_update_vp_kernel<T, false><<<c->grid, c->nwarps * 32, c->smem_bytes, c->stream>>>(v);
  } else if (sizeof(T) == 4 && c->W == 256) {   // the work capacity as a compile-time constant (fp32 kernels)
    if constexpr (sizeof(T) == 4) {
      if (mf) phd_update_kernel<T, true, 256><<<c->grid, c->nwarps * 32, c->smem_bytes, c->stream>>>(p);
      else phd_update_kernel<T, false, 256><<<c->grid, c->nwarps * 32, c->smem_bytes, c->stream>>>(p);
    }
  } else if (mf) phd_update_kernel<T, true, 0><<<c->grid, c->nwarps * 32, c->smem_bytes, c->stream>>>(p);
  else phd_update_kernel<T, false, 0><<<c->grid, c->nwarps * 32, c->smem_bytes, c->stream>>>(p);
  CU(c, cudaGetLastError());
  if (flags & RFSB200_UPDATE_FUSED_ALLREDUCE) c->comm_epoch++;
  if (prof) {
    CU(c, cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], c->stream));
    c->prof_n++;
  }
  return RFSB200_OK;
}

// Page-locked ranges this library handed out (rfsb200_host_alloc, the ctx staging area) with their device aliases:
// the per-step buffers of rfsb200_update_host are looked up here instead of asking the driver seven times per step.
struct PinnedRange { uintptr_t host, dev; size_t bytes; };
std::mutex g_pinned_mu;
std::map<uintptr_t, PinnedRange> g_pinned;

void* query_device_alias(const void* host_ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return a.devicePointer;
}
void remember_pinned(void* host_ptr, size_t bytes) {
  void* dev = query_device_alias(host_ptr);
  if (!dev) return;
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned[(uintptr_t)host_ptr] = PinnedRange{(uintptr_t)host_ptr, (uintptr_t)dev, bytes};
}
void forget_pinned(void* host_ptr) {
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned.erase((uintptr_t)host_ptr);
}

// device-accessible alias of a pinned / registered host pointer (unified addressing), or NULL for pageable memory
template <typename P>
P* device_view(P* host_ptr) {
  if (!host_ptr) return nullptr;
  {
    const uintptr_t h = (uintptr_t)host_ptr;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    auto it = g_pinned.upper_bound(h);
    if (it != g_pinned.begin()) {
      --it;
      if (h < it->second.host + it->second.bytes) return (P*)(it->second.dev + (h - it->second.host));
    }
  }
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, (const void*)host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return (P*)a.devicePointer;
}

int ensure_pinned(rfsb200_ctx* c, size_t bytes) {
  if (c->hpin_bytes >= bytes) return RFSB200_OK;
  if (c->hpin) { forget_pinned(c->hpin); cudaFreeHost(c->hpin); }
  c->hpin = nullptr;
  c->hpin_bytes = 0;
  CU(c, cudaMallocHost((void**)&c->hpin, bytes));
  memset(c->hpin, 0, bytes);   // (the completion word of rfsb200_update_host lives here)
  c->hpin_bytes = bytes;
  remember_pinned(c->hpin, bytes);
  return RFSB200_OK;
}

}  // namespace

namespace {
template <typename T>
int do_predict(rfsb200_ctx* c, const double* Q, int add_births, double birth_w) {
  PredictParams<T> p{};
  const StateBuf& st = c->st[c->front];
  p.gm = (T*)st.gm; p.cnt = st.cnt; p.unused = c->unused; p.flags = c->flags;
  p.pose = (const T*)c->pose; p.Z = (const T*)c->Zdev;
  p.N = c->N; p.cap = c->cap; p.nZ = c->last_nZ;
  p.add_births = (add_births && c->last_nZ > 0) ? 1 : 0;
  p.add_q = Q ? 1 : 0;
  const double* R = c->model.R;
  p.R00 = (T)R[0]; p.R01 = (T)R[1]; p.R10 = (T)R[2]; p.R11 = (T)R[3];
  if (Q) { p.q00 = (T)Q[0]; p.q01 = (T)Q[1]; p.q11 = (T)Q[2]; }
  p.birth_w = (T)birth_w;
  if (!p.add_births && !p.add_q) return RFSB200_OK;
  predict_maps_kernel<T><<<(c->N * 32 + 127) / 128, 128, 0, c->stream>>>(p);
  CU(c, cudaGetLastError());
  return RFSB200_OK;
}
template <typename T>
int do_predict_vp(rfsb200_ctx* c, const double* Q, int add_births, double birth_w) {
  VPPredictParams<T> p{};
  const StateBuf& st = c->st[c->front];
  p.gm = (T*)st.gm; p.cnt = st.cnt; p.unused = c->unused; p.flags = c->flags;
  p.pose = (const T*)c->pose; p.Z = (const T*)c->Zdev;
  p.N = c->N; p.cap = c->cap; p.nZ = c->last_nZ;
  p.add_births = (add_births && c->last_nZ > 0) ? 1 : 0;
  p.add_q = Q ? 1 : 0;
  const double* R = c->model.R;
  p.R00 = (T)R[0]; p.R01 = (T)R[1]; p.R10 = (T)R[3]; p.R11 = (T)R[4]; p.R22 = (T)R[8];
  if (Q) for (int k = 0; k < 6; k++) p.q[k] = (T)Q[k];
  p.birth_w = (T)birth_w;
  if (!p.add_births && !p.add_q) return RFSB200_OK;
  predict_maps_vp_kernel<T><<<(c->N * 32 + 127) / 128, 128, 0, c->stream>>>(p);
  CU(c, cudaGetLastError());
  return RFSB200_OK;
}
}  // namespace


// ================================================================================================
extern "C" {

static int resolve_pending(rfsb200_ctx* c);

int rfsb200_abi_version(void) { return RFSB200_ABI_VERSION; }

int rfsb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* rfsb200_last_error(const rfsb200_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_last_error.c_str();
}

int rfsb200_create(rfsb200_ctx** out, const rfsb200_dims* d) {
  if (!out || !d) return fail(nullptr, RFSB200_EINVAL, "rfsb200_create: NULL argument");
  *out = nullptr;
  if (d->n_particles <= 0 || d->gm_capacity <= 0 || d->gm_capacity > 1024 || d->work_capacity > 1024 ||
      d->z_capacity <= 0 || d->z_capacity > MAX_Z)
    return fail(nullptr, RFSB200_EINVAL, "rfsb200_create: sizes out of range");
  if ((d->lmk_dim != 2 && d->lmk_dim != 3) || d->meas_dim != d->lmk_dim || d->pose_dim != 3)
    return fail(nullptr, RFSB200_EUNSUPPORTED, "implemented: 2-D landmarks + 2-D measurements (RngBrg) or 3-D + 3-D (VictoriaPark), 3-D poses");
  if (d->precision != 32 && d->precision != 64)
    return fail(nullptr, RFSB200_EINVAL, "precision must be 32 or 64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, RFSB200_ENODEVICE, "no CUDA device: the PHD update path has no CPU fallback");
  if (d->device < 0 || d->device >= ndev) return fail(nullptr, RFSB200_EINVAL, "bad device ordinal %d", d->device);
  rfsb200_ctx* c = new (std::nothrow) rfsb200_ctx();
  if (!c) return fail(nullptr, RFSB200_ENOMEM, "out of host memory");
  c->dims = *d;
  c->N = d->n_particles;
  c->ld = d->lmk_dim;
  c->nc = c->ld * (c->ld + 1) / 2;
  c->npl = c->ld + c->nc + 1;
  c->cap = (d->gm_capacity + 7) & ~7;
  c->W = round_pow2(std::max(d->work_capacity, c->cap));
  // the fp32 Victoria Park kernels take any multiple of 32: their planes are 11 x W words per warp, and W = 192 instead
  // of 256 is the difference between 11 and 14 resident warps per SM
  if (d->lmk_dim == 3 && d->precision == 32) c->W = (std::max(d->work_capacity, c->cap) + 31) & ~31;
  c->prec = d->precision;
  c->tsize = d->precision == 32 ? 4 : 8;
  c->device = d->device;
  int rc = RFSB200_OK;
  auto body = [&]() -> int {
    CU(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CU(c, cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10)
      return fail(c, RFSB200_EUNSUPPORTED, "device sm_%d%d is not Blackwell (built for sm_100a only)", prop.major, prop.minor);
    CU(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(c, cudaEventCreate(&c->ev0));
    CU(c, cudaEventCreate(&c->ev1));
    const size_t gm_bytes = (size_t)c->N * c->npl * c->cap * c->tsize;
    for (int k = 0; k < 2; k++) {
      CU(c, cudaMalloc(&c->st[k].gm, gm_bytes));
      CU(c, cudaMemset(c->st[k].gm, 0, gm_bytes));
      CU(c, cudaMalloc((void**)&c->st[k].cnt, (size_t)c->N * 4));
      CU(c, cudaMemset(c->st[k].cnt, 0, (size_t)c->N * 4));
      CU(c, cudaMalloc((void**)&c->st[k].weight, (size_t)c->N * 8));
      CU(c, cudaMemset(c->st[k].weight, 0, (size_t)c->N * 8));
    }
    CU(c, cudaMalloc(&c->pose, (size_t)c->N * 4 * c->tsize));
    CU(c, cudaMalloc(&c->pose_cov, (size_t)c->N * 8 * c->tsize));
    CU(c, cudaMemset(c->pose_cov, 0, (size_t)c->N * 8 * c->tsize));
    CU(c, cudaMalloc(&c->pose_alt, (size_t)c->N * 4 * c->tsize));
    CU(c, cudaMalloc(&c->pose_cov_alt, (size_t)c->N * 8 * c->tsize));
    CU(c, cudaMemset(c->pose, 0, (size_t)c->N * 4 * c->tsize));
    CU(c, cudaMemset(c->pose_alt, 0, (size_t)c->N * 4 * c->tsize));
    CU(c, cudaMemset(c->pose_cov_alt, 0, (size_t)c->N * 8 * c->tsize));
    CU(c, cudaMalloc(&c->Zdev, (size_t)MAX_Z * 4 * c->tsize + MAX_Z * 4));
    CU(c, cudaMalloc((void**)&c->unused, (size_t)c->N * 8));
    CU(c, cudaMalloc((void**)&c->nfov, (size_t)c->N * 4));
    CU(c, cudaMalloc((void**)&c->flags, (size_t)c->N * 4));
    CU(c, cudaMalloc(&c->comm_mail, COMM_BANKS * 8 * sizeof(CommSlot)));
    CU(c, cudaMemset(c->comm_mail, 0xff, COMM_BANKS * 8 * sizeof(CommSlot)));   // epoch = ~0: never matches
    if (const char* e = getenv("RFSB200_COMM_TIMEOUT_MS")) c->comm_timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    CU(c, cudaMalloc((void**)&c->hin_ready, 8192 * 8));   // one flag per CTA of the update kernel (grid <= SMs x 32 CTAs per SM)
    CU(c, cudaMemset(c->hin_ready, 0, 8192 * 8));
    CU(c, cudaMalloc((void**)&c->comm_error, 4));
    CU(c, cudaMemset(c->comm_error, 0, 4));
    CU(c, cudaMalloc((void**)&c->unused_alt, (size_t)c->N * 8));
    CU(c, cudaMalloc((void**)&c->nfov_alt, (size_t)c->N * 4));
    CU(c, cudaMalloc((void**)&c->src_dev, (size_t)c->N * 8));
    CU(c, cudaMemset(c->unused_alt, 0, (size_t)c->N * 8));
    CU(c, cudaMemset(c->nfov_alt, 0, (size_t)c->N * 4));
    CU(c, cudaMemset(c->unused, 0, (size_t)c->N * 8));
    CU(c, cudaMemset(c->nfov, 0, (size_t)c->N * 4));
    CU(c, cudaMemset(c->flags, 0, (size_t)c->N * 4));
    CU(c, cudaMalloc((void**)&c->sums, 16));
    CU(c, cudaMalloc((void**)&c->totals, 16));
    CU(c, cudaMalloc((void**)&c->istats, 16));
    CU(c, cudaMalloc((void**)&c->ticket, 4));
    CU(c, cudaMalloc((void**)&c->mstats, 32));
    CU(c, cudaMemset(c->mstats, 0, 32));
    CU(c, cudaMalloc((void**)&c->work_counter, 4));
    CU(c, cudaMalloc((void**)&c->stats_out, 128));
    CU(c, cudaMemset(c->work_counter, 0, 4));
    CU(c, cudaMemset(c->stats_out, 0, 128));
    CU(c, cudaMemset(c->totals, 0, 16));
    CU(c, cudaMemset(c->istats, 0, 16));
    CU(c, cudaMemset(c->sums, 0, 16));
    CU(c, cudaMemset(c->ticket, 0, 4));
    CU(c, cudaMalloc((void**)&c->stg, (size_t)c->N * c->cap * c->npl * 8));
    CU(c, cudaMalloc((void**)&c->scan_dev, VP_SCAN_MAX * 8));
    CU(c, cudaMemset(c->scan_dev, 0, VP_SCAN_MAX * 8));
    CU(c, cudaMalloc((void**)&c->offs, (size_t)(c->N + 1) * 8));
    CU(c, cudaMalloc((void**)&c->stg_small, (size_t)c->N * 16 * 8));
    if (const char* e = getenv("RFSB200_ZERO_COPY")) c->zero_copy = atoi(e) != 0;   // 0: always stage through copies
    if (const char* e = getenv("RFSB200_DP_ONCHIP_MAXB")) c->dp_onchip = std::max(0, std::min(DP_MAXB, atoi(e)));   // test aid
    int r = ensure_pinned(c, 1 << 16);
    if (r) return r;
    if (c->prec == 32) r = configure_launch<float>(c, 0);
    else r = configure_launch<double>(c, 0);
    if (r) return r;
    CU(c, cudaDeviceSynchronize());
    return RFSB200_OK;
  };
  rc = body();
  if (rc != RFSB200_OK) {
    g_last_error = c->err;
    rfsb200_destroy(c);
    return rc;
  }
  *out = c;
  return RFSB200_OK;
}

int rfsb200_destroy(rfsb200_ctx* c) {
  if (!c) return RFSB200_OK;
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  for (int k = 0; k < 2; k++) {
    cudaFree(c->st[k].gm); cudaFree(c->st[k].cnt); cudaFree(c->st[k].weight);
  }
  cudaFree(c->pose); cudaFree(c->pose_cov); cudaFree(c->Zdev);
  cudaFree(c->pose_alt); cudaFree(c->pose_cov_alt);
  cudaFree(c->unused); cudaFree(c->nfov); cudaFree(c->flags);
  cudaFree(c->unused_alt); cudaFree(c->nfov_alt); cudaFree(c->src_dev);
  for (int r = 0; r < c->comm_world; r++)
    if (r != c->comm_rank && c->comm_peer[r] && !c->comm_local) cudaIpcCloseMemHandle(c->comm_peer[r]);
  cudaFree(c->comm_mail); cudaFree(c->comm_error);
  cudaFree(c->sums); cudaFree(c->totals); cudaFree(c->istats); cudaFree(c->ticket);
  cudaFree(c->work_counter); cudaFree(c->stats_out); cudaFree(c->mstats);
  cudaFree(c->stg); cudaFree(c->offs); cudaFree(c->stg_small); cudaFree(c->scan_dev); cudaFree(c->dp_scratch); cudaFree(c->prof_dev); cudaFree(c->hin_ready);
  cudaFree(c->murty_buf); cudaFree(c->murty_count); cudaFree(c->murty_idx); cudaFree(c->murty_ratio);
  for (int k = 0; k < 2; k++) { cudaFree(c->cand[k]); cudaFree(c->cand_n[k]); }
  if (c->hpin) { forget_pinned(c->hpin); cudaFreeHost(c->hpin); }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return RFSB200_OK;
}

int rfsb200_set_stream(rfsb200_ctx* c, void* s, int external) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  c->stream = external ? (cudaStream_t)s : c->own_stream;   // s == NULL && external: the legacy default stream
  return RFSB200_OK;
}

int rfsb200_synchronize(rfsb200_ctx* c) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

int rfsb200_set_model(rfsb200_ctx* c, const rfsb200_model_desc* m) {
  if (!c || !m) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (m->model_id != RFSB200_MODEL_RNGBRG && m->model_id != RFSB200_MODEL_VICTORIAPARK)
    return fail(c, RFSB200_EUNSUPPORTED, "model id %d has no device descriptor (MeasurementModel_RngBrg, MeasurementModel_VictoriaPark)", m->model_id);
  if ((m->model_id == RFSB200_MODEL_RNGBRG) != (c->ld == 2))
    return fail(c, RFSB200_EINVAL, "model id %d does not match the landmark dimension %d of the ctx", m->model_id, c->ld);
  if (!(m->clutter_intensity > 0) || !(m->clutter_integral > 0))
    return fail(c, RFSB200_EINVAL, "clutter intensity / integral must be > 0");
  if (m->model_id == RFSB200_MODEL_VICTORIAPARK) {
    if (m->pd_table_n < 1 || m->pd_table_n > VP_PD_MAX) return fail(c, RFSB200_EINVAL, "pd_table_n %d outside [1,%d]", m->pd_table_n, VP_PD_MAX);
    if (m->scan_n < 0 || m->scan_n > VP_SCAN_MAX || (m->scan_n > 0 && !m->scan))
      return fail(c, RFSB200_EINVAL, "scan_n %d outside [0,%d] or NULL scan", m->scan_n, VP_SCAN_MAX);
    CU(c, cudaSetDevice(c->device));
    if (m->scan_n > 0) {   // the scan changes before every update (src/rbphdslam_VictoriaPark.cpp:582): staged copy
      int r = ensure_pinned(c, 1 << 16);
      if (r) return r;
      CU(c, cudaStreamSynchronize(c->stream));   // the previous copy out of the staging area has finished
      memcpy(c->hpin + 49152, m->scan, (size_t)m->scan_n * 8);
      CU(c, cudaMemcpyAsync(c->scan_dev, c->hpin + 49152, (size_t)m->scan_n * 8, cudaMemcpyHostToDevice, c->stream));
    }
  }
  c->model = *m;
  c->model.scan = nullptr;   // the host pointer is not kept
  c->have_model = true;
  return RFSB200_OK;
}

int rfsb200_set_filter_cfg(rfsb200_ctx* c, const rfsb200_filter_cfg* f) {
  if (!c || !f) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (f->eval_point_count < 0 || f->eval_point_count > MAX_EVAL)
    return fail(c, RFSB200_EUNSUPPORTED, "eval_point_count %d outside [0,%d]", f->eval_point_count, MAX_EVAL);
  if (!(f->pruning_threshold > 0)) return fail(c, RFSB200_EINVAL, "pruning_threshold must be > 0");
  c->cfg = *f;
  c->merge_algo = (f->reserved_i[0] == 1) ? 0 : 1;  // reserved_i[0] == 1 selects the brute-force merge (debug)
  c->have_cfg = true;
  return RFSB200_OK;
}

int rfsb200_upload_maps(rfsb200_ctx* c, const int32_t* count, const double* mean, const double* cov, const double* w) {
  if (!c || !count) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  long long total = 0;
  for (int i = 0; i < c->N; i++) {
    if (count[i] < 0 || count[i] > c->dims.gm_capacity)
      return fail(c, RFSB200_ECAPACITY, "particle %d has %d Gaussians (gm_capacity %d)", i, count[i], c->dims.gm_capacity);
    total += count[i];
  }
  if (total > 0 && (!mean || !cov || !w)) return fail(c, RFSB200_EINVAL, "NULL map arrays");
  StateBuf& s = c->st[c->front];
  CU(c, cudaMemcpyAsync(s.cnt, count, (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
  scan_counts_kernel<<<1, 1024, 0, c->stream>>>(s.cnt, c->offs, c->N);
  double* dm = c->stg;
  double* dc = dm + (size_t)c->N * c->cap * c->ld;
  double* dw = dc + (size_t)c->N * c->cap * c->nc;
  if (total > 0) {
    CU(c, cudaMemcpyAsync(dm, mean, (size_t)total * c->ld * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(dc, cov, (size_t)total * c->nc * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(dw, w, (size_t)total * 8, cudaMemcpyHostToDevice, c->stream));
    const int blocks = (c->N * 32 + 255) / 256;
    if (c->prec == 32) pack_soa_kernel<float><<<blocks, 256, 0, c->stream>>>(s.cnt, c->offs, dm, dc, dw, (float*)s.gm, c->N, c->cap, c->ld, c->nc);
    else pack_soa_kernel<double><<<blocks, 256, 0, c->stream>>>(s.cnt, c->offs, dm, dc, dw, (double*)s.gm, c->N, c->cap, c->ld, c->nc);
  }
  CU(c, cudaGetLastError());
  c->last_out = c->front;
  c->have_maps = true;
  return RFSB200_OK;
}

int rfsb200_append_gaussians(rfsb200_ctx* c, const int32_t* count, const double* mean, const double* cov, const double* w) {
  if (!c || !count) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (!c->have_maps) return fail(c, RFSB200_ESTATE, "append_gaussians before upload_maps");
  CU(c, cudaSetDevice(c->device));
  long long total = 0;
  for (int i = 0; i < c->N; i++) {
    if (count[i] < 0) return fail(c, RFSB200_EINVAL, "negative count for particle %d", i);
    total += count[i];
  }
  if (total == 0) return RFSB200_OK;
  if (total > (long long)c->N * c->cap) return fail(c, RFSB200_ECAPACITY, "%lld Gaussians exceed the staging capacity", total);
  if (!mean || !cov || !w) return fail(c, RFSB200_EINVAL, "NULL map arrays");
  StateBuf& s = c->st[c->front];
  int* add_dev = c->src_dev;   // [2N] ints of scratch; the resample sources are not live between calls
  CU(c, cudaMemcpyAsync(add_dev, count, (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
  scan_counts_kernel<<<1, 1024, 0, c->stream>>>(add_dev, c->offs, c->N);
  double* dm = c->stg;
  double* dc = dm + (size_t)c->N * c->cap * c->ld;
  double* dw = dc + (size_t)c->N * c->cap * c->nc;
  CU(c, cudaMemcpyAsync(dm, mean, (size_t)total * c->ld * 8, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(dc, cov, (size_t)total * c->nc * 8, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(dw, w, (size_t)total * 8, cudaMemcpyHostToDevice, c->stream));
  const int blocks = (c->N * 32 + 255) / 256;
  if (c->prec == 32) append_soa_kernel<float><<<blocks, 256, 0, c->stream>>>(add_dev, c->offs, dm, dc, dw, (float*)s.gm, s.cnt, c->flags, c->N, c->cap, c->ld, c->nc);
  else append_soa_kernel<double><<<blocks, 256, 0, c->stream>>>(add_dev, c->offs, dm, dc, dw, (double*)s.gm, s.cnt, c->flags, c->N, c->cap, c->ld, c->nc);
  CU(c, cudaGetLastError());
  CU(c, cudaStreamSynchronize(c->stream));   // the caller's (pageable) buffers may go away
  c->last_out = c->front;
  return RFSB200_OK;
}

int rfsb200_set_poses(rfsb200_ctx* c, const double* pose, const double* pose_cov, int mode, const double* weight) {
  if (!c || !pose) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (mode < 0 || mode > 2 || (mode > 0 && !pose_cov)) return fail(c, RFSB200_EINVAL, "bad pose_cov_mode");
  CU(c, cudaSetDevice(c->device));
  double* dp = c->stg_small;
  if (weight) { if (int rcp = resolve_pending(c)) return rcp; }   // (the new weights replace normalised ones)
  double* dcov = dp + (size_t)c->N * 3;
  CU(c, cudaMemcpyAsync(dp, pose, (size_t)c->N * 3 * 8, cudaMemcpyHostToDevice, c->stream));
  if (mode == 1) CU(c, cudaMemcpyAsync(dcov, pose_cov, 6 * 8, cudaMemcpyHostToDevice, c->stream));
  if (mode == 2) CU(c, cudaMemcpyAsync(dcov, pose_cov, (size_t)c->N * 6 * 8, cudaMemcpyHostToDevice, c->stream));
  const int blocks = (c->N + 255) / 256;
  if (c->prec == 32) pose_convert_kernel<float><<<blocks, 256, 0, c->stream>>>(dp, dcov, mode, (float*)c->pose, (float*)c->pose_cov, c->N);
  else pose_convert_kernel<double><<<blocks, 256, 0, c->stream>>>(dp, dcov, mode, (double*)c->pose, (double*)c->pose_cov, c->N);
  CU(c, cudaGetLastError());
  c->pose_cov_mode = mode;
  if (weight) CU(c, cudaMemcpyAsync(c->st[c->front].weight, weight, (size_t)c->N * 8, cudaMemcpyHostToDevice, c->stream));
  c->have_poses = true;
  return RFSB200_OK;
}

// queue one update on the ctx stream (Z copy + kernels); no synchronisation
// Murty compatibility, after the update kernel: fetch the partitions it wrote out, replace each exact sum by the sum of
// the 200 best assignments (murty_compat.hpp), patch the weights of the particles concerned and add the weights up again.
static int murty_postprocess(rfsb200_ctx* c, int out_idx) {
  unsigned int cnt[2] = {0, 0};
  CU(c, cudaMemcpyAsync(cnt, c->murty_count, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->last_murty_terms = 0;
  if (cnt[1] == 0) return RFSB200_OK;
  const unsigned words = std::min(cnt[0], MURTY_WORDS);
  std::vector<unsigned long long> rec(words);
  CU(c, cudaMemcpy(rec.data(), c->murty_buf, (size_t)words * 8, cudaMemcpyDeviceToHost));
  const double log_kappa = log(c->model.clutter_intensity);
  std::map<int, double> log_ratio;   // particle -> sum over its large partitions of log(truncated / exact)
  std::vector<double> pd, L;
  size_t at = 0;
  for (unsigned k = 0; k < cnt[1] && at + 3 <= words; k++) {
    const int pi = (int)(rec[at] & 0xffffffffull), nR = (int)(rec[at] >> 32), nC = (int)rec[at + 1];
    double plog;
    memcpy(&plog, &rec[at + 2], 8);
    const size_t need = 3 + (size_t)nR + (size_t)nR * nC;
    if (nR <= 0 || nC <= 0 || at + need > words || pi < 0 || pi >= c->N) break;   // (a record cut off by the buffer's end)
    pd.resize(nR);
    L.resize((size_t)nR * nC);
    memcpy(pd.data(), &rec[at + 3], (size_t)nR * 8);
    memcpy(L.data(), &rec[at + 3 + nR], (size_t)nR * nC * 8);
    int terms = 0;
    const double s200 = murty::k_best_sum(L.data(), pd.data(), nR, nC, log_kappa, 200, &terms);
    c->last_murty_terms += terms;
    log_ratio[pi] += log(s200) - plog;
    at += need;
  }
  std::vector<int> idx;
  std::vector<double> ratio;
  for (auto& kv : log_ratio) { idx.push_back(kv.first); ratio.push_back(exp(kv.second)); }
  const int n = (int)idx.size();
  if (n > 0) {
    CU(c, cudaMemcpyAsync(c->murty_idx, idx.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->murty_ratio, ratio.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    murty_patch_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->st[out_idx].weight, c->murty_idx, c->murty_ratio, n);
    CU(c, cudaGetLastError());
    weight_sums_kernel<<<1, 512, 0, c->stream>>>(c->st[out_idx].weight, c->N, c->sums);
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));   // idx / ratio are pageable host vectors
  }
  return RFSB200_OK;
}

// the normalisation a RFSB200_UPDATE_DEFER_NORMALIZE step left open (see KParams::comm_defer)
static int resolve_pending(rfsb200_ctx* c) {
  if (!c->comm_pending) return RFSB200_OK;
  CommPeers peers{};
  for (int r = 0; r < 8; r++) peers.p[r] = c->comm_peer[r];
  comm_resolve_kernel<<<1, 1024, 0, c->stream>>>(peers, c->comm_rank, c->comm_world, c->comm_pending_epoch,
                                                 c->st[c->comm_pending_buf].weight, c->N, c->sums, c->comm_error, c->comm_timeout_ns);
  CU(c, cudaGetLastError());
  c->comm_pending = false;
  return RFSB200_OK;
}

static int enqueue_update(rfsb200_ctx* c, const double* Z, int32_t nZ, uint32_t flags, bool timed, int* launches_out) {
  if (!c->have_model || !c->have_cfg || !c->have_maps || !c->have_poses)
    return fail(c, RFSB200_ESTATE, "update before set_model / set_filter_cfg / upload_maps / set_poses");
  if (nZ < 0 || nZ > c->dims.z_capacity) return fail(c, RFSB200_ECAPACITY, "nZ %d > z_capacity %d", nZ, c->dims.z_capacity);
  if (!Z) return fail(c, RFSB200_EINVAL, "NULL Z");
  // the measurement batch travels by value in the kernel's parameter block (fp64; converted by the kernel, whose first
  // CTA also leaves the device copy the births of the next predict read): no copy, no staging slot
  for (int k = 0; k < c->ld * nZ; k++) c->Zval[k] = Z[k];
  int launches = 0;
  {  // launch configuration (occupancy queries on the first call of a mode): host work, kept out of the timed span
    const int mf = c->cfg.use_cluster_process ? 0 : 1;
    const int rc0 = (c->prec == 32) ? configure_launch<float>(c, mf) : configure_launch<double>(c, mf);
    if (rc0) return rc0;
  }
  if ((flags & RFSB200_UPDATE_DEFER_NORMALIZE) && !(flags & RFSB200_UPDATE_FUSED_ALLREDUCE))
    return fail(c, RFSB200_EINVAL, "RFSB200_UPDATE_DEFER_NORMALIZE needs RFSB200_UPDATE_FUSED_ALLREDUCE");
  // Sums a deferred step left open.  The update kernels pick the pairs up during set-up (that wait is what keeps the ranks
  // within one step of each other) and, if the open weights are the ones this step reads (the committed state), divide
  // them by the total on the way in; if they are those of the back buffer (a NO_COMMIT step) this launch overwrites them
  // and there is nothing to apply.  Everything else closes the open normalisation first.
  const bool open_front = c->comm_pending && c->comm_pending_buf == c->front;
  c->consume_pending = c->comm_pending && !c->hin_weight &&
                       !(open_front && (flags & RFSB200_UPDATE_NO_COMMIT) && (flags & RFSB200_UPDATE_DEFER_NORMALIZE));   // (two would be open)
  c->consume_scale = c->consume_pending && open_front;
  if (c->comm_pending && !c->consume_pending) {
    const int rc0 = resolve_pending(c);
    if (rc0) return rc0;
    launches++;
  }
  if (timed) CU(c, cudaEventRecord(c->ev0, c->stream));
  const int out_idx = c->front ^ 1;
  if ((flags & RFSB200_UPDATE_FUSED_ALLREDUCE) && c->comm_world > 1 && !c->comm_peer[0])
    return fail(c, RFSB200_ESTATE, "RFSB200_UPDATE_FUSED_ALLREDUCE before rfsb200_comm_connect");
  if ((flags & RFSB200_UPDATE_FUSED_ALLREDUCE) && c->comm_world > 1 && murty_active(c))
    return fail(c, RFSB200_EUNSUPPORTED, "murty_compat needs the weights on the host before they are summed: use "
                                         "RFSB200_UPDATE_NO_NORMALIZE + all-reduce + rfsb200_normalize across ranks");
  int rc = (c->prec == 32) ? launch_update<float>(c, nZ, out_idx, flags) : launch_update<double>(c, nZ, out_idx, flags);
  const bool consumed = c->consume_pending;
  c->consume_pending = false;
  if (rc) return rc;
  if (consumed && !(c->consume_scale && (flags & RFSB200_UPDATE_NO_COMMIT))) c->comm_pending = false;   // (an uncommitted step leaves the state open)
  if ((flags & RFSB200_UPDATE_FUSED_ALLREDUCE) && (flags & RFSB200_UPDATE_DEFER_NORMALIZE) && c->comm_world > 1 && !murty_active(c)) {
    c->comm_pending = true;
    c->comm_pending_epoch = c->comm_epoch;   // (already advanced by the launch)
    c->comm_pending_buf = out_idx;
  }
  launches++;
  c->last_out = out_idx;
  c->last_nZ = nZ;
  if (murty_active(c)) {
    rc = murty_postprocess(c, out_idx);
    if (rc) return rc;
    flags &= ~RFSB200_UPDATE_FUSED_ALLREDUCE;   // (one rank: the normalisation below replaces the fused one)
  }
  if (!(flags & (RFSB200_UPDATE_NO_NORMALIZE | RFSB200_UPDATE_FUSED_ALLREDUCE))) {
    normalize_kernel<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->st[out_idx].weight, c->sums, c->N, c->w_host);
    CU(c, cudaGetLastError());
    launches++;
  }
  if (!(flags & RFSB200_UPDATE_NO_COMMIT)) c->front = out_idx;
  if (timed) CU(c, cudaEventRecord(c->ev1, c->stream));
  *launches_out = launches;
  return RFSB200_OK;
}

// queue the D2H copy of the step scalars; fill `out` after the caller has synchronised
static int enqueue_stats(rfsb200_ctx* c) {
  unsigned char* h = c->hpin + 8192;
  CU(c, cudaMemcpyAsync(h, c->sums, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h + 16, c->stats_out, 104, cudaMemcpyDeviceToHost, c->stream));
  return RFSB200_OK;
}
static int fill_stats(rfsb200_ctx* c, int launches, rfsb200_step_out* out) {
  unsigned char* h = c->hpin + 8192;
  const double* s = (const double*)h;
  const unsigned long long* t = (const unsigned long long*)(h + 16);
  out->sum_w = s[0];
  out->sum_w2 = s[1];
  out->n_eff = s[1] > 0 ? s[0] * s[0] / s[1] : 0;
  out->gm_total_in = (int64_t)t[0];
  out->gm_total_out = (int64_t)t[1];
  out->gm_max_out = (int32_t)t[2];
  out->n_overflow = (int32_t)t[3];
  out->n_murty = (int32_t)t[4];
  out->n_merge_redo = (int32_t)t[5];
  for (int k = 0; k < 6; k++) out->reserved[k] = (int32_t)t[6 + k];   // merge diagnostics
  out->n_launches = launches;
  float ms = 0;
  CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  out->elapsed_us = ms * 1000.f;
  return RFSB200_OK;
}

int rfsb200_update(rfsb200_ctx* c, const double* Z, int32_t nZ, uint32_t flags, rfsb200_step_out* out) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  if (out) memset(out, 0, sizeof(*out));
  if (nZ == 0) {   // include/RBPHDFilter.hpp:451-452 (Q11)
    if (!c->have_model || !c->have_cfg || !c->have_maps || !c->have_poses)
      return fail(c, RFSB200_ESTATE, "update before set_model / set_filter_cfg / upload_maps / set_poses");
    return RFSB200_OK;
  }
  CU(c, cudaSetDevice(c->device));
  int launches = 0;
  int rc = enqueue_update(c, Z, nZ, flags, out != nullptr, &launches);
  if (rc) return rc;
  if (out) {
    rc = enqueue_stats(c);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    return fill_stats(c, launches, out);
  }
  return RFSB200_OK;
}

int rfsb200_update_host(rfsb200_ctx* c, const double* pose, const double* pose_cov, int mode, const double* weight,
                        const double* Z, int32_t nZ, uint32_t flags, double* w_out, uint64_t* unused_out,
                        int32_t* nfov_out, rfsb200_step_out* out) {
  if (!c || !pose) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (out) memset(out, 0, sizeof(*out));
  if (flags & RFSB200_UPDATE_DEFER_NORMALIZE) return fail(c, RFSB200_EINVAL, "the host-facing step returns normalised weights: no RFSB200_UPDATE_DEFER_NORMALIZE");
  if (c->comm_pending) {
    CU(c, cudaSetDevice(c->device));
    if (int rcp = resolve_pending(c)) return rcp;
  }
  if (c->zero_copy && !murty_active(c) && nZ > 0 && nZ <= c->dims.z_capacity && Z && mode >= 0 && mode <= 2 && (mode == 0 || pose_cov)) {
    // Pinned (device-accessible) caller buffers: ONE small kernel reads poses / weights / covariance / Z from host
    // memory, and the update kernel (or normalize_kernel) stores weights, unused masks, in-FOV counts and the step
    // scalars straight into the caller's buffers: no copies on either side of the update, one synchronisation.
    // Any pageable buffer (or RFSB200_ZERO_COPY=0) takes the staged path below.
    CU(c, cudaSetDevice(c->device));
    const double* d_pose = device_view(pose);
    const double* d_w = device_view(weight);
    const double* d_cov = mode == 2 ? device_view(pose_cov) : nullptr;
    double* d_wout = device_view(w_out);
    unsigned long long* d_unused = device_view((unsigned long long*)unused_out);
    int* d_nfov = device_view((int*)nfov_out);
    const bool ok = d_pose && (!weight || d_w) && (mode != 2 || d_cov) && (!w_out || d_wout) && (!unused_out || d_unused) &&
                    (!nfov_out || d_nfov);
    if (ok) {
      int rc = ensure_pinned(c, 1 << 16);
      if (rc) return rc;
      unsigned long long* stats_h = (unsigned long long*)(c->hpin + 8192);   // where fill_stats reads the step scalars
      volatile unsigned long long* done_h = (volatile unsigned long long*)(c->hpin + 8192 + 256);
      const bool fused_in = c->ld == 2;   // the 2-D kernels read the inputs themselves; the Victoria Park kernels get them converted
      if (fused_in) {
        c->hin_pose = d_pose; c->hin_weight = d_w; c->hin_pcov = d_cov;
        if (mode == 1) for (int k = 0; k < 6; k++) c->hin_cov6[k] = pose_cov[k];
      } else {
        HostInParams h{};
        h.pose = d_pose; h.weight = d_w; h.pcov = d_cov; h.mode = mode; h.N = c->N;
        if (mode == 1) for (int k = 0; k < 6; k++) h.cov6[k] = pose_cov[k];
        double* w_dev = c->st[c->front].weight;
        if (c->prec == 32) host_in_kernel<float><<<(c->N + 255) / 256, 256, 0, c->stream>>>(h, c->stg_small, (float*)c->pose, (float*)c->pose_cov, w_dev);
        else host_in_kernel<double><<<(c->N + 255) / 256, 256, 0, c->stream>>>(h, c->stg_small, (double*)c->pose, (double*)c->pose_cov, w_dev);
        CU(c, cudaGetLastError());
      }
      c->pose_cov_mode = mode;
      c->have_poses = true;
      c->w_host = d_wout; c->unused_host = d_unused; c->nfov_host = d_nfov;
      c->stats_host = out ? device_view(stats_h) : nullptr;
      const bool stats_direct = out && c->stats_host;
      // completion word: the last CTA of the update kernel sets it when every result is in host memory — possible when
      // that kernel is the last launch of the step (no separate normalisation) and nothing else has to be copied back
      const bool poll = (flags & (RFSB200_UPDATE_NO_NORMALIZE | RFSB200_UPDATE_FUSED_ALLREDUCE)) && (!out || stats_direct) &&
                        !(flags & RFSB200_UPDATE_STAGE_TIMES);
      c->done_host = poll ? device_view((unsigned long long*)done_h) : nullptr;
      const unsigned long long seq = ++c->done_seq;
      int l = 0;
      rc = enqueue_update(c, Z, nZ, flags, out != nullptr, &l);
      const bool polling = c->done_host != nullptr;
      c->w_host = nullptr; c->unused_host = nullptr; c->nfov_host = nullptr; c->stats_host = nullptr;
      c->hin_pose = nullptr; c->hin_weight = nullptr; c->hin_pcov = nullptr; c->done_host = nullptr;
      if (rc) return rc;
      if (out && !stats_direct) {
        rc = enqueue_stats(c);
        if (rc) return rc;
      }
      if (polling) {
        // wait on the word instead of the stream (a stream synchronisation costs several microseconds of wake-up);
        // the stream is queried now and then so that a failed launch cannot hang the caller
        unsigned spins = 0;
        while (*done_h != seq) {
          if ((++spins & 0x3fffu) == 0u) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) { CU(c, q); }
          }
#if defined(__x86_64__) || defined(_M_X64)
          __builtin_ia32_pause();
#endif
        }
        if (out) {   // the events of the timed span: recorded long ago, complete with the kernel
          CU(c, cudaEventSynchronize(c->ev1));
        }
      } else {
        CU(c, cudaStreamSynchronize(c->stream));
      }
      if (out) return fill_stats(c, (fused_in ? 0 : 1) + l, out);
      return RFSB200_OK;
    }
  }
  int rc = rfsb200_set_poses(c, pose, pose_cov, mode, weight);   // queued, no synchronisation
  if (rc) return rc;
  int launches = 1;   // pose_convert_kernel
  if (nZ > 0) {
    int l = 0;
    rc = enqueue_update(c, Z, nZ, flags, out != nullptr, &l);
    if (rc) return rc;
    launches += l;
  } else if (nZ < 0) {
    return fail(c, RFSB200_EINVAL, "negative nZ");
  }
  const int which = c->last_out;
  if (w_out) CU(c, cudaMemcpyAsync(w_out, c->st[nZ > 0 ? which : c->front].weight, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (unused_out) CU(c, cudaMemcpyAsync(unused_out, c->unused, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (nfov_out) CU(c, cudaMemcpyAsync(nfov_out, c->nfov, (size_t)c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  if (out && nZ > 0) {
    rc = enqueue_stats(c);
    if (rc) return rc;
  }
  CU(c, cudaStreamSynchronize(c->stream));
  if (out && nZ > 0) return fill_stats(c, launches, out);
  return RFSB200_OK;
}

int rfsb200_predict_maps(rfsb200_ctx* c, const double* Q, int32_t add_births, double birth_w) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  if (!c->have_maps) return fail(c, RFSB200_ESTATE, "predict_maps before upload_maps");
  if (add_births < 0) {   // the host consumed the unused measurements itself (candidate-list births)
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemsetAsync(c->unused, 0, (size_t)c->N * 8, c->stream));
    add_births = 0;
  }
  if (c->last_nZ == 0) add_births = 0;   // no update yet: there is no unused measurement to give birth from
  if (add_births && (!c->have_model || !c->have_poses))
    return fail(c, RFSB200_ESTATE, "births need set_model and set_poses (pose and R of the last update)");
  CU(c, cudaSetDevice(c->device));
  c->last_out = c->front;
  if (c->ld == 3) return c->prec == 32 ? do_predict_vp<float>(c, Q, add_births, birth_w) : do_predict_vp<double>(c, Q, add_births, birth_w);
  return c->prec == 32 ? do_predict<float>(c, Q, add_births, birth_w) : do_predict<double>(c, Q, add_births, birth_w);
}

static int ensure_candidates(rfsb200_ctx* c) {
  if (c->cand[0]) return RFSB200_OK;
  const size_t rec = (size_t)cand_rec(c->ld);
  for (int k = 0; k < 2; k++) {
    CU(c, cudaMalloc(&c->cand[k], (size_t)c->N * RFSB200_BIRTH_CAND_CAP * rec * 8));
    CU(c, cudaMalloc(&c->cand_n[k], (size_t)c->N * 4));
    CU(c, cudaMemsetAsync(c->cand_n[k], 0, (size_t)c->N * 4, c->stream));
  }
  c->cand_front = 0;
  return RFSB200_OK;
}

int rfsb200_birth_candidates(rfsb200_ctx* c, const rfsb200_birth_cfg* b, const int32_t* parent) {
  if (!c || !b) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (!c->have_maps) return fail(c, RFSB200_ESTATE, "birth_candidates before upload_maps");
  if (c->last_nZ > 0 && !c->have_model) return fail(c, RFSB200_ESTATE, "birth_candidates before set_model");
  if (!(b->support_dist >= 0.0)) return fail(c, RFSB200_EINVAL, "support_dist must be >= 0");
  CU(c, cudaSetDevice(c->device));
  if (int rc = ensure_candidates(c)) return rc;
  if (c->last_nZ > 0 && !c->have_poses) return fail(c, RFSB200_ESTATE, "birth_candidates needs the poses of the last update");
  int n_pass = 1;
  if (parent) {
    // level[i] = how many lower parent slots have to be finished before particle i can take over its list
    std::vector<int> level(c->N);
    for (int i = 0; i < c->N; i++) {
      if (parent[i] < 0 || parent[i] >= c->N) return fail(c, RFSB200_EINVAL, "parent %d of particle %d out of range", parent[i], i);
      level[i] = parent[i] < i ? level[parent[i]] + 1 : 0;
      n_pass = std::max(n_pass, level[i] + 1);
    }
    CU(c, cudaMemcpyAsync(c->src_dev, parent, (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->src_dev + c->N, level.data(), (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));   // parent[] and level[] are pageable host buffers
  }
  BirthCandParams p{};
  const StateBuf& st = c->st[c->front];
  p.N = c->N; p.cap = c->cap; p.cand_cap = RFSB200_BIRTH_CAND_CAP; p.nZ = c->last_nZ;
  p.pcov_mode = c->ld == 2 ? c->pose_cov_mode : 0;
  p.parent = parent ? c->src_dev : nullptr;
  p.level = parent ? c->src_dev + c->N : nullptr;
  p.pose64 = c->stg_small; p.pcov = c->pose_cov;
  p.unused = c->unused; p.nfov = c->nfov;
  p.cand_in = c->cand[c->cand_front]; p.cand_n_in = c->cand_n[c->cand_front];
  p.cand_out = c->cand[c->cand_front ^ 1]; p.cand_n_out = c->cand_n[c->cand_front ^ 1];
  p.gm = st.gm; p.cnt = st.cnt; p.flags = c->flags;
  for (int k = 0; k < 9; k++) p.R[k] = c->model.R[k];
  p.Slb = c->model.Slb; p.range_min = c->model.range_min; p.range_max = c->model.range_max;
  p.thr_r = c->model.innov_thr_range; p.thr_b = c->model.innov_thr_bearing;
  p.support_d2 = b->support_dist * b->support_dist; p.birth_w = b->birth_weight;
  p.count_thr = b->count_threshold; p.check_thr = b->check_threshold; p.cur_thr = b->current_count_threshold;
  const int md = c->ld;   // measurement dimension of both built-in plugin sets
  for (int k = 0; k < c->last_nZ * md; k++) p.Z[k] = c->Zval[k];
  const int blocks = (c->N + 127) / 128;
  for (int pass = 0; pass < n_pass; pass++) {
    p.pass = pass;
    if (c->ld == 3) {
      if (c->prec == 32) birth_candidates_kernel<float, 3><<<blocks, 128, 0, c->stream>>>(p);
      else birth_candidates_kernel<double, 3><<<blocks, 128, 0, c->stream>>>(p);
    } else {
      if (c->prec == 32) birth_candidates_kernel<float, 2><<<blocks, 128, 0, c->stream>>>(p);
      else birth_candidates_kernel<double, 2><<<blocks, 128, 0, c->stream>>>(p);
    }
    CU(c, cudaGetLastError());
  }
  c->cand_front ^= 1;
  c->last_out = c->front;
  return RFSB200_OK;
}

int rfsb200_get_birth_candidates(rfsb200_ctx* c, int32_t* n, double* mean, double* cov, int32_t* support, int32_t* checks) {
  if (!c || !n) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  if (int rc = ensure_candidates(c)) return rc;
  const int rec = cand_rec(c->ld), cap = RFSB200_BIRTH_CAND_CAP;
  std::vector<double> h((size_t)c->N * cap * rec);
  CU(c, cudaMemcpyAsync(n, c->cand_n[c->cand_front], (size_t)c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h.data(), c->cand[c->cand_front], h.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < c->N; i++)
    for (int k = 0; k < n[i] && k < cap; k++) {
      const size_t s = (size_t)i * cap + k;
      const double* r = &h[s * rec];
      if (mean) for (int d = 0; d < c->ld; d++) mean[s * c->ld + d] = r[d];
      if (cov) for (int q = 0; q < c->nc; q++) cov[s * c->nc + q] = r[c->ld + q];
      if (support) support[s] = (int32_t)r[rec - 2];
      if (checks) checks[s] = (int32_t)r[rec - 1];
    }
  return RFSB200_OK;
}

int rfsb200_set_birth_candidates(rfsb200_ctx* c, const int32_t* n, const double* mean, const double* cov, const int32_t* support,
                                 const int32_t* checks) {
  if (!c || !n) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  if (int rc = ensure_candidates(c)) return rc;
  const int rec = cand_rec(c->ld), cap = RFSB200_BIRTH_CAND_CAP;
  std::vector<double> h((size_t)c->N * cap * rec, 0.0);
  for (int i = 0; i < c->N; i++) {
    if (n[i] < 0 || n[i] > cap) return fail(c, RFSB200_ECAPACITY, "%d candidates for particle %d (capacity %d)", n[i], i, cap);
    if (n[i] > 0 && (!mean || !cov || !support || !checks)) return fail(c, RFSB200_EINVAL, "NULL candidate arrays");
    for (int k = 0; k < n[i]; k++) {
      const size_t s = (size_t)i * cap + k;
      double* r = &h[s * rec];
      for (int d = 0; d < c->ld; d++) r[d] = mean[s * c->ld + d];
      for (int q = 0; q < c->nc; q++) r[c->ld + q] = cov[s * c->nc + q];
      r[rec - 2] = (double)support[s];
      r[rec - 1] = (double)checks[s];
    }
  }
  CU(c, cudaMemcpyAsync(c->cand_n[c->cand_front], n, (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->cand[c->cand_front], h.data(), h.size() * 8, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

// lower Cholesky factor as Eigen::LLT computes it (n x n, row-major in / out, zero padded to 3x3)
static void cholesky_lower(const double* A, int n, double* L) {
  for (int k = 0; k < 9; k++) L[k] = 0;
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * 3 + k] * L[j * 3 + k];
    const double d = sqrt(s);
    L[j * 3 + j] = d;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * 3 + k] * L[j * 3 + k];
      L[i * 3 + j] = t / d;
    }
  }
}

int rfsb200_propagate(rfsb200_ctx* c, const rfsb200_motion_desc* m) {
  if (!c || !m) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (m->model_id != RFSB200_MOTION_ODOMETRY2D && m->model_id != RFSB200_MOTION_ACKERMAN2D)
    return fail(c, RFSB200_EUNSUPPORTED, "motion model id %d (MotionModel_Odometry2d = 1, MotionModel_Ackerman2d = 2)", m->model_id);
  if (!c->have_poses) return fail(c, RFSB200_ESTATE, "propagate before set_poses");
  if (m->model_id == RFSB200_MOTION_ACKERMAN2D && !(m->ackerman_l != 0)) return fail(c, RFSB200_EINVAL, "ackerman_l must be non-zero");
  CU(c, cudaSetDevice(c->device));
  MotionParams p{};
  p.model_id = m->model_id;
  p.n_in = m->model_id == RFSB200_MOTION_ODOMETRY2D ? 3 : 2;
  bool q_nonzero = false;
  for (int k = 0; k < 9; k++) q_nonzero = q_nonzero || (m->Q[k] != 0.0);
  p.use_model_noise = (m->use_model_noise && q_nonzero) ? 1 : 0;   // ProcessModel.hpp:145: only if Q_ != 0
  p.use_input_noise = m->use_input_noise ? 1 : 0;
  if (p.use_model_noise) cholesky_lower(m->Q, 3, p.LQ);
  if (p.use_input_noise) cholesky_lower(m->input_cov, p.n_in, p.Lu);
  for (int k = 0; k < 3; k++) p.u[k] = m->input[k];
  p.dt = m->dt; p.h = m->ackerman_h; p.l = m->ackerman_l; p.pdx = m->ackerman_dx; p.pdy = m->ackerman_dy;
  p.seed = m->seed; p.step = m->step_counter;
  const int blocks = (c->N + 255) / 256;
  if (c->prec == 32) propagate_kernel<float><<<blocks, 256, 0, c->stream>>>(c->stg_small, (float*)c->pose, p, c->N);
  else propagate_kernel<double><<<blocks, 256, 0, c->stream>>>(c->stg_small, (double*)c->pose, p, c->N);
  CU(c, cudaGetLastError());
  // the pose covariance after sample(): Q for every particle with model noise, none otherwise (Q1)
  if (p.use_model_noise) {
    unsigned char* h = c->hpin + 57344;
    CU(c, cudaStreamSynchronize(c->stream));   // staging area reuse
    const double q6[6] = {m->Q[0], m->Q[1], m->Q[2], m->Q[4], m->Q[5], m->Q[8]};
    if (c->prec == 32) { float* f = (float*)h; for (int k = 0; k < 6; k++) f[k] = (float)q6[k]; f[6] = f[7] = 0.f; }
    else { double* d = (double*)h; for (int k = 0; k < 6; k++) d[k] = q6[k]; d[6] = d[7] = 0.0; }
    CU(c, cudaMemcpyAsync(c->pose_cov, h, 8 * c->tsize, cudaMemcpyHostToDevice, c->stream));
    c->pose_cov_mode = 1;
  } else {
    c->pose_cov_mode = 0;
  }
  return RFSB200_OK;
}

int rfsb200_get_poses(rfsb200_ctx* c, double* pose) {
  if (!c || !pose) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (!c->have_poses) return fail(c, RFSB200_ESTATE, "get_poses before set_poses");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(pose, c->stg_small, (size_t)c->N * 3 * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

int rfsb200_resample(rfsb200_ctx* c, const int32_t* map_src, const int32_t* aux_src, const double* weight) {
  if (!c || !map_src) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (!c->have_maps) return fail(c, RFSB200_ESTATE, "resample before upload_maps");
  for (int i = 0; i < c->N; i++)
    if (map_src[i] < 0 || map_src[i] >= c->N || (aux_src && (aux_src[i] < -1 || aux_src[i] >= c->N)))
      return fail(c, RFSB200_EINVAL, "resample source %d of particle %d out of range", map_src[i], i);
  CU(c, cudaSetDevice(c->device));
  if (int rcp = resolve_pending(c)) return rcp;
  CU(c, cudaMemcpyAsync(c->src_dev, map_src, (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
  if (aux_src) CU(c, cudaMemcpyAsync(c->src_dev + c->N, aux_src, (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
  const StateBuf& in = c->st[c->front];
  const StateBuf& out = c->st[c->front ^ 1];
  const int blocks = (c->N * 32 + 127) / 128;
  const int* asrc = aux_src ? c->src_dev + c->N : nullptr;
  if (c->prec == 32)
    resample_gather_kernel<float><<<blocks, 128, 0, c->stream>>>((const float*)in.gm, in.cnt, c->unused, c->nfov, in.weight, c->src_dev, asrc,
                                                                 (float*)out.gm, out.cnt, c->unused_alt, c->nfov_alt, out.weight,
                                                                 weight ? 1 : 0, weight ? *weight : 0.0,
                                                                 (const float*)c->pose, (const float*)c->pose_cov, (float*)c->pose_alt,
                                                                 (float*)c->pose_cov_alt, c->pose_cov_mode, c->stg_small, c->stg_small + (size_t)c->N * 9, c->N, c->cap, c->npl);
  else
    resample_gather_kernel<double><<<blocks, 128, 0, c->stream>>>((const double*)in.gm, in.cnt, c->unused, c->nfov, in.weight, c->src_dev, asrc,
                                                                  (double*)out.gm, out.cnt, c->unused_alt, c->nfov_alt, out.weight,
                                                                  weight ? 1 : 0, weight ? *weight : 0.0,
                                                                  (const double*)c->pose, (const double*)c->pose_cov, (double*)c->pose_alt,
                                                                  (double*)c->pose_cov_alt, c->pose_cov_mode, c->stg_small, c->stg_small + (size_t)c->N * 9, c->N, c->cap, c->npl);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(c->stg_small, c->stg_small + (size_t)c->N * 9, (size_t)c->N * 3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));   // map_src / aux_src are the caller's (pageable) buffers
  std::swap(c->unused, c->unused_alt);
  std::swap(c->nfov, c->nfov_alt);
  std::swap(c->pose, c->pose_alt);
  if (c->pose_cov_mode == 2) std::swap(c->pose_cov, c->pose_cov_alt);
  c->front ^= 1;
  c->last_out = c->front;
  return RFSB200_OK;
}

static long long record_bytes(const rfsb200_ctx* c) {
  return (long long)c->npl * c->cap * (long long)c->tsize + (long long)sizeof(ParticleHeader);
}

int rfsb200_particle_record_bytes(rfsb200_ctx* c, int64_t* bytes) {
  if (!c || !bytes) return fail(c, RFSB200_EINVAL, "NULL argument");
  *bytes = record_bytes(c);
  return RFSB200_OK;
}

int rfsb200_export_particles(rfsb200_ctx* c, const int32_t* idx, int32_t n, void* dev_buf) {
  if (!c || (n > 0 && (!idx || !dev_buf))) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (n < 0) return fail(c, RFSB200_EINVAL, "negative n");
  if (!c->have_maps || !c->have_poses) return fail(c, RFSB200_ESTATE, "export before upload_maps / set_poses");
  if (n == 0) return RFSB200_OK;
  for (int k = 0; k < n; k++)
    if (idx[k] < 0 || idx[k] >= c->N) return fail(c, RFSB200_EINVAL, "particle index %d out of range", idx[k]);
  CU(c, cudaSetDevice(c->device));
  if (int rcp = resolve_pending(c)) return rcp;
  const StateBuf& s = c->st[c->front];
  // a heavy shard may have to export more copies than it holds particles: in chunks of the 2N-entry index scratch
  const int chunk = 2 * c->N;
  for (int k0 = 0; k0 < n; k0 += chunk) {
    const int m = std::min(chunk, n - k0);
    CU(c, cudaStreamSynchronize(c->stream));   // src_dev is reused as index scratch
    CU(c, cudaMemcpyAsync(c->src_dev, idx + k0, (size_t)m * 4, cudaMemcpyHostToDevice, c->stream));
    unsigned char* dst = (unsigned char*)dev_buf + (size_t)k0 * record_bytes(c);
    const int blocks = (m * 32 + 127) / 128;
    if (c->prec == 32)
      export_particles_kernel<float><<<blocks, 128, 0, c->stream>>>(c->src_dev, m, (const float*)s.gm, s.cnt, s.weight, (const float*)c->pose,
                                                                    (const float*)c->pose_cov, c->pose_cov_mode, c->stg_small, c->unused, c->nfov,
                                                                    dst, record_bytes(c), c->cap, c->npl, c->N);
    else
      export_particles_kernel<double><<<blocks, 128, 0, c->stream>>>(c->src_dev, m, (const double*)s.gm, s.cnt, s.weight, (const double*)c->pose,
                                                                     (const double*)c->pose_cov, c->pose_cov_mode, c->stg_small, c->unused, c->nfov,
                                                                     dst, record_bytes(c), c->cap, c->npl, c->N);
    CU(c, cudaGetLastError());
  }
  CU(c, cudaStreamSynchronize(c->stream));   // idx is the caller's buffer; the records are complete on return
  return RFSB200_OK;
}

int rfsb200_import_particles(rfsb200_ctx* c, const int32_t* slot, int32_t n, const void* dev_buf, double weight) {
  if (!c || (n > 0 && (!slot || !dev_buf))) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (n < 0 || n > c->N) return fail(c, RFSB200_EINVAL, "n %d outside [0, N]", n);
  if (!c->have_maps) return fail(c, RFSB200_ESTATE, "import before upload_maps");
  if (n == 0) return RFSB200_OK;
  for (int k = 0; k < n; k++)
    if (slot[k] < 0 || slot[k] >= c->N) return fail(c, RFSB200_EINVAL, "slot %d out of range", slot[k]);
  CU(c, cudaSetDevice(c->device));
  if (int rcp = resolve_pending(c)) return rcp;
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaMemcpyAsync(c->src_dev, slot, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  StateBuf& s = c->st[c->front];
  const int blocks = (n * 32 + 127) / 128;
  if (c->prec == 32)
    import_particles_kernel<float><<<blocks, 128, 0, c->stream>>>(c->src_dev, n, (float*)s.gm, s.cnt, s.weight, (float*)c->pose, (float*)c->pose_cov,
                                                                  c->pose_cov_mode, c->stg_small, c->unused, c->nfov, (const unsigned char*)dev_buf,
                                                                  record_bytes(c), c->cap, c->npl, c->N, weight);
  else
    import_particles_kernel<double><<<blocks, 128, 0, c->stream>>>(c->src_dev, n, (double*)s.gm, s.cnt, s.weight, (double*)c->pose, (double*)c->pose_cov,
                                                                   c->pose_cov_mode, c->stg_small, c->unused, c->nfov, (const unsigned char*)dev_buf,
                                                                   record_bytes(c), c->cap, c->npl, c->N, weight);
  CU(c, cudaGetLastError());
  CU(c, cudaStreamSynchronize(c->stream));
  c->last_out = c->front;
  c->have_poses = true;
  return RFSB200_OK;
}

int rfsb200_comm_export(rfsb200_ctx* c, void* handle64) {
  if (!c || !handle64) return fail(c, RFSB200_EINVAL, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  CU(c, cudaSetDevice(c->device));
  // a (re)connection starts from epoch 0 with an empty mailbox; cleared HERE, before any peer can hold the handle
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaMemset(c->comm_mail, 0xff, COMM_BANKS * 8 * sizeof(CommSlot)));
  CU(c, cudaMemset(c->comm_error, 0, 4));
  c->comm_epoch = 0;
  c->comm_bar_epoch = 0;
  c->comm_pending = false;
  cudaIpcMemHandle_t h;
  CU(c, cudaIpcGetMemHandle(&h, c->comm_mail));
  memcpy(handle64, &h, 64);
  return RFSB200_OK;
}

int rfsb200_comm_connect(rfsb200_ctx* c, int32_t rank, int32_t world, const void* handles) {
  if (!c || !handles) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(c, RFSB200_EINVAL, "rank %d / world %d out of range (world <= 8)", rank, world);
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r < world; r++) {
    if (r == rank) { c->comm_peer[r] = c->comm_mail; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    CU(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->comm_peer[r] = ptr;
  }
  c->comm_local = false;
  c->comm_rank = rank;
  c->comm_world = world;
  return RFSB200_OK;
}

int rfsb200_comm_connect_local(rfsb200_ctx* c, int32_t rank, int32_t world, rfsb200_ctx* const* peers) {
  if (!c || !peers) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(c, RFSB200_EINVAL, "rank %d / world %d out of range (world <= 8)", rank, world);
  if (peers[rank] != c) return fail(c, RFSB200_EINVAL, "peers[rank] must be this ctx");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  // a (re)connection starts from epoch 0 with an empty mailbox (as rfsb200_comm_export does for the IPC route); every
  // ctx of the group must be connected before the first of them runs a fused update
  CU(c, cudaMemset(c->comm_mail, 0xff, COMM_BANKS * 8 * sizeof(CommSlot)));
  CU(c, cudaMemset(c->comm_error, 0, 4));
  c->comm_epoch = 0;
  c->comm_bar_epoch = 0;
  c->comm_pending = false;
  for (int r = 0; r < world; r++) {
    if (!peers[r]) return fail(c, RFSB200_EINVAL, "NULL peer %d", r);
#if !defined(RFSB200_SIMT_HOST)   // (the host interpreter of tests/simt has one "device")
    if (peers[r]->device != c->device) {
      int can = 0;
      CU(c, cudaDeviceCanAccessPeer(&can, c->device, peers[r]->device));
      if (!can) return fail(c, RFSB200_EUNSUPPORTED, "no peer access from device %d to device %d", c->device, peers[r]->device);
      const cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(c, e);
      (void)cudaGetLastError();
    }
#endif
    c->comm_peer[r] = peers[r]->comm_mail;
  }
  c->comm_local = true;
  c->comm_rank = rank;
  c->comm_world = world;
  return RFSB200_OK;
}

int rfsb200_comm_resolve(rfsb200_ctx* c) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  CU(c, cudaSetDevice(c->device));
  return resolve_pending(c);
}

int rfsb200_comm_barrier(rfsb200_ctx* c) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  if (c->comm_world <= 1) return RFSB200_OK;
  if (!c->comm_peer[0]) return fail(c, RFSB200_ESTATE, "rfsb200_comm_barrier before rfsb200_comm_connect");
  CU(c, cudaSetDevice(c->device));
  CommPeers peers{};
  for (int r = 0; r < 8; r++) peers.p[r] = c->comm_peer[r];
  comm_barrier_kernel<<<1, 32, 0, c->stream>>>(peers, c->comm_rank, c->comm_world, c->comm_bar_epoch + 1, c->comm_error, c->comm_timeout_ns);
  CU(c, cudaGetLastError());
  c->comm_bar_epoch++;
  return RFSB200_OK;
}

int rfsb200_comm_error(rfsb200_ctx* c, int32_t* flag) {
  if (!c || !flag) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  int f = 0;
  CU(c, cudaMemcpyAsync(&f, c->comm_error, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemsetAsync(c->comm_error, 0, 4, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *flag = f;
  return RFSB200_OK;
}

int rfsb200_weight_sums_device(rfsb200_ctx* c, void** p) {
  if (!c || !p) return fail(c, RFSB200_EINVAL, "NULL argument");
  *p = c->sums;
  return RFSB200_OK;
}

int rfsb200_normalize(rfsb200_ctx* c) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  CU(c, cudaSetDevice(c->device));
  if (int rcp = resolve_pending(c)) return rcp;
  normalize_kernel<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->st[c->last_out].weight, c->sums, c->N, nullptr);
  CU(c, cudaGetLastError());
  return RFSB200_OK;
}

static int which_buf(rfsb200_ctx* c, int which) { return which == 0 ? c->front : c->last_out; }

int rfsb200_get_weights(rfsb200_ctx* c, int which, double* w) {
  if (!c || !w) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  if (int rcp = resolve_pending(c)) return rcp;
  CU(c, cudaMemcpyAsync(w, c->st[which_buf(c, which)].weight, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

int rfsb200_get_gm_sizes(rfsb200_ctx* c, int which, int32_t* n) {
  if (!c || !n) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(n, c->st[which_buf(c, which)].cnt, (size_t)c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

int rfsb200_get_map(rfsb200_ctx* c, int which, int32_t i, int32_t cap, int32_t* n, double* mean, double* cov, double* w) {
  if (!c || !n) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (i < 0 || i >= c->N) return fail(c, RFSB200_EINVAL, "particle index %d out of range", i);
  CU(c, cudaSetDevice(c->device));
  const StateBuf& s = c->st[which_buf(c, which)];
  int cnt = 0;
  CU(c, cudaMemcpyAsync(&cnt, s.cnt + i, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *n = cnt;
  if (cnt == 0) return RFSB200_OK;
  if (cnt > cap) return fail(c, RFSB200_ECAPACITY, "particle %d has %d Gaussians, caller capacity %d", i, cnt, cap);
  if (!mean || !cov || !w) return fail(c, RFSB200_EINVAL, "NULL output arrays");
  std::vector<unsigned char> tmp((size_t)c->npl * c->cap * c->tsize);
  CU(c, cudaMemcpyAsync(tmp.data(), (const unsigned char*)s.gm + (size_t)i * c->npl * c->cap * c->tsize, tmp.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < cnt; k++) {
    auto get = [&](int plane) -> double {
      size_t idx = (size_t)plane * c->cap + k;
      return c->prec == 32 ? (double)((const float*)tmp.data())[idx] : ((const double*)tmp.data())[idx];
    };
    for (int d = 0; d < c->ld; d++) mean[c->ld * k + d] = get(d);
    for (int d = 0; d < c->nc; d++) cov[c->nc * k + d] = get(c->ld + d);
    w[k] = get(c->ld + c->nc);
  }
  return RFSB200_OK;
}

int rfsb200_download_maps(rfsb200_ctx* c, int which, int64_t cap_total, int32_t* count, double* mean, double* cov, double* w) {
  if (!c || !count) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  const StateBuf& s = c->st[which_buf(c, which)];
  scan_counts_kernel<<<1, 1024, 0, c->stream>>>(s.cnt, c->offs, c->N);
  double* dm = c->stg;
  double* dc = dm + (size_t)c->N * c->cap * c->ld;
  double* dw = dc + (size_t)c->N * c->cap * c->nc;
  const int blocks = (c->N * 32 + 255) / 256;
  if (c->prec == 32) unpack_soa_kernel<float><<<blocks, 256, 0, c->stream>>>(s.cnt, c->offs, (const float*)s.gm, dm, dc, dw, c->N, c->cap, c->ld, c->nc);
  else unpack_soa_kernel<double><<<blocks, 256, 0, c->stream>>>(s.cnt, c->offs, (const double*)s.gm, dm, dc, dw, c->N, c->cap, c->ld, c->nc);
  CU(c, cudaGetLastError());
  long long total = 0;
  CU(c, cudaMemcpyAsync(count, s.cnt, (size_t)c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(&total, c->offs + c->N, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (total > cap_total) return fail(c, RFSB200_ECAPACITY, "%lld Gaussians do not fit caller capacity %lld", total, (long long)cap_total);
  if (total > 0) {
    if (!mean || !cov || !w) return fail(c, RFSB200_EINVAL, "NULL output arrays");
    CU(c, cudaMemcpyAsync(mean, dm, (size_t)total * c->ld * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(cov, dc, (size_t)total * c->nc * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(w, dw, (size_t)total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return RFSB200_OK;
}

int rfsb200_get_unused(rfsb200_ctx* c, uint64_t* mask, int32_t* nfov) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  CU(c, cudaSetDevice(c->device));
  if (mask) CU(c, cudaMemcpyAsync(mask, c->unused, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (nfov) CU(c, cudaMemcpyAsync(nfov, c->nfov, (size_t)c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

int rfsb200_get_flags(rfsb200_ctx* c, int32_t* flags) {
  if (!c || !flags) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(flags, c->flags, (size_t)c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return RFSB200_OK;
}

int rfsb200_permanent(rfsb200_ctx* c, const double* A, int32_t n, int32_t batch, double* out) {
  if (!c || !A || !out) return fail(c, RFSB200_EINVAL, "NULL argument");
  if (n < 1 || n > 24 || batch < 1) return fail(c, RFSB200_EINVAL, "n must be in [1,24], batch >= 1");
  CU(c, cudaSetDevice(c->device));
  double* dA = nullptr;
  double* dO = nullptr;
  const size_t bytes = (size_t)batch * n * n * 8;
  CU(c, cudaMalloc((void**)&dA, bytes));
  CU(c, cudaMalloc((void**)&dO, (size_t)batch * 8));
  cudaError_t e = cudaMemcpyAsync(dA, A, bytes, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    permanent_kernel<<<(batch * 32 + 127) / 128, 128, 0, c->stream>>>(dA, n, batch, dO);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dO, (size_t)batch * 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(dA);
  cudaFree(dO);
  if (e != cudaSuccess) return fail(c, RFSB200_ECUDA, "rfsb200_permanent: %s", cudaGetErrorString(e));
  return RFSB200_OK;
}

int rfsb200_profile_begin(rfsb200_ctx* c, int32_t max_updates) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  if (max_updates < 0 || max_updates > 4096) return fail(c, RFSB200_EINVAL, "max_updates must be in [0,4096]");
  CU(c, cudaSetDevice(c->device));
  while ((int)c->prof_ev.size() < 2 * max_updates) {
    cudaEvent_t e;
    CU(c, cudaEventCreate(&e));
    c->prof_ev.push_back(e);
  }
  c->prof_cap = max_updates;
  c->prof_n = 0;
  return RFSB200_OK;
}

int rfsb200_get_stage_times(rfsb200_ctx* c, rfsb200_stage_times* out) {
  if (!c) return fail(nullptr, RFSB200_EINVAL, "NULL ctx");
  if (!out) return fail(c, RFSB200_EINVAL, "rfsb200_get_stage_times: NULL argument");
  if (!c->prof_valid || !c->prof_dev) return fail(c, RFSB200_ESTATE, "no update ran with RFSB200_UPDATE_STAGE_TIMES");
  CU(c, cudaSetDevice(c->device));
  unsigned long long h[16];
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaMemcpy(h, c->prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
  *out = rfsb200_stage_times{};
  const double t0 = (double)h[12];
  out->kernel_us = ((double)h[15] - t0) * 1e-3;
  out->setup_us = ((double)h[13] - t0) * 1e-3;
  out->particles_us = ((double)h[14] - (double)h[13]) * 1e-3;
  out->epilogue_us = ((double)h[15] - (double)h[14]) * 1e-3;
  double tot = 0;
  for (int k = 0; k < N_STAGES; k++) tot += (double)h[k];
  out->warp_cycles = tot;
  if (tot > 0) {
    out->share_load = (double)h[STAGE_LOAD] / tot;
    out->share_map_update_kf = (double)h[STAGE_CORRECT] / tot;
    out->share_weighting = ((double)h[STAGE_WEIGHT] + (double)h[STAGE_MFWEIGHT]) / tot;
    out->share_merge = ((double)h[STAGE_MERGE] + (double)h[STAGE_M1] + (double)h[STAGE_M2] + (double)h[STAGE_M3] + (double)h[STAGE_M4]) / tot;
    out->reserved[0] = (double)h[STAGE_M1] / tot;   // inside the merge: cell sort, pair search, clusters, per-cluster greedy loops
    out->reserved[1] = (double)h[STAGE_M2] / tot;
    out->reserved[2] = (double)h[STAGE_M3] / tot;
    out->reserved[3] = (double)h[STAGE_M4] / tot;
    out->share_prune = (double)h[STAGE_PRUNE] / tot;
  }
  out->warps_per_cta = c->prof_nwarps;
  return RFSB200_OK;
}

int rfsb200_profile_read(rfsb200_ctx* c, float* us, int32_t cap, int32_t* n) {
  if (!c || !n) return fail(c, RFSB200_EINVAL, "NULL argument");
  CU(c, cudaSetDevice(c->device));
  *n = c->prof_n;
  for (int k = 0; k < c->prof_n; k++) {
    CU(c, cudaEventSynchronize(c->prof_ev[2 * k + 1]));
    float ms = 0;
    CU(c, cudaEventElapsedTime(&ms, c->prof_ev[2 * k], c->prof_ev[2 * k + 1]));
    if (us && k < cap) us[k] = ms * 1000.f;
  }
  c->prof_cap = 0;
  c->prof_n = 0;
  return RFSB200_OK;
}

int rfsb200_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return fail(nullptr, RFSB200_EINVAL, "NULL argument");
  cudaError_t e = cudaMallocHost(ptr, bytes);
  if (e != cudaSuccess) return fail(nullptr, RFSB200_ENOMEM, "cudaMallocHost(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
  remember_pinned(*ptr, (size_t)bytes);   // rfsb200_update_host finds its device alias without a driver query
  return RFSB200_OK;
}

int rfsb200_host_free(void* ptr) {
  if (ptr) {
    forget_pinned(ptr);
    cudaFreeHost(ptr);
  }
  return RFSB200_OK;
}

}  // extern "C"
