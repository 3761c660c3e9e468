// common.cuh — sm_100a building blocks used by the PHD update kernels:
// 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier, L2 bulk prefetch,
// proxy fences, warp reductions / scans.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfsb200 {

constexpr unsigned FULL = 0xffffffffu;

#if defined(RFSB200_SIMT_HOST)
// tests/simt/ — TEST INFRASTRUCTURE: the kernel sources of this directory interpreted lane by lane on the
// host so that the CPU test suite can check their logic.  The product library is never built this way;
// the wrappers below then come from tests/simt/simt_ptx.h (same names, host emulation of mbarrier / bulk copy).
}  // namespace rfsb200
#include "simt_ptx.h"
namespace rfsb200 {
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA 1-D bulk copies ------------------------------------------------------------------
// global -> shared, completion counted in bytes on an mbarrier. 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// global -> L2 only (no destination, no completion): warms the lines a later bulk load will fetch
__device__ __forceinline__ void tma_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
// order generic-proxy shared-memory accesses against the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- system-scope release / acquire and the global timer (fused cross-GPU sum, step_epilogue) -----------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// relaxed system-scope 8-byte accesses of the flag-in-word mailbox protocol (comm_send / comm_recv in phd_kernels.cuh)
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// device-scope release / acquire (flags in device memory between CTAs of one launch)
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

// ---- warp helpers -------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(FULL, v, o);
    v = u > v ? u : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(FULL, v, o);
    v = u < v ? u : v;
  }
  return v;
}
// inclusive scan of an int across the warp
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// ---- scalar math dispatch (fp32 product / fp64 verification) -----------------------------------
template <typename T> struct M;
template <> struct M<float> {
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __device__ __forceinline__ float log_(float x) { return logf(x); }
  // atan2 to ~1 ulp of pi/2 (1.1e-7 abs): min/max reduction to [0,1], degree-8 minimax in a^2
  static __device__ __forceinline__ float atan2_(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = __fdividef(mn, mx);
    if (mx == 0.f) a = 0.f;
    const float s = a * a;
    float q = 0.0028340641874819994f;
    q = fmaf(q, s, -0.016005029901862144f);
    q = fmaf(q, s, 0.042587608098983765f);
    q = fmaf(q, s, -0.07495445758104324f);
    q = fmaf(q, s, 0.10636754333972931f);
    q = fmaf(q, s, -0.14202570915222168f);
    q = fmaf(q, s, 0.19992484152317047f);
    q = fmaf(q, s, -0.3333306610584259f);
    q = fmaf(q, s, 1.0f);
    float r = q * a;
    if (ay > ax) r = 1.5707963267948966f - r;
    if (x < 0.f) r = 3.14159265358979323846f - r;
    return copysignf(r, y);
  }
  static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  static constexpr float PI = 3.14159265358979323846f;
  static constexpr float TWO_PI = 6.28318530717958647692f;
};
template <> struct M<double> {
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  static constexpr double PI = 3.14159265358979323846;
  static constexpr double TWO_PI = 6.28318530717958647692;
};

}  // namespace rfsb200
