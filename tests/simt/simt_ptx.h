// simt_ptx.h — TEST INFRASTRUCTURE ONLY (see simt.h).  Host emulation of the PTX wrappers of
// rfs-slam_b200/csrc/common.cuh (same names and signatures): mbarrier with transaction count, 1-D bulk copies
// (complete at once), bulk-group stores, proxy fences, system-scope release / acquire, the global timer.
#pragma once
#include <time.h>

namespace rfsb200 {

// the 8 bytes of an mbarrier, as this emulation keeps them
struct SimtMbar {
  uint32_t phase : 1, count : 15, pending : 16;
  int32_t tx;
};
static_assert(sizeof(SimtMbar) == 8, "an mbarrier is one 64-bit word");

inline void simt_mbar_check(SimtMbar* b) {
  if (b->pending == 0 && b->tx == 0) {   // phase complete: flip and re-arm
    b->phase ^= 1u;
    b->pending = b->count;
  }
}
inline uint32_t smem_u32(const void* p) { return (uint32_t)reinterpret_cast<uintptr_t>(p); }
inline void mbar_init(uint64_t* bar, uint32_t count) {
  SimtMbar* b = reinterpret_cast<SimtMbar*>(bar);
  b->phase = 0; b->count = count; b->pending = count; b->tx = 0;
}
inline void fence_mbar_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {   // arrive + expect_tx
  SimtMbar* b = reinterpret_cast<SimtMbar*>(bar);
  b->tx += (int32_t)bytes;
  if (b->pending == 0) { fprintf(stderr, "simt: mbarrier arrive beyond the expected count\n"); abort(); }
  b->pending--;
  simt_mbar_check(b);
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  SimtMbar* b = reinterpret_cast<SimtMbar*>(bar);
  if (b->phase != (parity & 1u)) return true;   // the phase with this parity has completed
  simt::spin_yield();
  return false;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
inline void simt_check_bulk(const void* a, const void* b, uint32_t bytes) {
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (bytes & 15)) {
    fprintf(stderr, "simt: bulk copy needs 16-byte aligned addresses and size (%p, %p, %u)\n", a, b, bytes);
    abort();
  }
}
inline void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  simt_check_bulk(smem_dst, gmem_src, bytes);
  memcpy(smem_dst, gmem_src, bytes);
  SimtMbar* b = reinterpret_cast<SimtMbar*>(bar);
  b->tx -= (int32_t)bytes;
  simt_mbar_check(b);
}
inline void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  simt_check_bulk(gmem_dst, smem_src, bytes);
  memcpy(gmem_dst, smem_src, bytes);
}
inline void tma_prefetch_l2(const void* gmem_src, uint32_t bytes) {   // no effect on the host; the alignment rules still hold
  simt_check_bulk(gmem_src, gmem_src, bytes);
  volatile unsigned char sink = reinterpret_cast<const volatile unsigned char*>(gmem_src)[bytes - 1];   // inside an allocation
  (void)sink;
}
inline void tma_store_commit() {}
inline void tma_store_wait_read() {}
inline void tma_store_wait_all() {}
inline void fence_proxy_async() {}
inline void st_release_sys_u64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
inline void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
inline unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(p);
  simt::spin_yield();
  return v;
}
inline void st_release_gpu_u64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
inline unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(p);
  simt::spin_yield();
  return v;
}
inline unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(p);
  simt::spin_yield();
  return v;
}
inline long long clock64() {   // the SM clock of the stage-timing kernels: host nanoseconds here
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000000000ll + (long long)ts.tv_nsec;
}
inline unsigned long long globaltimer_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

}  // namespace rfsb200
