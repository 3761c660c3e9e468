// simt.h — TEST INFRASTRUCTURE ONLY.
//
// A small host-side interpreter for CUDA kernel *sources*: every thread of a CTA is a fiber with its own stack, the
// warp- and CTA-level primitives (__syncwarp, __shfl_*_sync, __ballot_sync, __syncthreads, ...) are rendezvous points
// between the fibers, "device memory" is host memory.  tests/simt/simt_build.py compiles rfs-slam_b200/csrc/rfsb200_abi.cu
// — the same kernel and host sources the product is built from — against this header with g++, so that the CPU test
// suite (`-m "not gpu"`) can execute the real kernel logic on small inputs and compare it with the oracle.
//
// It is not a fallback: the package (rfs-slam_b200/capi.py) only ever loads csrc/librfsb200.so, nothing here is built
// or imported by the product, and nothing measured with it is ever reported.  A lane runs alone until it reaches a
// rendezvous, i.e. the interleaving is the most adversarial one independent thread scheduling allows: code that
// relies on implicit lock-step between lanes (a missing __syncwarp) fails here even where the hardware forgives it.
// Dynamic shared memory is filled with a NaN pattern before every CTA, so reads of uninitialised shared memory show.
#pragma once
#include <stdint.h>
#include <ucontext.h>
#include <sys/mman.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// ---- fiber contexts: a 20-instruction switch on x86-64 (no signal-mask system call), ucontext elsewhere --------
#if defined(__x86_64__) && !defined(SIMT_USE_UCONTEXT)
namespace simt { struct Context { void* sp = nullptr; }; }
extern "C" void simt_switch_context(simt::Context* from, simt::Context* to) __attribute__((visibility("hidden")));
asm(R"(
.text
.type simt_switch_context,@function
simt_switch_context:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq (%rsi), %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch_context,.-simt_switch_context
)");
namespace simt {
inline void switch_context(Context* from, Context* to) { simt_switch_context(from, to); }
inline void make_context(Context* c, void* stack, size_t bytes, void (*entry)()) {
  // entered by the `ret` of the switch: the entry address sits 16-byte aligned, six register slots below it
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
  void** x = reinterpret_cast<void**>(top - 16);
  x[0] = reinterpret_cast<void*>(entry);
  x[1] = nullptr;   // a return address the entry function never uses
  for (int k = 1; k <= 6; k++) x[-k] = nullptr;
  c->sp = x - 6;
}
}  // namespace simt
#else
namespace simt {
struct Context { ucontext_t uc; };
inline void switch_context(Context* from, Context* to) { swapcontext(&from->uc, &to->uc); }
inline void make_context(Context* c, void* stack, size_t bytes, void (*entry)()) {
  getcontext(&c->uc);
  c->uc.uc_stack.ss_sp = stack;
  c->uc.uc_stack.ss_size = bytes;
  c->uc.uc_link = nullptr;
  makecontext(&c->uc, entry, 0);
}
}  // namespace simt
#endif

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
// the CTAs of a launch run one after the other, so one static instance per variable; per host thread, because
// every host thread is one "GPU" (two ranks of the fused cross-GPU sum run on two threads)
#define __shared__ static thread_local

namespace simt {

struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
  Dim3() {}
  Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

struct Rendezvous {
  uint32_t key = 0;       // the mask the participants named
  uint32_t arrived = 0;
  uint32_t gen = 0;
  uint32_t part[2] = {0, 0};   // participants of generation g, by parity
  uint64_t vals[2][32];
};

struct Warp {
  uint32_t exited = 0;
  std::vector<std::unique_ptr<Rendezvous>> rv;   // one per mask in use (almost always only the full mask)
  Rendezvous* get(uint32_t key) {
    for (auto& r : rv) if (r->key == key) return r.get();
    rv.emplace_back(new Rendezvous());
    rv.back()->key = key;
    return rv.back().get();
  }
};

enum WaitKind { RUNNABLE = 0, WAIT_WARP = 1, WAIT_CTA = 2 };

struct Fiber {
  Context ctx;
  int tid = 0;
  bool done = false;
  int wait = RUNNABLE;
  Rendezvous* rv = nullptr;
  uint32_t gen = 0;
};

struct Cta {
  int nthreads = 0;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  int bar_arrived = 0, n_exited = 0;
  uint32_t bar_gen = 0;
  Context sched;
  Fiber* cur = nullptr;
  unsigned char* dyn = nullptr;
  void (*entry)(void*) = nullptr;
  void* entry_arg = nullptr;
};

// interpreter state is per host thread: a host thread is one "GPU" that runs its launches one after the other.
// Hidden visibility lets the compiler address the hot variables relative to the library's own TLS block (one
// descriptor call per function, -mtls-dialect=gnu2 in simt_build.py, instead of one __tls_get_addr per access).
#define SIMT_HOT_TLS __attribute__((visibility("hidden")))
inline thread_local Cta* g_cta SIMT_HOT_TLS = nullptr;
inline thread_local unsigned char* g_stacks = nullptr;
inline size_t g_stack_bytes = 256 * 1024;
inline thread_local int g_stack_count = 0;
inline thread_local uint64_t g_switches SIMT_HOT_TLS = 0;   // statistics
inline thread_local const char* g_kernel_name = "";

}  // namespace simt

// the CUDA built-in coordinates (set by the scheduler whenever a fiber is resumed)
using dim3 = simt::Dim3;
inline thread_local simt::Dim3 threadIdx SIMT_HOT_TLS, blockIdx SIMT_HOT_TLS, blockDim SIMT_HOT_TLS, gridDim SIMT_HOT_TLS;

namespace simt {

inline void yield_to_scheduler() {
  Cta* c = g_cta;
  Fiber* f = c->cur;
  g_switches++;
  switch_context(&f->ctx, &c->sched);
}

inline void check_release_cta(Cta* c) {
  if (c->bar_arrived > 0 && c->bar_arrived >= c->nthreads - c->n_exited) {
    c->bar_arrived = 0;
    c->bar_gen++;
  }
}
inline void check_release_warp(Warp& w) {
  for (auto& r : w.rv) {
    const uint32_t need = r->key & ~w.exited;
    if (r->arrived && (r->arrived & need) == need) {
      r->part[r->gen & 1] = r->arrived;
      r->arrived = 0;
      r->gen++;
    }
  }
}

inline void fiber_main() {
  Cta* c = g_cta;
  Fiber* f = c->cur;
  c->entry(c->entry_arg);
  f->done = true;
  const int warp = f->tid >> 5, lane = f->tid & 31;
  c->warps[warp].exited |= 1u << lane;
  c->n_exited++;
  check_release_warp(c->warps[warp]);   // an exited lane no longer holds anybody up
  check_release_cta(c);
  switch_context(&f->ctx, &c->sched);
  abort();   // a finished fiber is never resumed
}

[[noreturn]] inline void deadlock(Cta* c) {
  fprintf(stderr, "simt: deadlock in kernel %s, block %u: no runnable thread\n", g_kernel_name, blockIdx.x);
  int shown = 0;
  for (Fiber& f : c->fibers) {
    if (f.done) continue;
    if (shown++ < 8)
      fprintf(stderr, "  thread %d waits at %s (mask %08x, arrived %08x)\n", f.tid, f.wait == WAIT_CTA ? "__syncthreads" : "a warp rendezvous",
              f.rv ? f.rv->key : 0u, f.rv ? f.rv->arrived : 0u);
  }
  abort();
}

// run one CTA to completion
inline void run_cta(unsigned block_idx, unsigned grid, int nthreads, size_t smem_bytes, void (*entry)(void*), void* arg) {
  Cta cta;
  cta.nthreads = nthreads;
  cta.fibers.resize(nthreads);
  cta.warps.resize((nthreads + 31) / 32);
  for (int w = 0; w < (int)cta.warps.size(); w++) {   // lanes beyond the block size never exist
    const int live = nthreads - 32 * w;
    if (live < 32) cta.warps[w].exited = ~((1u << live) - 1u);
  }
  cta.entry = entry;
  cta.entry_arg = arg;
  // dynamic shared memory, 128-byte aligned, NaN / garbage pattern
  std::vector<unsigned char> dyn(smem_bytes + 256 + 512);
  memset(dyn.data(), 0xff, dyn.size());
  cta.dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 127) & ~uintptr_t(127));
  memset(cta.dyn + smem_bytes, 0xc7, 256);   // canary behind the launch's dynamic shared memory, checked below
  if (g_stack_count < nthreads) {
    if (g_stacks) munmap(g_stacks, (size_t)g_stack_count * g_stack_bytes);
    g_stacks = (unsigned char*)mmap(nullptr, (size_t)nthreads * g_stack_bytes, PROT_READ | PROT_WRITE,
                                    MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_stacks == MAP_FAILED) { perror("simt: mmap"); abort(); }
    g_stack_count = nthreads;
  }
  g_cta = &cta;
  for (int t = 0; t < nthreads; t++) {
    Fiber& f = cta.fibers[t];
    f.tid = t;
    make_context(&f.ctx, g_stacks + (size_t)t * g_stack_bytes, g_stack_bytes, &fiber_main);
  }
  gridDim = Dim3(grid);
  blockDim = Dim3((unsigned)nthreads);
  blockIdx = Dim3(block_idx);
  blockIdx.y = blockIdx.z = 0;
  int remaining = nthreads;
  while (remaining > 0) {
    bool progress = false;
    for (int t = 0; t < nthreads; t++) {
      Fiber& f = cta.fibers[t];
      if (f.done) continue;
      if (f.wait == WAIT_WARP) { if (f.rv->gen == f.gen) continue; }
      else if (f.wait == WAIT_CTA) { if (cta.bar_gen == f.gen) continue; }
      f.wait = RUNNABLE;
      cta.cur = &f;
      threadIdx.x = (unsigned)t; threadIdx.y = threadIdx.z = 0;
      switch_context(&cta.sched, &f.ctx);
      progress = true;
      if (f.done) remaining--;
    }
    if (!progress && remaining > 0) deadlock(&cta);
  }
  for (int k = 0; k < 256; k++)
    if (cta.dyn[smem_bytes + k] != 0xc7) {
      fprintf(stderr, "simt: kernel %s, block %u wrote behind its %zu bytes of dynamic shared memory\n", g_kernel_name, block_idx, smem_bytes);
      abort();
    }
  g_cta = nullptr;
}

// ---- rendezvous of the lanes named in `mask`; returns the values every participant contributed ----------------
struct Exchange { const uint64_t* vals; uint32_t part; };

inline Exchange exchange(uint32_t mask, uint64_t v) {
  Cta* c = g_cta;
  Fiber* f = c->cur;
  const int lane = f->tid & 31;
  Warp& w = c->warps[f->tid >> 5];
  if (!(mask & (1u << lane))) {
    fprintf(stderr, "simt: lane %d calls a *_sync primitive with mask %08x that does not name it (kernel %s)\n", lane, mask, g_kernel_name);
    abort();
  }
  Rendezvous* r = w.get(mask);
  const uint32_t g = r->gen;
  r->vals[g & 1][lane] = v;
  r->arrived |= 1u << lane;
  const uint32_t need = mask & ~w.exited;
  if ((r->arrived & need) == need) {
    r->part[g & 1] = r->arrived;
    r->arrived = 0;
    r->gen++;
  } else {
    f->wait = WAIT_WARP;
    f->rv = r;
    f->gen = g;
    yield_to_scheduler();
  }
  return Exchange{r->vals[g & 1], r->part[g & 1]};
}

inline void syncthreads() {
  Cta* c = g_cta;
  Fiber* f = c->cur;
  const uint32_t g = c->bar_gen;
  c->bar_arrived++;
  if (c->bar_arrived >= c->nthreads - c->n_exited) {
    c->bar_arrived = 0;
    c->bar_gen++;
  } else {
    f->wait = WAIT_CTA;
    f->gen = g;
    f->rv = nullptr;
    yield_to_scheduler();
  }
}

// a polite spin: lets the other threads of the CTA run
inline void spin_yield() { yield_to_scheduler(); }

inline int lane_id() { return g_cta->cur->tid & 31; }
inline unsigned char* dyn_smem() { return g_cta->dyn; }

template <typename T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle operand wider than 64 bits");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

// ---- kernel launch ------------------------------------------------------------------------------------------
template <typename F, typename... Args>
struct Thunk {
  F f;
  std::tuple<Args...>* args;
  static void call(void* self) {
    Thunk* t = static_cast<Thunk*>(self);
    std::apply(t->f, *t->args);
  }
};

struct Launcher {
  unsigned grid;
  int block;
  size_t smem;
  const char* name;
  template <typename S>
  Launcher(Dim3 g, Dim3 b, size_t s, S, const char* n) : grid(g.x), block((int)b.x), smem(s), name(n) {}
  template <typename... P, typename... Args>
  void operator()(void (*kernel)(P...), Args&&... args) {
    std::tuple<std::decay_t<P>...> a(static_cast<std::decay_t<P>>(args)...);
    Thunk<void (*)(P...), std::decay_t<P>...> th{kernel, &a};
    const char* saved = g_kernel_name;
    g_kernel_name = name;
    for (unsigned b = 0; b < grid; b++) run_cta(b, grid, block, smem, &decltype(th)::call, &th);
    g_kernel_name = saved;
  }
};

}  // namespace simt

// ---- warp / CTA primitives --------------------------------------------------------------------------------------
inline void __syncthreads() { simt::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::exchange(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = simt::lane_id();
  const simt::Exchange x = simt::exchange(mask, simt::to_bits(v));
  const int s = (lane & ~(width - 1)) | (src & (width - 1));
  return simt::from_bits<T>(x.vals[s]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  const int lane = simt::lane_id();
  const simt::Exchange x = simt::exchange(mask, simt::to_bits(v));
  const int s = lane ^ lanemask;
  if ((s & ~(width - 1)) != (lane & ~(width - 1))) return v;
  return simt::from_bits<T>(x.vals[s]);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = simt::lane_id();
  const simt::Exchange x = simt::exchange(mask, simt::to_bits(v));
  const int s = lane - (int)delta;
  if (s < (lane & ~(width - 1))) return v;
  return simt::from_bits<T>(x.vals[s]);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = simt::lane_id();
  const simt::Exchange x = simt::exchange(mask, simt::to_bits(v));
  const int s = lane + (int)delta;
  if (s > (lane | (width - 1))) return v;
  return simt::from_bits<T>(x.vals[s]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  const simt::Exchange x = simt::exchange(mask, pred ? 1u : 0u);
  unsigned b = 0;
  for (int l = 0; l < 32; l++)
    if ((x.part >> l) & 1u) b |= (unsigned)(x.vals[l] & 1u) << l;
  return b;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) {
  const simt::Exchange x = simt::exchange(mask, pred ? 1u : 0u);
  for (int l = 0; l < 32; l++)
    if (((x.part >> l) & 1u) && !(x.vals[l] & 1u)) return 0;
  return 1;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  const simt::Exchange x = simt::exchange(mask, v);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if ((x.part >> l) & 1u) r |= (unsigned)x.vals[l];
  return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  const simt::Exchange x = simt::exchange(mask, v);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if ((x.part >> l) & 1u) r += (unsigned)x.vals[l];
  return r;
}

// ---- atomics (one host thread: plain read-modify-write) -----------------------------------------------------------
template <typename T, typename U> inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U> inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U> inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U> inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U> inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename U> inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename U, typename V> inline T atomicCAS(T* p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---- bit and conversion intrinsics --------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
// n-th set bit of mask counted from base (CUDA __fns): offset > 0 upwards (offset 1 = the first set bit at or above
// base), offset < 0 downwards, offset 0 = base itself if set; 0xffffffff if there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  if (offset == 0) return ((mask >> base) & 1u) ? base : 0xffffffffu;
  if (offset > 0) {
    for (unsigned b = base; b < 32; b++)
      if (((mask >> b) & 1u) && --offset == 0) return b;
  } else {
    for (int b = (int)base; b >= 0; b--)
      if (((mask >> b) & 1u) && ++offset == 0) return (unsigned)b;
  }
  return 0xffffffffu;
}
inline unsigned __float_as_uint(float f) { return simt::from_bits<unsigned>(simt::to_bits(f)); }
inline int __float_as_int(float f) { return simt::from_bits<int>(simt::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return simt::from_bits<float>(simt::to_bits(u)); }
inline float __int_as_float(int u) { return simt::from_bits<float>(simt::to_bits(u)); }
inline double __longlong_as_double(long long v) { return simt::from_bits<double>(simt::to_bits(v)); }
inline long long __double_as_longlong(double v) { return simt::from_bits<long long>(simt::to_bits(v)); }
inline float __fdividef(float a, float b) { return a / b; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

// CUDA's global min / max overloads (mixed integer types included)
template <typename A, typename B> inline typename std::common_type<A, B>::type min(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)b < (C)a ? (C)b : (C)a;
}
template <typename A, typename B> inline typename std::common_type<A, B>::type max(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a < (C)b ? (C)b : (C)a;
}

// CUDA math functions glibc does not have
inline void sincospi(double x, double* s, double* c) { sincos(3.14159265358979323846 * x, s, c); }
inline void sincospif(float x, float* s, float* c) { sincosf(3.14159265358979323846f * x, s, c); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline void __sincosf(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
