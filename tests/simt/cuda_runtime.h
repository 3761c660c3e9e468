// cuda_runtime.h — TEST INFRASTRUCTURE ONLY (see simt.h).  The slice of the CUDA runtime API that
// rfs-slam_b200/csrc/rfsb200_abi.cu uses, on host memory: "device" allocations are malloc'ed, copies are memcpy,
// streams are in order and synchronous, events carry no time.  Found instead of the real header only by the build of
// tests/simt/simt_build.py (-I tests/simt); the product build never sees it.
#pragma once
#include "simt.h"

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
typedef struct simt_stream_s* cudaStream_t;
typedef struct simt_event_s* cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1;

inline const char* cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotSupported: return "operation not supported by the host interpreter";
    default: return "invalid value";
  }
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  p->major = 10;
  p->minor = 0;
  const char* e = getenv("SIMT_SM_COUNT");   // few "SMs": several particles per warp, several CTAs per launch
  p->multiProcessorCount = e ? atoi(e) : 148;
  if (p->multiProcessorCount < 1) p->multiProcessorCount = 1;
  snprintf(p->name, sizeof(p->name), "simt host interpreter");
  return cudaSuccess;
}
// Allocations carry a 256-byte canary on either side, checked when they are freed: a kernel that writes outside its
// buffers aborts the test instead of corrupting the heap quietly.  Contents start as garbage (0xa5), not zeros.
namespace simt {
constexpr size_t GUARD = 256;
inline void* guarded_alloc(size_t bytes) {
  void* q = nullptr;
  const size_t body = (bytes + 255) & ~size_t(255);
  if (posix_memalign(&q, 256, GUARD + body + GUARD)) return nullptr;
  unsigned char* b = static_cast<unsigned char*>(q);
  memset(b, 0xc7, GUARD);
  memset(b + GUARD, 0xa5, body);
  memset(b + GUARD + bytes, 0xc7, body - bytes + GUARD);
  memcpy(b, &bytes, sizeof(bytes));   // the size lives in the first bytes of the front guard
  return b + GUARD;
}
inline void guarded_free(void* p) {
  if (!p) return;
  unsigned char* b = static_cast<unsigned char*>(p) - GUARD;
  size_t bytes;
  memcpy(&bytes, b, sizeof(bytes));
  const size_t body = (bytes + 255) & ~size_t(255);
  bool ok = true;
  for (size_t k = sizeof(bytes); k < GUARD; k++) ok = ok && b[k] == 0xc7;
  for (size_t k = GUARD + bytes; k < GUARD + body + GUARD; k++) ok = ok && b[k] == 0xc7;
  if (!ok) {
    fprintf(stderr, "simt: a kernel or copy wrote outside a %zu-byte allocation (guard bytes damaged)\n", bytes);
    abort();
  }
  free(b);
}
}  // namespace simt
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = simt::guarded_alloc(bytes);
  if (!q) return cudaErrorMemoryAllocation;
  *p = static_cast<T*>(q);
  return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFree(void* p) { simt::guarded_free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { simt::guarded_free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
// shared memory is the only limit modelled: 228 KB per SM, 1 KB reserved per CTA, at most 2048 threads
template <typename K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* occ, K, int block, size_t smem) {
  const long by_smem = (228L * 1024) / (long)(smem + 1024);
  const long by_threads = 2048 / (block > 0 ? block : 1);
  long o = by_smem < by_threads ? by_smem : by_threads;
  *occ = (int)(o > 32 ? 32 : o);
  return cudaSuccess;
}
// every pointer is "pinned host memory" here (SIMT_PAGEABLE=1: none is, the staged path of rfsb200_update_host runs)
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  const char* e = getenv("SIMT_PAGEABLE");
  const bool pageable = e && atoi(e) != 0;
  a->type = pageable ? cudaMemoryTypeUnregistered : cudaMemoryTypeHost;
  a->device = 0;
  a->devicePointer = pageable ? nullptr : const_cast<void*>(p);
  a->hostPointer = const_cast<void*>(p);
  return cudaSuccess;
}
// "peers" are other contexts of the same process, driven by other host threads: an IPC handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return *p ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
