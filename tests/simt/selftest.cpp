// selftest.cpp — TEST INFRASTRUCTURE ONLY: checks the host interpreter of simt.h / simt_ptx.h itself on a few
// small kernels (shuffles, votes, partial masks, early exits, __syncthreads, atomics, mbarrier + bulk copy, and that
// a lane really runs ahead of its warp between rendezvous points).  Built and run by tests/test_simt_kernels.py.
#define RFSB200_SIMT_HOST 1
#include "cuda_runtime.h"
namespace rfsb200 {}
#include "simt_ptx.h"

#define CHECK(c) do { if (!(c)) { printf("selftest FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

static void k_shuffles(int* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v = 100 * warp + lane;
  int s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int up = __shfl_up_sync(0xffffffffu, v, 1);
  const int dn = __shfl_down_sync(0xffffffffu, v, 3);
  const int b5 = __shfl_sync(0xffffffffu, v, 5);
  const unsigned odd = __ballot_sync(0xffffffffu, lane & 1);
  int* o = out + 8 * threadIdx.x;
  o[0] = s; o[1] = up; o[2] = dn; o[3] = b5; o[4] = (int)odd;
  o[5] = __any_sync(0xffffffffu, lane == 31);
  o[6] = __all_sync(0xffffffffu, lane < 31);
  o[7] = (int)__reduce_or_sync(0xffffffffu, 1u << (lane & 7));
}

// lanes 0..15 and 16..31 rendezvous under different masks; lanes >= 24 leave early
static void k_partial(int* out) {
  const int lane = threadIdx.x & 31;
  if (lane >= 24) { out[threadIdx.x] = -1; return; }
  const unsigned mask = lane < 16 ? 0x0000ffffu : 0x00ff0000u;
  const unsigned b = __ballot_sync(mask, 1);
  __syncwarp(mask);
  out[threadIdx.x] = (int)b;
}

// a full-mask rendezvous after a quarter of the warp has exited (the hardware counts exited lanes as arrived)
static void k_exit_then_sync(int* out) {
  const int lane = threadIdx.x & 31;
  if (lane < 8) return;
  out[threadIdx.x] = (int)__ballot_sync(0xffffffffu, 1);
}

// without the second __syncwarp a lane reads a slot its owner has not written yet: the interpreter runs lanes one by one
static void k_needs_sync(int* out, int with_sync) {
  __shared__ int box[32];
  const int lane = threadIdx.x & 31;
  box[lane] = -7;
  __syncwarp();
  box[lane] = lane;
  if (with_sync) __syncwarp();
  out[lane] = box[31 - lane];
}

static void k_block(int* out, unsigned* counter) {
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  atomicAdd(&total, (int)threadIdx.x);
  __syncthreads();
  if (threadIdx.x == blockDim.x - 1) out[blockIdx.x] = total + 1000 * (int)atomicAdd(counter, 1u);
}

static void k_bulk(const float* src, float* dst, int n) {
  unsigned char* smem = simt::dyn_smem();
  float* buf = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  const int lane = threadIdx.x;
  if (lane == 0) { rfsb200::mbar_init(bar, 1); rfsb200::fence_mbar_init(); }
  __syncwarp();
  uint32_t phase = 0;
  for (int rep = 0; rep < 3; rep++) {
    if (lane == 0) {
      rfsb200::mbar_expect_tx(bar, 2 * n * 4);
      rfsb200::tma_load_1d(buf, src, n * 4, bar);
      rfsb200::tma_load_1d(buf + n, src + n, n * 4, bar);
    }
    rfsb200::mbar_wait(bar, phase);
    phase ^= 1;
    for (int k = lane; k < 2 * n; k += 32) dst[rep * 2 * n + k] = buf[k] + rep;
    __syncwarp();
  }
}

static void k_oob_smem(int n) {
  unsigned char* smem = simt::dyn_smem();
  if (threadIdx.x == 0) smem[n] = 1;   // one byte behind the launch's dynamic shared memory
}
static void k_oob_global(float* buf, int n) {
  if (threadIdx.x == 0) buf[n] = 1.f;  // one element behind the allocation
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "oob-smem")) {   // must abort
    simt::Launcher(1, 32, 1024, 0, "k_oob_smem")(k_oob_smem, 1024);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "oob-global")) {   // must abort at cudaFree
    float* b = nullptr;
    cudaMalloc(&b, 100 * 4);
    simt::Launcher(1, 32, 0, 0, "k_oob_global")(k_oob_global, b, 100);
    cudaFree(b);
    return 0;
  }
  {
    std::vector<int> out(8 * 96, 0);
    simt::Launcher(1, 96, 0, 0, "k_shuffles")(k_shuffles, out.data());
    for (int t = 0; t < 96; t++) {
      const int lane = t & 31, warp = t >> 5;
      const int* o = &out[8 * t];
      CHECK(o[0] == 32 * 100 * warp + 496);
      CHECK(o[1] == (lane == 0 ? 100 * warp : 100 * warp + lane - 1));
      CHECK(o[2] == (lane + 3 > 31 ? 100 * warp + lane : 100 * warp + lane + 3));
      CHECK(o[3] == 100 * warp + 5);
      CHECK((unsigned)o[4] == 0xaaaaaaaau);
      CHECK(o[5] == 1 && o[6] == 0 && o[7] == 0xff);
    }
  }
  {
    std::vector<int> out(32, 0);
    simt::Launcher(1, 32, 0, 0, "k_partial")(k_partial, out.data());
    for (int l = 0; l < 32; l++) CHECK(out[l] == (l < 16 ? 0xffff : (l < 24 ? 0x00ff0000 : -1)));
    std::vector<int> o2(32, 0);
    simt::Launcher(1, 32, 0, 0, "k_exit_then_sync")(k_exit_then_sync, o2.data());
    for (int l = 8; l < 32; l++) CHECK((unsigned)o2[l] == 0xffffff00u);
  }
  {
    std::vector<int> a(32, 0), b(32, 0);
    simt::Launcher(1, 32, 0, 0, "k_needs_sync")(k_needs_sync, a.data(), 1);
    simt::Launcher(1, 32, 0, 0, "k_needs_sync")(k_needs_sync, b.data(), 0);
    for (int l = 0; l < 32; l++) CHECK(a[l] == 31 - l);
    CHECK(b[31] == -7);   // the missing __syncwarp shows: lane 31 left the first rendezvous before lane 0 ran again
  }
  {
    std::vector<int> out(5, 0);
    unsigned counter = 0;
    simt::Launcher(5, 200, 0, 0, "k_block")(k_block, out.data(), &counter);
    for (int b = 0; b < 5; b++) CHECK(out[b] == 199 * 200 / 2 + 1000 * b);
    CHECK(counter == 5);
  }
  {
    const int n = 64;
    std::vector<float> src(2 * n), dst(6 * n, -1.f);
    float* s16 = nullptr; float* d16 = nullptr;
    cudaMalloc(&s16, 2 * n * 4); cudaMalloc(&d16, 6 * n * 4);
    for (int k = 0; k < 2 * n; k++) s16[k] = (float)k;
    simt::Launcher(1, 32, 8192, 0, "k_bulk")(k_bulk, (const float*)s16, d16, n);
    for (int rep = 0; rep < 3; rep++)
      for (int k = 0; k < 2 * n; k++) CHECK(d16[rep * 2 * n + k] == (float)(k + rep));
    cudaFree(s16); cudaFree(d16);
  }
  CHECK(__fns(0xb0u, 0, 1) == 4 && __fns(0xb0u, 0, 3) == 7 && __fns(0xb0u, 0, 4) == 0xffffffffu && __fns(0xb0u, 7, -2) == 5);
  CHECK(__ffsll(0x100000000LL) == 33 && __clzll(1LL) == 63 && __popcll(~0ull) == 64 && __ffs(0) == 0);
  CHECK(min(3, 5u) == 3u && max(2.5, 1) == 2.5);
  printf("selftest OK (%llu fiber switches)\n", (unsigned long long)simt::g_switches);
  return 0;
}
