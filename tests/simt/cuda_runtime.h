// TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> in the SIMT host emulation build (tests/simt).  It lets g++ compile a
// kernel translation unit of curobo_b200/csrc as ordinary C++: kernels become plain functions, the built-in index variables
// are thread-local, __syncthreads() is a real barrier across the std::threads that play the CTA's threads, shared memory is a
// static buffer (CTAs run one after another), atomicAdd is a real atomic.  Races between barriers are therefore real races.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static   // a kernel-local __shared__ array is one array per CTA; CTAs are emulated one at a time
#define __align__(n)
#define __launch_bounds__(...)
#define __grid_constant__
#ifndef __restrict__
#define __restrict__ __restrict
#endif

using std::max;
using std::min;

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorInvalidConfiguration = 9 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
// a small imaginary device: 2 SMs, B200-sized shared memory, two resident CTAs per SM for every kernel
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaFuncAttributes {
  size_t sharedSizeBytes = 0, constSizeBytes = 0, localSizeBytes = 0;
  int maxThreadsPerBlock = 1024, numRegs = 64, maxDynamicSharedSizeBytes = 232448;
};
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrMultiProcessorCount ? 2 : 232448;
  return cudaSuccess;
}
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K>
inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, K) { *a = cudaFuncAttributes(); return cudaSuccess; }
template <class K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 3; return cudaSuccess; }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }

struct float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 {
  unsigned x, y, z;
};

namespace simt {
inline thread_local uint3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local std::barrier<> *t_barrier = nullptr;       // the CTA
struct Warp {                                                  // one per 32 consecutive CTA threads
  std::barrier<> bar;
  unsigned long long slot[32];
  explicit Warp(int lanes) : bar(lanes) {}
};
inline thread_local Warp *t_warp = nullptr;
inline thread_local int t_lane = 0;
// named barriers (PTX bar.sync id, nthreads): one std::barrier per id, created by the first thread that arrives
struct NamedBarriers {
  std::mutex mu;
  std::unique_ptr<std::barrier<>> bar[16];
};
inline thread_local NamedBarriers *t_named = nullptr;
inline void named_barrier(int id, int nthreads) {
#ifndef CB200_SIMT_DROP_BARRIERS
  std::barrier<> *b;
  {
    std::lock_guard<std::mutex> g(t_named->mu);
    if (!t_named->bar[id]) t_named->bar[id].reset(new std::barrier<>(nthreads));
    b = t_named->bar[id].get();
  }
  b->arrive_and_wait();
#endif
}
}  // namespace simt
#define threadIdx (simt::t_threadIdx)
#define blockIdx (simt::t_blockIdx)
#define blockDim (simt::t_blockDim)
#define gridDim (simt::t_gridDim)

#ifdef CB200_SIMT_DROP_BARRIERS  // mutation switch for the race-detector self-test: barriers do nothing
inline void __syncthreads() {}
#else
inline void __syncthreads() { simt::t_barrier->arrive_and_wait(); }
#endif
#ifdef CB200_SIMT_DROP_BARRIERS
inline void __syncwarp(unsigned = 0xffffffffu) {}
#else
inline void __syncwarp(unsigned = 0xffffffffu) { simt::t_warp->bar.arrive_and_wait(); }
#endif

// Warp collectives, full-mask and convergent (the only form the kernels of this repository use): every lane publishes its
// value, the warp meets, every lane reads, the warp meets again so the slots can be reused.
namespace simt {
template <class T>
inline unsigned long long to_bits(T v) {
  static_assert(sizeof(T) <= 8, "collective payload");
  unsigned long long b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(unsigned long long b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
template <class T, class F>
inline T exchange(T v, F pick) {
  Warp &w = *t_warp;
  w.slot[t_lane] = to_bits(v);
  w.bar.arrive_and_wait();
  const T r = pick(w.slot);
  w.bar.arrive_and_wait();
  return r;
}
}  // namespace simt
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int base = simt::t_lane & ~(width - 1);
  return simt::exchange(v, [&](const unsigned long long *s) { return simt::from_bits<T>(s[base + (src & (width - 1))]); });
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32) {
  const int src = simt::t_lane ^ lane_mask;
  const bool ok = (src & ~(width - 1)) == (simt::t_lane & ~(width - 1));
  return simt::exchange(v, [&](const unsigned long long *s) { return simt::from_bits<T>(s[ok ? src : simt::t_lane]); });
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
  const int src = simt::t_lane + (int)delta;
  const bool ok = (src & ~(width - 1)) == (simt::t_lane & ~(width - 1));
  return simt::exchange(v, [&](const unsigned long long *s) { return simt::from_bits<T>(s[ok ? src : simt::t_lane]); });
}
inline unsigned __ballot_sync(unsigned, int pred) {
  return simt::exchange((unsigned)(pred != 0), [&](const unsigned long long *s) {
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] & 1ull) << i;
    return m;
  });
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  return simt::exchange(v, [&](const unsigned long long *s) { unsigned r = 0; for (int i = 0; i < 32; ++i) r += (unsigned)s[i]; return r; });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  return simt::exchange(v, [&](const unsigned long long *s) { unsigned r = ~0u; for (int i = 0; i < 32; ++i) r = std::min(r, (unsigned)s[i]); return r; });
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  return simt::exchange(v, [&](const unsigned long long *s) { unsigned r = 0; for (int i = 0; i < 32; ++i) r = std::max(r, (unsigned)s[i]); return r; });
}
inline float atomicAdd(float *p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
inline int atomicAdd(int *p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_acq_rel); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T>
inline T __ldg(const T *p) { return *p; }

// the CTA's dynamic shared memory: `extern __shared__ float smem[];` / `extern __shared__ int tile[];` in the kernels resolve
// to these (CTAs are emulated one at a time)
// (a block-scope `extern` inside a kernel in an unnamed namespace names a member of that namespace: the kernels of this
// repository all live in unnamed namespaces, so the buffers do too)
namespace {
#ifdef CB200_SIMT_SMEM_DECL
CB200_SIMT_SMEM_DECL
#else
alignas(16) float smem[96 * 1024];
alignas(16) int tile[96 * 1024];
#endif
}  // namespace

namespace simt {
// run `grid` CTAs of `block` threads one CTA after another; every CTA thread is a std::thread.  Inactive slots of the last
// (partial) warp hold zero in collectives: the kernels of this repository launch whole warps whenever they use collectives.
template <class Kernel, class... Args>
void launch(Kernel kern, int grid, int block, const Args &...args) {
  for (int b = 0; b < grid; ++b) {
    std::barrier<> bar(block);
    NamedBarriers named;
    std::vector<std::unique_ptr<Warp>> warps;
    for (int w = 0; w * 32 < block; ++w) {
      warps.emplace_back(new Warp(std::min(32, block - w * 32)));
      std::memset(warps.back()->slot, 0, sizeof(warps.back()->slot));
    }
    std::vector<std::thread> ts;
    ts.reserve(block);
    for (int t = 0; t < block; ++t)
      ts.emplace_back([&, t, b] {
        t_threadIdx = uint3{(unsigned)t, 0, 0};
        t_blockIdx = uint3{(unsigned)b, 0, 0};
        t_blockDim = uint3{(unsigned)block, 1, 1};
        t_gridDim = uint3{(unsigned)grid, 1, 1};
        t_barrier = &bar;
        t_named = &named;
        t_warp = warps[t / 32].get();
        t_lane = t % 32;
        kern(args...);
        t_warp->bar.arrive_and_drop();  // a thread that has left the kernel no longer takes part in barriers
        bar.arrive_and_drop();
      });
    for (auto &th : ts) th.join();
  }
}
}  // namespace simt
