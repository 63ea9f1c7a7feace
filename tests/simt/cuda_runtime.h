// TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> in the SIMT host emulation build (tests/simt).  It lets g++ compile a
// kernel translation unit of curobo_b200/csrc as ordinary C++: kernels become plain functions, the built-in index variables
// are thread-local, __syncthreads() is a real barrier across the std::threads that play the CTA's threads, shared memory is a
// static buffer (CTAs run one after another), atomicAdd is a real atomic.  Races between barriers are therefore real races.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__
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

struct float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 {
  unsigned x, y, z;
};

namespace simt {
inline thread_local uint3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local std::barrier<> *t_barrier = nullptr;
inline thread_local int t_lane_group = 0;
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
inline void __syncwarp() {}
#else
inline void __syncwarp() { simt::t_barrier->arrive_and_wait(); }  // conservative: a CTA-wide barrier (used with 1-warp CTAs only)
#endif
inline float atomicAdd(float *p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
template <class T>
inline T __ldg(const T *p) { return *p; }

// the CTA's dynamic shared memory: `extern __shared__ float smem[];` / `extern __shared__ int tile[];` in the kernels resolve
// to these (CTAs are emulated one at a time)
// (a block-scope `extern` inside a kernel in an unnamed namespace names a member of that namespace: the kernels of this
// repository all live in unnamed namespaces, so the buffers do too)
namespace {
alignas(16) float smem[96 * 1024];
alignas(16) int tile[96 * 1024];
}  // namespace

namespace simt {
// run `grid` CTAs of `block` threads one CTA after another; every CTA thread is a std::thread
template <class Kernel, class Args>
void launch(Kernel kern, int grid, int block, const Args &args) {
  for (int b = 0; b < grid; ++b) {
    std::barrier<> bar(block);
    std::vector<std::thread> ts;
    ts.reserve(block);
    for (int t = 0; t < block; ++t)
      ts.emplace_back([&, t, b] {
        t_threadIdx = uint3{(unsigned)t, 0, 0};
        t_blockIdx = uint3{(unsigned)b, 0, 0};
        t_blockDim = uint3{(unsigned)block, 1, 1};
        t_gridDim = uint3{(unsigned)grid, 1, 1};
        t_barrier = &bar;
        kern(args);
        bar.arrive_and_drop();  // a thread that has left the kernel no longer takes part in barriers
      });
    for (auto &th : ts) th.join();
  }
}
}  // namespace simt
