// cb200_dynamics.cu -- RNEA inverse dynamics + adjoint kernels, C ABI.
//
// Replaces rnea_forward_kernel / rnea_backward_kernel (curobo/_src/curobolib/kernels/dynamics/, launched by
// backends/cuda_core_backend/dynamics.py:24-250).  One thread per (seed x waypoint) row -- rows are independent and
// there are tens of thousands of them, so tree-level parallelism inside a row (the reference's threads_per_batch > 1
// path with shared-memory atomics) is not needed to fill the machine, and the serial order makes every sum
// deterministic.  A row's per-link spatial vectors live in a TRANSPOSED shared-memory tile
// [array][link][component][row-in-CTA]: consecutive threads touch consecutive words, no bank conflicts.
// HBM traffic per row: forward 3*D*4 in, D*4 + nl*80 (cache) out; adjoint D*4*3 + nl*80 in, 3*D*4 out.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/curobo_b200.h"
#include "cb200_dynamics.cuh"

namespace {
using namespace cb200::dyn;

inline int status(cudaError_t e) {
  if (e != cudaSuccess) (void)cudaGetLastError();
  return (int)e;
}

struct TileStore {  // [arr][link][comp][rows] floats in shared memory
  float *base;
  int rows, t, nl;
  __device__ __forceinline__ float get(int arr, int k, int c) const { return base[((size_t)(arr * nl + k) * 6 + c) * rows + t]; }
  __device__ __forceinline__ void set(int arr, int k, int c, float v) { base[((size_t)(arr * nl + k) * 6 + c) * rows + t] = v; }
};

struct FwdArgs {
  Model M;
  float *tau, *cache;
  const float *q, *qd, *qdd, *f_ext;
  int B;
};
__global__ void rnea_forward_rows(const __grid_constant__ FwdArgs a) {
  extern __shared__ float smem[];
  TileStore S{smem, (int)blockDim.x, (int)threadIdx.x, a.M.nl};
  const int D = a.M.D, nl = a.M.nl;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < a.B; row += (long long)gridDim.x * blockDim.x)
    rnea_forward_row(a.M, S, a.q + row * D, a.qd + row * D, a.qdd + row * D, a.f_ext ? a.f_ext + row * nl * 6 : nullptr,
                     a.tau + row * D, a.cache + row * nl * kCacheFloatsPerLink);
}

struct BwdArgs {
  Model M;
  float *gq, *gqd, *gqdd, *grad_f_ext;
  const float *grad_tau, *q, *qd, *cache;
  int B;
};
__global__ void rnea_backward_rows(const __grid_constant__ BwdArgs a) {
  extern __shared__ float smem[];
  TileStore S{smem, (int)blockDim.x, (int)threadIdx.x, a.M.nl};
  const int D = a.M.D, nl = a.M.nl;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < a.B; row += (long long)gridDim.x * blockDim.x)
    rnea_backward_row(a.M, S, a.grad_tau + row * D, a.q + row * D, a.qd + row * D, a.cache + row * nl * kCacheFloatsPerLink,
                      a.gq + row * D, a.gqd + row * D, a.gqdd + row * D, a.grad_f_ext ? a.grad_f_ext + row * nl * 6 : nullptr);
}

// rows per CTA: the largest of 128 / 64 / 32 whose tile leaves room for two CTAs per SM
template <class K>
int pick_rows(K kern, int floats_per_row, int &smem_out) {
  int dev = 0, max_smem = 227 * 1024;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  for (int rows = 128; rows >= 32; rows >>= 1) {
    const int smem = floats_per_row * rows * (int)sizeof(float);
    if (smem * 2 + 4096 <= max_smem || rows == 32) {
      if (smem > max_smem) return 0;
      if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
        (void)cudaGetLastError();
        continue;
      }
      smem_out = smem;
      return rows;
    }
  }
  return 0;
}

int grid_for(int B, int rows) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = ((long long)B + rows - 1) / rows;
  if (g > (long long)sms * 8) g = (long long)sms * 8;
  return (int)(g < 1 ? 1 : g);
}

bool model_ok(const Model &M) {
  return M.fixed_transforms && M.masses_com && M.inertias && M.joint_type && M.joint_map && M.link_map && M.joint_offset &&
         M.gravity && M.level_starts && M.level_links && M.nl >= 1 && M.nl <= 1024 && M.D >= 1 && M.n_levels >= 1;
}
}  // namespace

extern "C" {

int cb200_rnea_forward(float *tau, const float *q, const float *qd, const float *qdd, const float *fixed_transforms,
                       const float *link_masses_com, const float *link_inertias, const int8_t *joint_map_type,
                       const int16_t *joint_map, const int16_t *link_map, const float *joint_offset_map,
                       const float *gravity, const int16_t *level_starts, const int16_t *level_links, float *forward_cache,
                       int batch_size, int num_links, int num_dof, int n_levels, const float *f_ext, cb200_stream_t stream) {
  Model M{fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map, link_map, joint_offset_map, gravity,
          level_starts,     level_links,     num_links,     num_dof,        n_levels};
  if (!model_ok(M) || tau == nullptr || q == nullptr || qd == nullptr || qdd == nullptr || forward_cache == nullptr ||
      batch_size < 0)
    return status(cudaErrorInvalidValue);
  if (batch_size == 0) return status(cudaSuccess);
  int smem = 0;
  const int rows = pick_rows(rnea_forward_rows, 2 * num_links * 6, smem);
  if (rows == 0) return status(cudaErrorInvalidConfiguration);
  FwdArgs a{M, tau, forward_cache, q, qd, qdd, f_ext, batch_size};
  rnea_forward_rows<<<grid_for(batch_size, rows), rows, smem, (cudaStream_t)stream>>>(a);
  return status(cudaGetLastError());
}

int cb200_rnea_backward(float *grad_q, float *grad_qd, float *grad_qdd, const float *grad_tau, const float *q,
                        const float *qd, const float *fixed_transforms, const float *link_masses_com,
                        const float *link_inertias, const int8_t *joint_map_type, const int16_t *joint_map,
                        const int16_t *link_map, const float *joint_offset_map, const float *gravity,
                        const int16_t *level_starts, const int16_t *level_links, const float *forward_cache, int batch_size,
                        int num_links, int num_dof, int n_levels, float *grad_f_ext, cb200_stream_t stream) {
  Model M{fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map, link_map, joint_offset_map, gravity,
          level_starts,     level_links,     num_links,     num_dof,        n_levels};
  if (!model_ok(M) || grad_q == nullptr || grad_qd == nullptr || grad_qdd == nullptr || grad_tau == nullptr || q == nullptr ||
      qd == nullptr || forward_cache == nullptr || batch_size < 0)
    return status(cudaErrorInvalidValue);
  if (batch_size == 0) return status(cudaSuccess);
  int smem = 0;
  const int rows = pick_rows(rnea_backward_rows, 5 * num_links * 6, smem);
  if (rows == 0) return status(cudaErrorInvalidConfiguration);
  BwdArgs a{M, grad_q, grad_qd, grad_qdd, grad_f_ext, grad_tau, q, qd, forward_cache, batch_size};
  rnea_backward_rows<<<grid_for(batch_size, rows), rows, smem, (cudaStream_t)stream>>>(a);
  return status(cudaGetLastError());
}

}  // extern "C"
