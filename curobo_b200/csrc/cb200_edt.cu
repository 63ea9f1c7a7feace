// cb200_edt.cu -- exact 3-D nearest-site transform kernels, C ABI.
//
// Replaces the reference's five PBA+ launches + final copy (backends/cuda_core_backend/pba.py:60-124) with three in-place
// passes and no transposes: z is the contiguous axis of the [nx, ny, nz] grid, so
//   pass 1  floods along z           rows are contiguous: a warp stages 32 rows through a padded shared-memory tile,
//   pass 2  envelopes along y        32 columns adjacent in z per warp: every row of the tile is one 128-byte line,
//   pass 3  envelopes along x        columns adjacent in (y, z): same.
// A warp owns a tile of 32 columns held entirely in shared memory; a lane runs the sequential column algorithm
// (cb200_edt.cuh) on its own column at shared-memory latency -- the reference's threads walk their columns through global
// memory, one dependent load per row -- and the only HBM traffic is one coalesced read and one coalesced write of the grid
// per pass: 3 x 8 B per voxel (the reference moves 6 x 8 B plus its stack look-ups).  Every pass is in place (a tile is
// fully staged before its first row is written back), so the scratch `buffer` of the reference interface is not touched.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/curobo_b200.h"
#include "cb200_edt.cuh"

namespace {
using namespace cb200::edt;

inline int status(cudaError_t e) {
  if (e != cudaSuccess) (void)cudaGetLastError();
  return (int)e;
}

struct TileCol {  // one column of a shared-memory tile: row r at base[r * stride]
  int *base;
  int stride;
  __device__ __forceinline__ int get(int r) const { return base[r * stride]; }
  __device__ __forceinline__ void set(int r, int v) const { base[r * stride] = v; }
};
struct GlobalCol {  // the same column in the grid
  int *base;
  long long stride;
  __device__ __forceinline__ void set(int r, int v) const { base[(long long)r * stride] = v; }
};

constexpr int kLanes = 32;
constexpr int kPad = 33;  // padded row length of the transposed tile of the z pass

// pass 1: nearest site along z.  Tile = 32 consecutive (x, y) rows of nz ints, stored transposed [nz][33]
__global__ void __launch_bounds__(kLanes) edt_flood_z_kernel(int *__restrict__ grid, int nz, long long nrows) {
  extern __shared__ int tile[];
  const int lane = threadIdx.x;
  const long long ntiles = (nrows + kLanes - 1) / kLanes;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long row0 = t * kLanes;
    const int live_rows = (int)((nrows - row0) < kLanes ? (nrows - row0) : kLanes);
    for (int rr = 0; rr < live_rows; ++rr) {
      const int *src = grid + (row0 + rr) * nz;
      for (int z = lane; z < nz; z += kLanes) tile[z * kPad + rr] = src[z];
    }
    __syncwarp();
    if (lane < live_rows) {
      TileCol c{tile + lane, kPad};
      flood_column<2>(c, nz);
    }
    __syncwarp();
    for (int rr = 0; rr < live_rows; ++rr) {
      int *dst = grid + (row0 + rr) * nz;
      for (int z = lane; z < nz; z += kLanes) dst[z] = tile[z * kPad + rr];
    }
    __syncwarp();
  }
}

// passes 2 and 3: lower envelope along AXIS (1 = y, 0 = x).  Columns are indexed by (outer, inner) with `inner` contiguous in
// memory: AXIS 1: outer = x, inner = z; AXIS 0: outer = 0, inner = y * nz + z.  Tile = [n][32] (row r of the tile = 32 ints
// adjacent in memory).
template <int AXIS>
__global__ void __launch_bounds__(kLanes) edt_envelope_kernel(int *__restrict__ grid, int n, long long row_stride, int inner,
                                                               int n_outer, long long outer_stride, int nz) {
  extern __shared__ int tile[];
  const int lane = threadIdx.x;
  const int tiles_per_outer = (inner + kLanes - 1) / kLanes;
  const long long ntiles = (long long)tiles_per_outer * n_outer;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int outer = (int)(t / tiles_per_outer);
    const int col = (int)(t - (long long)outer * tiles_per_outer) * kLanes + lane;
    const bool live = col < inner;
    int *base = grid + (long long)outer * outer_stride + (live ? col : 0);
#pragma unroll 8
    for (int r = 0; r < n; ++r) tile[r * kLanes + lane] = live ? base[(long long)r * row_stride] : kEmpty;
    if (live) {
      Voxel q;
      if (AXIS == 1) {
        q.x = outer, q.y = 0, q.z = col;
      } else {
        q.x = 0, q.y = col / nz, q.z = col - (col / nz) * nz;
      }
      TileCol c{tile + lane, kLanes};
      GlobalCol o{base, row_stride};
      envelope_column<AXIS>(c, o, n, q);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) edt_distance_kernel(const int *__restrict__ sites, __half *__restrict__ out, int ny, int nz,
                                                            long long total, float voxel_size, float empty_value) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = sites[i];
    const int z = (int)(i % nz), y = (int)((i / nz) % ny), x = (int)(i / ((long long)nz * ny));
    const float d = v < 0 ? empty_value : sqrtf((float)site_distance_sq(v, x, y, z)) * voxel_size;
    out[i] = __float2half_rn(d);
  }
}

template <class K>
bool allow_smem(K kern, int smem) {
  if (smem <= 48 * 1024) return true;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess) return true;
  (void)cudaGetLastError();
  return false;
}
int grid_for(long long tiles) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 32;
  return (int)(tiles < 1 ? 1 : (tiles > cap ? cap : tiles));
}
bool dims_ok(int nx, int ny, int nz) {
  return nx >= 1 && ny >= 1 && nz >= 1 && nx <= kMaxDim && ny <= kMaxDim && nz <= kMaxDim &&
         (long long)nx * ny * nz <= 2147483647LL;
}
}  // namespace

extern "C" {

int cb200_pba3d(int32_t *site_index, int32_t *buffer, int nx, int ny, int nz, int m3, cb200_stream_t stream) {
  (void)buffer;  // the reference's ping-pong scratch: every pass here is in place
  (void)m3;      // the reference's colour-kernel block height
  if (site_index == nullptr || !dims_ok(nx, ny, nz)) return status(cudaErrorInvalidValue);
  const cudaStream_t st = (cudaStream_t)stream;
  const int smem_z = nz * kPad * (int)sizeof(int), smem_y = ny * kLanes * (int)sizeof(int),
            smem_x = nx * kLanes * (int)sizeof(int);
  if (!allow_smem(edt_flood_z_kernel, smem_z) || !allow_smem(edt_envelope_kernel<1>, smem_y) ||
      !allow_smem(edt_envelope_kernel<0>, smem_x))
    return status(cudaErrorInvalidConfiguration);
  const long long nrows = (long long)nx * ny;
  edt_flood_z_kernel<<<grid_for((nrows + kLanes - 1) / kLanes), kLanes, smem_z, st>>>(site_index, nz, nrows);
  const long long plane = (long long)ny * nz;
  edt_envelope_kernel<1><<<grid_for((long long)((nz + kLanes - 1) / kLanes) * nx), kLanes, smem_y, st>>>(
      site_index, ny, (long long)nz, nz, nx, plane, nz);
  edt_envelope_kernel<0><<<grid_for((plane + kLanes - 1) / kLanes), kLanes, smem_x, st>>>(site_index, nx, plane, (int)plane, 1, 0,
                                                                                         nz);
  return status(cudaGetLastError());
}

int cb200_edt_unsigned_distance(const int32_t *site_index, uint16_t *distance_fp16, int nx, int ny, int nz, float voxel_size,
                                float empty_value, cb200_stream_t stream) {
  if (site_index == nullptr || distance_fp16 == nullptr || !dims_ok(nx, ny, nz) || !(voxel_size > 0.0f))
    return status(cudaErrorInvalidValue);
  const long long total = (long long)nx * ny * nz;
  const long long blocks = (total + 255) / 256;
  edt_distance_kernel<<<grid_for(blocks), 256, 0, (cudaStream_t)stream>>>(site_index, reinterpret_cast<__half *>(distance_fp16), ny,
                                                                        nz, total, voxel_size, empty_value);
  return status(cudaGetLastError());
}

}  // extern "C"
