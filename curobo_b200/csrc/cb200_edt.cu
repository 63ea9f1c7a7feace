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
#include "cb200_launch.h"
#include "cb200_edt.cuh"

namespace {
using namespace cb200::edt;

inline int status(cudaError_t e) {
  if (e != cudaSuccess) (void)cudaGetLastError();
  return (int)e;
}

// The per-lane work of every pass lives in cb200_edt.cuh (FloodZ, Envelope<AXIS>: also executed lane by lane by the host
// emulation in tests/hostmath); a kernel is the grid-stride loop over tiles plus the warp barriers.  One warp per CTA.
__global__ void __launch_bounds__(kLanes) edt_flood_z_kernel(const __grid_constant__ FloodZ pass) {
  CB200_EXTERN_SHARED int tile[];
  const int lane = threadIdx.x;
  const long long ntiles = pass.ntiles();
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    pass.load(tile, t, lane);
    __syncwarp();
    pass.compute(tile, t, lane);
    __syncwarp();
    pass.store(tile, t, lane);
    __syncwarp();
  }
}

template <int AXIS>
__global__ void __launch_bounds__(kLanes) edt_envelope_kernel(const __grid_constant__ Envelope<AXIS> pass) {
  CB200_EXTERN_SHARED int tile[];
  const int lane = threadIdx.x;
  const long long ntiles = pass.ntiles();
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    pass.run(tile, t, lane);
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
  CB200_DEVICE_GUARD(site_index);
  (void)buffer;  // the reference's ping-pong scratch: every pass here is in place
  (void)m3;      // the reference's colour-kernel block height
  if (site_index == nullptr || !dims_ok(nx, ny, nz)) return status(cudaErrorInvalidValue);
  const cudaStream_t st = (cudaStream_t)stream;
  const Plan p = make_plan(site_index, nx, ny, nz);
  const int smem_z = p.z.tile_ints() * (int)sizeof(int), smem_y = p.y.tile_ints() * (int)sizeof(int),
            smem_x = p.x.tile_ints() * (int)sizeof(int);
  if (!allow_smem(edt_flood_z_kernel, smem_z) || !allow_smem(edt_envelope_kernel<1>, smem_y) ||
      !allow_smem(edt_envelope_kernel<0>, smem_x))
    return status(cudaErrorInvalidConfiguration);
  CB200_LAUNCH(edt_flood_z_kernel, grid_for(p.z.ntiles()), kLanes, smem_z, st, p.z);
  CB200_LAUNCH(edt_envelope_kernel<1>, grid_for(p.y.ntiles()), kLanes, smem_y, st, p.y);
  CB200_LAUNCH(edt_envelope_kernel<0>, grid_for(p.x.ntiles()), kLanes, smem_x, st, p.x);
  return status(cudaGetLastError());
}

int cb200_edt_unsigned_distance(const int32_t *site_index, uint16_t *distance_fp16, int nx, int ny, int nz, float voxel_size,
                                float empty_value, cb200_stream_t stream) {
  CB200_DEVICE_GUARD(distance_fp16);
  if (site_index == nullptr || distance_fp16 == nullptr || !dims_ok(nx, ny, nz) || !(voxel_size > 0.0f))
    return status(cudaErrorInvalidValue);
  const long long total = (long long)nx * ny * nz;
  const long long blocks = (total + 255) / 256;
  CB200_LAUNCH(edt_distance_kernel, grid_for(blocks), 256, 0, (cudaStream_t)stream, site_index, reinterpret_cast<__half *>(distance_fp16),
               ny, nz, total, voxel_size, empty_value);
  return status(cudaGetLastError());
}

}  // extern "C"
