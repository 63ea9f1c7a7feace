// TEST INFRASTRUCTURE: the nearest-site transform kernels of curobo_b200/csrc/cb200_edt.cu compiled as ordinary C++ and executed
// by std::threads (tests/simt/cuda_runtime.h): the three launches of cb200_pba3d with a chosen number of one-warp CTAs.
#define CB200_SIMT_EMULATION 1
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_edt.cu"

extern "C" int em_pba3d(int32_t *site_index, int nx, int ny, int nz, int n_ctas) {
  const Plan p = make_plan(site_index, nx, ny, nz);
  const BandedEnvelope<1> by{p.y};
  const BandedEnvelope<0> bx{p.x};
  if ((size_t)by.smem_ints() > sizeof(tile) / sizeof(int) || (size_t)bx.smem_ints() > sizeof(tile) / sizeof(int)) return 1;
  auto grid = [&](long long tiles) { return (int)(tiles < n_ctas ? (tiles < 1 ? 1 : tiles) : n_ctas); };
  const long long nrows = (long long)nx * ny;
  if (nz <= 128) {
    simt::launch(edt_flood_z_kernel<4>, grid((nrows + 15) / 16), 256, site_index, nz, nrows);
  } else if (nz <= 256) {
    simt::launch(edt_flood_z_kernel<8>, grid((nrows + 15) / 16), 256, site_index, nz, nrows);
  } else if (nz <= 512) {
    simt::launch(edt_flood_z_kernel<16>, grid((nrows + 15) / 16), 256, site_index, nz, nrows);
  } else {
    simt::launch(edt_flood_z_kernel<32>, grid((nrows + 15) / 16), 256, site_index, nz, nrows);
  }
  simt::launch(edt_envelope_kernel<1>, grid(by.e.ntiles()), kBands * kLanes, by);
  simt::launch(edt_envelope_kernel<0>, grid(bx.e.ntiles()), kBands * kLanes, bx);
  return 0;
}
