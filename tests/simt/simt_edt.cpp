// TEST INFRASTRUCTURE: the nearest-site transform kernels of curobo_b200/csrc/cb200_edt.cu compiled as ordinary C++ and executed
// by std::threads (tests/simt/cuda_runtime.h): the three launches of cb200_pba3d with a chosen number of one-warp CTAs.
#define CB200_SIMT_EMULATION 1
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_edt.cu"

extern "C" int em_pba3d(int32_t *site_index, int nx, int ny, int nz, int n_ctas) {
  const Plan p = make_plan(site_index, nx, ny, nz);
  if ((size_t)p.z.tile_ints() > sizeof(tile) / sizeof(int) || (size_t)p.y.tile_ints() > sizeof(tile) / sizeof(int) ||
      (size_t)p.x.tile_ints() > sizeof(tile) / sizeof(int))
    return 1;
  auto grid = [&](long long tiles) { return (int)(tiles < n_ctas ? (tiles < 1 ? 1 : tiles) : n_ctas); };
  simt::launch(edt_flood_z_kernel, grid(p.z.ntiles()), kLanes, p.z);
  simt::launch(edt_envelope_kernel<1>, grid(p.y.ntiles()), kLanes, p.y);
  simt::launch(edt_envelope_kernel<0>, grid(p.x.ntiles()), kLanes, p.x);
  return 0;
}
