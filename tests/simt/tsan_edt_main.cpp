// TEST INFRASTRUCTURE: race detector run of the nearest-site transform kernels (see tsan_dynamics_main.cpp): a seeded grid is
// built here, transformed by the emulated kernels under -fsanitize=thread, and checked against a brute-force distance.
#include <cstdio>
#include <cstdlib>

#include "simt_edt.cpp"

int main() {
  const int nx = 21, ny = 40, nz = 37;  // partial 32-column tiles in every pass
  std::vector<int32_t> g((size_t)nx * ny * nz, -1), sites;
  unsigned s = 12345;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z) {
        s = s * 1664525u + 1013904223u;
        if ((s >> 8) % 97 == 0) {
          g[((size_t)x * ny + y) * nz + z] = (z << 20) | (y << 10) | x;
          sites.push_back((z << 20) | (y << 10) | x);
        }
      }
  if (em_pba3d(g.data(), nx, ny, nz, 5)) return 3;
  long long bad = 0;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z) {
        const int v = g[((size_t)x * ny + y) * nz + z];
        auto d2 = [&](int p) {
          const int dx = (p & 1023) - x, dy = ((p >> 10) & 1023) - y, dz = ((p >> 20) & 1023) - z;
          return dx * dx + dy * dy + dz * dz;
        };
        int best = 1 << 30;
        for (int p : sites) best = std::min(best, d2(p));
        if (v < 0 || d2(v) != best) ++bad;
      }
  printf("%s %lld\n", bad == 0 ? "ok" : "WRONG", bad);
  return bad == 0 ? 0 : 4;
}
