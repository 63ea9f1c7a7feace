// cb200_edt.cuh -- exact 3-D nearest-site transform (Euclidean distance transform), one column at a time.
//
// SURVEY.md section 8(f) rank 4, the producer side of the ESDF wire format: replaces the reference's PBA+ pipeline
// (curobo/_src/curobolib/kernels/parallel_banding/pba3d_kernel.cuh: kernel_flood_z :67-104, kernel_maurer_axis :160-222,
// kernel_color_axis :235-350; launched five times by backends/cuda_core_backend/pba.py:60-124).
// Same data: int32 per voxel, site = (z << 20) | (y << 10) | x in cuRobo indices (site_encoding.cuh:13-19 with the kernels'
// (sx, sy, sz) = (nz, ny, nx); perception/mapper/util/utils_quantization.py:40-54), negative = no site, grid [nx, ny, nz]
// with z contiguous; result = packed coordinates of a nearest site, 0x80000000 where the grid has no site at all.
// Same mathematics: a 1-D nearest-site flood along the first axis, then the lower envelope of the parabolas
// (t - r)^2 + h_r along each remaining axis with Maurer's integer dominance test (is_voronoi_dominated :119-142) -- exact.
// The structure is ours: a column lives behind an accessor (`Col`) so the same routines run on a shared-memory tile on the GPU
// and on plain arrays on the host (tests/hostmath); the stack of dominant sites is built IN PLACE over the rows already
// consumed (a stack never holds more entries than rows read), and the fill writes through a second accessor (`Out`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef CB_HD
#define CB_HD __host__ __device__ __forceinline__
#endif

namespace cb200 {
namespace edt {

constexpr int kEmpty = static_cast<int>(0x80000000u);
constexpr int kMask = 0x3ff;
constexpr int kMaxDim = 1023;  // 10-bit coordinates

// AXIS: 0 = x (bits 9-0, slowest in memory), 1 = y (bits 19-10), 2 = z (bits 29-20, contiguous in memory)
template <int AXIS>
CB_HD int coord(int v) {
  return (v >> (10 * AXIS)) & kMask;
}
CB_HD int pack(int x, int y, int z) { return (z << 20) | (y << 10) | x; }

struct Voxel {  // cuRobo indices of the query column; the along-axis member is ignored
  int x, y, z;
};

// squared distance from site v to the column through q, measured off-axis (the h of the parabola (t - r)^2 + h)
template <int AXIS>
CB_HD int off_axis_sq(int v, const Voxel &q) {
  int d = 0;
  if (AXIS != 0) {
    const int e = coord<0>(v) - q.x;
    d += e * e;
  }
  if (AXIS != 1) {
    const int e = coord<1>(v) - q.y;
    d += e * e;
  }
  if (AXIS != 2) {
    const int e = coord<2>(v) - q.z;
    d += e * e;
  }
  return d;
}

// First axis: nearest site along the column only (kernel_flood_z), in place.  Ties keep the site at the larger row, as the
// reference's backward sweep does.
template <int AXIS, class Col>
CB_HD void flood_column(Col &c, int n) {
  int carry = kEmpty;
  for (int i = 0; i < n; ++i) {
    const int v = c.get(i);
    if (v >= 0) carry = v;
    c.set(i, carry < 0 ? kEmpty : carry);
  }
  for (int i = n - 2; i >= 0; --i) {
    const int f = c.get(i);
    const int db = carry < 0 ? 0x7fffffff : (coord<AXIS>(carry) > i ? coord<AXIS>(carry) - i : i - coord<AXIS>(carry));
    const int df = f < 0 ? 0x7fffffff : (coord<AXIS>(f) > i ? coord<AXIS>(f) - i : i - coord<AXIS>(f));
    if (df < db) carry = f;
    c.set(i, carry);
  }
}

// Remaining axes.  c.get(r) = nearest site of row r within the axes already processed (negative = none); the site's own
// coordinate along AXIS is r.  Pass 1 builds the stack of Voronoi-dominant sites in place (entries 0..m-1 of the column),
// pass 2 walks rows upward and emits the nearest stack entry of every row through `out`.
// Returns the stack height (0 = the column sees no site: every row gets kEmpty).
template <int AXIS, class Col, class Out>
CB_HD int envelope_column(Col &c, Out &out, int n, const Voxel &q) {
  int m = 0;
  // registers mirror the two top entries: (row, g = row^2 + h)
  int r_top = 0, g_top = 0, r_below = 0, g_below = 0;
  for (int r = 0; r < n; ++r) {
    const int v = c.get(r);
    if (v < 0) continue;
    const int g = r * r + off_axis_sq<AXIS>(v, q);
    while (m >= 2) {
      // top is dominated when the parabolas of `below` and the new site meet before top's range (Maurer):
      //   (g_top - g_below) (r - r_top) > (g - g_top) (r_top - r_below)        [64-bit: up to 3.2e6 * 1023]
      const long long lhs = (long long)(g_top - g_below) * (long long)(r - r_top);
      const long long rhs = (long long)(g - g_top) * (long long)(r_top - r_below);
      if (!(lhs > rhs)) break;
      --m;  // pop: `below` becomes the top; reload the entry under it
      r_top = r_below;
      g_top = g_below;
      if (m >= 2) {
        const int u = c.get(m - 2);
        r_below = coord<AXIS>(u);
        g_below = r_below * r_below + off_axis_sq<AXIS>(u, q);
      }
    }
    c.set(m, v);  // push (m <= number of sites seen so far <= r: never overwrites an unread row)
    ++m;
    r_below = r_top;
    g_below = g_top;
    r_top = r;
    g_top = g;
  }
  if (m == 0) {
    for (int t = 0; t < n; ++t) out.set(t, kEmpty);
    return 0;
  }
  int k = 0;
  int cur = c.get(0);
  int h_cur = off_axis_sq<AXIS>(cur, q), r_cur = coord<AXIS>(cur);
  int nxt = m > 1 ? c.get(1) : kEmpty;
  int h_nxt = m > 1 ? off_axis_sq<AXIS>(nxt, q) : 0, r_nxt = m > 1 ? coord<AXIS>(nxt) : 0;
  for (int t = 0; t < n; ++t) {
    int best = (r_cur - t) * (r_cur - t) + h_cur;
    while (k + 1 < m) {
      const int cand = (r_nxt - t) * (r_nxt - t) + h_nxt;
      if (cand > best) break;  // ties move on to the later site (the envelope is sorted: it stays nearest longer)
      best = cand;
      ++k;
      cur = nxt;
      h_cur = h_nxt;
      r_cur = r_nxt;
      if (k + 1 < m) {
        nxt = c.get(k + 1);
        h_nxt = off_axis_sq<AXIS>(nxt, q);
        r_nxt = coord<AXIS>(nxt);
      }
    }
    out.set(t, cur);
  }
  return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-tile schedule of the three passes.  These are the functions a lane executes on the GPU (cb200_edt.cu) AND the
// functions the host emulation executes lane by lane (tests/hostmath: hm_pba3d_tiles), so the index arithmetic of the
// kernels is covered by the CPU tests; the kernels themselves only add the grid-stride loop and the warp barriers.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLanes = 32;
constexpr int kPad = 33;  // padded row length of the transposed tile of the z pass

struct TileCol {  // one column of a tile: row r at base[r * stride]
  int *base;
  int stride;
  CB_HD int get(int r) const { return base[r * stride]; }
  CB_HD void set(int r, int v) const { base[r * stride] = v; }
};
struct GridCol {  // the same column in the grid
  int *base;
  long long stride;
  CB_HD void set(int r, int v) const { base[(long long)r * stride] = v; }
};

// pass 1: nearest site along z.  Tile t = 32 consecutive (x, y) rows of nz ints, staged transposed as [nz][33].
struct FloodZ {
  int *grid;
  int nz;
  long long nrows;  // nx * ny
  CB_HD long long ntiles() const { return (nrows + kLanes - 1) / kLanes; }
  CB_HD int tile_ints() const { return nz * kPad; }
  CB_HD int live_rows(long long t) const {
    const long long left = nrows - t * kLanes;
    return (int)(left < kLanes ? left : kLanes);
  }
  CB_HD void load(int *tile, long long t, int lane) const {  // lane copies column-slices z = lane, lane+32, .. of every row
    const int rows = live_rows(t);
    for (int rr = 0; rr < rows; ++rr) {
      const int *src = grid + (t * kLanes + rr) * nz;
      for (int z = lane; z < nz; z += kLanes) tile[z * kPad + rr] = src[z];
    }
  }
  CB_HD void compute(int *tile, long long t, int lane) const {  // lane owns row `lane` of the tile
    if (lane < live_rows(t)) {
      TileCol c{tile + lane, kPad};
      flood_column<2>(c, nz);
    }
  }
  CB_HD void store(const int *tile, long long t, int lane) const {
    const int rows = live_rows(t);
    for (int rr = 0; rr < rows; ++rr) {
      int *dst = grid + (t * kLanes + rr) * nz;
      for (int z = lane; z < nz; z += kLanes) dst[z] = tile[z * kPad + rr];
    }
  }
};

// passes 2 and 3: lower envelope along AXIS (1 = y, 0 = x).  Columns are indexed by (outer, inner) with `inner` contiguous in
// memory -- AXIS 1: outer = x, inner = z; AXIS 0: outer = 0, inner = y * nz + z.  Tile = [n][32]: row r of the tile is 32 ints
// adjacent in memory.  A lane stages its own column, builds the stack in place and writes its rows back: no lane reads
// another lane's column, so the pass needs no barrier.
template <int AXIS>
struct Envelope {
  int *grid;
  int n;                 // extent along AXIS
  long long row_stride;  // elements between consecutive rows of a column
  int inner, n_outer;
  long long outer_stride;
  int nz;
  CB_HD int tiles_per_outer() const { return (inner + kLanes - 1) / kLanes; }
  CB_HD long long ntiles() const { return (long long)tiles_per_outer() * n_outer; }
  CB_HD int tile_ints() const { return n * kLanes; }
  CB_HD void run(int *tile, long long t, int lane) const {
    const int outer = (int)(t / tiles_per_outer());
    const int col = (int)(t - (long long)outer * tiles_per_outer()) * kLanes + lane;
    if (col >= inner) return;
    int *base = grid + (long long)outer * outer_stride + col;
#ifdef __CUDA_ARCH__
#pragma unroll 8
#endif
    for (int r = 0; r < n; ++r) tile[r * kLanes + lane] = base[(long long)r * row_stride];
    Voxel q;
    if (AXIS == 1) {
      q.x = outer, q.y = 0, q.z = col;
    } else {
      q.x = 0, q.y = col / nz, q.z = col - (col / nz) * nz;
    }
    TileCol c{tile + lane, kLanes};
    GridCol o{base, row_stride};
    envelope_column<AXIS>(c, o, n, q);
  }
};

struct Plan {  // the three passes of one transform of a [nx, ny, nz] grid
  FloodZ z;
  Envelope<1> y;
  Envelope<0> x;
};
CB_HD Plan make_plan(int *grid, int nx, int ny, int nz) {
  const long long plane = (long long)ny * nz;
  Plan p;
  p.z = FloodZ{grid, nz, (long long)nx * ny};
  p.y = Envelope<1>{grid, ny, (long long)nz, nz, nx, plane, nz};
  p.x = Envelope<0>{grid, nx, plane, (int)plane, 1, 0, nz};
  return p;
}

// |voxel - site| * voxel_size as fp16 bits are produced by the caller; this is the integer part
CB_HD int site_distance_sq(int v, int x, int y, int z) {
  const int dx = coord<0>(v) - x, dy = coord<1>(v) - y, dz = coord<2>(v) - z;
  return dx * dx + dy * dy + dz * dz;
}

}  // namespace edt
}  // namespace cb200
