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
// Tiles of the envelope passes.  The per-thread phase functions of the banded schedule (BandedEnvelope below) are what a thread
// executes on the GPU (cb200_edt.cu) AND what the host emulation executes thread by thread (tests/hostmath: hm_pba3d_tiles),
// so the index arithmetic of the kernels is covered by the CPU tests; the kernels only add the grid-stride loop and the
// barriers.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLanes = 32;

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

// pass 1: nearest site along z, one z-row of nz contiguous ints at a time (warp-scan kernel in cb200_edt.cu; flood_column
// above is its sequential statement)
struct FloodZ {
  int *grid;
  int nz;
  long long nrows;  // nx * ny
};

// passes 2 and 3: lower envelope along AXIS (1 = y, 0 = x).  Columns are indexed by (outer, inner) with `inner` contiguous in
// memory -- AXIS 1: outer = x, inner = z; AXIS 0: outer = 0, inner = y * nz + z.  Tile = [n][32]: row r of the tile is 32 ints
// adjacent in memory.
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

// ---------------------------------------------------------------------------------------------------------------------
// Banded schedule of the envelope passes (round 2).  The column algorithm above is sequential in the rows, and a 256^3 grid
// has only 65,536 columns: one thread per column leaves a B200 three quarters empty and every thread walks 2 x 256 dependent
// steps (measured: 0.18 + 0.35 ms for the two passes, 9 % of the HBM bound).  The stack of dominant sites is the lower convex
// hull of the points (r, g_r = r^2 + h_r) -- (t - r)^2 + h = t^2 - 2 t r + g_r -- so the classic hull decomposition applies:
//   1. every BAND of rows builds the hull of its own points (Maurer's test, in place over the band's rows),
//   2. the band hulls are joined left to right by walking the common tangent (entries cut off are dropped from the END of the
//      earlier bands' ranges and from the START of the new band's range: no data moves, a column's hull is the concatenation
//      of the ranges [lo_b, hi_b) of its bands),
//   3. every band fills its own rows: binary search for the hull entry of its first row, then the same walk as above.
// Steps 1 and 3 run one thread per (column, band); step 2 one thread per column, touching only entries that are cut.
// Same dominance test, same tie rule in the walk as envelope_column, hence the same nearest sites.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBands = 8;

struct HullPoint {
  int r, g;  // row and r^2 + off-axis squared distance
};
template <int AXIS, class Col>
CB_HD HullPoint hull_point(const Col &c, int pos, const Voxel &q) {
  const int u = c.get(pos);
  const int r = coord<AXIS>(u);
  return HullPoint{r, r * r + off_axis_sq<AXIS>(u, q)};
}
// u is cut off by its neighbours a (left) and c (right): Maurer's integer test, as in envelope_column
CB_HD bool hull_dominated(const HullPoint &a, const HullPoint &u, const HullPoint &c) {
  return (long long)(u.g - a.g) * (long long)(c.r - u.r) > (long long)(c.g - u.g) * (long long)(u.r - a.r);
}

// The same test inside one band: row differences are at most ceil(1023 / kBands) = 128 and g differences below 3.2e6
// (g = r^2 + h <= 3 * 1023^2), so both products stay under 2^31 and plain 32-bit multiplies are exact.
CB_HD bool hull_dominated_in_band(const HullPoint &a, const HullPoint &u, const HullPoint &c) {
  return (u.g - a.g) * (c.r - u.r) > (c.g - u.g) * (u.r - a.r);
}

// step 1: hull of rows [r0, r1), written to positions r0 .. r0 + m - 1; returns m
template <int AXIS, class Col>
CB_HD int band_hull(Col &c, int r0, int r1, const Voxel &q) {
  int m = 0;
  HullPoint top{0, 0}, below{0, 0};
  for (int r = r0; r < r1; ++r) {
    const int v = c.get(r);
    if (v < 0) continue;
    const HullPoint p{r, r * r + off_axis_sq<AXIS>(v, q)};
    while (m >= 2 && hull_dominated_in_band(below, top, p)) {
      --m;
      top = below;
      if (m >= 2) below = hull_point<AXIS>(c, r0 + m - 2, q);
    }
    c.set(r0 + m, v);
    ++m;
    below = top;
    top = p;
  }
  return m;
}

// Per-column band ranges live behind an accessor too: R.lo(b), R.hi(b) are references (shared memory on the GPU).
// step 2: join the band hulls.  In: R.hi(b) = size of band b's hull, R.lo(b) = 0.  Out: the surviving range of every band.
template <int AXIS, class Col, class Ranges>
CB_HD void join_band_hulls(const Col &c, Ranges &R, int nb, int band_rows, const Voxel &q) {
  for (int b = 1; b < nb; ++b) {
    const int mb = R.hi(b);
    if (mb == 0) continue;
    int j = 0;
    HullPoint pj = hull_point<AXIS>(c, b * band_rows, q);
    for (;;) {
      int tb = b - 1;
      while (tb >= 0 && R.hi(tb) <= R.lo(tb)) --tb;
      if (tb < 0) break;  // nothing left of the earlier bands
      bool changed = false;
      HullPoint top = hull_point<AXIS>(c, tb * band_rows + R.hi(tb) - 1, q);
      for (;;) {  // cut entries off the end of the joined hull while the new band's first entry dominates them
        int bb = tb, bi = R.hi(tb) - 2;
        if (bi < R.lo(tb)) {
          bb = tb - 1;
          while (bb >= 0 && R.hi(bb) <= R.lo(bb)) --bb;
          if (bb < 0) break;
          bi = R.hi(bb) - 1;
        }
        const HullPoint below = hull_point<AXIS>(c, bb * band_rows + bi, q);
        if (!hull_dominated(below, top, pj)) break;
        R.hi(tb) -= 1;
        tb = bb;
        top = below;
        changed = true;
      }
      while (j + 1 < mb) {  // cut entries off the start of the new band while the joined hull's last entry dominates them
        const HullPoint pn = hull_point<AXIS>(c, b * band_rows + j + 1, q);
        if (!hull_dominated(top, pj, pn)) break;
        ++j;
        pj = pn;
        changed = true;
      }
      if (!changed) break;
    }
    R.lo(b) = j;
  }
}

// step 3: nearest hull entry of rows [t0, t1) -> out.  The hull is the concatenation of the band ranges.
template <int AXIS, class Col, class Ranges, class Out>
CB_HD void fill_band(const Col &c, const Ranges &R, Out &out, int t0, int t1, int nb, int band_rows, const Voxel &q) {
  int b = 0;
  while (b < nb && R.hi(b) <= R.lo(b)) ++b;
  if (b == nb) {
    for (int t = t0; t < t1; ++t) out.set(t, kEmpty);
    return;
  }
  auto first_after = [&](int bb) {  // first non-empty band after bb, or nb
    int x = bb + 1;
    while (x < nb && R.hi(x) <= R.lo(x)) ++x;
    return x;
  };
  auto dist = [&](int pos, int t) {
    const int u = c.get(pos);
    const int r = coord<AXIS>(u);
    return (r - t) * (r - t) + off_axis_sq<AXIS>(u, q);
  };
  // the entry of row t0: the first k whose successor is farther from t0 (successor nearer or equal -> move on; the
  // predicate is monotone along the hull).  Bands first, then a binary search inside the band.
  for (;;) {
    const int nb2 = first_after(b);
    if (nb2 == nb) break;
    if (dist(nb2 * band_rows + R.lo(nb2), t0) > dist(b * band_rows + R.hi(b) - 1, t0)) break;
    b = nb2;
  }
  int lo = R.lo(b), hi = R.hi(b) - 1;  // answer in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (dist(b * band_rows + mid + 1, t0) > dist(b * band_rows + mid, t0)) {
      hi = mid;
    } else {
      lo = mid + 1;
    }
  }
  int i = lo;
  int cur = c.get(b * band_rows + i);
  int h_cur = off_axis_sq<AXIS>(cur, q), r_cur = coord<AXIS>(cur);
  // successor cursor
  int nb_b = b, nb_i = i + 1;
  if (nb_i >= R.hi(b)) {
    nb_b = first_after(b);
    nb_i = nb_b < nb ? R.lo(nb_b) : 0;
  }
  bool has_next = nb_b < nb;
  int nxt = has_next ? c.get(nb_b * band_rows + nb_i) : kEmpty;
  int h_nxt = has_next ? off_axis_sq<AXIS>(nxt, q) : 0, r_nxt = has_next ? coord<AXIS>(nxt) : 0;
  for (int t = t0; t < t1; ++t) {
    int best = (r_cur - t) * (r_cur - t) + h_cur;
    while (has_next) {
      const int cand = (r_nxt - t) * (r_nxt - t) + h_nxt;
      if (cand > best) break;  // ties move on to the later site, as in envelope_column
      best = cand;
      cur = nxt;
      h_cur = h_nxt;
      r_cur = r_nxt;
      b = nb_b;
      i = nb_i;
      nb_i = i + 1;
      if (nb_i >= R.hi(b)) {
        nb_b = first_after(b);
        nb_i = nb_b < nb ? R.lo(nb_b) : 0;
      }
      has_next = nb_b < nb;
      if (has_next) {
        nxt = c.get(nb_b * band_rows + nb_i);
        h_nxt = off_axis_sq<AXIS>(nxt, q);
        r_nxt = coord<AXIS>(nxt);
      }
    }
    out.set(t, cur);
  }
}

struct TileRanges {  // lo / hi of the bands of one column of a tile: [band][32] ints each
  int *lo_, *hi_;
  CB_HD int &lo(int b) const { return lo_[b * kLanes]; }
  CB_HD int &hi(int b) const { return hi_[b * kLanes]; }
};

// A tile of the banded passes = the same 32 adjacent columns as Envelope<AXIS>; a CTA of kBands warps owns it: thread
// (band = warp, column = lane).  The three phases are separated by CTA barriers in the kernel (and in the host emulation).
template <int AXIS>
struct BandedEnvelope {
  Envelope<AXIS> e;
  CB_HD int band_rows() const { return (e.n + kBands - 1) / kBands; }
  CB_HD int smem_ints() const { return e.n * kLanes + 2 * kBands * kLanes; }
  struct Ctx {
    bool valid;
    int *base;
    Voxel q;
    int r0, r1;
  };
  CB_HD Ctx ctx(long long t, int band, int lane) const {
    Ctx x;
    const int outer = (int)(t / e.tiles_per_outer());
    const int col = (int)(t - (long long)outer * e.tiles_per_outer()) * kLanes + lane;
    x.valid = col < e.inner;
    x.base = e.grid + (long long)outer * e.outer_stride + col;
    if (AXIS == 1) {
      x.q.x = outer, x.q.y = 0, x.q.z = col;
    } else {
      x.q.x = 0, x.q.y = col / e.nz, x.q.z = col - (col / e.nz) * e.nz;
    }
    x.r0 = band * band_rows();
    x.r1 = x.r0 + band_rows() < e.n ? x.r0 + band_rows() : e.n;
    if (x.r0 > e.n) x.r0 = x.r1 = e.n;
    return x;
  }
  CB_HD TileRanges ranges(int *tile, int lane) const {
    int *lo = tile + e.n * kLanes + lane;
    return TileRanges{lo, lo + kBands * kLanes};
  }
  CB_HD void phase_load_and_hull(int *tile, long long t, int band, int lane) const {
    const Ctx x = ctx(t, band, lane);
    const TileRanges R = ranges(tile, lane);
    int m = 0;
    if (x.valid) {
#ifdef __CUDA_ARCH__
#pragma unroll 8
#endif
      for (int r = x.r0; r < x.r1; ++r) tile[r * kLanes + lane] = x.base[(long long)r * e.row_stride];
      TileCol c{tile + lane, kLanes};
      m = band_hull<AXIS>(c, x.r0, x.r1, x.q);
    }
    R.lo(band) = 0;
    R.hi(band) = m;
  }
  CB_HD void phase_join(int *tile, long long t, int band, int lane) const {
    if (band != 0) return;
    const Ctx x = ctx(t, band, lane);
    if (!x.valid) return;
    TileRanges R = ranges(tile, lane);
    const TileCol c{tile + lane, kLanes};
    join_band_hulls<AXIS>(c, R, kBands, band_rows(), x.q);
  }
  CB_HD void phase_fill(int *tile, long long t, int band, int lane) const {
    const Ctx x = ctx(t, band, lane);
    if (!x.valid || x.r0 >= x.r1) return;
    const TileRanges R = ranges(tile, lane);
    const TileCol c{tile + lane, kLanes};
    GridCol o{x.base, e.row_stride};
    fill_band<AXIS>(c, R, o, x.r0, x.r1, kBands, band_rows(), x.q);
  }
};

// |voxel - site| * voxel_size as fp16 bits are produced by the caller; this is the integer part
CB_HD int site_distance_sq(int v, int x, int y, int z) {
  const int dx = coord<0>(v) - x, dy = coord<1>(v) - y, dz = coord<2>(v) - z;
  return dx * dx + dy * dy + dz * dz;
}

}  // namespace edt
}  // namespace cb200
