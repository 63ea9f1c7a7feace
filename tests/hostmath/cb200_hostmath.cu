// TEST-ONLY shared library: exposes the scalar __host__ __device__ building blocks of
// curobo_b200/csrc/cb200_math.cuh to the CPU test-suite so the arithmetic is checked against oracle/
// without a GPU.  It is NOT part of the product library (libcurobo_b200.so has no host compute path).
#include <stdint.h>

#include "../../curobo_b200/csrc/cb200_math.cuh"

using namespace cb200;

extern "C" {

void hm_local_transform(const float *fixed12, int jt, float theta, float *out12) {
  local_link_transform(fixed12, jt, theta, out12);
}

void hm_quat_from_transform(const float *t12, float *wxyz) {
  Q4 q = quat_from_transform(t12);
  wxyz[0] = q.w;
  wxyz[1] = q.x;
  wxyz[2] = q.y;
  wxyz[3] = q.z;
}

// spheres [B,H,S,4]; sweep/speed as in the product kernel; single env 0
void hm_scene(const float *spheres, int B, int H, int S, float weight, float eta, int sweep, int speed, float dt,
              const float *cub_dims, const float *cub_inv_pose, const uint8_t *cub_enable, const int32_t *cub_count,
              int cub_max_n, const float *vox_params, const float *vox_inv_pose, const uint8_t *vox_enable,
              const int32_t *vox_count, const uint16_t *vox_feat, int vox_nvox, int vox_max_n, float vox_max_dist,
              float *out_cost, float *out_grad, const uint16_t *vox_mip, int vox_mip_stride) {
  CuboidSet cs{};
  VoxelSet vs{};
  if (cub_inv_pose) cs = CuboidSet{cub_dims, cub_inv_pose, cub_enable, cub_count, cub_max_n, 1};
  if (vox_inv_pose) vs = VoxelSet{vox_params, vox_inv_pose, vox_enable, vox_count, vox_feat, vox_nvox, vox_max_n, 1, vox_max_dist, vox_mip, vox_mip_stride};
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int s = 0; s < S; ++s) {
        const long long i = ((long long)b * H + h) * S + s;
        const float *sp = spheres + 4 * i;
        V3 c = mk3(sp[0], sp[1], sp[2]);
        V3 g = mk3(0, 0, 0);
        float cost;
        if (!sweep) {
          cost = sphere_scene_discrete(c, sp[3], eta, weight, cs, vs, 0, g);
        } else {
          bool hp = h > 0, hn = h < H - 1;
          V3 pv = c, nx = c;
          if (hp) pv = mk3(sp[-4 * S], sp[-4 * S + 1], sp[-4 * S + 2]);
          if (hn) nx = mk3(sp[4 * S], sp[4 * S + 1], sp[4 * S + 2]);
          cost = sphere_scene_swept(c, sp[3], eta, weight, hp, pv, hn, nx, cs, vs, 0, g);
          if (speed && hp && hn) speed_metric(pv, c, nx, dt, cost, g);
        }
        out_cost[i] = cost;
        out_grad[4 * i] = g.x;
        out_grad[4 * i + 1] = g.y;
        out_grad[4 * i + 2] = g.z;
        out_grad[4 * i + 3] = 0.f;
      }
}

// one tool frame
void hm_tool_pose(const float *pos3, const float *quat_wxyz, const float *goal_pos, const float *goal_quat, int n_goalset,
                  float w_pos, float w_rot, const float *axes6, float tol_p, float tol_r, int method, float *out /*12*/,
                  int *goal_idx) {
  PoseOut o = tool_pose_cost(mk3(pos3[0], pos3[1], pos3[2]), Q4{quat_wxyz[1], quat_wxyz[2], quat_wxyz[3], quat_wxyz[0]},
                             goal_pos, goal_quat, n_goalset, w_pos, w_rot, axes6, 0, tol_p, tol_r, method);
  out[0] = o.pos_cost;
  out[1] = o.rot_cost;
  out[2] = o.pos_err;
  out[3] = o.rot_err;
  out[4] = o.g_pos.x;
  out[5] = o.g_pos.y;
  out[6] = o.g_pos.z;
  out[7] = o.gq_w;
  out[8] = o.gq_x;
  out[9] = o.gq_y;
  out[10] = o.gq_z;
  *goal_idx = o.goal_idx;
}
}

// ------------------------------------------------------------------------------------------------
// B-spline knot -> state and adjoint, host build of curobo_b200/csrc/cb200_bspline.cuh
// ------------------------------------------------------------------------------------------------
#include "../../curobo_b200/csrc/cb200_bspline.cuh"

namespace bs = cb200::bspline;

template <int DEG>
static void hm_bspline_forward_t(float *op, float *ov, float *oa, float *oj, const float *u, const float *sp, const float *sv,
                                 const float *sa, const float *sj, const float *gp, const float *gv, const float *ga,
                                 const float *gj, const int32_t *sidx, const int32_t *gidx, const float *traj_dt,
                                 const uint8_t *implicit, const int32_t *interp_h, int B, int T, int D, int nk) {
  for (int b = 0; b < B; ++b) {
    int padded = T;
    float dt = interp_h ? traj_dt[0] : traj_dt[gidx[b]];
    if (interp_h) padded = (interp_h[b] < T - 1 ? interp_h[b] : T - 1) + 1;
    const int steps = (padded - 1) / (nk + DEG + 1);
    for (int d = 0; d < D; ++d) {
      bs::ControlPolygon<DEG> cp = bs::make_polygon<DEG>(u, b, d, D, nk, dt, steps, implicit[gidx[b]] != 0, sp, sv, sa, sj,
                                                        sidx[b], gp, gv, ga, gj, gidx[b]);
      for (int h = 0; h < T; ++h) {
        bs::State4 s = bs::evaluate<DEG>(cp, h, steps);
        const size_t o = ((size_t)b * T + h) * D + d;
        op[o] = s.p;
        ov[o] = s.v;
        oa[o] = s.a;
        oj[o] = s.j;
      }
    }
  }
}

template <int DEG>
static void hm_bspline_backward_t(float *out, const float *gp, const float *gv, const float *ga, const float *gj,
                                  const float *traj_dt, const int32_t *dt_idx, const uint8_t *implicit, int B, int T, int D,
                                  int nk) {
  const int horizon = T - 1, steps = horizon / (nk + DEG + 1);
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < nk; ++k)
      for (int d = 0; d < D; ++d) {
        const size_t base = (size_t)b * T * D + d;
        auto G = [=](int h, int which) -> float {
          const float *src = which == 0 ? gp : which == 1 ? gv : which == 2 ? ga : gj;
          return src[base + (size_t)h * D];
        };
        out[((size_t)b * nk + k) * D + d] =
            bs::knot_gradient<DEG>(k, steps, nk, horizon, implicit[dt_idx[b]] != 0, traj_dt[dt_idx[b]], G);
      }
}

extern "C" {
int hm_bspline_forward(float *op, float *ov, float *oa, float *oj, const float *u, const float *sp, const float *sv,
                       const float *sa, const float *sj, const float *gp, const float *gv, const float *ga, const float *gj,
                       const int32_t *sidx, const int32_t *gidx, const float *traj_dt, const uint8_t *implicit,
                       const int32_t *interp_h, int B, int T, int D, int nk, int degree) {
  if (degree == 3) hm_bspline_forward_t<3>(op, ov, oa, oj, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, interp_h, B, T, D, nk);
  else if (degree == 4) hm_bspline_forward_t<4>(op, ov, oa, oj, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, interp_h, B, T, D, nk);
  else if (degree == 5) hm_bspline_forward_t<5>(op, ov, oa, oj, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, interp_h, B, T, D, nk);
  else return 1;
  return 0;
}

int hm_bspline_backward(float *out, const float *gp, const float *gv, const float *ga, const float *gj, const float *traj_dt,
                        const int32_t *dt_idx, const uint8_t *implicit, int B, int T, int D, int nk, int degree) {
  if (degree == 3) hm_bspline_backward_t<3>(out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, T, D, nk);
  else if (degree == 4) hm_bspline_backward_t<4>(out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, T, D, nk);
  else if (degree == 5) hm_bspline_backward_t<5>(out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, T, D, nk);
  else return 1;
  return 0;
}
}

// ------------------------------------------------------------------------------------------------
// RNEA inverse dynamics + adjoint, host build of curobo_b200/csrc/cb200_dynamics.cuh
// ------------------------------------------------------------------------------------------------
#include <vector>

#include "../../curobo_b200/csrc/cb200_dynamics.cuh"

namespace dy = cb200::dyn;

struct HostStore {
  std::vector<float> buf;
  int nl;
  HostStore(int arrays, int nl_) : buf((size_t)arrays * nl_ * 6, 0.0f), nl(nl_) {}
  float get(int arr, int k, int c) const { return buf[((size_t)arr * nl + k) * 6 + c]; }
  void set(int arr, int k, int c, float v) { buf[((size_t)arr * nl + k) * 6 + c] = v; }
};

extern "C" {
void hm_rnea_forward(float *tau, float *cache, const float *q, const float *qd, const float *qdd, const float *fixed,
                     const float *mc, const float *inertia, const int8_t *jtype, const int16_t *jmap, const int16_t *lmap,
                     const float *joff, const float *gravity, const int16_t *lstarts, const int16_t *llinks, int B, int nl,
                     int D, int n_levels) {
  dy::Model M{fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, nl, D, n_levels};
  for (int b = 0; b < B; ++b) {
    HostStore S(2, nl);
    dy::rnea_forward_row(M, S, q + (size_t)b * D, qd + (size_t)b * D, qdd + (size_t)b * D, nullptr, tau + (size_t)b * D,
                         cache + (size_t)b * nl * dy::kCacheFloatsPerLink);
  }
}

void hm_rnea_backward(float *gq, float *gqd, float *gqdd, const float *grad_tau, const float *q, const float *qd,
                      const float *cache, const float *fixed, const float *mc, const float *inertia, const int8_t *jtype,
                      const int16_t *jmap, const int16_t *lmap, const float *joff, const float *gravity,
                      const int16_t *lstarts, const int16_t *llinks, int B, int nl, int D, int n_levels) {
  dy::Model M{fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, nl, D, n_levels};
  for (int b = 0; b < B; ++b) {
    HostStore S(5, nl);
    dy::rnea_backward_row(M, S, grad_tau + (size_t)b * D, q + (size_t)b * D, qd + (size_t)b * D,
                          cache + (size_t)b * nl * dy::kCacheFloatsPerLink, gq + (size_t)b * D, gqd + (size_t)b * D,
                          gqdd + (size_t)b * D, nullptr);
  }
}
}

// ---- exact nearest-site transform (cb200_edt.cuh): the three passes of cb200_pba3d with the same column routines, the
// shared-memory tile replaced by a scratch column and the coalesced global column by a strided view of the grid ---------
#include <algorithm>

#include "../../curobo_b200/csrc/cb200_edt.cuh"
namespace ed = cb200::edt;
namespace {
struct HostCol {
  int *base;
  long long stride;
  int get(int r) const { return base[(long long)r * stride]; }
  void set(int r, int v) const { base[(long long)r * stride] = v; }
};
}  // namespace
// The kernels' own schedule, emulated: `n_ctas` CTAs stride over the tiles of each envelope pass exactly like
// edt_envelope_kernel does; the kBands x 32 threads of a CTA run one after the other between the barriers (phase by phase).
// The z pass is the sequential statement of the warp-scan kernel (the kernel itself runs under tests/simt).
template <int AXIS>
static void hm_banded_pass(const ed::BandedEnvelope<AXIS> &bp, std::vector<int> &tile, int n_ctas) {
  tile.assign((size_t)bp.smem_ints(), 0);
  for (int cta = 0; cta < n_ctas; ++cta)
    for (long long t = cta; t < bp.e.ntiles(); t += n_ctas) {
      for (int band = 0; band < ed::kBands; ++band)
        for (int lane = 0; lane < ed::kLanes; ++lane) bp.phase_load_and_hull(tile.data(), t, band, lane);
      for (int band = 0; band < ed::kBands; ++band)
        for (int lane = 0; lane < ed::kLanes; ++lane) bp.phase_join(tile.data(), t, band, lane);
      for (int band = 0; band < ed::kBands; ++band)
        for (int lane = 0; lane < ed::kLanes; ++lane) bp.phase_fill(tile.data(), t, band, lane);
    }
}
extern "C" {
void hm_pba3d(int32_t *grid, int nx, int ny, int nz) {
  std::vector<int> scratch((size_t)(nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz)));
  for (long long row = 0; row < (long long)nx * ny; ++row) {  // pass 1: flood along z, in place
    HostCol c{grid + row * nz, 1};
    ed::flood_column<2>(c, nz);
  }
  const long long plane = (long long)ny * nz;
  for (int x = 0; x < nx; ++x)  // pass 2: envelope along y; the column is staged (the stack is built in place over it)
    for (int z = 0; z < nz; ++z) {
      int *base = grid + x * plane + z;
      for (int r = 0; r < ny; ++r) scratch[r] = base[(long long)r * nz];
      HostCol c{scratch.data(), 1}, o{base, nz};
      ed::envelope_column<1>(c, o, ny, ed::Voxel{x, 0, z});
    }
  for (long long col = 0; col < plane; ++col) {  // pass 3: envelope along x
    int *base = grid + col;
    for (int r = 0; r < nx; ++r) scratch[r] = base[(long long)r * plane];
    HostCol c{scratch.data(), 1}, o{base, plane};
    ed::envelope_column<0>(c, o, nx, ed::Voxel{0, (int)(col / nz), (int)(col % nz)});
  }
}

void hm_pba3d_tiles(int32_t *grid, int nx, int ny, int nz, int n_ctas) {
  const ed::Plan p = ed::make_plan(grid, nx, ny, nz);
  for (long long row = 0; row < p.z.nrows; ++row) {
    HostCol c{grid + row * nz, 1};
    ed::flood_column<2>(c, nz);
  }
  std::vector<int> tile;
  hm_banded_pass(ed::BandedEnvelope<1>{p.y}, tile, n_ctas);
  hm_banded_pass(ed::BandedEnvelope<0>{p.x}, tile, n_ctas);
}
}

// ---- mesh obstacles (cb200_mesh.cuh): the device routine over a host-built BVH -----------------------------------------
#include "../../curobo_b200/csrc/cb200_mesh.cuh"
extern "C" void hm_mesh_sdf(const float *nodes, const float *tris, const float *points, int n, float max_distance, float *out) {
  for (int i = 0; i < n; ++i) {
    const cb200::SdfGrad r = cb200::mesh_sdf_grad(reinterpret_cast<const float4 *>(nodes), reinterpret_cast<const float4 *>(tris),
                                                  cb200::mk3(points[3 * i], points[3 * i + 1], points[3 * i + 2]), max_distance);
    out[4 * i] = r.sdf;
    out[4 * i + 1] = r.n.x;
    out[4 * i + 2] = r.n.y;
    out[4 * i + 3] = r.n.z;
  }
}
