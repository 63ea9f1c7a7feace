// cb200_dynamics_tile.cuh -- RNEA inverse dynamics + adjoint for a TILE of rows held entirely in shared memory, executed by a
// whole CTA: the in-kernel form used by the dynamics-aware trajectory rollout.
//
// Round 1 walked the recursion with lane 0 of the row's warp (31 lanes idle, ~12 k dependent instructions per row: 3.8 x the
// plain trajectory kernel).  Here the CTA switches mapping for the dynamics of R consecutive rows of a seed: a thread is
// (row r = tid % R, worker w = tid / R), i.e. the lanes of a warp are DIFFERENT ROWS walking the same link -- the mapping of
// the stand-alone kernels in cb200_dynamics.cu (rnea_forward_cta / rnea_backward_cta), whose phase structure this follows:
// work that does not depend on the tree recursion runs over (link, row) pairs on all workers; the two recursions run
// level-synchronously, one CTA barrier per depth level, the leaf -> root passes PULL (a link sums its children in level order,
// so every sum keeps the serial order of rnea_forward_kernel.cuh:225-270 / rnea_backward_kernel.cuh:296-460).
// Differences from the stand-alone kernels: no HBM cache -- v, a, f of the forward pass stay in the tile for the adjoint --
// and the joint-space inputs / outputs are exchanged with the caller through the IO rows of the tile.
//
// Tile layout (floats), RS = R + 1 (odd stride: consecutive rows hit consecutive banks):
//   T  [5][nl][6][RS]   0 v, 1 a, 2 f -> f_bar, 3 a_bar, 4 v_bar
//   SC [nl][2][RS]      sin / cos (revolute) or travel (prismatic) of every link's joint
//   IO [8][D][RS]       0 q, 1 qd, 2 qdd, 3 tau -> d cost / d tau, 4 grad_q, 5 grad_qd, 6 grad_qdd, 7 effort cost
#pragma once
#include "cb200_dynamics.cuh"

namespace cb200 {
namespace dyn {

constexpr int kTileIo = 8;
__host__ __device__ inline int tile_floats(int nl, int D, int R) { return (5 * 6 * nl + 2 * nl + kTileIo * D) * (R + 1); }

#if defined(__CUDACC__) || defined(CB200_SIMT_EMULATION)
// who meets at the phase boundaries: the whole CTA (every thread of the CTA runs the tile functions)
struct SyncCta {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <class L>
struct Tile {
  float *T, *SC, *IO;
  int nl, D, RS, r, w, W;
  const Model &M;
  __device__ __forceinline__ float &at(int arr, int k, int c) const { return T[((arr * nl + k) * 6 + c) * RS + r]; }
  __device__ __forceinline__ void load(int arr, int k, float *o) const {
#pragma unroll
    for (int c = 0; c < 6; ++c) o[c] = at(arr, k, c);
  }
  __device__ __forceinline__ void store(int arr, int k, const float *x) const {
#pragma unroll
    for (int c = 0; c < 6; ++c) at(arr, k, c) = x[c];
  }
  __device__ __forceinline__ float &io(int arr, int d) const { return IO[(arr * D + d) * RS + r]; }
  __device__ __forceinline__ Rp rp(int k, int jt) const {
    const float x = SC[(k * 2 + 0) * RS + r], c = SC[(k * 2 + 1) * RS + r];
    return local_Rp_sc<L>(M.fixed_transforms + 12 * k, jt, x, c, x);
  }
};

// tau = RNEA(q, qd, qdd) of every row of the tile.  In: IO 0..2; IO 3 must be zero.  Out: IO 3 = tau; T 0/1/2 = v, a, f.
// Called by every thread of the Sync group (CTA or warp); ends with a barrier of that group.
template <class Sync, class L>
__device__ __forceinline__ void tile_rnea_forward(const Tile<L> &S) {
  const Model &M = S.M;
  const int nl = S.nl, W = S.W;
  float g[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = L::f(M.gravity + i);
  for (int k = S.w; k < nl; k += W) {  // sin / cos of every joint, off the serial chain
    const int jt = M.joint_type[k], ji = M.joint_map[k];
    float qe = 0.0f, sn = 0.0f, cs = 1.0f;
    if (jt >= 0 && ji >= 0) qe = L::f(M.joint_offset + 2 * k) * S.io(0, ji) + L::f(M.joint_offset + 2 * k + 1);
    if (jt >= 3) sincosf(qe, &sn, &cs);
    S.SC[(k * 2 + 0) * S.RS + S.r] = jt >= 3 ? sn : qe;
    S.SC[(k * 2 + 1) * S.RS + S.r] = cs;
  }
  Sync::sync();
  for (int lv = 0; lv < M.n_levels; ++lv) {  // root -> leaves: v, a
    for (int idx = M.level_starts[lv] + S.w; idx < M.level_starts[lv + 1]; idx += W) {
      const int k = M.level_links[idx], jt = M.joint_type[k], ji = M.joint_map[k], par = M.link_map[k];
      const bool root = (par < 0) || (par == k), moving = (jt >= 0) && (ji >= 0);
      const float mul = moving ? L::f(M.joint_offset + 2 * k) : 1.0f;
      const float qd_eff = moving ? mul * S.io(1, ji) : 0.0f, qdd_eff = moving ? mul * S.io(2, ji) : 0.0f;
      const Rp t = S.rp(k, jt);
      float v[6], ac[6], tmp[6];
      if (root) {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = 0.0f;
        Xv(t, g, ac);
      } else {
        S.load(0, par, tmp);
        Xv(t, tmp, v);
        S.load(1, par, tmp);
        Xv(t, tmp, ac);
      }
      if (jt >= 0) {
        const int s = s_index(jt);
        add6(v, s, qd_eff);
        add6(ac, s, qdd_eff);
        motion_cross_S_add(ac, v, s, qd_eff);
      }
      S.store(0, k, v);
      S.store(1, k, ac);
    }
    Sync::sync();
  }
  for (int k = S.w; k < nl; k += W) {  // every (link, row): f = I a + v x* (I v)
    float v[6], ac[6], Ia[6], Iv[6], x[6];
    S.load(0, k, v);
    S.load(1, k, ac);
    inertia_times<L>(M.masses_com + 4 * k, M.inertias + 8 * k, ac, Ia);
    inertia_times<L>(M.masses_com + 4 * k, M.inertias + 8 * k, v, Iv);
    force_cross(v, Iv, x);
#pragma unroll
    for (int i = 0; i < 6; ++i) Ia[i] += x[i];
    S.store(2, k, Ia);
  }
  Sync::sync();
  for (int lv = M.n_levels - 2; lv >= 0; --lv) {  // leaves -> root: a link pulls its children's wrenches
    const int c0 = M.level_starts[lv + 1], c1 = M.level_starts[lv + 2];
    for (int idx = M.level_starts[lv] + S.w; idx < c0; idx += W) {
      const int k = M.level_links[idx];
      float f[6];
      S.load(2, k, f);
      bool any = false;
      for (int ci = c0; ci < c1; ++ci) {
        const int c = M.level_links[ci];
        if (M.link_map[c] != k) continue;
        const Rp t = S.rp(c, M.joint_type[c]);
        float fc[6], x[6];
        S.load(2, c, fc);
        XTf(t, fc, x);
#pragma unroll
        for (int i = 0; i < 6; ++i) f[i] += x[i];
        any = true;
      }
      if (any) S.store(2, k, f);
    }
    Sync::sync();
  }
  for (int k = S.w; k < nl; k += W) {  // joint torques (several links share a joint only through mimic joints)
    const int jt = M.joint_type[k], ji = M.joint_map[k];
    if (jt >= 0 && ji >= 0) atomicAdd(&S.io(3, ji), L::f(M.joint_offset + 2 * k) * S.at(2, k, s_index(jt)));
  }
  Sync::sync();
}

// Adjoint.  In: IO 0 (q), 1 (qd), 3 (d cost / d tau), T 0/1/2 from tile_rnea_forward; IO 4..6 hold the values the three
// gradients are ADDED to (zero, or terms the caller already owns).  Out: IO 4 / 5 / 6 += grad_q / grad_qd / grad_qdd.
template <class Sync, class L>
__device__ __forceinline__ void tile_rnea_backward(const Tile<L> &S) {
  const Model &M = S.M;
  const int nl = S.nl, W = S.W;
  float g[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = L::f(M.gravity + i);
  for (int lv = 0; lv < M.n_levels; ++lv) {  // root -> leaves: f_bar; the dX^T/dq term of grad_q
    for (int idx = M.level_starts[lv] + S.w; idx < M.level_starts[lv + 1]; idx += W) {
      const int k = M.level_links[idx], jt = M.joint_type[k], ji = M.joint_map[k], par = M.link_map[k];
      const bool root = (par < 0) || (par == k), moving = (jt >= 0) && (ji >= 0);
      const float mul = moving ? L::f(M.joint_offset + 2 * k) : 1.0f;
      const int s = jt >= 0 ? s_index(jt) : 0;
      float fk[6], fbar[6] = {0, 0, 0, 0, 0, 0}, gq1 = 0.0f;
      S.load(2, k, fk);
      if (moving) add6(fbar, s, mul * S.io(3, ji));
      if (!root) {
        const Rp t = S.rp(k, jt);
        float fp[6], X[6];
        S.load(2, par, fp);
        Xv(t, fp, X);
#pragma unroll
        for (int i = 0; i < 6; ++i) fbar[i] += X[i];
        if (moving) gq1 = mul * dot_crf_S(X, fk, s);
      }
      S.store(2, k, fbar);
      S.at(3, k, 0) = gq1;
    }
    Sync::sync();
  }
  for (int k = S.w; k < nl; k += W) {  // every (link, row): the adjoint terms that do not involve the children
    const float *mc = M.masses_com + 4 * k, *in = M.inertias + 8 * k;
    float v[6], fbar[6], ab[6], t1[6], t2[6], vb[6];
    S.load(0, k, v);
    S.load(2, k, fbar);
    const float gq1 = S.at(3, k, 0);
    inertia_times<L>(mc, in, fbar, ab);  // a_bar += I f_bar
    inertia_times<L>(mc, in, v, t1);
    force_cross(fbar, t1, t2);  // v_bar -= crf(f_bar) I v
#pragma unroll
    for (int i = 0; i < 6; ++i) vb[i] = 0.0f - t2[i];
    motion_cross(v, fbar, t1);
    inertia_times<L>(mc, in, t1, t2);  // v_bar -= I crm(v) f_bar
#pragma unroll
    for (int i = 0; i < 6; ++i) vb[i] -= t2[i];
    S.store(3, k, ab);
    S.store(4, k, vb);
    S.at(2, k, 0) = gq1;
  }
  Sync::sync();
  for (int lv = M.n_levels - 1; lv >= 0; --lv) {  // leaves -> root: a link pulls its children's a_bar, v_bar
    const int c0 = M.level_starts[lv + 1], c1 = (lv + 1 < M.n_levels) ? M.level_starts[lv + 2] : c0;
    for (int idx = M.level_starts[lv] + S.w; idx < c0; idx += W) {
      const int k = M.level_links[idx], jt = M.joint_type[k], ji = M.joint_map[k];
      float ab[6] = {0, 0, 0, 0, 0, 0}, vb[6] = {0, 0, 0, 0, 0, 0}, x[6], y[6];
      for (int ci = c0; ci < c1; ++ci) {
        const int c = M.level_links[ci];
        if (M.link_map[c] != k) continue;
        const Rp t = S.rp(c, M.joint_type[c]);
        S.load(3, c, x);
        XTf(t, x, y);
#pragma unroll
        for (int i = 0; i < 6; ++i) ab[i] += y[i];
        S.load(4, c, x);
        XTf(t, x, y);
#pragma unroll
        for (int i = 0; i < 6; ++i) vb[i] += y[i];
      }
      S.load(3, k, x);
      S.load(4, k, y);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        ab[i] += x[i];
        vb[i] += y[i];
      }
      if (jt >= 0 && ji >= 0) force_cross_S_add(vb, s_index(jt), L::f(M.joint_offset + 2 * k) * S.io(1, ji), ab);
      S.store(3, k, ab);
      S.store(4, k, vb);
    }
    Sync::sync();
  }
  for (int k = S.w; k < nl; k += W) {  // every moving (link, row): the three joint-space gradients
    const int jt = M.joint_type[k], ji = M.joint_map[k], par = M.link_map[k];
    if (jt < 0 || ji < 0) continue;
    const bool root = (par < 0) || (par == k);
    const float mul = L::f(M.joint_offset + 2 * k);
    const int s = s_index(jt);
    float v[6], ab[6], vb[6], fx[6], X[6], u[6];
    S.load(0, k, v);
    S.load(3, k, ab);
    S.load(4, k, vb);
    const Rp t = S.rp(k, jt);
    force_cross(v, ab, fx);
    const float gqdd = mul * pick6(ab, s);
    const float gqd = (0.0f - mul * pick6(fx, s)) + mul * pick6(vb, s);
    float gq;
    if (!root) {
      S.load(1, par, u);
      Xv(t, u, X);
      gq = S.at(2, k, 0) - mul * dot_crm_S(ab, X, s);
      S.load(0, par, u);
      Xv(t, u, X);
      gq -= mul * dot_crm_S(vb, X, s);
    } else {
      Xv(t, g, X);
      gq = 0.0f - mul * dot_crm_S(ab, X, s);
    }
    atomicAdd(&S.io(4, ji), gq);
    atomicAdd(&S.io(5, ji), gqd);
    atomicAdd(&S.io(6, ji), gqdd);
  }
  Sync::sync();
}
#endif  // __CUDACC__ || CB200_SIMT_EMULATION

}  // namespace dyn
}  // namespace cb200
