// cb200_dynamics.cuh -- RNEA inverse dynamics (tau from q, qd, qdd) and its adjoint, one row at a time.
//
// SURVEY.md section 8(f) rank 3.  Arithmetic follows the reference kernels in their serial order
// (curobo/_src/curobolib/kernels/dynamics/rnea_forward_kernel.cuh:54-285, rnea_backward_kernel.cuh:60-460,
// spatial_algebra.cuh, rnea_helpers.cuh) including their joint-axis specialised operators entry by entry (the
// prismatic case of motion_cross_S lands in the angular slots there; so it does here).
// The structure is ours: a row's per-link spatial vectors live behind a small accessor (`Store`) so the same code
// runs with a transposed shared-memory tile on the GPU (thread per row, conflict-free) and with plain arrays on the
// host (tests/hostmath).  Spatial vectors are Featherstone-ordered: [angular(3); linear(3)].
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifndef CB_HD
#define CB_HD __host__ __device__ __forceinline__
#endif

namespace cb200 {
namespace dyn {

constexpr int kCacheFloatsPerLink = 20;  // dynamics_constants.h: v(6) a(6) f(6) pad(2)
constexpr int kCacheF = 12;

struct Model {  // robot constants (device or host pointers)
  const float *fixed_transforms;  // [nl,12]
  const float *masses_com;        // [nl,4]  cx cy cz m
  const float *inertias;          // [nl,8]  ixx iyy izz ixy ixz iyz pad pad (at the CoM)
  const int8_t *joint_type;       // [nl]
  const int16_t *joint_map;       // [nl]
  const int16_t *link_map;        // [nl]
  const float *joint_offset;      // [nl,2]
  const float *gravity;           // [6]
  const int16_t *level_starts;    // [n_levels+1]
  const int16_t *level_links;     // [nl]
  int nl, D, n_levels;
};

// How the robot's constants are read: through the read-only (non-coherent) path when they sit in global memory -- the
// compiler may then hoist and reorder those loads across the shared-memory traffic of the recursion -- or as plain loads
// when a CTA has staged them into shared memory.
struct LdNc {
  static CB_HD float f(const float *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
  }
};
struct LdPlain {
  static CB_HD float f(const float *p) { return *p; }
};
CB_HD float ld(const float *p) { return LdNc::f(p); }

CB_HD int s_index(int jt) { return jt >= 3 ? jt - 3 : 3 + jt; }

struct Rp {
  float R[9], p[3];
};

// fixed * J(angle): rotation (row-major) and translation  (rnea_helpers.cuh:25-95).  sn / cs = sin / cos of the angle for a
// revolute joint (computed by the caller: the CTA kernels take them off the serial chain); angle is the prismatic travel.
template <class L = LdNc>
CB_HD Rp local_Rp_sc(const float *ft, int jt, float sn, float cs, float angle) {
  Rp o;
  float f[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) f[i] = L::f(ft + i);
  o.p[0] = f[3];
  o.p[1] = f[7];
  o.p[2] = f[11];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o.R[3 * r + 0] = f[4 * r + 0];
    o.R[3 * r + 1] = f[4 * r + 1];
    o.R[3 * r + 2] = f[4 * r + 2];
  }
  if (jt < 0) return o;
  if (jt >= 3) {
    const int ax = jt - 3;
    const float s = sn, c = cs;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float x = f[4 * r + 0], y = f[4 * r + 1], z = f[4 * r + 2];
      if (ax == 0) {
        o.R[3 * r + 1] = c * y + s * z;
        o.R[3 * r + 2] = -s * y + c * z;
      } else if (ax == 1) {
        o.R[3 * r + 0] = c * x - s * z;
        o.R[3 * r + 2] = s * x + c * z;
      } else {
        o.R[3 * r + 0] = c * x + s * y;
        o.R[3 * r + 1] = -s * x + c * y;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float col = jt == 0 ? f[4 * r + 0] : (jt == 1 ? f[4 * r + 1] : f[4 * r + 2]);
      o.p[r] += col * angle;
    }
  }
  return o;
}
template <class L = LdNc>
CB_HD Rp local_Rp(const float *ft, int jt, float angle) {
  float s = 0.0f, c = 1.0f;
  if (jt >= 3) sincosf(angle, &s, &c);
  return local_Rp_sc<L>(ft, jt, s, c, angle);
}

// motion transform parent -> child:  w' = R^T w,  v' = R^T (v + w x p)
CB_HD void Xv(const Rp &t, const float *v, float *o) {
  const float w0 = v[0], w1 = v[1], w2 = v[2];
  const float u0 = v[3] + (w1 * t.p[2] - w2 * t.p[1]);
  const float u1 = v[4] + (w2 * t.p[0] - w0 * t.p[2]);
  const float u2 = v[5] + (w0 * t.p[1] - w1 * t.p[0]);
  const float *R = t.R;
  o[0] = R[0] * w0 + R[3] * w1 + R[6] * w2;
  o[1] = R[1] * w0 + R[4] * w1 + R[7] * w2;
  o[2] = R[2] * w0 + R[5] * w1 + R[8] * w2;
  o[3] = R[0] * u0 + R[3] * u1 + R[6] * u2;
  o[4] = R[1] * u0 + R[4] * u1 + R[7] * u2;
  o[5] = R[2] * u0 + R[5] * u1 + R[8] * u2;
}

// force transform child -> parent:  f' = R f,  n' = R n + p x (R f)
CB_HD void XTf(const Rp &t, const float *f, float *o) {
  const float *R = t.R;
  const float a0 = R[0] * f[3] + R[1] * f[4] + R[2] * f[5];
  const float a1 = R[3] * f[3] + R[4] * f[4] + R[5] * f[5];
  const float a2 = R[6] * f[3] + R[7] * f[4] + R[8] * f[5];
  const float n0 = R[0] * f[0] + R[1] * f[1] + R[2] * f[2];
  const float n1 = R[3] * f[0] + R[4] * f[1] + R[5] * f[2];
  const float n2 = R[6] * f[0] + R[7] * f[1] + R[8] * f[2];
  o[0] = n0 + (t.p[1] * a2 - t.p[2] * a1);
  o[1] = n1 + (t.p[2] * a0 - t.p[0] * a2);
  o[2] = n2 + (t.p[0] * a1 - t.p[1] * a0);
  o[3] = a0;
  o[4] = a1;
  o[5] = a2;
}

CB_HD void motion_cross(const float *a, const float *b, float *r) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  r[3] = a[4] * b[2] - a[5] * b[1] + a[1] * b[5] - a[2] * b[4];
  r[4] = a[5] * b[0] - a[3] * b[2] + a[2] * b[3] - a[0] * b[5];
  r[5] = a[3] * b[1] - a[4] * b[0] + a[0] * b[4] - a[1] * b[3];
}

CB_HD void force_cross(const float *v, const float *f, float *r) {
  r[0] = -v[2] * f[1] + v[1] * f[2] - v[5] * f[4] + v[4] * f[5];
  r[1] = v[2] * f[0] - v[0] * f[2] + v[5] * f[3] - v[3] * f[5];
  r[2] = -v[1] * f[0] + v[0] * f[1] - v[4] * f[3] + v[3] * f[4];
  r[3] = -v[2] * f[4] + v[1] * f[5];
  r[4] = v[2] * f[3] - v[0] * f[5];
  r[5] = -v[1] * f[3] + v[0] * f[4];
}

// Joint-axis operators (spatial_algebra.cuh:68-110,200-252), one case per motion-subspace slot s (0..2 revolute x/y/z,
// 3..5 prismatic x/y/z).  Written as switches over compile-time indices so every spatial vector stays in registers.
CB_HD void motion_cross_S_add(float *acc, const float *v, int s, float alpha) {  // acc += crm(v) S alpha
  switch (s) {
    case 0:
      acc[1] += v[2] * alpha;
      acc[2] += -v[1] * alpha;
      acc[4] += v[5] * alpha;
      acc[5] += -v[4] * alpha;
      break;
    case 1:
      acc[0] += -v[2] * alpha;
      acc[2] += v[0] * alpha;
      acc[3] += -v[5] * alpha;
      acc[5] += v[3] * alpha;
      break;
    case 2:
      acc[0] += v[1] * alpha;
      acc[1] += -v[0] * alpha;
      acc[3] += v[4] * alpha;
      acc[4] += -v[3] * alpha;
      break;
    case 3:
      acc[1] += v[2] * alpha;
      acc[2] += -v[1] * alpha;
      break;
    case 4:
      acc[0] += -v[2] * alpha;
      acc[2] += v[0] * alpha;
      break;
    case 5:
      acc[0] += v[1] * alpha;
      acc[1] += -v[0] * alpha;
      break;
    default:
      break;
  }
}
CB_HD float dot_crf_S(const float *a, const float *b, int s) {  // a . crf(S) b
  float r = 0.0f;
  switch (s) {
    case 0:
      r += -a[1] * b[2];
      r += a[2] * b[1];
      r += -a[4] * b[5];
      r += a[5] * b[4];
      break;
    case 1:
      r += a[0] * b[2];
      r += -a[2] * b[0];
      r += a[3] * b[5];
      r += -a[5] * b[3];
      break;
    case 2:
      r += -a[0] * b[1];
      r += a[1] * b[0];
      r += -a[3] * b[4];
      r += a[4] * b[3];
      break;
    case 3:
      r += -a[1] * b[5];
      r += a[2] * b[4];
      break;
    case 4:
      r += a[0] * b[5];
      r += -a[2] * b[3];
      break;
    case 5:
      r += -a[0] * b[4];
      r += a[1] * b[3];
      break;
    default:
      break;
  }
  return r;
}
CB_HD float dot_crm_S(const float *a, const float *b, int s) {  // a . crm(S) b
  float r = 0.0f;
  switch (s) {
    case 0:
      r += -a[1] * b[2];
      r += a[2] * b[1];
      r += -a[4] * b[5];
      r += a[5] * b[4];
      break;
    case 1:
      r += a[0] * b[2];
      r += -a[2] * b[0];
      r += a[3] * b[5];
      r += -a[5] * b[3];
      break;
    case 2:
      r += -a[0] * b[1];
      r += a[1] * b[0];
      r += -a[3] * b[4];
      r += a[4] * b[3];
      break;
    case 3:
      r += -a[4] * b[2];
      r += a[5] * b[1];
      break;
    case 4:
      r += a[3] * b[2];
      r += -a[5] * b[0];
      break;
    case 5:
      r += -a[3] * b[1];
      r += a[4] * b[0];
      break;
    default:
      break;
  }
  return r;
}
CB_HD float pick6(const float *a, int s) {  // a[s] without dynamic register indexing
  float r = a[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) r = (s == i) ? a[i] : r;
  return r;
}
CB_HD void add6(float *a, int s, float x) {  // a[s] += x
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (s == i) a[i] += x;
}
CB_HD void force_cross_S_add(float *res, int s, float alpha, const float *b) {  // res += crf(S alpha) b
  float e[6], t[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) e[i] = (s == i) ? alpha : 0.0f;
  force_cross(e, b, t);
#pragma unroll
  for (int i = 0; i < 6; ++i) res[i] += t[i];
}

// spatial inertia about the link origin times a motion vector (spatial_algebra.cuh:129-163)
template <class L = LdNc>
CB_HD void inertia_times(const float *mc, const float *in, const float *u, float *r) {
  const float cx = L::f(mc), cy = L::f(mc + 1), cz = L::f(mc + 2), m = L::f(mc + 3);
  const float ixx = L::f(in), iyy = L::f(in + 1), izz = L::f(in + 2), ixy = L::f(in + 3), ixz = L::f(in + 4), iyz = L::f(in + 5);
  const float w0 = u[0], w1 = u[1], w2 = u[2];
  const float h0 = u[3] + w1 * cz - w2 * cy;
  const float h1 = u[4] + w2 * cx - w0 * cz;
  const float h2 = u[5] + w0 * cy - w1 * cx;
  r[3] = m * h0;
  r[4] = m * h1;
  r[5] = m * h2;
  const float I0 = ixx * w0 + ixy * w1 + ixz * w2;
  const float I1 = ixy * w0 + iyy * w1 + iyz * w2;
  const float I2 = ixz * w0 + iyz * w1 + izz * w2;
  r[0] = I0 + m * (cy * h2 - cz * h1);
  r[1] = I1 + m * (cz * h0 - cx * h2);
  r[2] = I2 + m * (cx * h1 - cy * h0);
}

// ---- per-row state behind an accessor: S.get(array, link, comp) / S.set(...) ---------------------------------------
// arrays: 0 v, 1 a|f (forward) ; backward: 0 v, 1 a, 2 f|fbar, 3 abar, 4 vbar
template <class Store>
CB_HD void load6(const Store &S, int arr, int k, float *o) {
#pragma unroll
  for (int c = 0; c < 6; ++c) o[c] = S.get(arr, k, c);
}
template <class Store>
CB_HD void store6(Store &S, int arr, int k, const float *x) {
#pragma unroll
  for (int c = 0; c < 6; ++c) S.set(arr, k, c, x[c]);
}

struct JointRef {
  int jt, ji, par;
  bool root, moving;
  float mul, q_eff;
};
template <class L = LdNc>
CB_HD JointRef joint_ref(const Model &M, int k, const float *q_row) {
  JointRef j;
  j.jt = M.joint_type[k];
  j.ji = M.joint_map[k];
  j.par = M.link_map[k];
  j.root = (j.par < 0) || (j.par == k);
  j.moving = (j.jt >= 0) && (j.ji >= 0);
  j.mul = 1.0f;
  j.q_eff = 0.0f;
  if (j.moving) {
    j.mul = L::f(M.joint_offset + 2 * k);
    j.q_eff = j.mul * q_row[j.ji] + L::f(M.joint_offset + 2 * k + 1);
  }
  return j;
}

// tau[D] (accumulated into; zeroed here) and the row's cache [nl*20].  q / qd / qdd: this row's [D].
template <class Store, class L = LdNc>
CB_HD void rnea_forward_row(const Model &M, Store &S, const float *q, const float *qd, const float *qdd, const float *f_ext,
                            float *tau, float *cache) {
  for (int d = 0; d < M.D; ++d) tau[d] = 0.0f;
  float g[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = L::f(M.gravity + i);
  for (int idx = 0; idx < M.nl; ++idx) {  // level order: parents before children
    const int k = M.level_links[idx];
    const JointRef j = joint_ref<L>(M, k, q);
    const float qd_eff = j.moving ? j.mul * qd[j.ji] : 0.0f, qdd_eff = j.moving ? j.mul * qdd[j.ji] : 0.0f;
    const Rp t = local_Rp<L>(M.fixed_transforms + 12 * k, j.jt, j.q_eff);
    float v[6], a[6], tmp[6];
    if (j.root) {
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = 0.0f;
      Xv(t, g, a);
    } else {
      load6(S, 0, j.par, tmp);
      Xv(t, tmp, v);
      load6(S, 1, j.par, tmp);
      Xv(t, tmp, a);
    }
    if (j.jt >= 0) {
      const int s = s_index(j.jt);
      add6(v, s, qd_eff);
      add6(a, s, qdd_eff);
      motion_cross_S_add(a, v, s, qd_eff);
    }
    store6(S, 0, k, v);
    store6(S, 1, k, a);
  }
  for (int k = 0; k < M.nl; ++k) {  // f = I a + v x* (I v) [- f_ext]; cache v, a
    float v[6], a[6], Ia[6], Iv[6], x[6];
    load6(S, 0, k, v);
    load6(S, 1, k, a);
    float *ck = cache + (size_t)k * kCacheFloatsPerLink;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      ck[i] = v[i];
      ck[6 + i] = a[i];
    }
    inertia_times<L>(M.masses_com + 4 * k, M.inertias + 8 * k, a, Ia);
    inertia_times<L>(M.masses_com + 4 * k, M.inertias + 8 * k, v, Iv);
    force_cross(v, Iv, x);
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = Ia[i] + x[i] - (f_ext ? f_ext[6 * k + i] : 0.0f);
    store6(S, 1, k, a);
  }
  for (int lv = M.n_levels - 1; lv >= 0; --lv) {  // leaves -> root: torques, wrench propagation
    for (int idx = M.level_starts[lv]; idx < M.level_starts[lv + 1]; ++idx) {
      const int k = M.level_links[idx];
      const JointRef j = joint_ref<L>(M, k, q);
      float f[6];
      load6(S, 1, k, f);
      if (j.moving) tau[j.ji] += j.mul * pick6(f, s_index(j.jt));
      if (!j.root) {
        const Rp t = local_Rp<L>(M.fixed_transforms + 12 * k, j.jt, j.q_eff);
        float c[6], fp[6];
        XTf(t, f, c);
        load6(S, 1, j.par, fp);
#pragma unroll
        for (int i = 0; i < 6; ++i) fp[i] += c[i];
        store6(S, 1, j.par, fp);
      }
    }
  }
  for (int k = 0; k < M.nl; ++k) {
    float *ck = cache + (size_t)k * kCacheFloatsPerLink + kCacheF;
#pragma unroll
    for (int c = 0; c < 6; ++c) ck[c] = S.get(1, k, c);
  }
}

// grad_q / grad_qd / grad_qdd [D] (overwritten) from grad_tau [D] and the row's forward cache.
template <class Store, class L = LdNc>
CB_HD void rnea_backward_row(const Model &M, Store &S, const float *grad_tau, const float *q, const float *qd,
                             const float *cache, float *gq, float *gqd, float *gqdd, float *grad_f_ext) {
  for (int d = 0; d < M.D; ++d) gq[d] = gqd[d] = gqdd[d] = 0.0f;
  float g[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = L::f(M.gravity + i);
  for (int k = 0; k < M.nl; ++k) {
    const float *ck = cache + (size_t)k * kCacheFloatsPerLink;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      S.set(0, k, c, ck[c]);
      S.set(1, k, c, ck[6 + c]);
      S.set(2, k, c, ck[kCacheF + c]);
      S.set(3, k, c, 0.0f);
      S.set(4, k, c, 0.0f);
    }
  }
  // pass 1, root -> leaves: f_bar
  for (int idx = 0; idx < M.nl; ++idx) {
    const int k = M.level_links[idx];
    const JointRef j = joint_ref<L>(M, k, q);
    float fk[6], fbar[6] = {0, 0, 0, 0, 0, 0};
    load6(S, 2, k, fk);
    if (j.moving) add6(fbar, s_index(j.jt), j.mul * grad_tau[j.ji]);
    if (!j.root) {
      const Rp t = local_Rp<L>(M.fixed_transforms + 12 * k, j.jt, j.q_eff);
      float fp[6], X[6];
      load6(S, 2, j.par, fp);
      Xv(t, fp, X);
#pragma unroll
      for (int i = 0; i < 6; ++i) fbar[i] += X[i];
      if (j.moving) gq[j.ji] += j.mul * dot_crf_S(X, fk, s_index(j.jt));
    }
    store6(S, 2, k, fbar);
  }
  if (grad_f_ext != nullptr)
    for (int k = 0; k < M.nl; ++k)
      for (int c = 0; c < 6; ++c) grad_f_ext[6 * k + c] = -S.get(2, k, c);
  // pass 2, leaves -> root
  for (int lv = M.n_levels - 1; lv >= 0; --lv) {
    for (int idx = M.level_starts[lv]; idx < M.level_starts[lv + 1]; ++idx) {
      const int k = M.level_links[idx];
      const JointRef j = joint_ref<L>(M, k, q);
      const float *mc = M.masses_com + 4 * k, *in = M.inertias + 8 * k;
      float v[6], fbar[6], ab[6], vb[6], t1[6], t2[6];
      load6(S, 0, k, v);
      load6(S, 2, k, fbar);
      load6(S, 3, k, ab);
      load6(S, 4, k, vb);
      inertia_times<L>(mc, in, fbar, t1);
#pragma unroll
      for (int i = 0; i < 6; ++i) ab[i] += t1[i];
      inertia_times<L>(mc, in, v, t1);
      force_cross(fbar, t1, t2);
#pragma unroll
      for (int i = 0; i < 6; ++i) vb[i] -= t2[i];
      motion_cross(v, fbar, t1);
      inertia_times<L>(mc, in, t1, t2);
#pragma unroll
      for (int i = 0; i < 6; ++i) vb[i] -= t2[i];
      const int s = j.jt >= 0 ? s_index(j.jt) : 0;
      if (j.moving) {
        const float qd_k = j.mul * qd[j.ji];
        gqdd[j.ji] += j.mul * pick6(ab, s);
        float fx[6];
        force_cross(v, ab, fx);
        gqd[j.ji] -= j.mul * pick6(fx, s);
        force_cross_S_add(vb, s, qd_k, ab);
      }
      const Rp t = local_Rp<L>(M.fixed_transforms + 12 * k, j.jt, j.q_eff);
      if (!j.root) {
        float c[6], pa[6];
        XTf(t, ab, c);
        load6(S, 3, j.par, pa);
#pragma unroll
        for (int i = 0; i < 6; ++i) pa[i] += c[i];
        store6(S, 3, j.par, pa);
        if (j.moving) {
          float ap[6], X[6];
          load6(S, 1, j.par, ap);
          Xv(t, ap, X);
          gq[j.ji] -= j.mul * dot_crm_S(ab, X, s);
        }
      } else if (j.moving) {
        float X[6];
        Xv(t, g, X);
        gq[j.ji] -= j.mul * dot_crm_S(ab, X, s);
      }
      if (j.moving) gqd[j.ji] += j.mul * pick6(vb, s);
      if (!j.root) {
        float c[6], pv[6];
        XTf(t, vb, c);
        load6(S, 4, j.par, pv);
#pragma unroll
        for (int i = 0; i < 6; ++i) pv[i] += c[i];
        store6(S, 4, j.par, pv);
        if (j.moving) {
          float vp[6], X[6];
          load6(S, 0, j.par, vp);
          Xv(t, vp, X);
          gq[j.ji] -= j.mul * dot_crm_S(vb, X, s);
        }
      }
    }
  }
}

}  // namespace dyn
}  // namespace cb200
