// cb200_mesh.cuh -- signed distance to a triangle mesh through a bounding-volume hierarchy (mesh obstacles).
//
// Replaces what the reference gets from Warp's mesh queries: compute_local_sdf_with_grad (curobo/_src/geom/data/data_mesh.py:
// 643-700) = wp.mesh_query_point (closest point within max_distance + inside / outside sign) + wp.mesh_eval_position, then
//   signed_dist = |cl - p| * sign,   grad_local = -(cl - p) / |cl - p|   (zero below 1e-6),
// and (max_distance, 0) when nothing lies within max_distance = max(half the bounding-box diagonal, query_distance).
// warp-lang is a third-party dependency that is not vendored (pyproject.toml:37); its BVH and sign test are not restated.
// The structure here is ours:
//   * a binary BVH built on the host (curobo_b200/mesh.py: median splits, <= 4 triangles per leaf) stored in depth-first order with
//     SKIP links: node i's subtree is [i, skip_i), so the traversal needs no stack -- "box nearer than the best hit: step to i + 1,
//     else jump to skip_i" -- and a thread keeps only (best distance, best triangle, feature) in registers;
//   * the exact closest point on a triangle with the feature it lies on (face / edge / vertex; Ericson, Real-Time Collision
//     Detection 5.1.5), and the sign from the ANGLE-WEIGHTED PSEUDO-NORMAL of that feature (Baerentzen & Aanaes 2005): exact
//     inside / outside for closed manifold meshes, no ray casts.
// Node = 2 float4: (box min, skip) (box max, leaf) with leaf = -1 (inner) or first_triangle * 16 + count.
// Triangle = 8 float4: a, b, c, face normal, pseudo-normals of edges ab, bc, ca; the three vertex pseudo-normals ride in the w
// lanes (layout in tri_normal below).  Host and device run the same code (tests/hostmath).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cb200_math.cuh"

namespace cb200 {

struct MeshSet {            // device view of all mesh obstacles (curobo_b200/mesh.py::MeshData)
  const float4 *nodes;      // concatenated BVH nodes, 2 float4 each
  const float4 *tris;       // concatenated triangles, 8 float4 each
  const int32_t *node_off;  // [num_envs * max_n] first node of mesh k
  const int32_t *tri_off;   // [num_envs * max_n] first triangle of mesh k
  const float *dims;        // [num_envs * max_n, 4] bounding-box extents (x, y, z, pad)
  const float *inv_pose;    // [num_envs * max_n, 8]
  const uint8_t *enable;
  const int32_t *count;     // [num_envs]
  int max_n, num_envs;
};

CB_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  union {
    float f;
    int i;
  } u;
  u.f = f;
  return u.i;
#endif
}

struct TriHit {
  V3 cl;     // closest point
  int feat;  // 0 face, 1 / 2 / 3 vertex a / b / c, 4 / 5 / 6 edge ab / bc / ca
};
CB_HD TriHit closest_point_on_triangle(V3 p, V3 a, V3 b, V3 c) {
  const V3 ab = b - a, ac = c - a, ap = p - a;
  const float d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0.0f && d2 <= 0.0f) return TriHit{a, 1};
  const V3 bp = p - b;
  const float d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0.0f && d4 <= d3) return TriHit{b, 2};
  const float vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) return TriHit{a + (d1 / (d1 - d3)) * ab, 4};
  const V3 cp = p - c;
  const float d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0.0f && d5 <= d6) return TriHit{c, 3};
  const float vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) return TriHit{a + (d2 / (d2 - d6)) * ac, 6};
  const float va = d3 * d6 - d5 * d4;
  if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) return TriHit{b + ((d4 - d3) / ((d4 - d3) + (d5 - d6))) * (c - b), 5};
  const float denom = 1.0f / (va + vb + vc);
  return TriHit{a + (vb * denom) * ab + (vc * denom) * ac, 0};
}

CB_HD float4 ld4(const float4 *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// pseudo-normal of feature `feat` of the triangle stored at t[0..7]
CB_HD V3 tri_normal(const float4 *t, int feat) {
  if (feat == 0 || feat >= 4) {
    const float4 n = ld4(t + (feat == 0 ? 3 : feat));  // 3 face, 4 ab, 5 bc, 6 ca
    return mk3(n.x, n.y, n.z);
  }
  if (feat == 1) return mk3(ld4(t + 0).w, ld4(t + 1).w, ld4(t + 2).w);
  if (feat == 2) return mk3(ld4(t + 3).w, ld4(t + 4).w, ld4(t + 5).w);
  const float4 l = ld4(t + 7);
  return mk3(ld4(t + 6).w, l.x, l.y);
}

// signed distance + local gradient, data_mesh.py:643-700.
// `need_below`: the caller only acts on sdf < need_below (penetration = radius + activation distance - sdf > 0).  A point whose
// distance to the mesh's bounding box is already >= need_below cannot get there (it is outside the mesh and farther from every
// triangle than from the box), so the traversal is skipped -- exact for the collision cost, and what makes a mesh cheap for the
// many spheres that are nowhere near it.
CB_HD SdfGrad mesh_sdf_grad(const float4 *nodes, const float4 *tris, V3 p, float max_distance, float need_below = 3.0e38f) {
  {
    const float4 lo = ld4(nodes), hi = ld4(nodes + 1);
    const float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.0f), dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.0f),
                dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.0f);
    if (need_below < 1.0e37f && dx * dx + dy * dy + dz * dz >= need_below * need_below) {
      SdfGrad far;
      far.sdf = need_below;
      far.n = mk3(0.f, 0.f, 0.f);
      return far;
    }
  }
  float best2 = max_distance * max_distance;
  int best_t = -1, best_f = 0;
  V3 best_c = p;
  const int end = f2i(ld4(nodes).w);  // the root's skip link = number of nodes
  int i = 0;
  while (i < end) {
    const float4 lo = ld4(nodes + 2 * i), hi = ld4(nodes + 2 * i + 1);
    const float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.0f), dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.0f),
                dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.0f);
    if (dx * dx + dy * dy + dz * dz < best2) {
      const int leaf = f2i(hi.w);
      if (leaf >= 0) {
        const int first = leaf >> 4, cnt = leaf & 15;
        for (int k = 0; k < cnt; ++k) {
          const float4 *t = tris + (size_t)(first + k) * 8;
          const float4 a = ld4(t), b = ld4(t + 1), c = ld4(t + 2);
          const TriHit h = closest_point_on_triangle(p, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z));
          const V3 d = h.cl - p;
          const float d2 = dot(d, d);
          if (d2 < best2) {
            best2 = d2;
            best_t = first + k;
            best_f = h.feat;
            best_c = h.cl;
          }
        }
      }
      ++i;
    } else {
      i = f2i(lo.w);
    }
  }
  SdfGrad o;
  if (best_t < 0) {  // nothing within max_distance
    o.sdf = max_distance;
    o.n = mk3(0.f, 0.f, 0.f);
    return o;
  }
  const V3 delta = best_c - p;
  const float len = sqrtf(dot(delta, delta));
  const V3 nf = tri_normal(tris + (size_t)best_t * 8, best_f);
  const float sign = dot(p - best_c, nf) >= 0.0f ? 1.0f : -1.0f;
  o.sdf = len * sign;
  o.n = len > 1e-6f ? (-1.0f / len) * delta : mk3(0.f, 0.f, 0.f);
  return o;
}

}  // namespace cb200
