// cb200_bspline.cuh -- uniform B-spline knot -> (position, velocity, acceleration, jerk) and its adjoint.
//
// SURVEY.md section 8(f) rank 1.  Arithmetic follows the reference's MATRIX basis backend
// (curobo/_src/curobolib/kernels/trajectory/bspline/basis/bspline_basis_matrix.cuh, instantiated by
// backends/cuda_core_backend/trajectory.py:71,163,252) so that results agree with the reference's kernels to the
// last bit where the compiler contracts the same way; the structure is ours:
//
//   * every control point of every spline segment is addressed by ONE "virtual knot index"
//     v = segment - S + i  in [-S, n_knots + S):  v < 0 -> fixed knot from the start state, 0 <= v < n_knots-1 -> user
//     knot, v >= n_knots-1 -> last user knot (replicate mode) or fixed knot from the goal state (implicit mode).
//     This one rule is the reference's three assignment patterns
//     (bspline_boundary_constraint.cuh:121-265, tables in bspline_interpolation.cuh:96-167).
//   * the adjoint walks (interpolation step j) x (support slot i) per knot in registers; no shuffles.
//
// Everything is __host__ __device__ so tests/hostmath can run it on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef CB_HD
#define CB_HD __host__ __device__ __forceinline__
#endif

namespace cb200 {
namespace bspline {

constexpr float kMinDt = 1e-6f;  // common/curobo_constants.h:28 (fp32Precision)

// basis coefficient tables: row = basis function, column = power t^DEG .. t^0
template <int DEG>
struct Coeffs;
template <>
struct Coeffs<3> {
  static CB_HD float m(int i, int j) {
    constexpr float t[4][4] = {{-1.0f / 6.0f, 3.0f / 6.0f, -3.0f / 6.0f, 1.0f / 6.0f},
                               {3.0f / 6.0f, -6.0f / 6.0f, 0.0f, 4.0f / 6.0f},
                               {-3.0f / 6.0f, 3.0f / 6.0f, 3.0f / 6.0f, 1.0f / 6.0f},
                               {1.0f / 6.0f, 0.0f, 0.0f, 0.0f}};
    return t[i][j];
  }
  // virtual knots that reproduce a boundary state (p, v*dt, a*dt^2, j*dt^3): row = derivative order
  static CB_HD float fixed(int r, int i) {
    constexpr float t[4][4] = {{1.0f, 1.0f, 1.0f, 1.0f},
                               {-1.0f, 0.0f, 1.0f, 2.0f},
                               {1.0f / 3.0f, -1.0f / 6.0f, 1.0f / 3.0f, 11.0f / 6.0f},
                               {0.0f, 0.0f, 0.0f, 0.0f}};
    return t[r][i];
  }
};
template <>
struct Coeffs<4> {
  static CB_HD float m(int i, int j) {
    constexpr float t[5][5] = {{1.0f / 24.0f, -4.0f / 24.0f, 6.0f / 24.0f, -4.0f / 24.0f, 1.0f / 24.0f},
                               {-4.0f / 24.0f, 12.0f / 24.0f, -6.0f / 24.0f, -12.0f / 24.0f, 11.0f / 24.0f},
                               {6.0f / 24.0f, -12.0f / 24.0f, -6.0f / 24.0f, 12.0f / 24.0f, 11.0f / 24.0f},
                               {-4.0f / 24.0f, 4.0f / 24.0f, 6.0f / 24.0f, 4.0f / 24.0f, 1.0f / 24.0f},
                               {1.0f / 24.0f, 0.0f, 0.0f, 0.0f, 0.0f}};
    return t[i][j];
  }
  static CB_HD float fixed(int r, int i) {
    constexpr float t[4][5] = {{1.0f, 1.0f, 1.0f, 1.0f, 1.0f},
                               {-3.0f / 2.0f, -1.0f / 2.0f, 1.0f / 2.0f, 3.0f / 2.0f, 5.0f / 2.0f},
                               {11.0f / 12.0f, -1.0f / 12.0f, -1.0f / 12.0f, 11.0f / 12.0f, 35.0f / 12.0f},
                               {-3.0f / 12.0f, 1.0f / 12.0f, -1.0f / 12.0f, 3.0f / 12.0f, 25.0f / 12.0f}};
    return t[r][i];
  }
};
template <>
struct Coeffs<5> {
  static CB_HD float m(int i, int j) {
    constexpr float t[6][6] = {
        {-1.0f / 120.0f, 5.0f / 120.0f, -10.0f / 120.0f, 10.0f / 120.0f, -5.0f / 120.0f, 1.0f / 120.0f},
        {5.0f / 120.0f, -20.0f / 120.0f, 20.0f / 120.0f, 20.0f / 120.0f, -50.0f / 120.0f, 26.0f / 120.0f},
        {-10.0f / 120.0f, 30.0f / 120.0f, -0.0f / 120.0f, -60.0f / 120.0f, 0.0f / 120.0f, 66.0f / 120.0f},
        {10.0f / 120.0f, -20.0f / 120.0f, -20.0f / 120.0f, 20.0f / 120.0f, 50.0f / 120.0f, 26.0f / 120.0f},
        {-5.0f / 120.0f, 5.0f / 120.0f, 10.0f / 120.0f, 10.0f / 120.0f, 5.0f / 120.0f, 1.0f / 120.0f},
        {1.0f / 120.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}};
    return t[i][j];
  }
  static CB_HD float fixed(int r, int i) {
    // the reference keeps six-digit decimals for the jerk row (bspline_boundary_constraint.cuh:66-70)
    constexpr float t[4][6] = {{1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f},
                               {-2.0f, -1.0f, 0.0f, 1.0f, 2.0f, 3.0f},
                               {1.75f, 0.25f, -0.25f, 0.25f, 1.75f, 4.25f},
                               {-0.833333f, 0.083333f, 0.0f, -0.083333f, 0.833333f, 3.75f}};
    return t[r][i];
  }
};

// Power vectors of the k-th derivative of (t^DEG, .., t, 1), written with the reference's association
// (bspline_basis_matrix.cuh:47-151) so products round identically.
template <int DEG, int K>
CB_HD void power_vector(float t, float *tp);
template <> CB_HD void power_vector<3, 0>(float t, float *tp) { tp[0] = t * t * t; tp[1] = t * t; tp[2] = t; tp[3] = 1.0f; }
template <> CB_HD void power_vector<3, 1>(float t, float *tp) { tp[0] = 3.0f * t * t; tp[1] = 2.0f * t; tp[2] = 1.0f; }
template <> CB_HD void power_vector<3, 2>(float t, float *tp) { tp[0] = 6.0f * t; tp[1] = 2.0f; }
template <> CB_HD void power_vector<3, 3>(float, float *tp) { tp[0] = 6.0f; }
template <> CB_HD void power_vector<4, 0>(float t, float *tp) { tp[0] = t * t * t * t; tp[1] = t * t * t; tp[2] = t * t; tp[3] = t; tp[4] = 1.0f; }
template <> CB_HD void power_vector<4, 1>(float t, float *tp) { tp[0] = 4.0f * t * t * t; tp[1] = 3.0f * t * t; tp[2] = 2.0f * t; tp[3] = 1.0f; }
template <> CB_HD void power_vector<4, 2>(float t, float *tp) { tp[0] = 12.0f * t * t; tp[1] = 6.0f * t; tp[2] = 2.0f; }
template <> CB_HD void power_vector<4, 3>(float t, float *tp) { tp[0] = 24.0f * t; tp[1] = 6.0f; }
template <> CB_HD void power_vector<5, 0>(float t, float *tp) { tp[0] = t * t * t * t * t; tp[1] = t * t * t * t; tp[2] = t * t * t; tp[3] = t * t; tp[4] = t; tp[5] = 1.0f; }
template <> CB_HD void power_vector<5, 1>(float t, float *tp) { tp[0] = 5.0f * t * t * t * t; tp[1] = 4.0f * t * t * t; tp[2] = 3.0f * t * t; tp[3] = 2.0f * t; tp[4] = 1.0f; }
template <> CB_HD void power_vector<5, 2>(float t, float *tp) { tp[0] = 20.0f * t * t * t; tp[1] = 12.0f * t * t; tp[2] = 6.0f * t; tp[3] = 2.0f; }
template <> CB_HD void power_vector<5, 3>(float t, float *tp) { tp[0] = 60.0f * t * t; tp[1] = 24.0f * t; tp[2] = 6.0f; }

// K-th derivative basis (in units of knot spacing): basis[i] = sum_{j < S-K} m(i,j) * tp[j], accumulated from 0
// left to right like curobo::common::partial_matrix_vector_product (common/math.cuh:88-104).
template <int DEG, int K>
CB_HD void basis(float t, float *b) {
  constexpr int S = DEG + 1;
  float tp[S];
  power_vector<DEG, K>(t, tp);
#pragma unroll
  for (int i = 0; i < S; ++i) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < S - K; ++j) acc += Coeffs<DEG>::m(i, j) * tp[j];
    b[i] = acc;
  }
}

// normalised time of row h inside its segment (bspline_interpolation.cuh:236-240)
CB_HD float segment_time(int h, int steps) {
  return (steps > 0) ? (float(h) / float(steps)) - float(int(h / steps)) : 0.0f;
}

struct BoundaryState {
  float p, v, a, j;
};

// One trajectory's (b, d) view of the control polygon.
template <int DEG>
struct ControlPolygon {
  static constexpr int S = DEG + 1;
  const float *knots;  // &u[b, 0, d], stride `stride` between knots
  int stride, n_knots;
  bool implicit_goal;
  float kdt, kdt2, kdt3;
  BoundaryState start, goal;

  CB_HD float fixed_knot(const BoundaryState &s, int i) const {
    // bspline_boundary_constraint.cuh:112-117 (same association)
    return Coeffs<DEG>::fixed(0, i) * s.p + Coeffs<DEG>::fixed(1, i) * s.v * kdt + Coeffs<DEG>::fixed(2, i) * s.a * kdt2 +
           Coeffs<DEG>::fixed(3, i) * s.j * kdt3;
  }

  // control point i of segment `seg`
  CB_HD float point(int seg, int i) const {
    const int v = seg - S + i;
    if (seg < S) {  // start boundary wins (bspline_boundary_constraint.cuh:345-348)
      if (v < 0) return fixed_knot(start, v + S);
      return (v < n_knots) ? knots[(size_t)v * stride] : 0.0f;
    }
    if (v < n_knots - 1) return knots[(size_t)v * stride];
    if (implicit_goal) return fixed_knot(goal, v - (n_knots - 1));
    return knots[(size_t)(n_knots - 1) * stride];
  }
};

struct State4 {
  float p, v, a, j;
};

// Spline state at padded-horizon row h.  horizon = padded_horizon - 1; rows past the last segment evaluate the
// last segment at t = 1 (bspline_interpolation.cuh:72-80,240).
template <int DEG>
CB_HD State4 evaluate(const ControlPolygon<DEG> &cp, int h, int steps) {
  constexpr int S = DEG + 1;
  const int n_seg = cp.n_knots + S;
  int seg = (steps > 0) ? h / steps : 0;
  float t = segment_time(h, steps);
  if (seg >= n_seg) {
    seg = n_seg - 1;
    t = 1.0f;
  }
  float c[S], b[S];
#pragma unroll
  for (int i = 0; i < S; ++i) c[i] = cp.point(seg, i);
  State4 o;
  float r = 0.0f;
  basis<DEG, 0>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) r += c[i] * b[i];
  o.p = r;
  r = 0.0f;
  basis<DEG, 1>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) r += c[i] * b[i];
  o.v = r / cp.kdt;
  r = 0.0f;
  basis<DEG, 2>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) r += c[i] * b[i];
  o.a = r / cp.kdt2;
  r = 0.0f;
  basis<DEG, 3>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) r += c[i] * b[i];
  o.j = r / cp.kdt3;
  return o;
}

template <int DEG>
CB_HD ControlPolygon<DEG> make_polygon(const float *u, int b, int d, int D, int n_knots, float dt, int steps, bool implicit_goal,
                                       const float *sp, const float *sv, const float *sa, const float *sj, int s_row,
                                       const float *gp, const float *gv, const float *ga, const float *gj, int g_row) {
  ControlPolygon<DEG> cp;
  cp.knots = u + (size_t)b * n_knots * D + d;
  cp.stride = D;
  cp.n_knots = n_knots;
  cp.implicit_goal = implicit_goal;
  cp.kdt = fmaxf(dt, kMinDt) * steps;  // bspline_interpolation.cuh:64
  cp.kdt2 = cp.kdt * cp.kdt;
  cp.kdt3 = cp.kdt * cp.kdt * cp.kdt;  // bspline_context.cuh:56
  const size_t si = (size_t)s_row * D + d, gi = (size_t)g_row * D + d;
  cp.start = BoundaryState{sp[si], sv[si], sa[si], sj[si]};
  cp.goal = implicit_goal ? BoundaryState{gp[gi], gv[gi], ga[gi], gj[gi]} : BoundaryState{0.0f, 0.0f, 0.0f, 0.0f};
  return cp;
}

// ---------------------------------------------------------------------------------------------------------
// Adjoint: d loss / d knot[k] from the four row gradients.  Row h = (k+1+i)*steps + j uses knot k at support slot
// S-1-i with t = j/steps (bspline_gradient_util.cuh:85-104, bspline_context.cuh:152-186).
// `G(h, which)` returns grad_{pos,vel,acc,jerk}[b, h, d].
// ---------------------------------------------------------------------------------------------------------
template <int DEG, class Load>
CB_HD float knot_gradient_step(int k, int j, int steps, int n_knots, int horizon, bool implicit_goal, float kdt, float kdt2,
                               float kdt3, Load G) {
  constexpr int S = DEG + 1;
  float g[4][S];
#pragma unroll
  for (int w = 0; w < 4; ++w)
#pragma unroll
    for (int i = 0; i < S; ++i) g[w][i] = 0.0f;
  const int ext = (n_knots + S) * steps;
  const bool dead = implicit_goal && k >= n_knots - 1;  // the knot is overwritten by the goal state
  if (!dead) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const int h = (k + 1 + i) * steps + j;
      if (h < ext) {
#pragma unroll
        for (int w = 0; w < 4; ++w) g[w][i] = G(h, w);
      }
    }
  }
  if (!implicit_goal && k == n_knots - 1) {
    // replicate mode: the last knot also fills every later support slot (bspline_gradient_util.cuh:106-125) and
    // the padded last row, whose POSITION gradient alone is collected (:127-147)
#pragma unroll
    for (int i = 1; i < S; ++i)
#pragma unroll
      for (int x = 0; x < i; ++x)
#pragma unroll
        for (int w = 0; w < 4; ++w) g[w][x] += g[w][i];
    if (j == 0) {
      const float last = G(horizon, 0);
#pragma unroll
      for (int x = 0; x < S; ++x) g[0][x] += last;
    }
  }
  const float t = segment_time((k + DEG) * steps + j, steps);  // bspline_common.cuh:166,175
  float b[S];
  float sp = 0.0f, sv = 0.0f, sa = 0.0f, sj = 0.0f;
  basis<DEG, 0>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) sp += g[0][i] * b[S - 1 - i];
  basis<DEG, 1>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) sv += g[1][i] * b[S - 1 - i];
  basis<DEG, 2>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) sa += g[2][i] * b[S - 1 - i];
  basis<DEG, 3>(t, b);
#pragma unroll
  for (int i = 0; i < S; ++i) sj += g[3][i] * b[S - 1 - i];
  return sp + (sv / kdt) + (sa / kdt2) + (sj / kdt3);
}

// Sum over the interpolation steps.  For power-of-two step counts the additions are done in the order of the
// reference's shuffle-down tree (bspline_gradient_util.cuh:34-55: v[j] += v[j + n/2], then n/4, ...), so results
// match it bit for bit; other step counts (where that tree mis-pairs lanes) use a plain left-to-right sum.
template <int DEG, class Load>
CB_HD float knot_gradient(int k, int steps, int n_knots, int horizon, bool implicit_goal, float dt, Load G) {
  const float kdt = dt * steps;  // not clamped in the adjoint (bspline_common.cuh:172-173)
  const float kdt2 = kdt * kdt, kdt3 = kdt * kdt * kdt;
  if (steps <= 32 && (steps & (steps - 1)) == 0) {
    float part[32];
    for (int j = 0; j < steps; ++j)
      part[j] = knot_gradient_step<DEG>(k, j, steps, n_knots, horizon, implicit_goal, kdt, kdt2, kdt3, G);
    for (int half = steps >> 1; half >= 1; half >>= 1)
      for (int j = 0; j < half; ++j) part[j] += part[j + half];
    return part[0];
  }
  float tot = 0.0f;
  for (int j = 0; j < steps; ++j) tot += knot_gradient_step<DEG>(k, j, steps, n_knots, horizon, implicit_goal, kdt, kdt2, kdt3, G);
  return tot;
}

}  // namespace bspline
}  // namespace cb200
