// cb200_trajectory.cu -- B-spline knot -> state kernels and their adjoint, C ABI.
//
// Replaces the reference's three trajectory launches
//   interpolate_bspline_kernel            (kernels/trajectory/bspline/bspline_kernel.cuh:87-149)
//   interpolate_bspline_single_dt_kernel  (:216-270)
//   bspline_backward_kernel               (:326-373)
// Both are pure HBM streams: the forward writes 4 x [B,T,D] floats from a [B,nk,D] read, the adjoint reads
// 4 x [B,T,D] (each row (DEG+1) times, from L1/L2) and writes [B,nk,D].  One thread per output element with the
// dof index fastest so every warp touches consecutive addresses; no shared memory, no shuffles.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/curobo_b200.h"
#include "cb200_bspline.cuh"
#include "cb200_launch.h"

namespace {
using namespace cb200::bspline;

inline int status(cudaError_t e) {
  if (e != cudaSuccess) (void)cudaGetLastError();  // never leave a sticky error behind for the caller's next CUDA call
  return (int)e;
}

struct FwdArgs {
  float *out_p, *out_v, *out_a, *out_j, *out_dt;
  const float *u;
  const float *sp, *sv, *sa, *sj, *gp, *gv, *ga, *gj;
  const int32_t *start_idx, *goal_idx;
  const float *traj_dt;                   // [n_goal] indexed by goal_idx, or a single value (single-dt mode)
  const uint8_t *implicit;                // [n_goal]
  const int32_t *interpolation_horizon;   // single-dt mode: per-batch horizon; nullptr otherwise
  int B, T, D, n_knots;
};

template <int DEG>
__global__ void __launch_bounds__(256) bspline_forward_kernel(const __grid_constant__ FwdArgs a) {
  const long long n = (long long)a.B * a.T * a.D;
  for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(tid % a.D);
    const int h = (int)((tid / a.D) % a.T);
    const int b = (int)(tid / ((long long)a.D * a.T));
    const int s_row = __ldg(a.start_idx + b), g_row = __ldg(a.goal_idx + b);
    int padded = a.T;
    float dt;
    if (a.interpolation_horizon != nullptr) {  // bspline_kernel.cuh:255-258
      padded = min(__ldg(a.interpolation_horizon + b), a.T - 1) + 1;
      dt = __ldg(a.traj_dt);
    } else {
      dt = __ldg(a.traj_dt + g_row);
    }
    const int steps = (padded - 1) / (a.n_knots + DEG + 1);
    const ControlPolygon<DEG> cp = make_polygon<DEG>(a.u, b, d, a.D, a.n_knots, dt, steps, a.implicit[g_row] != 0, a.sp, a.sv,
                                                     a.sa, a.sj, s_row, a.gp, a.gv, a.ga, a.gj, g_row);
    const State4 s = evaluate<DEG>(cp, h, steps);
    a.out_p[tid] = s.p;
    a.out_v[tid] = s.v;
    a.out_a[tid] = s.a;
    a.out_j[tid] = s.j;
    if (h == 0 && d == 0) a.out_dt[b] = dt;
  }
}

struct BwdArgs {
  float *out;
  const float *gp, *gv, *ga, *gj;
  const float *traj_dt;
  const int32_t *dt_idx;
  const uint8_t *implicit;
  int B, T, D, n_knots;
};

template <int DEG>
__global__ void __launch_bounds__(256) bspline_backward_kernel(const __grid_constant__ BwdArgs a) {
  const long long n = (long long)a.B * a.n_knots * a.D;
  const int horizon = a.T - 1;
  const int steps = horizon / (a.n_knots + DEG + 1);
  for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(tid % a.D);
    const int k = (int)((tid / a.D) % a.n_knots);
    const int b = (int)(tid / ((long long)a.D * a.n_knots));
    const int row = __ldg(a.dt_idx + b);
    const float dt = __ldg(a.traj_dt + row);
    const bool implicit = a.implicit[row] != 0;
    const size_t base = (size_t)b * a.T * a.D + d;
    const float *gp = a.gp, *gv = a.gv, *ga = a.ga, *gj = a.gj;
    const int D = a.D;
    auto G = [=](int h, int which) -> float {
      const float *src = which == 0 ? gp : which == 1 ? gv : which == 2 ? ga : gj;
      return __ldg(src + base + (size_t)h * D);
    };
    a.out[tid] = knot_gradient<DEG>(k, steps, a.n_knots, horizon, implicit, dt, G);
  }
}

int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 8;  // 8 x 256 threads / SM resident: a whole number of waves
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_forward(const FwdArgs &a, int degree, cudaStream_t stream) {
  if (a.B <= 0 || a.T <= 0 || a.D <= 0 || a.n_knots <= 0) return status(cudaErrorInvalidValue);
  const long long n = (long long)a.B * a.T * a.D;
  const int grid = grid_for(n, 256);
  switch (degree) {
    case 3: CB200_LAUNCH(bspline_forward_kernel<3>, grid, 256, 0, stream, a); break;
    case 4: CB200_LAUNCH(bspline_forward_kernel<4>, grid, 256, 0, stream, a); break;
    case 5: CB200_LAUNCH(bspline_forward_kernel<5>, grid, 256, 0, stream, a); break;
    default: return status(cudaErrorInvalidValue);
  }
  return status(cudaGetLastError());
}
}  // namespace

extern "C" {

int cb200_bspline_forward(float *out_position, float *out_velocity, float *out_acceleration, float *out_jerk, float *out_dt,
                          const float *u_position, const float *start_position, const float *start_velocity,
                          const float *start_acceleration, const float *start_jerk, const float *goal_position,
                          const float *goal_velocity, const float *goal_acceleration, const float *goal_jerk,
                          const int32_t *start_idx, const int32_t *goal_idx, const float *traj_dt,
                          const uint8_t *use_implicit_goal_state, int batch_size, int padded_horizon, int dof, int n_knots,
                          int bspline_degree, cb200_stream_t stream) {
  CB200_DEVICE_GUARD(out_position);
  FwdArgs a{out_position, out_velocity, out_acceleration, out_jerk, out_dt, u_position, start_position, start_velocity,
            start_acceleration, start_jerk, goal_position, goal_velocity, goal_acceleration, goal_jerk, start_idx, goal_idx,
            traj_dt, use_implicit_goal_state, nullptr, batch_size, padded_horizon, dof, n_knots};
  return launch_forward(a, bspline_degree, (cudaStream_t)stream);
}

int cb200_bspline_single_dt(float *out_position, float *out_velocity, float *out_acceleration, float *out_jerk, float *out_dt,
                            const float *u_position, const float *knot_dt, const float *start_position,
                            const float *start_velocity, const float *start_acceleration, const float *start_jerk,
                            const float *goal_position, const float *goal_velocity, const float *goal_acceleration,
                            const float *goal_jerk, const int32_t *start_idx, const int32_t *goal_idx,
                            const float *interpolation_dt, const uint8_t *use_implicit_goal_state,
                            const int32_t *interpolation_horizon, int batch_size, int max_out_tsteps, int dof, int n_knots,
                            int bspline_degree, cb200_stream_t stream) {
  CB200_DEVICE_GUARD(out_position);
  (void)knot_dt;  // carried by the reference signature, never read by its kernel (bspline_kernel.cuh:216-270)
  if (interpolation_horizon == nullptr) return status(cudaErrorInvalidValue);
  FwdArgs a{out_position, out_velocity, out_acceleration, out_jerk, out_dt, u_position, start_position, start_velocity,
            start_acceleration, start_jerk, goal_position, goal_velocity, goal_acceleration, goal_jerk, start_idx, goal_idx,
            interpolation_dt, use_implicit_goal_state, interpolation_horizon, batch_size, max_out_tsteps, dof, n_knots};
  return launch_forward(a, bspline_degree, (cudaStream_t)stream);
}

int cb200_bspline_backward(float *out_grad_knots, const float *grad_position, const float *grad_velocity,
                           const float *grad_acceleration, const float *grad_jerk, const float *traj_dt,
                           const int32_t *dt_idx, const uint8_t *use_implicit_goal_state, int batch_size, int padded_horizon,
                           int dof, int n_knots, int bspline_degree, cb200_stream_t stream) {
  CB200_DEVICE_GUARD(out_grad_knots);
  const int horizon = padded_horizon - 1;
  // same argument checks as the reference launcher (trajectory_kernel_launch.cu:592-627)
  if (batch_size <= 0 || dof <= 0 || n_knots <= 0 || horizon < 5) return status(cudaErrorInvalidValue);
  if (bspline_degree < 3 || bspline_degree > 5) return status(cudaErrorInvalidValue);
  const int steps = horizon / (n_knots + bspline_degree + 1);
  if (steps <= 0 || steps > 32) return status(cudaErrorInvalidValue);
  BwdArgs a{out_grad_knots, grad_position, grad_velocity, grad_acceleration, grad_jerk, traj_dt, dt_idx, use_implicit_goal_state,
            batch_size, padded_horizon, dof, n_knots};
  const long long n = (long long)batch_size * n_knots * dof;
  const int grid = grid_for(n, 128);
  switch (bspline_degree) {
    case 3: CB200_LAUNCH(bspline_backward_kernel<3>, grid, 128, 0, (cudaStream_t)stream, a); break;
    case 4: CB200_LAUNCH(bspline_backward_kernel<4>, grid, 128, 0, (cudaStream_t)stream, a); break;
    default: CB200_LAUNCH(bspline_backward_kernel<5>, grid, 128, 0, (cudaStream_t)stream, a); break;
  }
  return status(cudaGetLastError());
}

}  // extern "C"
