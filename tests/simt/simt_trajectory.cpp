// TEST INFRASTRUCTURE: the B-spline kernels of curobo_b200/csrc/cb200_trajectory.cu compiled as ordinary C++ and executed by
// std::threads (tests/simt/cuda_runtime.h).
#define CB200_SIMT_EMULATION 1
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_trajectory.cu"

extern "C" {
int em_bspline_forward(int grid, float *op, float *ov, float *oa, float *oj, float *odt, const float *u, const float *sp,
                       const float *sv, const float *sa, const float *sj, const float *gp, const float *gv, const float *ga,
                       const float *gj, const int32_t *start_idx, const int32_t *goal_idx, const float *traj_dt,
                       const uint8_t *implicit, const int32_t *interpolation_horizon, int B, int T, int D, int n_knots, int degree) {
  FwdArgs a{op, ov, oa, oj, odt, u, sp, sv, sa, sj, gp, gv, ga, gj, start_idx, goal_idx, traj_dt, implicit, interpolation_horizon,
            B, T, D, n_knots};
  if (degree == 3) simt::launch(bspline_forward_kernel<3>, grid, 64, a);
  else if (degree == 4) simt::launch(bspline_forward_kernel<4>, grid, 64, a);
  else if (degree == 5) simt::launch(bspline_forward_kernel<5>, grid, 64, a);
  else return 1;
  return 0;
}

int em_bspline_backward(int grid, float *out, const float *gp, const float *gv, const float *ga, const float *gj,
                        const float *traj_dt, const int32_t *dt_idx, const uint8_t *implicit, int B, int T, int D, int n_knots,
                        int degree) {
  BwdArgs a{out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, T, D, n_knots};
  if (degree == 3) simt::launch(bspline_backward_kernel<3>, grid, 64, a);
  else if (degree == 4) simt::launch(bspline_backward_kernel<4>, grid, 64, a);
  else if (degree == 5) simt::launch(bspline_backward_kernel<5>, grid, 64, a);
  else return 1;
  return 0;
}
}
