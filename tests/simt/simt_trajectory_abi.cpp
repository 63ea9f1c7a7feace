// TEST INFRASTRUCTURE: cb200_trajectory.cu (B-spline kernels + their C ABI) as a second translation unit of the emulated main
// library: cb200_rollout_cost_grad calls cb200_bspline_forward / _backward for the expanded knots schedule.
#define CB200_SIMT_EMULATION 1
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_trajectory.cu"
