// oracle/ref_kernels_launcher.cu -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Thin extern "C" launchers around the REFERENCE's own CUDA kernels, which are #included from the
// reference tree where they lie (-I /root/reference/curobo/_src/curobolib/kernels; nothing is copied
// into this repository).  Built by curobo_b200/build.py::build_reference_kernels into
// oracle/_ref/libcurobo_ref.so (git-ignored; travels to the GPU box with the snapshot).
//
// Used ONLY by tests/ (to pin the oracle and to compare our kernels against the reference's on the
// same inputs) and by bench.py's "reference kernels, unfused" side measurement.
//
// Launch math follows the reference launchers:
//   forward : backends/cuda_core_backend/kinematics.py:90-177 + kinematics_config.py:53-95
//             (threads_per_batch 32, max_threads 256, smem = bpb*nl*96 B)
//   backward: kinematics.py:282-379 + kinematics_config.py:97-175 (warp-reduce path)
//   self    : backends/cuda_core_backend/geometry.py:63-227
#include <cstdint>
#include <cuda_runtime.h>

#include "kinematics/kinematics_forward_kernel.cuh"
#include "kinematics/kinematics_backward_kernel.cuh"
#include "geometry/self_collision/self_collision_kernel.cuh"

namespace ck = curobo::kinematics;
namespace cs = curobo::geometry::self_collision;

extern "C" {

int ref_kinematics_forward_spheres(float *link_pos, float *link_quat, float *batch_robot_spheres, float *batch_com,
                                   float *global_cumul_mat, const float *q, const float *fixed_transform,
                                   const float *robot_spheres, const float *link_masses_com, const int8_t *joint_map_type,
                                   const int16_t *joint_map, const int16_t *link_map, const int16_t *tool_frame_map,
                                   const int16_t *link_sphere_map, const float *joint_offset_map,
                                   const int32_t *env_query_idx, int num_envs, int batch_size, int horizon, int n_joints,
                                   int num_spheres, int num_links, int n_tool_frames, cudaStream_t stream) {
  const int tpb = 32, max_threads = 256;
  int bpb = 8;
  const int smem_per = num_links * 12 * 4 * 2;
  if (bpb > 48 * 1024 / smem_per) bpb = 48 * 1024 / smem_per;
  if (bpb * tpb > max_threads) bpb = max_threads / tpb;
  if (bpb < 1) bpb = 1;
  if (bpb > batch_size) bpb = batch_size;
  const int threads = bpb * tpb;
  const int blocks = (batch_size + bpb - 1) / bpb;
  const size_t smem = (size_t)bpb * smem_per;
  ck::kinematics_forward_spheres_kernel<-1, 32, true, false><<<blocks, threads, smem, stream>>>(
      link_pos, link_quat, batch_robot_spheres, batch_com, global_cumul_mat, q, fixed_transform, robot_spheres,
      link_masses_com, joint_map_type, joint_map, link_map, tool_frame_map, link_sphere_map, joint_offset_map,
      env_query_idx, batch_size, horizon, num_spheres, num_envs, num_links, n_joints, n_tool_frames);
  return (int)cudaGetLastError();
}

}  // extern "C"

template <int MAXJ>
static void launch_bwd(int blocks, int threads, size_t smem, cudaStream_t stream, float *grad_out, const float *g_pos,
                       const float *g_quat, const float *g_sph, const float *g_com, const float *b_com,
                       const float *g_jac, const float *cumul, const float *robot_spheres, const float *masses,
                       const int8_t *jtype, const int16_t *jmap, const int16_t *lmap, const int16_t *tool,
                       const int16_t *sphl, const int32_t *envq, const int16_t *cd, const int16_t *co, const int16_t *jd,
                       const int16_t *jo, const bool *jae, const float *joff, int batch, int horizon, int ns, int nl,
                       int nj, int nt, int nenv, int tpb) {
  ck::kinematics_backward_kernel<float, float, MAXJ, true, false, false><<<blocks, threads, smem, stream>>>(
      grad_out, g_pos, g_quat, g_sph, g_com, b_com, g_jac, cumul, robot_spheres, masses, jtype, jmap, lmap, tool, sphl,
      envq, cd, co, jd, jo, jae, joff, batch, horizon, ns, nl, nj, nt, nenv, tpb);
}

extern "C" {

int ref_kinematics_backward(float *grad_out, const float *g_pos, const float *g_quat, const float *g_sph,
                            const float *g_com, const float *b_com, const float *g_jac, const float *cumul,
                            const float *robot_spheres, const float *masses, const int16_t *lmap, const int16_t *jmap,
                            const int8_t *jtype, const int16_t *tool, const int16_t *sphl, const int16_t *cd,
                            const int16_t *co, const int16_t *jd, const int16_t *jo, const uint8_t *jae,
                            const float *joff, const int32_t *envq, int num_envs, int batch_size, int horizon,
                            int n_joints, int num_spheres, int num_links, int n_tool_frames, cudaStream_t stream) {
  const int max_threads = 128, tpb = 32;
  int bpb = 32;
  const int smem_per = num_links * 12 * 4;
  if (bpb > 48 * 1024 / smem_per) bpb = 48 * 1024 / smem_per;
  if (bpb * tpb > max_threads) bpb = max_threads / tpb;
  if (bpb < 1) bpb = 1;
  if (bpb > batch_size) bpb = batch_size;
  const int threads = bpb * tpb;
  const int blocks = (batch_size * tpb + threads - 1) / threads;
  const size_t smem = (size_t)bpb * smem_per;
  const bool *jb = reinterpret_cast<const bool *>(jae);
  if (n_joints < 16)
    launch_bwd<16>(blocks, threads, smem, stream, grad_out, g_pos, g_quat, g_sph, g_com, b_com, g_jac, cumul,
                   robot_spheres, masses, jtype, jmap, lmap, tool, sphl, envq, cd, co, jd, jo, jb, joff, batch_size,
                   horizon, num_spheres, num_links, n_joints, n_tool_frames, num_envs, tpb);
  else if (n_joints < 64)
    launch_bwd<64>(blocks, threads, smem, stream, grad_out, g_pos, g_quat, g_sph, g_com, b_com, g_jac, cumul,
                   robot_spheres, masses, jtype, jmap, lmap, tool, sphl, envq, cd, co, jd, jo, jb, joff, batch_size,
                   horizon, num_spheres, num_links, n_joints, n_tool_frames, num_envs, tpb);
  else
    launch_bwd<128>(blocks, threads, smem, stream, grad_out, g_pos, g_quat, g_sph, g_com, b_com, g_jac, cumul,
                    robot_spheres, masses, jtype, jmap, lmap, tool, sphl, envq, cd, co, jd, jo, jb, joff, batch_size,
                    horizon, num_spheres, num_links, n_joints, n_tool_frames, num_envs, tpb);
  return (int)cudaGetLastError();
}

int ref_self_collision_distance(float *out_distance, float *out_vec, float *pair_distance, uint8_t *sparse_index,
                                const float *robot_spheres, const float *sphere_padding, const float *weight,
                                int16_t *pair_locations, float *block_batch_max_value, int16_t *block_batch_max_index,
                                int num_blocks_per_batch, int max_threads_per_block, int batch_size, int horizon,
                                int nspheres, int num_collision_pairs, int compute_grad, cudaStream_t stream) {
  const size_t smem = (size_t)16 * nspheres;
  if (num_blocks_per_batch == 1) {
    int threads = max_threads_per_block < num_collision_pairs ? max_threads_per_block : num_collision_pairs;
    threads = ((threads + 31) / 32) * 32;
    if (threads > max_threads_per_block) threads = max_threads_per_block;
    cs::self_collision_max_distance_kernel<false><<<batch_size * horizon, threads, smem, stream>>>(
        out_distance, out_vec, pair_distance, sparse_index, robot_spheres, sphere_padding, weight, pair_locations,
        batch_size, horizon, nspheres, num_collision_pairs, compute_grad != 0);
  } else {
    cs::self_collision_max_block_kernel<false>
        <<<batch_size * horizon * num_blocks_per_batch, max_threads_per_block, smem, stream>>>(
            out_vec, pair_distance, sparse_index, robot_spheres, sphere_padding, pair_locations, block_batch_max_value,
            block_batch_max_index, num_blocks_per_batch, batch_size, horizon, nspheres, num_collision_pairs);
    const int t2 = num_blocks_per_batch < 512 ? num_blocks_per_batch : 512;
    cs::self_collision_max_reduce_kernel<<<batch_size * horizon, t2, 0, stream>>>(
        out_distance, out_vec, pair_distance, sparse_index, robot_spheres, sphere_padding, weight, pair_locations,
        block_batch_max_value, block_batch_max_index, num_blocks_per_batch, batch_size, horizon, nspheres,
        num_collision_pairs, compute_grad != 0);
  }
  return (int)cudaGetLastError();
}
}

// ------------------------------------------------------------------------------------------------
// B-spline knot -> state kernels.  Launch math follows
// backends/pybind/trajectory_kernel_launch.cu:263-404 (forward, 128 threads), :406-566 (single dt, 256 threads),
// :571-683 (backward, 128 threads, batch*dof*threads_for_n_knots); MATRIX basis backend like :259.
// ------------------------------------------------------------------------------------------------
#include "trajectory/bspline/bspline_common.cuh"
#include "trajectory/bspline/bspline_kernel.cuh"

namespace cbs = curobo::trajectory::bspline;

template <int DEG>
static void ref_bspline_fwd_t(int blocks, int threads, cudaStream_t stream, float *op, float *ov, float *oa, float *oj,
                              float *odt, const float *u, const float *sp, const float *sv, const float *sa,
                              const float *sj, const float *gp, const float *gv, const float *ga, const float *gj,
                              const int32_t *sidx, const int32_t *gidx, const float *traj_dt, const uint8_t *implicit,
                              int B, int T, int D, int nk) {
  cbs::interpolate_bspline_kernel<float, DEG, cbs::BasisBackend::MATRIX><<<blocks, threads, 0, stream>>>(
      op, ov, oa, oj, odt, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, B, T, D, nk);
}

template <int DEG>
static void ref_bspline_sdt_t(int blocks, int threads, cudaStream_t stream, float *op, float *ov, float *oa, float *oj,
                              float *odt, const float *u, const float *knot_dt, const float *sp, const float *sv,
                              const float *sa, const float *sj, const float *gp, const float *gv, const float *ga,
                              const float *gj, const int32_t *sidx, const int32_t *gidx, const float *interp_dt,
                              const uint8_t *implicit, const int32_t *interp_h, int B, int T, int D, int nk) {
  cbs::interpolate_bspline_single_dt_kernel<float, DEG, cbs::BasisBackend::MATRIX><<<blocks, threads, 0, stream>>>(
      op, ov, oa, oj, odt, u, knot_dt, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, interp_dt, implicit, interp_h, B, T,
      D, nk);
}

template <int DEG>
static int ref_bspline_bwd_t(cudaStream_t stream, float *out, const float *gp, const float *gv, const float *ga,
                             const float *gj, const float *traj_dt, const int32_t *dt_idx, const uint8_t *implicit,
                             int B, int horizon, int D, int nk) {
  cbs::BSplineBackwardLayout layout = cbs::compute_bspline_backward_layout<DEG>(horizon, D, nk);
  if (layout.interpolation_steps <= 0 || layout.interpolation_steps > 32) return 1;
  const int k_size = B * D * layout.threads_for_n_knots;
  const int threads = k_size > 128 ? 128 : k_size;
  const int blocks = (k_size + threads - 1) / threads;
  cbs::bspline_backward_kernel<DEG, float, cbs::BasisBackend::MATRIX><<<blocks, threads, 0, stream>>>(
      out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, horizon, D, nk);
  return 0;
}

extern "C" {

int ref_bspline_forward(float *op, float *ov, float *oa, float *oj, float *odt, const float *u, const float *sp,
                        const float *sv, const float *sa, const float *sj, const float *gp, const float *gv,
                        const float *ga, const float *gj, const int32_t *sidx, const int32_t *gidx, const float *traj_dt,
                        const uint8_t *implicit, int B, int padded_horizon, int D, int nk, int degree,
                        cudaStream_t stream) {
  const int k_size = B * padded_horizon * D;
  const int threads = k_size > 128 ? 128 : k_size;
  const int blocks = (k_size + threads - 1) / threads;
  if (degree == 3) ref_bspline_fwd_t<3>(blocks, threads, stream, op, ov, oa, oj, odt, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, B, padded_horizon, D, nk);
  else if (degree == 4) ref_bspline_fwd_t<4>(blocks, threads, stream, op, ov, oa, oj, odt, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, B, padded_horizon, D, nk);
  else if (degree == 5) ref_bspline_fwd_t<5>(blocks, threads, stream, op, ov, oa, oj, odt, u, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, traj_dt, implicit, B, padded_horizon, D, nk);
  else return 1;
  return (int)cudaGetLastError();
}

int ref_bspline_single_dt(float *op, float *ov, float *oa, float *oj, float *odt, const float *u, const float *knot_dt,
                          const float *sp, const float *sv, const float *sa, const float *sj, const float *gp,
                          const float *gv, const float *ga, const float *gj, const int32_t *sidx, const int32_t *gidx,
                          const float *interp_dt, const uint8_t *implicit, const int32_t *interp_h, int B,
                          int max_out_tsteps, int D, int nk, int degree, cudaStream_t stream) {
  const int k_size = B * max_out_tsteps * D;
  const int threads = k_size > 256 ? 256 : k_size;
  const int blocks = (k_size + threads - 1) / threads;
  if (degree == 3) ref_bspline_sdt_t<3>(blocks, threads, stream, op, ov, oa, oj, odt, u, knot_dt, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, interp_dt, implicit, interp_h, B, max_out_tsteps, D, nk);
  else if (degree == 4) ref_bspline_sdt_t<4>(blocks, threads, stream, op, ov, oa, oj, odt, u, knot_dt, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, interp_dt, implicit, interp_h, B, max_out_tsteps, D, nk);
  else if (degree == 5) ref_bspline_sdt_t<5>(blocks, threads, stream, op, ov, oa, oj, odt, u, knot_dt, sp, sv, sa, sj, gp, gv, ga, gj, sidx, gidx, interp_dt, implicit, interp_h, B, max_out_tsteps, D, nk);
  else return 1;
  return (int)cudaGetLastError();
}

int ref_bspline_backward(float *out, const float *gp, const float *gv, const float *ga, const float *gj,
                         const float *traj_dt, const int32_t *dt_idx, const uint8_t *implicit, int B,
                         int padded_horizon, int D, int nk, int degree, cudaStream_t stream) {
  const int horizon = padded_horizon - 1;  // trajectory_kernel_launch.cu:592
  int rc = 1;
  if (degree == 3) rc = ref_bspline_bwd_t<3>(stream, out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, horizon, D, nk);
  else if (degree == 4) rc = ref_bspline_bwd_t<4>(stream, out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, horizon, D, nk);
  else if (degree == 5) rc = ref_bspline_bwd_t<5>(stream, out, gp, gv, ga, gj, traj_dt, dt_idx, implicit, B, horizon, D, nk);
  if (rc) return rc;
  return (int)cudaGetLastError();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Optimizer kernels.  Launch math follows
// backends/cuda_core_backend/optimization_config.py:54-70 (line search: one CTA of opt_dim threads per problem) and
// :76-127 (L-BFGS: one CTA of v_dim threads per problem; shared-memory variant when it fits 64 KB).
// ------------------------------------------------------------------------------------------------
#include "optimization/lbfgs/lbfgs_step_kernel.cuh"
#include "optimization/line_search/line_search_kernel.cuh"

namespace cop = curobo::optimization;

template <int M>
static int ref_lbfgs_t(float *step_vec, float *rho, float *y, float *s, float *q, float *x_0, float *grad_0,
                       const float *grad_q, float epsilon, int B, int V, int stable, int use_shared, cudaStream_t stream) {
  const size_t basic = (size_t)M * 4, shared = (size_t)(((2 * V) + 2) * M + 32 + 1) * 4;
  if (use_shared && shared <= 65536) {
    if (shared > 48000) cudaFuncSetAttribute(cop::kernel_lbfgs_step_shared_memory<float, false, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared);
    cop::kernel_lbfgs_step_shared_memory<float, false, M><<<B, V, shared, stream>>>(step_vec, rho, y, s, q, x_0, grad_0, grad_q,
                                                                                  epsilon, B, M, V, stable != 0);
  } else {
    cop::kernel_lbfgs_step<float, false, M><<<B, V, basic, stream>>>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B,
                                                                     M, V, stable != 0);
  }
  return (int)cudaGetLastError();
}

extern "C" {

int ref_lbfgs_step(float *step_vec, float *rho, float *y, float *s, float *q, float *x_0, float *grad_0,
                   const float *grad_q, float epsilon, int B, int m, int V, int stable, int use_shared,
                   cudaStream_t stream) {
  switch (m) {  // the reference JIT-compiles FIXED_M = history (optimization.py:178-183); the histories the tests use
    case 3: return ref_lbfgs_t<3>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B, V, stable, use_shared, stream);
    case 5: return ref_lbfgs_t<5>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B, V, stable, use_shared, stream);
    case 7: return ref_lbfgs_t<7>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B, V, stable, use_shared, stream);
    case 15: return ref_lbfgs_t<15>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B, V, stable, use_shared, stream);
    case 27: return ref_lbfgs_t<27>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B, V, stable, use_shared, stream);
    case 31: return ref_lbfgs_t<31>(step_vec, rho, y, s, q, x_0, grad_0, grad_q, epsilon, B, V, stable, use_shared, stream);
    default: return -1;
  }
}

int ref_line_search(float *best_cost, float *best_action, int16_t *best_iteration, int16_t *current_iteration,
                    uint8_t *converged, int convergence_iteration, float cost_delta_threshold,
                    float cost_relative_threshold, float *exploration_cost, float *exploration_action,
                    float *exploration_gradient, int32_t *exploration_idx, float *selected_cost, float *selected_action,
                    float *selected_gradient, int32_t *selected_idx, const float *search_cost, const float *search_action,
                    const float *search_gradient, const float *step_direction, const float *search_magnitudes, float c_1,
                    float c_2, int strong_wolfe, int approx_wolfe, int n, int V, int B, cudaStream_t stream) {
  if (n == 4)
    cop::line_search::kernel_line_search<float, 4><<<B, V, 0, stream>>>(
        best_cost, best_action, best_iteration, current_iteration, converged, convergence_iteration, cost_delta_threshold,
        cost_relative_threshold, exploration_cost, exploration_action, exploration_gradient, exploration_idx, selected_cost,
        selected_action, selected_gradient, selected_idx, search_cost, search_action, search_gradient, step_direction,
        search_magnitudes, c_1, c_2, strong_wolfe != 0, approx_wolfe != 0, n, V, B);
  else
    cop::line_search::kernel_line_search<float, -1><<<B, V, 0, stream>>>(
        best_cost, best_action, best_iteration, current_iteration, converged, convergence_iteration, cost_delta_threshold,
        cost_relative_threshold, exploration_cost, exploration_action, exploration_gradient, exploration_idx, selected_cost,
        selected_action, selected_gradient, selected_idx, search_cost, search_action, search_gradient, step_direction,
        search_magnitudes, c_1, c_2, strong_wolfe != 0, approx_wolfe != 0, n, V, B);
  return (int)cudaGetLastError();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// RNEA inverse dynamics kernels, serial path (threads_per_batch = 1) like
// backends/cuda_core_backend/dynamics_config.py:118-215: shared memory per batch element = N_LINKS*12 floats (forward),
// N_LINKS*30 floats + one N_LINKS*12 block of inertial parameters (backward); <= 48 KB per block.
// The kernels are templated on (N_LINKS, N_DOF); the three robots of the test-suite are instantiated.
// ------------------------------------------------------------------------------------------------
#include "dynamics/rnea_forward_kernel.cuh"
#include "dynamics/rnea_backward_kernel.cuh"

namespace cdy = curobo::dynamics;

template <int NL, int ND>
static int ref_rnea_fwd_t(float *tau, const float *q, const float *qd, const float *qdd, const float *fixed, const float *mc,
                          const float *inertia, const int8_t *jtype, const int16_t *jmap, const int16_t *lmap,
                          const float *joff, const float *gravity, const int16_t *lstarts, const int16_t *llinks, float *cache,
                          int B, int n_levels, cudaStream_t stream) {
  const int per = NL * 12 * 4;
  int bpb = 48 * 1024 / per;
  if (bpb > 64) bpb = 64;
  if (bpb >= 32) bpb = bpb / 32 * 32;
  if (bpb < 1) return 1;
  const int blocks = (B + bpb - 1) / bpb;
  cdy::rnea_forward_kernel<NL, ND, 1, false><<<blocks, bpb, (size_t)bpb * per, stream>>>(
      tau, q, qd, qdd, fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, cache, nullptr, B, n_levels);
  return (int)cudaGetLastError();
}

template <int NL, int ND>
static int ref_rnea_bwd_t(float *gq, float *gqd, float *gqdd, const float *gtau, const float *q, const float *qd,
                          const float *fixed, const float *mc, const float *inertia, const int8_t *jtype, const int16_t *jmap,
                          const int16_t *lmap, const float *joff, const float *gravity, const int16_t *lstarts,
                          const int16_t *llinks, const float *cache, int B, int n_levels, cudaStream_t stream) {
  const int shared_block = NL * 12 * 4, per = NL * 30 * 4;
  int bpb = (48 * 1024 - shared_block) / per;
  if (bpb > 64) bpb = 64;
  if (bpb < 1) return 1;
  // The reference kernel returns its out-of-range threads BEFORE the block-wide load of the inertial parameters into shared
  // memory (rnea_backward_kernel.cuh:94 vs :142-161), so a partially filled last block reads unloaded parameters for some
  // links. The checker therefore only launches full blocks: the largest block size that divides the batch.
  while (B % bpb) bpb--;
  const int blocks = (B + bpb - 1) / bpb;
  cdy::rnea_backward_kernel<NL, ND, 1, false><<<blocks, bpb, (size_t)shared_block + (size_t)bpb * per, stream>>>(
      gq, gqd, gqdd, nullptr, gtau, q, qd, fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, cache, B,
      n_levels);
  return (int)cudaGetLastError();
}

extern "C" {

int ref_rnea_forward(float *tau, const float *q, const float *qd, const float *qdd, const float *fixed, const float *mc,
                     const float *inertia, const int8_t *jtype, const int16_t *jmap, const int16_t *lmap, const float *joff,
                     const float *gravity, const int16_t *lstarts, const int16_t *llinks, float *cache, int B, int nl, int D,
                     int n_levels, cudaStream_t stream) {
#define CB_FWD(NL, ND) \
  if (nl == NL && D == ND) return ref_rnea_fwd_t<NL, ND>(tau, q, qd, qdd, fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, cache, B, n_levels, stream);
  CB_FWD(13, 7)
  CB_FWD(42, 35)
  CB_FWD(56, 49)
#undef CB_FWD
  return -1;
}

int ref_rnea_backward(float *gq, float *gqd, float *gqdd, const float *gtau, const float *q, const float *qd,
                      const float *fixed, const float *mc, const float *inertia, const int8_t *jtype, const int16_t *jmap,
                      const int16_t *lmap, const float *joff, const float *gravity, const int16_t *lstarts,
                      const int16_t *llinks, const float *cache, int B, int nl, int D, int n_levels, cudaStream_t stream) {
#define CB_BWD(NL, ND) \
  if (nl == NL && D == ND) return ref_rnea_bwd_t<NL, ND>(gq, gqd, gqdd, gtau, q, qd, fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, cache, B, n_levels, stream);
  CB_BWD(13, 7)
  CB_BWD(42, 35)
  CB_BWD(56, 49)
#undef CB_BWD
  return -1;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// PBA+ 3-D EDT kernels: the five launches + final copy of
// backends/cuda_core_backend/pba.py:86-124 with the launch shapes of pba_config.py:51-83.
// ------------------------------------------------------------------------------------------------
#include "parallel_banding/pba3d_kernel.cuh"

namespace cpb = curobo::parallel_banding;

extern "C" int ref_pba3d(int *site_index, int *buffer, int nx, int ny, int nz, int m3, cudaStream_t stream) {
  const int sx = nz, sy = ny, sz = nx;
  auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
  int *buf0 = site_index, *buf1 = buffer;
  cpb::kernel_flood_z<<<dim3(cdiv(sx, 32), cdiv(sy, 4)), dim3(32, 4), 0, stream>>>(buf0, buf1, sx, sy, sz);
  cpb::kernel_maurer_axis<<<dim3(cdiv(sx, 32), cdiv(sz, 4)), dim3(32, 4), 0, stream>>>(buf1, buf0, sx, sy, sz);
  cpb::kernel_color_axis<<<dim3(cdiv(sx, 32), sz), dim3(32, m3), 0, stream>>>(buf0, buf1, sx, sy, sz);
  cpb::kernel_maurer_axis<<<dim3(cdiv(sy, 32), cdiv(sz, 4)), dim3(32, 4), 0, stream>>>(buf1, buf0, sy, sx, sz);
  cpb::kernel_color_axis<<<dim3(cdiv(sy, 32), sz), dim3(32, m3), 0, stream>>>(buf0, buf1, sy, sx, sz);
  cudaMemcpyAsync(site_index, buffer, (size_t)nx * ny * nz * sizeof(int), cudaMemcpyDeviceToDevice, stream);
  return (int)cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// FK with the centre-of-mass output and its gradient (COMPUTE_COM = true instantiations of the same reference kernels),
// launch shapes as in ref_kinematics_forward_spheres / ref_kinematics_backward above.
// ------------------------------------------------------------------------------------------------
template <int MAXJ>
static void launch_bwd_com(int blocks, int threads, size_t smem, cudaStream_t stream, float *grad_out, const float *g_pos,
                           const float *g_quat, const float *g_sph, const float *g_com, const float *b_com, const float *cumul,
                           const float *robot_spheres, const float *masses, const int8_t *jtype, const int16_t *jmap,
                           const int16_t *lmap, const int16_t *tool, const int16_t *sphl, const int32_t *envq, const int16_t *cd,
                           const int16_t *co, const int16_t *jd, const int16_t *jo, const bool *jae, const float *joff, int batch,
                           int horizon, int ns, int nl, int nj, int nt, int nenv, int tpb) {
  ck::kinematics_backward_kernel<float, float, MAXJ, true, true, false><<<blocks, threads, smem, stream>>>(
      grad_out, g_pos, g_quat, g_sph, g_com, b_com, nullptr, cumul, robot_spheres, masses, jtype, jmap, lmap, tool, sphl, envq, cd,
      co, jd, jo, jae, joff, batch, horizon, ns, nl, nj, nt, nenv, tpb);
}

extern "C" {

int ref_kinematics_forward_spheres_com(float *link_pos, float *link_quat, float *batch_robot_spheres, float *batch_com,
                                       float *global_cumul_mat, const float *q, const float *fixed_transform,
                                       const float *robot_spheres, const float *link_masses_com, const int8_t *joint_map_type,
                                       const int16_t *joint_map, const int16_t *link_map, const int16_t *tool_frame_map,
                                       const int16_t *link_sphere_map, const float *joint_offset_map,
                                       const int32_t *env_query_idx, int num_envs, int batch_size, int horizon, int n_joints,
                                       int num_spheres, int num_links, int n_tool_frames, cudaStream_t stream) {
  const int tpb = 32, max_threads = 256;
  int bpb = 8;
  const int smem_per = num_links * 12 * 4 * 2;
  if (bpb > 48 * 1024 / smem_per) bpb = 48 * 1024 / smem_per;
  if (bpb * tpb > max_threads) bpb = max_threads / tpb;
  if (bpb < 1) bpb = 1;
  if (bpb > batch_size) bpb = batch_size;
  const int threads = bpb * tpb;
  const int blocks = (batch_size + bpb - 1) / bpb;
  ck::kinematics_forward_spheres_kernel<-1, 32, true, true><<<blocks, threads, (size_t)bpb * smem_per, stream>>>(
      link_pos, link_quat, batch_robot_spheres, batch_com, global_cumul_mat, q, fixed_transform, robot_spheres,
      link_masses_com, joint_map_type, joint_map, link_map, tool_frame_map, link_sphere_map, joint_offset_map,
      env_query_idx, batch_size, horizon, num_spheres, num_envs, num_links, n_joints, n_tool_frames);
  return (int)cudaGetLastError();
}

int ref_kinematics_backward_com(float *grad_out, const float *g_pos, const float *g_quat, const float *g_sph, const float *g_com,
                                const float *b_com, const float *cumul, const float *robot_spheres, const float *masses,
                                const int16_t *lmap, const int16_t *jmap, const int8_t *jtype, const int16_t *tool,
                                const int16_t *sphl, const int16_t *cd, const int16_t *co, const int16_t *jd, const int16_t *jo,
                                const uint8_t *jae, const float *joff, const int32_t *envq, int num_envs, int batch_size,
                                int horizon, int n_joints, int num_spheres, int num_links, int n_tool_frames,
                                cudaStream_t stream) {
  const int max_threads = 128, tpb = 32;
  int bpb = 32;
  const int smem_per = num_links * 12 * 4;
  if (bpb > 48 * 1024 / smem_per) bpb = 48 * 1024 / smem_per;
  if (bpb * tpb > max_threads) bpb = max_threads / tpb;
  if (bpb < 1) bpb = 1;
  if (bpb > batch_size) bpb = batch_size;
  const int threads = bpb * tpb;
  const int blocks = (batch_size * tpb + threads - 1) / threads;
  const size_t smem = (size_t)bpb * smem_per;
  const bool *jb = reinterpret_cast<const bool *>(jae);
  if (n_joints < 16)
    launch_bwd_com<16>(blocks, threads, smem, stream, grad_out, g_pos, g_quat, g_sph, g_com, b_com, cumul, robot_spheres, masses,
                       jtype, jmap, lmap, tool, sphl, envq, cd, co, jd, jo, jb, joff, batch_size, horizon, num_spheres, num_links,
                       n_joints, n_tool_frames, num_envs, tpb);
  else if (n_joints < 64)
    launch_bwd_com<64>(blocks, threads, smem, stream, grad_out, g_pos, g_quat, g_sph, g_com, b_com, cumul, robot_spheres, masses,
                       jtype, jmap, lmap, tool, sphl, envq, cd, co, jd, jo, jb, joff, batch_size, horizon, num_spheres, num_links,
                       n_joints, n_tool_frames, num_envs, tpb);
  else
    launch_bwd_com<128>(blocks, threads, smem, stream, grad_out, g_pos, g_quat, g_sph, g_com, b_com, cumul, robot_spheres, masses,
                        jtype, jmap, lmap, tool, sphl, envq, cd, co, jd, jo, jb, joff, batch_size, horizon, num_spheres, num_links,
                        n_joints, n_tool_frames, num_envs, tpb);
  return (int)cudaGetLastError();
}

}  // extern "C"
