"""Tool-pose and c-space cost operators (drop-in at the kernel-launch level of the reference's
Warp autograd Functions: cost/wp_tool_pose.py:698-855, cost/wp_cspace_state.py:288-420,
cost/wp_cspace_position.py:20-225).  Outputs are caller-allocated buffers, written in place."""
from __future__ import annotations

import torch

from . import lib as _lib
from .backends.tensor_checks import check_tensors, stream_ptr


def tool_pose_distance(current_position, current_quat, goal_position, goal_quat, idxs_goal,
                       position_orientation_weight, terminal_pose_axes_weight_factor,
                       non_terminal_pose_axes_weight_factor, terminal_pose_convergence_tolerance,
                       non_terminal_pose_convergence_tolerance, project_distance_to_goal, out_distance,
                       out_position_distance, out_rotation_distance, out_position_gradient, out_rotation_gradient,
                       out_goalset_idx, use_lie_group: bool = False):
    """Argument order of ToolPoseDistance.forward (cost/wp_tool_pose.py:700-721).

    Shapes: current_* [B,H,L,3/4]; goal_* [G,L,n_goalset,3/4]; idxs_goal [B,1] int32."""
    if current_position.ndim != 4 or current_quat.ndim != 4:
        raise ValueError("current_position / current_quat must be 4D tensors")
    if goal_position.ndim != 4 or goal_quat.ndim != 4:
        raise ValueError("goal_position / goal_quat must be 4D tensors (-1, num_links, num_goalset, 3|4)")
    b, h, nl, _ = current_position.shape
    if idxs_goal.shape != (b, 1):
        raise ValueError(f"idxs_goal must have shape ({b}, 1) but got {tuple(idxs_goal.shape)}")
    if out_distance.shape != (b, h, nl * 2):
        raise ValueError("out_distance must have shape (b, h, num_links*2)")
    # "Only 0 is supported for now" (wp_tool_pose.py:743-744).  The check reads the tensor back, which is illegal during
    # CUDA-graph capture, so it is done on eager calls only (a captured call was validated by its warm-up run).
    if (project_distance_to_goal is not None and not torch.cuda.is_current_stream_capturing()
            and bool(project_distance_to_goal.any())):
        raise ValueError("b200 tool-pose cost: project_distance_to_goal is not supported")
    dev = current_position.device
    check_tensors(dev, torch.float32, current_position=current_position, current_quat=current_quat,
                  goal_position=goal_position, goal_quat=goal_quat,
                  position_orientation_weight=position_orientation_weight,
                  terminal_pose_axes_weight_factor=terminal_pose_axes_weight_factor,
                  non_terminal_pose_axes_weight_factor=non_terminal_pose_axes_weight_factor,
                  terminal_pose_convergence_tolerance=terminal_pose_convergence_tolerance,
                  non_terminal_pose_convergence_tolerance=non_terminal_pose_convergence_tolerance,
                  out_distance=out_distance, out_position_distance=out_position_distance,
                  out_rotation_distance=out_rotation_distance, out_position_gradient=out_position_gradient,
                  out_rotation_gradient=out_rotation_gradient)
    check_tensors(dev, torch.int32, idxs_goal=idxs_goal, out_goalset_idx=out_goalset_idx)
    err = _lib.load().cb200_tool_pose_distance(
        out_distance.data_ptr(), out_position_distance.data_ptr(), out_rotation_distance.data_ptr(),
        out_position_gradient.data_ptr(), out_rotation_gradient.data_ptr(), out_goalset_idx.data_ptr(),
        current_position.data_ptr(), current_quat.data_ptr(), goal_position.data_ptr(), goal_quat.data_ptr(),
        idxs_goal.data_ptr(), position_orientation_weight.data_ptr(), terminal_pose_axes_weight_factor.data_ptr(),
        non_terminal_pose_axes_weight_factor.data_ptr(), terminal_pose_convergence_tolerance.data_ptr(),
        non_terminal_pose_convergence_tolerance.data_ptr(), b, h, nl, int(goal_position.shape[2]),
        1 if use_lie_group else 0, stream_ptr(dev))
    _lib.check(err, "tool_pose_distance")
    return out_distance, out_position_distance, out_rotation_distance, out_goalset_idx


def cspace_state_cost(pos, vel, acc, jerk, joint_torque, state_dt, target_joint_position,
                      idxs_target_joint_position, p_b, v_b, a_b, j_b, effort_b, weight, activation_distance,
                      squared_l2_regularization_weights, cspace_target_weight, cspace_non_terminal_weight_factor,
                      cspace_target_dof_weight, out_cost, out_grad_p, out_grad_v, out_grad_a, out_grad_j,
                      out_grad_tau, retime_weights: bool, retime_regularization_weights: bool, write_grad: bool = True):
    """Launch-argument order of forward_cspace_state_warp (cost/wp_cspace_state.py:21-53)."""
    b, h, d = pos.shape
    dev = pos.device
    check_tensors(dev, torch.float32, pos=pos, vel=vel, acc=acc, jerk=jerk, joint_torque=joint_torque,
                  state_dt=state_dt, target_joint_position=target_joint_position, p_b=p_b, v_b=v_b, a_b=a_b,
                  j_b=j_b, effort_b=effort_b, weight=weight, activation_distance=activation_distance,
                  squared_l2_regularization_weights=squared_l2_regularization_weights,
                  cspace_target_weight=cspace_target_weight,
                  cspace_non_terminal_weight_factor=cspace_non_terminal_weight_factor,
                  cspace_target_dof_weight=cspace_target_dof_weight, out_cost=out_cost, out_grad_p=out_grad_p,
                  out_grad_v=out_grad_v, out_grad_a=out_grad_a, out_grad_j=out_grad_j, out_grad_tau=out_grad_tau)
    check_tensors(dev, torch.int32, idxs_target_joint_position=idxs_target_joint_position)
    err = _lib.load().cb200_cspace_state_cost(
        out_cost.data_ptr(), out_grad_p.data_ptr(), out_grad_v.data_ptr(), out_grad_a.data_ptr(),
        out_grad_j.data_ptr(), out_grad_tau.data_ptr(), pos.data_ptr(), vel.data_ptr(), acc.data_ptr(),
        jerk.data_ptr(), joint_torque.data_ptr(), state_dt.data_ptr(), target_joint_position.data_ptr(),
        idxs_target_joint_position.data_ptr(), p_b.data_ptr(), v_b.data_ptr(), a_b.data_ptr(), j_b.data_ptr(),
        effort_b.data_ptr(), weight.data_ptr(), activation_distance.data_ptr(),
        squared_l2_regularization_weights.data_ptr(), cspace_target_weight.data_ptr(),
        cspace_non_terminal_weight_factor.data_ptr(), cspace_target_dof_weight.data_ptr(), int(bool(write_grad)),
        b, h, d, int(bool(retime_weights)), int(bool(retime_regularization_weights)), stream_ptr(dev))
    _lib.check(err, "cspace_state_cost")
    return out_cost


def cspace_position_cost(pos, effort, cspace_target, cspace_target_idx, p_b, effort_b, weight, activation_distance,
                         cspace_target_weight, cspace_target_dof_weight, squared_l2_reg_weight, current_position,
                         current_velocity, idxs_current_state, v_b, state_dt, out_cost, out_grad_p, out_grad_tau,
                         write_grad: bool = True):
    """Launch-argument order of forward_cspace_position_warp (cost/wp_cspace_position.py:232-256)."""
    b, h, d = pos.shape
    dev = pos.device
    check_tensors(dev, torch.float32, pos=pos, effort=effort, cspace_target=cspace_target, p_b=p_b, effort_b=effort_b,
                  weight=weight, activation_distance=activation_distance, cspace_target_weight=cspace_target_weight,
                  cspace_target_dof_weight=cspace_target_dof_weight, squared_l2_reg_weight=squared_l2_reg_weight,
                  current_position=current_position, current_velocity=current_velocity, v_b=v_b, state_dt=state_dt,
                  out_cost=out_cost, out_grad_p=out_grad_p, out_grad_tau=out_grad_tau)
    check_tensors(dev, torch.int32, cspace_target_idx=cspace_target_idx, idxs_current_state=idxs_current_state)
    err = _lib.load().cb200_cspace_position_cost(
        out_cost.data_ptr(), out_grad_p.data_ptr(), out_grad_tau.data_ptr(), pos.data_ptr(), effort.data_ptr(),
        cspace_target.data_ptr(), cspace_target_idx.data_ptr(), p_b.data_ptr(), effort_b.data_ptr(),
        weight.data_ptr(), activation_distance.data_ptr(), cspace_target_weight.data_ptr(),
        cspace_target_dof_weight.data_ptr(), squared_l2_reg_weight.data_ptr(), current_position.data_ptr(),
        current_velocity.data_ptr(), idxs_current_state.data_ptr(), v_b.data_ptr(), state_dt.data_ptr(),
        int(bool(write_grad)), b, h, d, stream_ptr(dev))
    _lib.check(err, "cspace_position_cost")
    return out_cost
