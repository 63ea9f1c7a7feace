"""`kinematics` kernel-backend module: same function names and argument order as
curobo/_src/curobolib/backends/cuda_core_backend/kinematics.py:90-115,282-312 (and the pybind twin,
backends/pybind/kinematics_bindings.cpp:14-120), so it can be registered as a third backend under
curobo._src.curobolib.backends and `KinematicsFusedFunction` works unchanged.

Tensors are validated like curobo/_src/curobolib/cuda_ops/tensor_checks.py (device, contiguity,
dtype) BEFORE launch -- no `.contiguous()` fix-ups (CUDA-graph hazard) -- and errors raise ValueError.
Launches go to `torch.cuda.current_stream()` (backends/cuda_core_backend/kinematics.py:130-131).
"""
from __future__ import annotations

import torch

from .. import lib as _lib
from .tensor_checks import check_tensors, stream_ptr


def launch_kinematics_forward(
    link_pos: torch.Tensor,
    link_quat: torch.Tensor,
    batch_center_of_mass: torch.Tensor,
    global_cumul_mat: torch.Tensor,
    joint_vec: torch.Tensor,
    fixed_transform: torch.Tensor,
    link_masses_com: torch.Tensor,
    joint_map_type: torch.Tensor,
    joint_map: torch.Tensor,
    link_map: torch.Tensor,
    tool_frame_map: torch.Tensor,
    joint_offset_map: torch.Tensor,
    batch_size: int,
    horizon: int,
    n_joints: int,
    compute_com: bool = False,
) -> None:
    """FK without sphere output (cuda_core_backend/kinematics.py:21-88): tool poses + cumulative transforms
    (+ batch_center_of_mass [B*H, 4] = world CoM xyz, total mass, when compute_com)."""
    dev = joint_vec.device
    if compute_com:
        check_tensors(dev, torch.float32, batch_center_of_mass=batch_center_of_mass, link_masses_com=link_masses_com)
    check_tensors(dev, torch.float32, link_pos=link_pos, link_quat=link_quat, global_cumul_mat=global_cumul_mat,
                  joint_vec=joint_vec, fixed_transform=fixed_transform, joint_offset_map=joint_offset_map)
    check_tensors(dev, torch.int8, joint_map_type=joint_map_type)
    check_tensors(dev, torch.int16, joint_map=joint_map, link_map=link_map, tool_frame_map=tool_frame_map)
    L = _lib.load()
    err = L.cb200_kinematics_forward_spheres(
        link_pos.data_ptr(), link_quat.data_ptr(), None, batch_center_of_mass.data_ptr() if compute_com else None,
        global_cumul_mat.data_ptr(), joint_vec.data_ptr(), fixed_transform.data_ptr(), None,
        link_masses_com.data_ptr() if compute_com else None, joint_map_type.data_ptr(), joint_map.data_ptr(), link_map.data_ptr(),
        tool_frame_map.data_ptr(), None, joint_offset_map.data_ptr(), None, 1, int(batch_size), int(horizon), int(n_joints),
        0, int(link_map.shape[0]), int(tool_frame_map.shape[0]), 1, int(bool(compute_com)), stream_ptr(dev))
    _lib.check(err, "launch_kinematics_forward")


def launch_kinematics_forward_spheres_jacobian(*args, **kwargs) -> None:
    """Jacobian-producing FK variant (cuda_core_backend/kinematics.py:180-280): not on the rollout hot path
    (SURVEY.md section 8a lists a2-a5 only).  Raises instead of silently doing nothing."""
    raise ValueError("b200 backend: launch_kinematics_forward_spheres_jacobian is outside the hot-path scope; "
                     "keep the reference backend for Jacobian queries")


def launch_kinematics_forward_spheres(
    link_pos: torch.Tensor,
    link_quat: torch.Tensor,
    batch_robot_spheres: torch.Tensor,
    batch_center_of_mass: torch.Tensor,
    global_cumul_mat: torch.Tensor,
    joint_vec: torch.Tensor,
    fixed_transform: torch.Tensor,
    robot_spheres: torch.Tensor,
    link_masses_com: torch.Tensor,
    joint_map_type: torch.Tensor,
    joint_map: torch.Tensor,
    link_map: torch.Tensor,
    tool_frame_map: torch.Tensor,
    link_sphere_map: torch.Tensor,
    joint_offset_map: torch.Tensor,
    env_query_idx: torch.Tensor,
    num_envs: int,
    batch_size: int,
    horizon: int,
    n_joints: int,
    num_spheres: int,
    output_threads_per_batch: int = 32,
    write_global_cumul: bool = True,
    compute_com: bool = False,
) -> None:
    """FK + spheres + tool poses; outputs are the passed-in buffers (written in place)."""
    if output_threads_per_batch not in (32, 64, 128):
        raise ValueError("output_threads_per_batch must be one of 32, 64, or 128")
    dev = joint_vec.device
    if compute_com:
        check_tensors(dev, torch.float32, batch_center_of_mass=batch_center_of_mass, link_masses_com=link_masses_com)
    check_tensors(dev, torch.float32, link_pos=link_pos, link_quat=link_quat,
                  batch_robot_spheres=batch_robot_spheres, global_cumul_mat=global_cumul_mat, joint_vec=joint_vec,
                  fixed_transform=fixed_transform, robot_spheres=robot_spheres, joint_offset_map=joint_offset_map)
    check_tensors(dev, torch.int8, joint_map_type=joint_map_type)
    check_tensors(dev, torch.int16, joint_map=joint_map, link_map=link_map, tool_frame_map=tool_frame_map,
                  link_sphere_map=link_sphere_map)
    check_tensors(dev, torch.int32, env_query_idx=env_query_idx)
    L = _lib.load()
    err = L.cb200_kinematics_forward_spheres(
        link_pos.data_ptr(), link_quat.data_ptr(), batch_robot_spheres.data_ptr(),
        batch_center_of_mass.data_ptr() if batch_center_of_mass is not None else None,
        global_cumul_mat.data_ptr(), joint_vec.data_ptr(), fixed_transform.data_ptr(), robot_spheres.data_ptr(),
        link_masses_com.data_ptr() if link_masses_com is not None else None,
        joint_map_type.data_ptr(), joint_map.data_ptr(), link_map.data_ptr(), tool_frame_map.data_ptr(),
        link_sphere_map.data_ptr(), joint_offset_map.data_ptr(), env_query_idx.data_ptr(),
        int(num_envs), int(batch_size), int(horizon), int(n_joints), int(num_spheres), int(link_map.shape[0]),
        int(tool_frame_map.shape[0]), int(bool(write_global_cumul)), int(bool(compute_com)), stream_ptr(dev))
    _lib.check(err, "launch_kinematics_forward_spheres")


def launch_kinematics_backward(
    grad_out: torch.Tensor,
    grad_nlinks_pos: torch.Tensor,
    grad_nlinks_quat: torch.Tensor,
    grad_spheres: torch.Tensor,
    grad_center_of_mass: torch.Tensor,
    batch_center_of_mass: torch.Tensor,
    grad_jacobian: torch.Tensor,
    global_cumul_mat: torch.Tensor,
    robot_spheres: torch.Tensor,
    link_masses_com: torch.Tensor,
    link_map: torch.Tensor,
    joint_map: torch.Tensor,
    joint_map_type: torch.Tensor,
    tool_frame_map: torch.Tensor,
    link_sphere_map: torch.Tensor,
    link_chain_data: torch.Tensor,
    link_chain_offsets: torch.Tensor,
    joint_links_data: torch.Tensor,
    joint_links_offsets: torch.Tensor,
    joint_affects_endeffector: torch.Tensor,
    joint_offset_map: torch.Tensor,
    env_query_idx: torch.Tensor,
    num_envs: int,
    batch_size: int,
    horizon: int,
    n_joints: int,
    num_spheres: int,
    compute_com: bool = False,
    compute_jacobian_grad: bool = False,
) -> None:
    """grad_out[B*H, D] (overwritten) from sphere / tool-pose gradients and the saved cumul matrices."""
    if compute_jacobian_grad:
        raise ValueError("b200 backend: Jacobian-output gradients are outside the hot-path scope")
    dev = grad_out.device
    if compute_com:
        check_tensors(dev, torch.float32, grad_center_of_mass=grad_center_of_mass, batch_center_of_mass=batch_center_of_mass,
                      link_masses_com=link_masses_com)
    check_tensors(dev, torch.float32, grad_out=grad_out, grad_nlinks_pos=grad_nlinks_pos,
                  grad_nlinks_quat=grad_nlinks_quat, grad_spheres=grad_spheres, global_cumul_mat=global_cumul_mat,
                  robot_spheres=robot_spheres, joint_offset_map=joint_offset_map)
    check_tensors(dev, torch.int16, link_map=link_map, joint_map=joint_map, tool_frame_map=tool_frame_map,
                  link_sphere_map=link_sphere_map)
    check_tensors(dev, torch.int8, joint_map_type=joint_map_type)
    check_tensors(dev, torch.int32, env_query_idx=env_query_idx)
    if grad_nlinks_quat.data_ptr() % 16 != 0:          # cuda_ops/kinematics.py:315-316
        raise ValueError("grad_nlinks_quat must be 16-byte aligned")

    def p(t):
        return t.data_ptr() if t is not None else None
    L = _lib.load()
    err = L.cb200_kinematics_backward(
        grad_out.data_ptr(), grad_nlinks_pos.data_ptr(), grad_nlinks_quat.data_ptr(), grad_spheres.data_ptr(),
        p(grad_center_of_mass), p(batch_center_of_mass), p(grad_jacobian), global_cumul_mat.data_ptr(),
        robot_spheres.data_ptr(), p(link_masses_com), link_map.data_ptr(), joint_map.data_ptr(),
        joint_map_type.data_ptr(), tool_frame_map.data_ptr(), link_sphere_map.data_ptr(), p(link_chain_data),
        p(link_chain_offsets), p(joint_links_data), p(joint_links_offsets), p(joint_affects_endeffector),
        joint_offset_map.data_ptr(), env_query_idx.data_ptr(), int(num_envs), int(batch_size), int(horizon),
        int(n_joints), int(num_spheres), int(link_map.shape[0]), int(tool_frame_map.shape[0]), int(bool(compute_com)), 0,
        stream_ptr(dev))
    _lib.check(err, "launch_kinematics_backward")
