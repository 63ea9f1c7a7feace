"""Host-side mirror of the reference's kinematics + self-collision operator layer on top of the b200
backend modules: same class / function names, argument meaning and buffer-ownership rules as

  KinematicsParams          curobo/_src/robot/types/kinematics_params.py:23-158
  KinematicsFusedFunction   curobo/_src/curobolib/cuda_ops/kinematics.py:25-356
  Kinematics                curobo/_src/robot/kinematics/kinematics.py:75-198
  SelfCollisionDistance     curobo/_src/curobolib/cuda_ops/geometry.py:18-104
  SelfCollisionCost         curobo/_src/cost/cost_self_collision.py:31-130

so the parity tests read like the reference's own tests.  Everything here is plumbing: every number
is produced by the CUDA kernels behind curobo_b200.backends (no CPU path).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .backends import geometry as geometry_cu
from .backends import kinematics as kinematics_cu
from .robot_model import RobotModel


@dataclass
class KinematicsParams:
    fixed_transforms: torch.Tensor
    link_map: torch.Tensor
    joint_map: torch.Tensor
    joint_map_type: torch.Tensor
    joint_offset_map: torch.Tensor
    tool_frame_map: torch.Tensor
    link_spheres: torch.Tensor            # [n_cfg, S, 4]
    link_sphere_idx_map: torch.Tensor
    link_chain_data: torch.Tensor
    link_chain_offsets: torch.Tensor
    joint_links_data: torch.Tensor
    joint_links_offsets: torch.Tensor
    joint_affects_endeffector: torch.Tensor
    link_masses_com: torch.Tensor
    num_dof: int
    num_envs: int = 1

    @property
    def num_links(self) -> int:
        return int(self.link_map.shape[0])

    @property
    def num_spheres(self) -> int:
        return int(self.link_spheres.shape[1])

    @property
    def num_pose_links(self) -> int:
        return int(self.tool_frame_map.shape[0])

    @classmethod
    def from_robot_model(cls, rm: RobotModel, device) -> "KinematicsParams":
        t = lambda a: torch.as_tensor(a).to(device).contiguous()  # noqa: E731
        ls = rm.link_spheres if rm.link_spheres.ndim == 3 else rm.link_spheres[None]
        return cls(t(rm.fixed_transforms), t(rm.link_map), t(rm.joint_map), t(rm.joint_map_type),
                   t(rm.joint_offset_map.reshape(-1)), t(rm.tool_frame_map), t(ls), t(rm.link_sphere_idx_map),
                   t(rm.link_chain_data), t(rm.link_chain_offsets), t(rm.joint_links_data),
                   t(rm.joint_links_offsets), t(rm.joint_affects_endeffector.astype("uint8")),
                   t(rm.link_masses_com), rm.num_dof, int(ls.shape[0]))


class KinematicsFusedFunction(torch.autograd.Function):
    """FK forward writes the passed-in buffers and returns them; backward launches the J^T kernel
    (cuda_ops/kinematics.py:93-356).  Absent incoming gradients use the caller's pre-zeroed buffers."""

    @staticmethod
    def create_buffers(batch: int, horizon: int, kp: KinematicsParams, device) -> dict:
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)  # noqa: E731
        L, S, D, nl = kp.num_pose_links, kp.num_spheres, kp.num_dof, kp.num_links
        return {"batch_link_position": z(batch, horizon, L, 3), "batch_link_quaternion": z(batch, horizon, L, 4),
                "batch_robot_spheres": z(batch, horizon, S, 4), "batch_com": z(batch, horizon, 4),
                "batch_cumul_mat": z(batch, horizon, nl, 3, 4), "grad_out_q": z(batch, horizon, D),
                "grad_in_link_pos": z(batch, horizon, L, 3), "grad_in_link_quat": z(batch, horizon, L, 4),
                "grad_in_robot_spheres": z(batch, horizon, S, 4)}

    @staticmethod
    def forward(ctx, joint_seq, batch_link_position, batch_link_quaternion, batch_robot_spheres, batch_com,
                batch_cumul_mat, kp: KinematicsParams, grad_out, grad_in_link_pos, grad_in_link_quat,
                grad_in_robot_spheres, env_query_idx, horizon: int):
        b_size = batch_link_position.shape[0] * batch_link_position.shape[1]
        ctx.set_materialize_grads(False)
        kinematics_cu.launch_kinematics_forward_spheres(
            batch_link_position, batch_link_quaternion, batch_robot_spheres, batch_com, batch_cumul_mat,
            joint_seq.detach(), kp.fixed_transforms, kp.link_spheres, kp.link_masses_com, kp.joint_map_type,
            kp.joint_map, kp.link_map, kp.tool_frame_map, kp.link_sphere_idx_map, kp.joint_offset_map, env_query_idx,
            kp.num_envs, b_size, horizon, joint_seq.shape[-1], batch_robot_spheres.shape[2], 32, True, False)
        ctx.kp, ctx.horizon, ctx.env_query_idx = kp, horizon, env_query_idx
        ctx.bufs = (grad_out, grad_in_link_pos, grad_in_link_quat, grad_in_robot_spheres, batch_com)
        ctx.save_for_backward(batch_cumul_mat)
        ctx.mark_non_differentiable(batch_cumul_mat)
        return batch_link_position, batch_link_quaternion, batch_robot_spheres, batch_cumul_mat

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_pos, g_quat, g_sph, g_cumul):
        (cumul,) = ctx.saved_tensors
        kp = ctx.kp
        grad_out, z_pos, z_quat, z_sph, batch_com = ctx.bufs
        g_pos = z_pos if g_pos is None else g_pos.contiguous()
        g_quat = z_quat if g_quat is None else g_quat.contiguous()
        g_sph = z_sph if g_sph is None else g_sph.contiguous()
        b_size = cumul.shape[0] * cumul.shape[1]
        kinematics_cu.launch_kinematics_backward(
            grad_out, g_pos, g_quat, g_sph, None, batch_com, None, cumul, kp.link_spheres, kp.link_masses_com,
            kp.link_map, kp.joint_map, kp.joint_map_type, kp.tool_frame_map, kp.link_sphere_idx_map,
            kp.link_chain_data, kp.link_chain_offsets, kp.joint_links_data, kp.joint_links_offsets,
            kp.joint_affects_endeffector, kp.joint_offset_map, ctx.env_query_idx, kp.num_envs, b_size, ctx.horizon,
            kp.num_dof, g_sph.shape[2], False, False)
        return (grad_out,) + (None,) * 12


class KinematicsComFunction(torch.autograd.Function):
    """KinematicsFusedFunction with `compute_com=True` (cuda_ops/kinematics.py:93-356): the forward pass also writes the
    centre of mass [B,H,4] (world xyz, total mass), the backward pass takes its gradient (xyz; the mass slot is ignored,
    kinematics_backward_helper.cuh:187-260).  A separate class so that the path without CoM stays exactly as it is."""

    @staticmethod
    def forward(ctx, joint_seq, batch_link_position, batch_link_quaternion, batch_robot_spheres, batch_com,
                batch_cumul_mat, kp: KinematicsParams, grad_out, grad_in_link_pos, grad_in_link_quat,
                grad_in_robot_spheres, grad_in_com, env_query_idx, horizon: int):
        b_size = batch_link_position.shape[0] * batch_link_position.shape[1]
        ctx.set_materialize_grads(False)
        kinematics_cu.launch_kinematics_forward_spheres(
            batch_link_position, batch_link_quaternion, batch_robot_spheres, batch_com, batch_cumul_mat,
            joint_seq.detach(), kp.fixed_transforms, kp.link_spheres, kp.link_masses_com, kp.joint_map_type,
            kp.joint_map, kp.link_map, kp.tool_frame_map, kp.link_sphere_idx_map, kp.joint_offset_map, env_query_idx,
            kp.num_envs, b_size, horizon, joint_seq.shape[-1], batch_robot_spheres.shape[2], 32, True, True)
        ctx.kp, ctx.horizon, ctx.env_query_idx = kp, horizon, env_query_idx
        ctx.bufs = (grad_out, grad_in_link_pos, grad_in_link_quat, grad_in_robot_spheres, grad_in_com)
        ctx.save_for_backward(batch_cumul_mat, batch_com)
        ctx.mark_non_differentiable(batch_cumul_mat)
        return batch_link_position, batch_link_quaternion, batch_robot_spheres, batch_cumul_mat, batch_com

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_pos, g_quat, g_sph, g_cumul, g_com):
        cumul, batch_com = ctx.saved_tensors
        kp = ctx.kp
        grad_out, z_pos, z_quat, z_sph, z_com = ctx.bufs
        g_pos = z_pos if g_pos is None else g_pos.contiguous()
        g_quat = z_quat if g_quat is None else g_quat.contiguous()
        g_sph = z_sph if g_sph is None else g_sph.contiguous()
        g_com = z_com if g_com is None else g_com.contiguous()
        b_size = cumul.shape[0] * cumul.shape[1]
        kinematics_cu.launch_kinematics_backward(
            grad_out, g_pos, g_quat, g_sph, g_com, batch_com, None, cumul, kp.link_spheres, kp.link_masses_com,
            kp.link_map, kp.joint_map, kp.joint_map_type, kp.tool_frame_map, kp.link_sphere_idx_map,
            kp.link_chain_data, kp.link_chain_offsets, kp.joint_links_data, kp.joint_links_offsets,
            kp.joint_affects_endeffector, kp.joint_offset_map, ctx.env_query_idx, kp.num_envs, b_size, ctx.horizon,
            kp.num_dof, g_sph.shape[2], True, False)
        return (grad_out,) + (None,) * 13


@dataclass
class KinematicsState:
    tool_pose_position: torch.Tensor      # [B,H,L,3]
    tool_pose_quaternion: torch.Tensor    # [B,H,L,4] wxyz
    robot_spheres: torch.Tensor           # [B,H,S,4]
    cumul_mat: torch.Tensor               # [B,H,nl,3,4]
    center_of_mass: Optional[torch.Tensor] = None   # [B,H,4] world xyz, total mass (Kinematics(compute_com=True))


class Kinematics:
    """`Kinematics.compute_kinematics` of the reference (robot/kinematics/kinematics.py:172-198)."""

    def __init__(self, robot: RobotModel, device="cuda:0", compute_com: bool = False):
        self.device = torch.device(device)
        self.robot = robot
        self.compute_com = bool(compute_com)
        self.params = KinematicsParams.from_robot_model(robot, self.device)
        self._shape = None
        self._env0 = torch.zeros(1, dtype=torch.int32, device=self.device)

    def update_batch_size(self, batch: int, horizon: int) -> None:
        if self._shape != (batch, horizon):
            self._bufs = KinematicsFusedFunction.create_buffers(batch, horizon, self.params, self.device)
            self._shape = (batch, horizon)

    def compute_kinematics(self, q: torch.Tensor, env_query_idx: Optional[torch.Tensor] = None) -> KinematicsState:
        if q.ndim == 2:
            q = q.unsqueeze(1)
        if q.ndim != 3 or q.shape[-1] != self.params.num_dof:
            raise ValueError(f"joint tensor must be [B, (H,) {self.params.num_dof}], got {tuple(q.shape)}")
        b, h, _ = q.shape
        self.update_batch_size(b, h)
        B = self._bufs
        eq = env_query_idx if env_query_idx is not None else self._env0
        if self.compute_com:
            if "grad_in_com" not in B:
                B["grad_in_com"] = torch.zeros((b, h, 4), dtype=torch.float32, device=self.device)
            pos, quat, sph, cum, com = KinematicsComFunction.apply(
                q, B["batch_link_position"], B["batch_link_quaternion"], B["batch_robot_spheres"], B["batch_com"],
                B["batch_cumul_mat"], self.params, B["grad_out_q"], B["grad_in_link_pos"], B["grad_in_link_quat"],
                B["grad_in_robot_spheres"], B["grad_in_com"], eq, h)
            return KinematicsState(pos, quat, sph, cum, com)
        pos, quat, sph, cum = KinematicsFusedFunction.apply(
            q, B["batch_link_position"], B["batch_link_quaternion"], B["batch_robot_spheres"], B["batch_com"],
            B["batch_cumul_mat"], self.params, B["grad_out_q"], B["grad_in_link_pos"], B["grad_in_link_quat"],
            B["grad_in_robot_spheres"], eq, h)
        return KinematicsState(pos, quat, sph, cum)


class SelfCollisionDistance(torch.autograd.Function):
    """cuda_ops/geometry.py:18-104: forward writes distance + gradient buffers, backward hands the
    gradient buffer out (scaled by the incoming gradient only if return_loss)."""

    @staticmethod
    def forward(ctx, robot_spheres, out_distance, out_vec, pair_distance, sparse_idx, weight, sphere_padding,
                pair_locations, block_batch_max_value, block_batch_max_index, num_blocks_per_batch,
                max_threads_per_block, store_pair_distance, return_loss):
        ctx.set_materialize_grads(False)
        b, h, n, _ = robot_spheres.shape
        geometry_cu.self_collision_distance(
            out_distance, out_vec, pair_distance, sparse_idx, robot_spheres.detach(), sphere_padding, weight,
            pair_locations, block_batch_max_value, block_batch_max_index, num_blocks_per_batch,
            max_threads_per_block, b, h, n, pair_locations.shape[0], store_pair_distance,
            robot_spheres.requires_grad)
        ctx.return_loss = return_loss
        ctx.save_for_backward(out_vec)
        return out_distance

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out_distance):
        g = None
        if grad_out_distance is not None and ctx.needs_input_grad[0]:
            (g,) = ctx.saved_tensors
            if ctx.return_loss:
                g = g * grad_out_distance
        return (g,) + (None,) * 13


class SelfCollisionCost:
    """cost/cost_self_collision.py:31-130 (buffers per (B,H); forward returns [B,H,1])."""

    def __init__(self, robot: RobotModel, weight: float, device="cuda:0"):
        self.device = torch.device(device)
        t = lambda a: torch.as_tensor(a).to(self.device).contiguous()  # noqa: E731
        self.robot = robot
        self.weight = torch.tensor([weight], dtype=torch.float32, device=self.device)
        self.sphere_padding = t(robot.sphere_padding)
        self.pairs = t(robot.collision_pairs)
        self._shape = None

    def setup_batch_tensors(self, batch: int, horizon: int) -> None:
        S, dev = self.robot.num_spheres, self.device
        nb = self.robot.num_blocks_per_batch
        self._out_distance = torch.zeros((batch, horizon, 1), dtype=torch.float32, device=dev)
        self._out_vec = torch.zeros((batch, horizon, S, 4), dtype=torch.float32, device=dev)
        self._sparse = torch.zeros((batch, horizon, S), dtype=torch.uint8, device=dev)
        self._pair_distance = torch.zeros((1,), dtype=torch.float32, device=dev)
        self._bbmv = torch.zeros((batch, horizon, nb), dtype=torch.float32, device=dev)
        self._bbmi = torch.zeros((batch, horizon, nb, 2), dtype=torch.int16, device=dev)
        self._shape = (batch, horizon)

    def forward(self, robot_spheres: torch.Tensor) -> torch.Tensor:
        b, h, _, _ = robot_spheres.shape
        if self._shape != (b, h):
            self.setup_batch_tensors(b, h)
        return SelfCollisionDistance.apply(
            robot_spheres, self._out_distance, self._out_vec, self._pair_distance, self._sparse, self.weight,
            self.sphere_padding, self.pairs, self._bbmv, self._bbmi, self.robot.num_blocks_per_batch,
            self.robot.max_threads_per_block, False, False)
