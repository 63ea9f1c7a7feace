"""Host mirror of the reference's inverse-dynamics operator.

  RNEAFunction  <- curobo/_src/curobolib/cuda_ops/dynamics.py (torch.autograd.Function around the two launches)
  Dynamics      <- curobo/_src/robot/dynamics/dynamics.py:40-330 (buffers per (batch, horizon), compute_inverse_dynamics)

CUDA only.  The inertial tensors (link_masses_com [nl,4], link_inertias [nl,8]) come from the caller, exactly as the
reference's KinematicsParams carries them; building them from a URDF is the reference loader's job and out of scope.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .backends import dynamics as dynamics_cu
from .robot_model import RobotModel


def tree_levels(link_map: np.ndarray):
    """CSR of links by tree depth (KinematicsParams.link_level_offsets / link_level_data)."""
    nl = len(link_map)
    depth = np.zeros(nl, np.int32)
    for k in range(1, nl):
        depth[k] = depth[int(link_map[k])] + 1
    order = np.argsort(depth, kind="stable").astype(np.int16)
    starts = np.zeros(int(depth.max()) + 2, np.int64)
    for d in depth:
        starts[d + 1] += 1
    return np.cumsum(starts).astype(np.int16), order


class RNEAFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, qd, qdd, dyn: "Dynamics"):
        B = q.shape[0]
        tau, cache = dyn._tau[:B], dyn._cache[:B]
        dynamics_cu.launch_rnea_forward(tau, q, qd, qdd, *dyn._model, cache, B, dyn.num_links, dyn.num_dof, dyn.n_levels)
        ctx.dyn = dyn
        ctx.save_for_backward(q, qd)
        return tau

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_tau):
        dyn = ctx.dyn
        q, qd = ctx.saved_tensors
        B = q.shape[0]
        gq, gqd, gqdd = dyn._gq[:B], dyn._gqd[:B], dyn._gqdd[:B]
        dynamics_cu.launch_rnea_backward(gq, gqd, gqdd, grad_tau.contiguous(), q, qd, *dyn._model, dyn._cache[:B], B,
                                         dyn.num_links, dyn.num_dof, dyn.n_levels)
        return gq, gqd, gqdd, None


class Dynamics:
    """tau = RNEA(q, qd, qdd) with gravity, differentiable.  q / qd / qdd: [batch, horizon, dof] or [N, dof]."""

    def __init__(self, robot: RobotModel, link_masses_com, link_inertias, gravity=(0.0, 0.0, -9.81), device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("Dynamics is CUDA-only; there is no CPU path")
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(self.device)  # noqa: E731
        self.num_links, self.num_dof = robot.num_links, robot.num_dof
        starts, order = tree_levels(robot.link_map)
        self.n_levels = len(starts) - 1
        # spatial gravity: the base accelerates with -g (robot/dynamics/dynamics.py get_gravity_spatial)
        g6 = np.array([0, 0, 0, -gravity[0], -gravity[1], -gravity[2]], np.float32)
        self._model = (t(robot.fixed_transforms, np.float32), t(link_masses_com, np.float32), t(link_inertias, np.float32),
                       t(robot.joint_map_type, np.int8), t(robot.joint_map, np.int16), t(robot.link_map, np.int16),
                       t(robot.joint_offset_map, np.float32), t(g6, np.float32), t(starts, np.int16), t(order, np.int16))
        self._n = 0

    def setup_batch_size(self, batch_size: int, horizon: int = 1) -> None:
        n = batch_size * horizon
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)  # noqa: E731
        self._tau, self._gq, self._gqd, self._gqdd = (z(n, self.num_dof) for _ in range(4))
        self._cache = z(n, self.num_links * 20)
        self._n = n

    def compute_inverse_dynamics(self, q: torch.Tensor, qd: torch.Tensor, qdd: torch.Tensor) -> torch.Tensor:
        shape = q.shape
        q2, qd2, qdd2 = (x.reshape(-1, self.num_dof).contiguous() for x in (q, qd, qdd))
        if q2.shape[0] > self._n:
            self.setup_batch_size(q2.shape[0])
        return RNEAFunction.apply(q2, qd2, qdd2, self).view(shape)
