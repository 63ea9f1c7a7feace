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
from .backends import tensor_checks as _tc
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
        _tc.require_cuda(self.device, "Dynamics is CUDA-only; there is no CPU path")
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


class DynamicsStateCost:
    """The dynamics-aware STATE c-space cost, unfused: what `RobotStateTransition.compute_augmented_state`
    (transition/robot_state_transition.py:380-389: joint_torque = robot_dynamics.compute_inverse_dynamics(state)) followed
    by `StateCSpaceFunction` (cost/wp_cspace_state.py:21-285, effort channel: bound hinge, squared-L2 and the energy term
    (tau qd dt)^2) and autograd's walk back through the RNEA adjoint compute -- as three launches and three in-place adds,
    no autograd graph, CUDA-graph capturable:

        tau = RNEA(q, qd, qdd)                                   cb200_rnea_forward
        cost, g_p, g_v, g_a, g_j, g_tau = cspace_state(...)      cb200_cspace_state_cost
        g_p, g_v, g_a += RNEA^T(g_tau)                           cb200_rnea_backward

    All tensors [batch, horizon, dof]; `limits` = dict p / v / a / j / tau -> [2, dof] device tensors; weight,
    activation_distance, reg_weights are the five-element vectors of the reference's cost config.
    """

    def __init__(self, dynamics: Dynamics, limits: dict, weight, activation_distance, reg_weights,
                 retime_weights: bool = True, retime_regularization_weights: bool = True):
        from . import cost as _cost
        self._cost = _cost
        self.dyn = dynamics
        dev = dynamics.device
        f = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, np.float32))).to(dev)  # noqa: E731
        self.limits = {k: f(v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in limits.items()}
        self.weight, self.act, self.reg = f(weight), f(activation_distance), f(reg_weights)
        self.retime_w, self.retime_r = bool(retime_weights), bool(retime_regularization_weights)
        D = dynamics.num_dof
        self._zero_w, self._one = f([0.0]), f([1.0])
        self._dof_w = f(np.ones(D))
        self._target = f(np.zeros((1, D)))
        self._shape = None

    def _setup(self, shape):
        B, H, D = shape
        dev = self.dyn.device
        z = lambda: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
        self.out_cost, self.g_p, self.g_v, self.g_a, self.g_j, self.g_tau = (z() for _ in range(6))
        self._idx = torch.zeros(B, dtype=torch.int32, device=dev)
        self.dyn.setup_batch_size(B, H)
        self._shape = tuple(shape)

    def evaluate(self, pos, vel, acc, jerk, state_dt, target_joint_position: Optional[torch.Tensor] = None,
                 idxs_target: Optional[torch.Tensor] = None, target_weight: Optional[torch.Tensor] = None,
                 non_terminal_factor: Optional[torch.Tensor] = None, target_dof_weight: Optional[torch.Tensor] = None):
        """-> (cost [B,H,D], grad_p, grad_v, grad_a, grad_j, tau); buffers are reused between calls."""
        if self._shape != tuple(pos.shape):
            self._setup(pos.shape)
        B, H, D = pos.shape
        dyn, n = self.dyn, B * H
        q2, qd2, qdd2 = (x.reshape(n, D) for x in (pos, vel, acc))
        tau = dyn._tau[:n]
        dynamics_cu.launch_rnea_forward(tau, q2, qd2, qdd2, *dyn._model, dyn._cache[:n], n, dyn.num_links, D, dyn.n_levels)
        lim = self.limits
        self._cost.cspace_state_cost(
            pos, vel, acc, jerk, tau.view(B, H, D), state_dt,
            self._target if target_joint_position is None else target_joint_position,
            self._idx if idxs_target is None else idxs_target, lim["p"], lim["v"], lim["a"], lim["j"], lim["tau"],
            self.weight, self.act, self.reg, self._zero_w if target_weight is None else target_weight,
            self._one if non_terminal_factor is None else non_terminal_factor,
            self._dof_w if target_dof_weight is None else target_dof_weight, self.out_cost, self.g_p, self.g_v, self.g_a,
            self.g_j, self.g_tau, self.retime_w, self.retime_r)
        gq, gqd, gqdd = dyn._gq[:n], dyn._gqd[:n], dyn._gqdd[:n]
        dynamics_cu.launch_rnea_backward(gq, gqd, gqdd, self.g_tau.view(n, D), q2, qd2, *dyn._model, dyn._cache[:n], n,
                                         dyn.num_links, D, dyn.n_levels)
        self.g_p.view(n, D).add_(gq)
        self.g_v.view(n, D).add_(gqd)
        self.g_a.view(n, D).add_(gqdd)
        return self.out_cost, self.g_p, self.g_v, self.g_a, self.g_j, tau.view(B, H, D)
