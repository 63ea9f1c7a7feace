"""Host side of the fused rollout cost+gradient evaluation.

`RolloutEngine.evaluate_action(q)` is what one optimizer iteration calls: it stands where the
reference has  RobotRollout.evaluate_action -> RobotCostManager.compute_costs -> sum -> backward
(curobo/_src/rollout/rollout_robot.py:252-263; rollout/cost_manager/cost_manager_robot.py:195-286;
optim/components/gradient_opt_core.py:445-480) and returns the per-row cost and d(cost)/dq from ONE
kernel launch.  All buffers are allocated once per (B, H) (reference ownership contract,
cuda_ops/kinematics.py:27-90) so the call is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from .backends.tensor_checks import check_tensors, stream_ptr
from .backends import tensor_checks as _tc
from .robot_model import RobotModel
from .scene import CuboidData, VoxelData, c_cuboid_set, c_voxel_set


@dataclass
class RolloutConfig:
    """Weights/switches of the cost terms (value semantics; a zero weight disables the term)."""
    self_weight: float = 0.0
    scene_weight: float = 0.0
    scene_activation: float = 0.0
    use_sweep: bool = False
    use_speed_metric: bool = False
    pose_weight: Optional[Sequence[float]] = None        # (position, rotation)
    pose_lie: bool = False
    cspace_type: Optional[str] = None                    # None | "position" | "state"
    cspace_weight: Sequence[float] = (0.0,) * 5
    cspace_activation: Sequence[float] = (0.0,) * 5
    cspace_reg: Sequence[float] = (0.0,) * 5
    retime_weights: bool = True
    retime_reg: bool = True
    cspace_target_weight: float = 0.0                    # needs RolloutEngine.update_cspace_target
    cspace_non_terminal_weight_factor: float = 1.0       # STATE cost only (waypoints h < H-1)

    # shipped task configs of the reference -------------------------------------------------
    @classmethod
    def ik(cls) -> "RolloutConfig":
        """content/configs/task/ik/lbfgs_ik.yml:4-37."""
        return cls(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.0, pose_weight=(10000.0, 500.0),
                   cspace_type="position", cspace_weight=(5000.0, 1000.0, 0, 0, 0),
                   cspace_activation=(0.01, 0.01, 0, 0, 0))

    @classmethod
    def trajopt(cls) -> "RolloutConfig":
        """content/configs/task/trajopt/lbfgs_bspline_trajopt.yml:40-90."""
        return cls(self_weight=10000.0, scene_weight=100000.0, scene_activation=0.0025, use_sweep=True,
                   use_speed_metric=True, pose_weight=(1000000.0, 100000.0), cspace_type="state",
                   cspace_weight=(10000.0, 10000.0, 100.0, 50.0, 100.0), cspace_activation=(0.01,) * 5,
                   cspace_reg=(1000.0, 10000.0, 5.0, 0.0, 10000.0))

    @classmethod
    def mpc(cls) -> "RolloutConfig":
        """content/configs/task/mpc/lbfgs_mpc.yml:4-52 (c-space target term at weight 1000, 0.05 on non-terminal
        waypoints; bound weights not retimed, regularisation weights retimed)."""
        return cls(self_weight=100000.0, scene_weight=10000.0, scene_activation=0.01, use_sweep=True,
                   use_speed_metric=True, pose_weight=(5000.0, 200.0), cspace_type="state",
                   cspace_weight=(1000.0, 1000.0, 1000.0, 100.0, 0.0), cspace_activation=(0.01,) * 5,
                   cspace_reg=(0.01, 10000.0, 10.0, 0.0, 0.0), retime_weights=False, retime_reg=True,
                   cspace_target_weight=1000.0, cspace_non_terminal_weight_factor=0.05)

    def to_oracle_cfg(self, num_tool_frames: int) -> dict:
        d = dict(self_weight=self.self_weight, scene_weight=self.scene_weight, scene_eta=self.scene_activation,
                 sweep=self.use_sweep, speed_metric=self.use_speed_metric, pose_lie=self.pose_lie,
                 cspace_type=self.cspace_type, cspace_weight=list(self.cspace_weight),
                 cspace_activation=list(self.cspace_activation), cspace_reg=list(self.cspace_reg),
                 retime_weights=self.retime_weights, retime_reg=self.retime_reg,
                 cspace_target_weight=self.cspace_target_weight,
                 cspace_non_terminal_weight_factor=self.cspace_non_terminal_weight_factor)
        if self.pose_weight is not None:
            d["pose_weight"] = list(self.pose_weight)
        return d


@dataclass
class RolloutOutput:
    cost: torch.Tensor                 # [B,H] sum of all terms per row
    grad_q: torch.Tensor               # [B,H,D]
    self_cost: torch.Tensor            # [B,H]
    scene_cost: torch.Tensor           # [B,H,S]
    pose_cost: torch.Tensor            # [B,H,2L]
    cspace_cost: torch.Tensor          # [B,H,D]
    grad_vel: Optional[torch.Tensor] = None
    grad_acc: Optional[torch.Tensor] = None
    grad_jerk: Optional[torch.Tensor] = None
    link_pos: Optional[torch.Tensor] = None
    link_quat: Optional[torch.Tensor] = None
    robot_spheres: Optional[torch.Tensor] = None
    pose_goalset_idx: Optional[torch.Tensor] = None
    grad_knots: Optional[torch.Tensor] = None   # [B,n_knots,D] (evaluate_knots only)


def pack_robot_blob(rm: RobotModel) -> np.ndarray:
    """Robot constants -> one byte blob (layout: curobo_b200/csrc/cb200_blob.h), packed by the C helper."""
    L = _lib.load()
    ls = rm.link_spheres if rm.link_spheres.ndim == 3 else rm.link_spheres[None]     # [n_cfg, S, 4]
    sz = _lib.RobotSizes(rm.num_links, rm.num_dof, rm.num_spheres, rm.num_tool_frames, int(rm.collision_pairs.shape[0]),
                         int(ls.shape[0]))
    nbytes = L.cb200_robot_blob_bytes(C.byref(sz))
    if nbytes <= 0:
        raise ValueError(f"robot does not fit the blob format (code {nbytes}); links <= 64 required")
    out = np.zeros(nbytes, np.uint8)

    def p(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data
    keep = []
    n = L.cb200_pack_robot_blob(
        out.ctypes.data, nbytes, C.byref(sz), p(rm.fixed_transforms, np.float32), p(rm.link_map, np.int16),
        p(rm.joint_map, np.int16), p(rm.joint_map_type, np.int8), p(rm.joint_offset_map, np.float32),
        p(rm.tool_frame_map, np.int16), p(ls, np.float32), p(rm.link_sphere_idx_map, np.int16),
        p(rm.sphere_padding, np.float32), p(rm.collision_pairs, np.int16), p(rm.position_limits, np.float32),
        p(rm.velocity_limits, np.float32), p(rm.acceleration_limits, np.float32), p(rm.jerk_limits, np.float32),
        p(rm.effort_limits, np.float32))
    if n <= 0:
        raise ValueError(f"cb200_pack_robot_blob rejected the robot model (code {n})")
    return out[:n]


class RolloutEngine:
    def __init__(self, robot: RobotModel, cfg: RolloutConfig, device="cuda:0",
                 cuboid: Optional[CuboidData] = None, voxel: Optional[VoxelData] = None,
                 store_fk_outputs: bool = False, use_voxel_mip: bool = False):
        self.robot, self.cfg, self.device = robot, cfg, torch.device(device)
        _tc.require_cuda(self.device, "RolloutEngine is CUDA-only (sm_100a); there is no CPU path")
        self._lib = _lib.load()
        self._blob_host = pack_robot_blob(robot)
        self._blob = torch.from_numpy(self._blob_host.copy()).to(self.device)
        # several link-sphere configurations (attached objects per environment): rows pick theirs through env_query_idx
        self._sphere_cfgs = None
        if robot.link_spheres.ndim == 3 and robot.link_spheres.shape[0] > 1:
            self._sphere_cfgs = torch.from_numpy(np.ascontiguousarray(robot.link_spheres, np.float32)).to(self.device)
        self._cs_target = None
        self.cuboid, self.voxel = cuboid, voxel
        self.use_voxel_mip = use_voxel_mip
        self.refresh_world()
        self.store_fk_outputs = store_fk_outputs
        self._B = self._H = -1
        self._goal = None
        self._ccfg = self._make_ccfg(1)
        self._effort_cost = None
        self._dyn_params = None
        # row ticket counter of the humanoid kernel (zero between launches; this engine's launches are stream ordered)
        self._work_counter = torch.zeros(2, dtype=torch.int32, device=self.device)

    def attach_dynamics(self, dynamics, effort_limits=None, fused: bool = False) -> None:
        """Make the STATE c-space cost dynamics-aware: after the fused launch, tau = RNEA(q, qd, qdd) is
        evaluated for every row and the effort channel of the STATE cost -- bound hinge (cspace_weight[4], cspace_activation[4]),
        squared-L2 (cspace_reg[3]) and the energy term (cspace_reg[4]) -- is added to cost / cspace_cost, its gradient to
        grad_q / grad_vel / grad_acc through the RNEA adjoint (curobo_b200.dynamics.DynamicsStateCost: three more launches).
        The fused kernel evaluates those terms with tau = 0, where they vanish (what the reference does without
        `compute_inverse_dynamics`, transition/robot_state_transition.py:380-396), so nothing is counted twice.
        Applies to evaluate_action with vel / acc / jerk / dt given; `dynamics` is a curobo_b200.dynamics.Dynamics."""
        from .dynamics import DynamicsStateCost
        if self.cfg.cspace_type != "state":
            raise ValueError("attach_dynamics needs the STATE c-space cost")
        rm = self.robot
        # fused = the trajectory kernel evaluates RNEA, the effort terms and the RNEA adjoint for its own rows (swept mode;
        # effort limits = the robot's, which are part of the packed robot blob); otherwise -- and for custom limits or
        # discrete mode -- the same terms are added by three more launches after the fused one.  Default: the host
        # composition, measured faster on B200 (MPC 1024 x 30: 0.47 ms vs 0.52 ms in-kernel, plain kernel 0.34 ms).
        self._dyn_params = None
        if fused and effort_limits is None and self.cfg.use_sweep:
            m = dynamics._model       # (fixed, masses_com, inertias, joint types, joint map, link map, offsets, gravity, ...)
            self._dyn_keepalive = (m[1], m[2], m[7])
            self._dyn_params = _lib.DynamicsParams(m[1].data_ptr(), m[2].data_ptr(), m[7].data_ptr())
            self._effort_cost = None
            return
        w = [0.0, 0.0, 0.0, 0.0, float(self.cfg.cspace_weight[4])]
        reg = [0.0, 0.0, 0.0, float(self.cfg.cspace_reg[3]), float(self.cfg.cspace_reg[4])]
        lim = dict(p=rm.position_limits, v=rm.velocity_limits, a=rm.acceleration_limits, j=rm.jerk_limits,
                   tau=rm.effort_limits if effort_limits is None else effort_limits)
        self._effort_cost = DynamicsStateCost(dynamics, lim, w, list(self.cfg.cspace_activation), reg,
                                              retime_weights=self.cfg.retime_weights,
                                              retime_regularization_weights=self.cfg.retime_reg)

    def refresh_world(self) -> None:
        """Re-read the obstacle holders; call after the ESDF values (or obstacle tensors) were replaced or updated in
        place.  With `use_voxel_mip=True` (opt-in) it rebuilds the ESDF lower-bound pyramid level (one tiny launch) that
        lets discrete collision skip the corner fetches of samples that are provably inactive.  Exact while the level
        matches the grid: torch-side updates of `features` are detected (version stamp; the level is rebuilt on the next
        call), updates through raw pointers by foreign kernels are not -- call this after each of those."""
        if self.voxel is not None and self.use_voxel_mip:
            from .scene import build_voxel_mip
            build_voxel_mip(self.voxel)
        self._cs = c_cuboid_set(self.cuboid, self.device)
        self._vs = c_voxel_set(self.voxel, self.device)
        if self.voxel is not None and not self.use_voxel_mip and self._vs is not None:
            self._vs.mip, self._vs.mip_stride = None, 0

    # -- configuration ------------------------------------------------------------------------
    def _make_ccfg(self, num_goalset: int) -> _lib.RolloutCfg:
        c = self.cfg
        cc = _lib.RolloutCfg()
        cc.self_weight, cc.scene_weight, cc.scene_activation = c.self_weight, c.scene_weight, c.scene_activation
        cc.use_sweep, cc.use_speed_metric = int(c.use_sweep), int(c.use_speed_metric)
        pw = c.pose_weight if c.pose_weight is not None else (0.0, 0.0)
        cc.pose_weight[0], cc.pose_weight[1] = float(pw[0]), float(pw[1])
        cc.pose_rotation_method = 1 if c.pose_lie else 0
        cc.cspace_type = {None: 0, "position": 1, "state": 2}[c.cspace_type]
        for i in range(5):
            cc.cspace_weight[i] = float(c.cspace_weight[i]) if i < len(c.cspace_weight) else 0.0
            cc.cspace_activation[i] = float(c.cspace_activation[i]) if i < len(c.cspace_activation) else 0.0
            cc.cspace_reg[i] = float(c.cspace_reg[i]) if i < len(c.cspace_reg) else 0.0
        cc.retime_weights, cc.retime_regularization_weights = int(c.retime_weights), int(c.retime_reg)
        cc.num_goalset = num_goalset
        cc.cspace_target_weight = float(c.cspace_target_weight)
        cc.cspace_non_terminal_weight_factor = float(c.cspace_non_terminal_weight_factor)
        return cc

    def update_cspace_target(self, target: torch.Tensor, idxs_target: Optional[torch.Tensor] = None,
                             dof_weight: Optional[torch.Tensor] = None) -> None:
        """C-space target of the cost (`target_joint_position [n, D]`, `idxs_target_joint_position [B]` int32,
        `cspace_target_dof_weight [D]`; cost/cost_cspace_state.py, wp_cspace_state.py:84-89,220-226).  Read when
        cfg.cspace_target_weight > 0."""
        dev, D = self.device, self.robot.num_dof
        check_tensors(dev, torch.float32, cspace_target=target)
        if target.ndim != 2 or target.shape[1] != D:
            raise ValueError(f"cspace target must be [n, {D}], got {tuple(target.shape)}")
        if idxs_target is not None:
            check_tensors(dev, torch.int32, idxs_cspace_target=idxs_target)
        if dof_weight is not None:
            check_tensors(dev, torch.float32, cspace_target_dof_weight=dof_weight)
            if tuple(dof_weight.shape) != (D,):
                raise ValueError(f"cspace_target_dof_weight must be [{D}]")
        self._cs_target = (target, idxs_target, dof_weight)

    def setup_batch_tensors(self, batch: int, horizon: int) -> None:
        """Allocate every output once per (B, H) -- never inside evaluate_action."""
        rm, dev = self.robot, self.device
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
        D, S, Lt = rm.num_dof, rm.num_spheres, rm.num_tool_frames
        self.out = RolloutOutput(cost=z(batch, horizon), grad_q=z(batch, horizon, D), self_cost=z(batch, horizon),
                                 scene_cost=z(batch, horizon, S), pose_cost=z(batch, horizon, 2 * Lt),
                                 cspace_cost=z(batch, horizon, D))
        if self.cfg.cspace_type == "state":
            self.out.grad_vel, self.out.grad_acc, self.out.grad_jerk = (z(batch, horizon, D) for _ in range(3))
        if self.store_fk_outputs:
            self.out.link_pos, self.out.link_quat = z(batch, horizon, Lt, 3), z(batch, horizon, Lt, 4)
            self.out.robot_spheres = z(batch, horizon, S, 4)
            self.out.pose_goalset_idx = z(batch, horizon, Lt, dt=torch.int32)
        self._B, self._H = batch, horizon

    def update_goal(self, goal_position: torch.Tensor, goal_quat: torch.Tensor, idxs_goal: torch.Tensor,
                    terminal_axes=None, non_terminal_axes=None, terminal_tol=None, non_terminal_tol=None) -> None:
        """goal_* [G, L, n_goalset, 3|4] (quat wxyz), idxs_goal [B] int32 (GoalRegistry rows)."""
        dev = self.device
        check_tensors(dev, torch.float32, goal_position=goal_position, goal_quat=goal_quat)
        check_tensors(dev, torch.int32, idxs_goal=idxs_goal)
        Lt = self.robot.num_tool_frames
        if goal_position.ndim != 4 or goal_position.shape[1] != Lt or goal_quat.shape[:3] != goal_position.shape[:3]:
            raise ValueError("goal tensors must be [G, num_tool_frames, n_goalset, 3|4]")
        extra = {}
        for name, t, shape in (("terminal_axes", terminal_axes, (Lt, 6)), ("non_terminal_axes", non_terminal_axes, (Lt, 6)),
                               ("terminal_tol", terminal_tol, (Lt, 2)), ("non_terminal_tol", non_terminal_tol, (Lt, 2))):
            if t is not None:
                check_tensors(dev, torch.float32, **{name: t})
                if tuple(t.shape) != shape:
                    raise ValueError(f"{name} must have shape {shape}")
            extra[name] = t
        self._goal = (goal_position, goal_quat, idxs_goal, extra)
        self._ccfg = self._make_ccfg(int(goal_position.shape[2]))

    # -- the hot call --------------------------------------------------------------------------
    def evaluate_action(self, q: torch.Tensor, vel=None, acc=None, jerk=None, dt=None,
                        env_query_idx: Optional[torch.Tensor] = None) -> RolloutOutput:
        if q.ndim != 3 or q.shape[2] != self.robot.num_dof:
            raise ValueError(f"q must be [B, H, {self.robot.num_dof}], got {tuple(q.shape)}")
        B, H, _ = q.shape
        if (B, H) != (self._B, self._H):
            self.setup_batch_tensors(B, H)
        dev = self.device
        check_tensors(dev, torch.float32, q=q)
        io = _lib.RolloutIO()
        io.q = q.data_ptr()
        for name, t in (("vel", vel), ("acc", acc), ("jerk", jerk), ("dt", dt)):
            if t is not None:
                check_tensors(dev, torch.float32, **{name: t})
                setattr(io, name, t.data_ptr())
        out = self._launch(io, B, H, env_query_idx)
        if self._effort_cost is not None and vel is not None and acc is not None and jerk is not None and dt is not None:
            c, gp, gv, ga, _, _ = self._effort_cost.evaluate(q, vel, acc, jerk, dt)
            out.cspace_cost.add_(c)
            out.cost.add_(c.sum(-1))
            out.grad_q.add_(gp)
            out.grad_vel.add_(gv)
            out.grad_acc.add_(ga)
        return out

    def evaluate_knots(self, knots: torch.Tensor, start_state, start_state_idx: torch.Tensor, goal_state,
                       goal_state_idx: torch.Tensor, use_implicit_goal_state: torch.Tensor, bspline_degree: int = 4,
                       interpolation_steps: int = 4, env_query_idx: Optional[torch.Tensor] = None,
                       store_state: bool = False, in_kernel_spline: bool = False) -> RolloutOutput:
        """B-spline action space: knots [B,n_knots,D] -> row costs and d cost / d knots in ONE C
        call.  Default schedule: spline kernel -> rollout kernel -> adjoint kernel on the same stream (3 launches, the
        state makes one round trip through L2).  `in_kernel_spline=True`: the rollout kernel evaluates its rows from the
        knots itself (2 launches; the state never leaves the SM unless `store_state`); identical results, currently
        ~15 % slower at MPC scale (profiles/r01_d).  start_state / goal_state
        carry position, velocity, acceleration, jerk [n, D]; goal_state.dt [n_goal] is the trajectory dt
        (same contract as curobo_b200.trajectory.StateFromBSplineKnot.forward)."""
        D = self.robot.num_dof
        if knots.ndim != 3 or knots.shape[2] != D:
            raise ValueError(f"knots must be [B, n_knots, {D}], got {tuple(knots.shape)}")
        if bspline_degree not in (3, 4, 5):
            raise RuntimeError(f"Unsupported B-spline degree: {bspline_degree}")
        if self.cfg.cspace_type != "state":
            raise ValueError("evaluate_knots needs the STATE c-space cost (velocity / acceleration / jerk gradients)")
        if goal_state.dt is None:
            raise ValueError("dt is None")
        B, nk, _ = knots.shape
        H = (nk + bspline_degree + 1) * interpolation_steps + 1
        if not 1 <= interpolation_steps <= 32:
            raise RuntimeError("interpolation_steps must be in [1, 32]")
        dev = self.device
        if (B, H) != (self._B, self._H):
            self.setup_batch_tensors(B, H)
        o = self.out
        if o.grad_knots is None or tuple(o.grad_knots.shape) != (B, nk, D):
            o.grad_knots = torch.zeros((B, nk, D), dtype=torch.float32, device=dev)
        f32 = dict(knots=knots, traj_dt=goal_state.dt)
        for pre, st in (("start", start_state), ("goal", goal_state)):
            for fld in ("position", "velocity", "acceleration", "jerk"):
                f32[f"{pre}_{fld}"] = getattr(st, fld)
        check_tensors(dev, torch.float32, **f32)
        check_tensors(dev, torch.int32, start_state_idx=start_state_idx, goal_state_idx=goal_state_idx)
        check_tensors(dev, torch.uint8, use_implicit_goal_state=use_implicit_goal_state)
        if goal_state_idx.shape[0] != B or start_state_idx.shape[0] != B:
            raise ValueError("start_state_idx / goal_state_idx need one entry per batch row")
        sp = _lib.SplineInput()
        sp.knots = knots.data_ptr()
        for pre, st in (("start", start_state), ("goal", goal_state)):
            for fld in ("position", "velocity", "acceleration", "jerk"):
                setattr(sp, f"{pre}_{fld}", getattr(st, fld).data_ptr())
        sp.start_idx, sp.goal_idx = start_state_idx.data_ptr(), goal_state_idx.data_ptr()
        sp.traj_dt, sp.use_implicit_goal_state = goal_state.dt.data_ptr(), use_implicit_goal_state.data_ptr()
        sp.n_knots, sp.degree = nk, bspline_degree
        sp.grad_knots = o.grad_knots.data_ptr()
        if store_state or not in_kernel_spline:
            if getattr(self, "_state", None) is None or tuple(self._state[0].shape) != (B, H, D):
                self._state = tuple(torch.zeros((B, H, D), dtype=torch.float32, device=dev) for _ in range(4))
                self._state_dt = torch.zeros((B,), dtype=torch.float32, device=dev)
            sp.out_position, sp.out_velocity, sp.out_acceleration, sp.out_jerk = (t.data_ptr() for t in self._state)
            if not in_kernel_spline:
                sp.out_dt = self._state_dt.data_ptr()
        io = _lib.RolloutIO()
        io.spline = C.pointer(sp)
        # dynamics-aware cost with the expanded schedule: the trajectory kernel reads the spline states like caller-provided
        # ones, its RNEA-adjoint gradients flow into grad_knots through the spline adjoint behind it
        if self._dyn_params is not None and not in_kernel_spline:
            io.dynamics = C.pointer(self._dyn_params)
        if self._effort_cost is not None:
            raise ValueError("evaluate_knots supports the dynamics-aware cost only inside the kernel (attach_dynamics(fused=True), "
                             "swept mode, in_kernel_spline=False)")
        if self._dyn_params is not None and in_kernel_spline:
            raise ValueError("the dynamics-aware cost needs the expanded spline schedule (in_kernel_spline=False)")
        return self._launch(io, B, H, env_query_idx)

    def _launch(self, io, B: int, H: int, env_query_idx) -> RolloutOutput:
        dev = self.device
        o = self.out
        io.robot_blob, io.robot_blob_host = self._blob.data_ptr(), self._blob_host.ctypes.data
        io.robot_blob_bytes = int(self._blob_host.shape[0])
        if self._cs is not None:
            io.cuboids = C.pointer(self._cs)
        if self._vs is not None:
            io.voxels = C.pointer(self._vs)
        if self.voxel is not None and self.use_voxel_mip:
            from .scene import voxel_mip_is_fresh
            if not voxel_mip_is_fresh(self.voxel):      # the ESDF tensor was updated or replaced since the level was built
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("the ESDF changed since refresh_world(); call it before capturing a graph")
                self.refresh_world()
                io.voxels = C.pointer(self._vs)
        if env_query_idx is not None:
            check_tensors(dev, torch.int32, env_query_idx=env_query_idx)
            io.env_query_idx = env_query_idx.data_ptr()
        if self._goal is not None and self.cfg.pose_weight is not None:
            gp, gq, ig, extra = self._goal
            if ig.shape[0] != B:
                raise ValueError("idxs_goal must have one entry per batch row")
            io.goal_position, io.goal_quat, io.idxs_goal = gp.data_ptr(), gq.data_ptr(), ig.data_ptr()
            for cname, key in (("pose_axes_terminal", "terminal_axes"), ("pose_axes_non_terminal", "non_terminal_axes"),
                               ("pose_tol_terminal", "terminal_tol"), ("pose_tol_non_terminal", "non_terminal_tol")):
                if extra[key] is not None:
                    setattr(io, cname, extra[key].data_ptr())
        if self.cfg.cspace_target_weight > 0.0 and self.cfg.cspace_type is not None:
            if self._cs_target is None:
                raise ValueError("cspace_target_weight > 0 needs update_cspace_target(...) first")
            tgt, tidx, tdw = self._cs_target
            if tidx is not None and tidx.shape[0] != B:
                raise ValueError("idxs_cspace_target must have one entry per batch row")
            io.cspace_target = tgt.data_ptr()
            if tidx is not None:
                io.idxs_cspace_target = tidx.data_ptr()
            if tdw is not None:
                io.cspace_target_dof_weight = tdw.data_ptr()
        if self._sphere_cfgs is not None:
            io.sphere_configs, io.num_sphere_configs = self._sphere_cfgs.data_ptr(), int(self._sphere_cfgs.shape[0])
        io.cost, io.grad_q = o.cost.data_ptr(), o.grad_q.data_ptr()
        io.self_cost, io.scene_cost = o.self_cost.data_ptr(), o.scene_cost.data_ptr()
        io.pose_cost, io.cspace_cost = o.pose_cost.data_ptr(), o.cspace_cost.data_ptr()
        for name in ("grad_vel", "grad_acc", "grad_jerk", "link_pos", "link_quat", "robot_spheres", "pose_goalset_idx"):
            t = getattr(o, name)
            if t is not None:
                setattr(io, name, t.data_ptr())
        io.batch_size, io.horizon = B, H
        io.work_counter = self._work_counter.data_ptr()
        if self._dyn_params is not None and io.vel and io.acc and not bool(io.spline):
            io.dynamics = C.pointer(self._dyn_params)
        err = self._lib.cb200_rollout_cost_grad(C.byref(self._ccfg), C.byref(io), stream_ptr(dev))
        _lib.check(err, "rollout_cost_grad")
        return o


class FusedRolloutFunction(torch.autograd.Function):
    """cost[B] = sum_h rollout cost; backward returns the gradient computed in the same launch
    (the reference's own pattern: forward writes the gradient buffer, backward hands it out,
    cuda_ops/geometry.py:95-104, wp_autograd.py:103-110; the optimizer's upstream gradient is all-ones,
    gradient_opt_core.py:478).  Unlike the reference's Functions with use_grad_input=False the upstream gradient IS
    applied (per seed), so `cost.mean()` or a weighted sum differentiates correctly, and the gradient is saved as a
    copy: a second forward before backward does not disturb the first one's gradient."""

    @staticmethod
    def forward(ctx, q: torch.Tensor, engine: RolloutEngine):
        out = engine.evaluate_action(q.detach())
        ctx.save_for_backward(out.grad_q.clone())
        return out.cost.sum(dim=1)

    @staticmethod
    def backward(ctx, grad_cost):
        (g,) = ctx.saved_tensors
        return g * grad_cost.reshape(-1, 1, 1), None


class HostRolloutPipeline:
    """`evaluate_action` for HOST-resident inputs, the way an optimizer that keeps its iterate on the host (or another
    process feeding joint batches) drives the kernel: per step  pinned q -> H2D -> fused rollout -> D2H of cost + grad_q.

    Each slot owns an engine (= one set of output buffers), pinned host buffers and ONE captured CUDA graph holding the
    three stages, replayed on the slot's own stream: a step costs the host a single graph launch, and with two slots the
    upload of step i+1 overlaps the kernel of step i (stream order inside a slot makes buffer reuse safe).

        pipe = HostRolloutPipeline([eng_a, eng_b], B, H, dt=dt)
        pipe.slots[k].q_host[...] = ...      # fill the pinned input of slot k
        pipe.submit(k)                       # H2D + kernel + D2H, asynchronous
        cost, grad = pipe.result(k)          # waits for slot k; pinned host tensors
    """

    class Slot:
        def __init__(self, engine, B, H, D, device):
            self.engine = engine
            self.stream = torch.cuda.Stream(device)
            self.q_host = torch.empty((B, H, D), dtype=torch.float32).pin_memory()
            self.cost_host = torch.empty((B, H), dtype=torch.float32).pin_memory()
            self.grad_host = torch.empty((B, H, D), dtype=torch.float32).pin_memory()
            self.q_dev = torch.empty((B, H, D), dtype=torch.float32, device=device)
            self.graph = None

    def __init__(self, engines: Sequence[RolloutEngine], batch: int, horizon: int, **eval_kwargs):
        if not engines:
            raise ValueError("at least one engine")
        self.device = engines[0].device
        D = engines[0].robot.num_dof
        self._kw = eval_kwargs
        self.slots = [HostRolloutPipeline.Slot(e, batch, horizon, D, self.device) for e in engines]
        self.h2d_bytes = int(self.slots[0].q_host.numel() * 4)
        self.d2h_bytes = int((self.slots[0].cost_host.numel() + self.slots[0].grad_host.numel()) * 4)
        for s in self.slots:
            s.q_host.zero_()
            self._stages(s)                                    # eager once: output allocation, launch-plan caches
        torch.cuda.synchronize(self.device)
        for s in self.slots:
            s.graph = torch.cuda.CUDAGraph()
            with torch.cuda.device(self.device), torch.cuda.graph(s.graph, stream=s.stream):
                self._stages(s)

    def _stages(self, s) -> None:
        s.q_dev.copy_(s.q_host, non_blocking=True)
        out = s.engine.evaluate_action(s.q_dev, **self._kw)
        s.cost_host.copy_(out.cost, non_blocking=True)
        s.grad_host.copy_(out.grad_q, non_blocking=True)

    def submit(self, k: int) -> None:
        s = self.slots[k]
        with torch.cuda.stream(s.stream):
            s.graph.replay()

    def result(self, k: int):
        s = self.slots[k]
        s.stream.synchronize()
        return s.cost_host, s.grad_host

    def wait_all(self) -> None:
        for s in self.slots:
            s.stream.synchronize()
