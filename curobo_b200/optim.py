"""Host mirror of the reference's L-BFGS optimizer iteration, built on the fused
rollout: one optimizer iteration = 3 kernel launches

    cb200_lbfgs_step (+ search points)  ->  cb200_rollout_cost_grad on [B * n_linesearch] rows  ->  cb200_line_search

where the reference runs ~30 (gradient_opt_core.py:334-400: line search strategy -> rollout 15-25 launches ->
wolfe kernel -> LBFGS kernel + torch glue).

  LBFGScu                 <- curobo/_src/curobolib/cuda_ops/optimization.py:192-252 (same positional arguments)
  wolfe_line_search       <- cuda_ops/optimization.py:22-189 (flat-tensor form of the same launch)
  QuasiNewtonBuffers      <- optim/components/quasi_newton_buffers.py:20-130
  LBFGSOpt.optimize       <- optim/gradient/lbfgs.py:157-240 + optim/components/gradient_opt_core.py:290-400 with
                             line_search_type approx_wolfe, CUDA-kernel step direction and line search
                             (content/configs/task/ik/lbfgs_ik.yml)

CUDA only; no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Callable, List, Optional, Tuple

import torch

from .backends import optimization as optimization_cu
from .backends import tensor_checks as _tc


class LBFGScu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step_vec, rho_buffer, y_buffer, s_buffer, q, grad_q, x_0, grad_0, epsilon=0.1, stable_mode=False,
                use_shared_buffers=True):
        m, b, v_dim, _ = y_buffer.shape
        R = optimization_cu.launch_lbfgs_step(step_vec, rho_buffer, y_buffer, s_buffer, q, grad_q, x_0, grad_0, epsilon, b, m,
                                              v_dim, stable_mode, use_shared_buffers)
        return R[0].view(step_vec.shape)

    @staticmethod
    def backward(ctx, grad_output):
        return (None,) * 11


@dataclass
class QuasiNewtonBuffers:
    """(s, y, rho) history + reference point, shapes as in the reference ([m,B,V,1], [m,B,1,1], [B,V,1])."""
    history: int
    device: torch.device
    s: Optional[torch.Tensor] = None
    y: Optional[torch.Tensor] = None
    rho: Optional[torch.Tensor] = None
    x_0: Optional[torch.Tensor] = None
    grad_0: Optional[torch.Tensor] = None
    step_q_buffer: Optional[torch.Tensor] = None

    def resize(self, num_problems: int, opt_dim: int) -> None:
        z = lambda *s: torch.zeros(s, device=self.device, dtype=torch.float32)  # noqa: E731
        b = num_problems
        self.x_0, self.grad_0 = z(b, opt_dim, 1), z(b, opt_dim, 1)
        self.y, self.s = z(self.history, b, opt_dim, 1), z(self.history, b, opt_dim, 1)
        self.rho = z(self.history, b, 1, 1)
        self.step_q_buffer = z(b, opt_dim)

    def clear(self) -> None:
        for t in (self.s, self.y, self.rho, self.step_q_buffer):
            t.fill_(0.0)

    def set_reference(self, x: torch.Tensor, grad: torch.Tensor) -> None:
        self.x_0.copy_(x.view_as(self.x_0))
        self.grad_0.copy_(grad.view_as(self.grad_0))


@dataclass
class LBFGSOptCfg:
    """The fields of optim/gradient/lbfgs.py:38-86 this loop uses (defaults = content/configs/task/ik/lbfgs_ik.yml)."""
    num_iters: int = 100
    history: int = 7
    epsilon: float = 0.01
    stable_mode: bool = True
    line_search_scale: List[float] = field(default_factory=lambda: [0.0, 0.1, 0.5, 1.0])
    line_search_wolfe_c_1: float = 1e-5
    line_search_wolfe_c_2: float = 0.9
    strong_wolfe: bool = False
    approx_wolfe: bool = True
    step_scale: float = 0.98
    fix_terminal_action: bool = False
    cost_delta_threshold: float = 0.0
    cost_relative_threshold: float = 0.0
    convergence_iteration: int = 10
    initial_step_scale: float = 0.001


class LBFGSOpt:
    """Batched L-BFGS over `num_problems` independent problems of dimension action_horizon * action_dim.

    `cost_grad_fn(x_set)` evaluates x_set [B * n_linesearch, opt_dim] and returns (cost [B * n], grad [B * n, opt_dim])
    -- with the fused rollout that is one kernel launch (see IKSolver below)."""

    def __init__(self, cfg: LBFGSOptCfg, num_problems: int, action_horizon: int, action_dim: int,
                 action_bound_lows: torch.Tensor, action_bound_highs: torch.Tensor,
                 cost_grad_fn: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], device="cuda:0"):
        cfg = replace(cfg, line_search_scale=list(cfg.line_search_scale))   # never mutate the caller's config
        self.cfg, self.device = cfg, torch.device(device)
        _tc.require_cuda(self.device, "LBFGSOpt is CUDA-only")
        self.B, self.H, self.D = num_problems, action_horizon, action_dim
        self.V = action_horizon * action_dim
        if cfg.history > self.V:
            cfg.history = self.V  # lbfgs.py:186-188
        if self.V > 1024 or cfg.history > 31:
            raise ValueError("opt_dim > 1024 or history > 31 is not supported by the step kernel")
        self.n = len(cfg.line_search_scale)
        dev, B, V, n = self.device, self.B, self.V, self.n
        z = lambda *s, dt=torch.float32: torch.zeros(s, device=dev, dtype=dt)  # noqa: E731
        self.qn = QuasiNewtonBuffers(cfg.history, dev)
        self.qn.resize(B, V)
        self.magnitudes = torch.tensor(cfg.line_search_scale, device=dev, dtype=torch.float32)
        # optim/components/action_bounds.py:31: step_max = step_scale * |high - low|
        self.step_max = (cfg.step_scale * (action_bound_highs - action_bound_lows).abs()).to(dev, torch.float32).contiguous()
        self.clamp_step = cfg.step_scale not in (0.0, 1.0)
        self.cost_grad_fn = cost_grad_fn
        self.x_set, self.step_scaled = z(B, n, V), z(B, V)
        self.best_cost, self.best_action = z(B), z(B, V)
        self.best_iteration, self.current_iteration = z(B, dt=torch.int16), z(B, dt=torch.int16)
        self.converged = z(B, dt=torch.uint8)
        self.exploration_cost, self.exploration_action, self.exploration_gradient = z(B), z(B, V), z(B, V)
        self.cost, self.action, self.gradient = z(B), z(B, V), z(B, V)
        self.exploration_idx, self.selected_idx = z(B, n, dt=torch.int32), z(B, n, dt=torch.int32)

    def reset(self, x0: torch.Tensor) -> None:
        """Initial evaluation (gradient_opt_core.py:400-470): cost/grad at x0, best = x0, first step = -initial_step_scale * grad."""
        B, V, n = self.B, self.V, self.n
        self.qn.clear()
        self.current_iteration.zero_()
        self.best_iteration.zero_()
        self.converged.zero_()
        x0 = x0.reshape(B, V).contiguous()
        self.x_set.copy_(x0[:, None, :].expand(B, n, V))
        c, g = self.cost_grad_fn(self.x_set.view(B * n, V))
        c, g = c.view(B, n), g.view(B, n, V)
        self.exploration_action.copy_(x0)
        self.exploration_gradient.copy_(g[:, 0])
        self.exploration_cost.copy_(c[:, 0])
        self.action.copy_(x0)
        self.gradient.copy_(g[:, 0])
        self.cost.copy_(c[:, 0])
        self.best_cost.copy_(c[:, 0])
        self.best_action.copy_(x0)
        self.qn.set_reference(x0, g[:, 0])
        self._first = True

    def _step_direction(self) -> None:
        cfg, qn = self.cfg, self.qn
        if self._first:
            # no curvature pair yet: steepest descent scaled by initial_step_scale, through the same search-point set-up
            self._first = False
            step = (-cfg.initial_step_scale * self.exploration_gradient)
            if self.clamp_step:
                ratio = (step.view(self.B, self.H, self.D).abs() / self.step_max.view(1, 1, -1)).reshape(self.B, -1).amax(dim=1)
                step = step / ratio.clamp(min=1.0)[:, None]
            if cfg.fix_terminal_action and self.H > 1:
                step.view(self.B, self.H, self.D)[:, -1] = 0.0
            self.step_scaled.copy_(step)
            self.x_set.copy_(self.exploration_action[:, None, :] + self.magnitudes.view(1, -1, 1) * step[:, None, :])
            return
        optimization_cu.launch_lbfgs_step(
            qn.step_q_buffer, qn.rho, qn.y, qn.s, self.exploration_action, self.exploration_gradient, qn.x_0, qn.grad_0,
            cfg.epsilon, self.B, cfg.history, self.V, cfg.stable_mode, True, x_set=self.x_set, step_scaled=self.step_scaled,
            search_magnitudes=self.magnitudes, action_step_max=self.step_max if self.clamp_step else None,
            fix_terminal_action=cfg.fix_terminal_action, action_dim=self.D)

    def step(self) -> None:
        """One optimizer iteration: step direction + search points, rollout, line search."""
        cfg, B, V, n = self.cfg, self.B, self.V, self.n
        self._step_direction()
        c, g = self.cost_grad_fn(self.x_set.view(B * n, V))
        optimization_cu.launch_line_search(
            self.best_cost, self.best_action, self.best_iteration, self.current_iteration, self.converged,
            cfg.convergence_iteration, cfg.cost_delta_threshold, cfg.cost_relative_threshold, self.exploration_cost,
            self.exploration_action, self.exploration_gradient, self.exploration_idx.view(-1), self.cost, self.action,
            self.gradient, self.selected_idx.view(-1), c.view(B, n).contiguous(), self.x_set, g.view(B, n, V).contiguous(),
            self.step_scaled, self.magnitudes, cfg.line_search_wolfe_c_1, cfg.line_search_wolfe_c_2, cfg.strong_wolfe,
            cfg.approx_wolfe, n, V, B)

    def optimize(self, x0: torch.Tensor, num_iters: Optional[int] = None) -> torch.Tensor:
        self.reset(x0)
        for _ in range(num_iters if num_iters is not None else self.cfg.num_iters):
            self.step()
        return self.best_action.view(self.B, self.H, self.D)

    def optimize_graphed(self, x0: torch.Tensor, num_iters: Optional[int] = None) -> torch.Tensor:
        """The whole solve -- initial evaluation + `num_iters` x (step direction, rollout, line search) -- as ONE CUDA-graph
        launch ("whole _opt_iters as one launch", gradient_opt_core.py:334-400, which the reference
        also replays from a graph, one graph per iteration block).  The first call runs one eager solve (allocations, launch-plan
        caches), captures the loop on a side stream and replays it; later calls copy x0 into the captured input and replay.
        Every buffer the loop touches is owned by this object or by the cost function's engine, so the replay is allocation
        free; `cost_grad_fn` must launch on the current stream and must not synchronise (RolloutEngine does neither)."""
        n_it = num_iters if num_iters is not None else self.cfg.num_iters
        g = getattr(self, "_graph", None)
        if g is None or self._graph_iters != n_it:
            self._graph_x0 = torch.empty((self.B, self.V), device=self.device, dtype=torch.float32)
            self._graph_x0.copy_(x0.reshape(self.B, self.V))
            self.optimize(self._graph_x0, n_it)                          # eager warm-up, same buffers
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.device(self.device), torch.cuda.graph(g):
                self.optimize(self._graph_x0, n_it)
            self._graph, self._graph_iters = g, n_it
        self._graph_x0.copy_(x0.reshape(self.B, self.V))
        self._graph.replay()
        return self.best_action.view(self.B, self.H, self.D)
