"""`optimization` kernel-backend module: same function names and argument order as
curobo/_src/curobolib/backends/cuda_core_backend/optimization.py:26-246 (pybind twins:
backends/pybind/line_search_kernel_launch.cu, lbfgs_step_kernel_launch.cu), SURVEY.md section 8f rank 2.
Tensors are validated (device, contiguity, dtype) BEFORE launch; errors raise; launches go to the current stream.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import lib as _lib
from .tensor_checks import check_tensors, stream_ptr


def launch_line_search(
    best_cost: torch.Tensor,
    best_action: torch.Tensor,
    best_iteration: torch.Tensor,
    current_iteration: torch.Tensor,
    converged_global: torch.Tensor,
    convergence_iteration: int,
    cost_delta_threshold: float,
    cost_relative_threshold: float,
    exploration_cost: torch.Tensor,
    exploration_action: torch.Tensor,
    exploration_gradient: torch.Tensor,
    exploration_idx: torch.Tensor,
    selected_cost: torch.Tensor,
    selected_action: torch.Tensor,
    selected_gradient: torch.Tensor,
    selected_idx: torch.Tensor,
    search_cost: torch.Tensor,
    search_action: torch.Tensor,
    search_gradient: torch.Tensor,
    step_direction: torch.Tensor,
    search_magnitudes: torch.Tensor,
    armijo_threshold_c_1: float,
    curvature_threshold_c_2: float,
    strong_wolfe: bool,
    approx_wolfe: bool,
    n_linesearch: int,
    opt_dim: int,
    batchsize: int,
) -> None:
    """Parallel Wolfe line search + best / convergence bookkeeping; every output is written in place."""
    if n_linesearch > 32:
        raise RuntimeError("n_linesearch greater than 32 is not supported")
    if opt_dim > 1024:
        raise RuntimeError("opt_dim greater than 1024 is not supported")
    dev = search_cost.device
    check_tensors(dev, torch.float32, best_cost=best_cost, best_action=best_action, exploration_cost=exploration_cost,
                  exploration_action=exploration_action, exploration_gradient=exploration_gradient,
                  selected_cost=selected_cost, selected_action=selected_action, selected_gradient=selected_gradient,
                  search_cost=search_cost, search_action=search_action, search_gradient=search_gradient,
                  step_direction=step_direction, search_magnitudes=search_magnitudes)
    check_tensors(dev, torch.int16, best_iteration=best_iteration, current_iteration=current_iteration)
    check_tensors(dev, torch.uint8, converged_global=converged_global)
    check_tensors(dev, torch.int32, exploration_idx=exploration_idx, selected_idx=selected_idx)
    L = _lib.load()
    err = L.cb200_line_search(
        best_cost.data_ptr(), best_action.data_ptr(), best_iteration.data_ptr(), current_iteration.data_ptr(),
        converged_global.data_ptr(), int(convergence_iteration), float(cost_delta_threshold),
        float(cost_relative_threshold), exploration_cost.data_ptr(), exploration_action.data_ptr(),
        exploration_gradient.data_ptr(), exploration_idx.data_ptr(), selected_cost.data_ptr(),
        selected_action.data_ptr(), selected_gradient.data_ptr(), selected_idx.data_ptr(), search_cost.data_ptr(),
        search_action.data_ptr(), search_gradient.data_ptr(), step_direction.data_ptr(), search_magnitudes.data_ptr(),
        float(armijo_threshold_c_1), float(curvature_threshold_c_2), int(bool(strong_wolfe)), int(bool(approx_wolfe)),
        int(n_linesearch), int(opt_dim), int(batchsize), stream_ptr(dev))
    _lib.check(err, "launch_line_search")


def launch_lbfgs_step(
    step_vec: torch.Tensor,
    rho_buffer: torch.Tensor,
    y_buffer: torch.Tensor,
    s_buffer: torch.Tensor,
    q: torch.Tensor,
    grad_q: torch.Tensor,
    x_0: torch.Tensor,
    grad_0: torch.Tensor,
    epsilon: float,
    batch_size: int,
    history_m: int,
    v_dim: int,
    stable_mode: bool,
    use_shared_buffers: bool = True,
    x_set: Optional[torch.Tensor] = None,
    step_scaled: Optional[torch.Tensor] = None,
    search_magnitudes: Optional[torch.Tensor] = None,
    action_step_max: Optional[torch.Tensor] = None,
    fix_terminal_action: bool = False,
    action_dim: int = 0,
) -> List[torch.Tensor]:
    """L-BFGS two-loop step + history roll, in place.  Returns [step_vec, rho_buffer, y_buffer, s_buffer, x_0, grad_0]
    like the reference.  `use_shared_buffers` is accepted for signature parity (the history is always staged on
    chip for v_dim <= 32).  The five keyword arguments after it are this library's optional fused line-search
    set-up (include/curobo_b200.h): x_set [B,n,V] = q + magnitudes * scale_action(step)."""
    if history_m > 31:
        raise RuntimeError("History_m greater than 31 is not supported")  # optimization.py:173-174
    if history_m < 0:
        raise RuntimeError("History_m less than 0 is not supported")
    dev = step_vec.device
    check_tensors(dev, torch.float32, step_vec=step_vec, rho_buffer=rho_buffer, y_buffer=y_buffer, s_buffer=s_buffer, q=q,
                  grad_q=grad_q, x_0=x_0, grad_0=grad_0)
    n_ls, adim = 0, 0
    if x_set is not None:
        check_tensors(dev, torch.float32, x_set=x_set, search_magnitudes=search_magnitudes)
        n_ls = int(search_magnitudes.numel())
        if step_scaled is not None:
            check_tensors(dev, torch.float32, step_scaled=step_scaled)
        if action_step_max is not None:
            check_tensors(dev, torch.float32, action_step_max=action_step_max)
            adim = int(action_step_max.numel())
        if action_dim > 0:   # explicit: the terminal action is frozen whether or not the step is clamped
            if action_step_max is not None and int(action_step_max.numel()) != int(action_dim):
                raise ValueError("action_step_max must have action_dim entries")
            adim = int(action_dim)
    L = _lib.load()
    p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    err = L.cb200_lbfgs_step(
        step_vec.data_ptr(), rho_buffer.data_ptr(), y_buffer.data_ptr(), s_buffer.data_ptr(), q.data_ptr(),
        grad_q.data_ptr(), x_0.data_ptr(), grad_0.data_ptr(), float(epsilon), int(batch_size), int(history_m), int(v_dim),
        int(bool(stable_mode)), p(x_set), p(step_scaled), p(search_magnitudes), n_ls, p(action_step_max), adim,
        int(bool(fix_terminal_action)), stream_ptr(dev))
    _lib.check(err, "launch_lbfgs_step")
    return [step_vec, rho_buffer, y_buffer, s_buffer, x_0, grad_0]
