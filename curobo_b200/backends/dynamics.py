"""`dynamics` kernel-backend module: same function names and argument order as
curobo/_src/curobolib/backends/cuda_core_backend/dynamics.py:24-250.
Tensors are validated (device, contiguity, dtype) BEFORE launch; errors raise; launches go to the current stream."""
from __future__ import annotations

from typing import Optional

import torch

from .. import lib as _lib
from .tensor_checks import check_tensors, stream_ptr


def _check_model(dev, fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map, link_map, joint_offset_map,
                 gravity, level_starts, level_links):
    check_tensors(dev, torch.float32, fixed_transforms=fixed_transforms, link_masses_com=link_masses_com,
                  link_inertias=link_inertias, joint_offset_map=joint_offset_map, gravity=gravity)
    check_tensors(dev, torch.int8, joint_map_type=joint_map_type)
    check_tensors(dev, torch.int16, joint_map=joint_map, link_map=link_map, level_starts=level_starts, level_links=level_links)


def launch_rnea_forward(
    tau: torch.Tensor,
    q: torch.Tensor,
    qd: torch.Tensor,
    qdd: torch.Tensor,
    fixed_transforms: torch.Tensor,
    link_masses_com: torch.Tensor,
    link_inertias: torch.Tensor,
    joint_map_type: torch.Tensor,
    joint_map: torch.Tensor,
    link_map: torch.Tensor,
    joint_offset_map: torch.Tensor,
    gravity: torch.Tensor,
    level_starts: torch.Tensor,
    level_links: torch.Tensor,
    forward_cache: torch.Tensor,
    batch_size: int,
    num_links: int,
    num_dof: int,
    n_levels: int,
    threads_per_batch: int = 1,
    f_ext: Optional[torch.Tensor] = None,
) -> None:
    """tau = RNEA(q, qd, qdd [, f_ext]); fills forward_cache [B, num_links * 20] for the adjoint.  `threads_per_batch` is
    accepted for signature parity and ignored (rows are processed one thread each, deterministic sums)."""
    dev = q.device
    check_tensors(dev, torch.float32, tau=tau, q=q, qd=qd, qdd=qdd, forward_cache=forward_cache)
    _check_model(dev, fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map, link_map, joint_offset_map,
                 gravity, level_starts, level_links)
    if f_ext is not None:
        check_tensors(dev, torch.float32, f_ext=f_ext)
    L = _lib.load()
    err = L.cb200_rnea_forward(
        tau.data_ptr(), q.data_ptr(), qd.data_ptr(), qdd.data_ptr(), fixed_transforms.data_ptr(), link_masses_com.data_ptr(),
        link_inertias.data_ptr(), joint_map_type.data_ptr(), joint_map.data_ptr(), link_map.data_ptr(),
        joint_offset_map.data_ptr(), gravity.data_ptr(), level_starts.data_ptr(), level_links.data_ptr(),
        forward_cache.data_ptr(), int(batch_size), int(num_links), int(num_dof), int(n_levels),
        f_ext.data_ptr() if f_ext is not None else None, stream_ptr(dev))
    _lib.check(err, "launch_rnea_forward")


def launch_rnea_backward(
    grad_q: torch.Tensor,
    grad_qd: torch.Tensor,
    grad_qdd: torch.Tensor,
    grad_tau: torch.Tensor,
    q: torch.Tensor,
    qd: torch.Tensor,
    fixed_transforms: torch.Tensor,
    link_masses_com: torch.Tensor,
    link_inertias: torch.Tensor,
    joint_map_type: torch.Tensor,
    joint_map: torch.Tensor,
    link_map: torch.Tensor,
    joint_offset_map: torch.Tensor,
    gravity: torch.Tensor,
    level_starts: torch.Tensor,
    level_links: torch.Tensor,
    forward_cache: torch.Tensor,
    batch_size: int,
    num_links: int,
    num_dof: int,
    n_levels: int,
    threads_per_batch: int = 1,
    grad_f_ext: Optional[torch.Tensor] = None,
) -> None:
    """(grad_q, grad_qd, grad_qdd) from grad_tau and the forward cache; the outputs are overwritten."""
    dev = q.device
    check_tensors(dev, torch.float32, grad_q=grad_q, grad_qd=grad_qd, grad_qdd=grad_qdd, grad_tau=grad_tau, q=q, qd=qd,
                  forward_cache=forward_cache)
    _check_model(dev, fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map, link_map, joint_offset_map,
                 gravity, level_starts, level_links)
    if grad_f_ext is not None:
        check_tensors(dev, torch.float32, grad_f_ext=grad_f_ext)
    L = _lib.load()
    err = L.cb200_rnea_backward(
        grad_q.data_ptr(), grad_qd.data_ptr(), grad_qdd.data_ptr(), grad_tau.data_ptr(), q.data_ptr(), qd.data_ptr(),
        fixed_transforms.data_ptr(), link_masses_com.data_ptr(), link_inertias.data_ptr(), joint_map_type.data_ptr(),
        joint_map.data_ptr(), link_map.data_ptr(), joint_offset_map.data_ptr(), gravity.data_ptr(), level_starts.data_ptr(),
        level_links.data_ptr(), forward_cache.data_ptr(), int(batch_size), int(num_links), int(num_dof), int(n_levels),
        grad_f_ext.data_ptr() if grad_f_ext is not None else None, stream_ptr(dev))
    _lib.check(err, "launch_rnea_backward")
