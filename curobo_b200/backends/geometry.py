"""`geometry` kernel-backend module: same name and argument order as
curobo/_src/curobolib/backends/cuda_core_backend/geometry.py:63-82 (pybind twin
backends/pybind/geometry_bindings.cpp:16-34), so `SelfCollisionDistance`
(curobo/_src/curobolib/cuda_ops/geometry.py:18-104) works unchanged on top of it."""
from __future__ import annotations

import torch

from .. import lib as _lib
from .tensor_checks import check_tensors, stream_ptr


def self_collision_distance(
    out_distance: torch.Tensor,
    out_vec: torch.Tensor,
    pair_distance: torch.Tensor,
    sparse_index: torch.Tensor,
    robot_spheres: torch.Tensor,
    sphere_padding: torch.Tensor,
    weight: torch.Tensor,
    pair_locations: torch.Tensor,
    block_batch_max_value: torch.Tensor,
    block_batch_max_index: torch.Tensor,
    num_blocks_per_batch: int,
    max_threads_per_block: int,
    batch_size: int,
    horizon: int,
    nspheres: int,
    num_collision_pairs: int,
    store_pair_distance: bool,
    compute_grad: bool,
) -> None:
    dev = robot_spheres.device
    check_tensors(dev, torch.float32, out_distance=out_distance, out_vec=out_vec, robot_spheres=robot_spheres,
                  sphere_padding=sphere_padding, weight=weight)
    check_tensors(dev, torch.uint8, sparse_index=sparse_index)
    check_tensors(dev, torch.int16, pair_locations=pair_locations)
    if store_pair_distance:
        check_tensors(dev, torch.float32, pair_distance=pair_distance)

    def p(t):
        return t.data_ptr() if t is not None else None
    L = _lib.load()
    err = L.cb200_self_collision_distance(
        out_distance.data_ptr(), out_vec.data_ptr(), p(pair_distance), sparse_index.data_ptr(),
        robot_spheres.data_ptr(), sphere_padding.data_ptr(), weight.data_ptr(), pair_locations.data_ptr(),
        p(block_batch_max_value), p(block_batch_max_index), int(num_blocks_per_batch), int(max_threads_per_block),
        int(batch_size), int(horizon), int(nspheres), int(num_collision_pairs), int(bool(store_pair_distance)),
        int(bool(compute_grad)), stream_ptr(dev))
    _lib.check(err, "self_collision_distance")
