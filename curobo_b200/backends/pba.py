"""Drop-in for curobo._src.curobolib.backends.cuda_core_backend.pba: same function name and
argument order as `launch_pba3d` (pba.py:60-124), forwarding to the C ABI on the current stream."""
from __future__ import annotations

import torch

from .. import lib as _lib
from .tensor_checks import check_tensors, stream_ptr


def launch_pba3d(site_index: torch.Tensor, buffer: torch.Tensor, nx: int, ny: int, nz: int, m3: int = 2) -> None:
    """Exact 3-D nearest-site transform of `site_index` in place ((nx*ny*nz,) or (nx, ny, nz) int32; sites >= 0 hold
    their packed coordinates, non-sites < 0).  `buffer` (same size, int32) is the reference's scratch: checked like the
    reference checks it, not used.  `m3` is accepted for signature parity."""
    dev = site_index.device
    check_tensors(dev, torch.int32, site_index=site_index, buffer=buffer)
    n = int(nx) * int(ny) * int(nz)
    if site_index.numel() != n or buffer.numel() != n:
        raise ValueError(f"site_index / buffer must hold nx*ny*nz = {n} elements, got {site_index.numel()} / {buffer.numel()}")
    err = _lib.load().cb200_pba3d(site_index.data_ptr(), buffer.data_ptr(), int(nx), int(ny), int(nz), int(m3), stream_ptr(dev))
    _lib.check(err, "pba3d")


def launch_edt_unsigned_distance(site_index: torch.Tensor, distance: torch.Tensor, nx: int, ny: int, nz: int,
                                 voxel_size: float, empty_value: float = 1e4) -> None:
    """distance (float16, nx*ny*nz) = |voxel - nearest site| * voxel_size; `empty_value` where the grid has no site."""
    dev = site_index.device
    check_tensors(dev, torch.int32, site_index=site_index)
    check_tensors(dev, torch.float16, distance=distance)
    n = int(nx) * int(ny) * int(nz)
    if site_index.numel() != n or distance.numel() != n:
        raise ValueError(f"site_index / distance must hold nx*ny*nz = {n} elements")
    err = _lib.load().cb200_edt_unsigned_distance(site_index.data_ptr(), distance.data_ptr(), int(nx), int(ny), int(nz),
                                                  float(voxel_size), float(empty_value), stream_ptr(dev))
    _lib.check(err, "edt_unsigned_distance")
