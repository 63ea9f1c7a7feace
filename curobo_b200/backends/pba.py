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


def launch_esdf_seed_sites(combined_sdf: torch.Tensor, site_index: torch.Tensor, nx: int, ny: int, nz: int, voxel_size: float,
                           truncation_distance: float) -> None:
    """site_index (int32, nx*ny*nz) from a dense SDF (float32; > 1e9 = unobserved): surface / truncation-boundary voxels become
    sites (builder_esdf.py:255-261), everything else -1."""
    dev = site_index.device
    check_tensors(dev, torch.float32, combined_sdf=combined_sdf)
    check_tensors(dev, torch.int32, site_index=site_index)
    n = int(nx) * int(ny) * int(nz)
    if site_index.numel() != n or combined_sdf.numel() != n:
        raise ValueError(f"combined_sdf / site_index must hold nx*ny*nz = {n} elements")
    err = _lib.load().cb200_esdf_seed_sites(combined_sdf.data_ptr(), site_index.data_ptr(), int(nx), int(ny), int(nz),
                                            float(voxel_size), float(truncation_distance), stream_ptr(dev))
    _lib.check(err, "esdf_seed_sites")


def launch_esdf_signed_distance(site_index: torch.Tensor, static_sdf, combined_sdf, distance: torch.Tensor, nx: int, ny: int,
                                nz: int, voxel_size: float, adjacent_skip_steps: float = 1.0) -> None:
    """distance (float16) = +-|voxel - nearest site| * voxel_size, signed like compute_esdf_from_min_tsdf_kernel
    (builder_esdf.py:410-503) from dense static / combined SDFs (float32, > 1e9 = unobserved; either may be None)."""
    dev = site_index.device
    check_tensors(dev, torch.int32, site_index=site_index)
    check_tensors(dev, torch.float16, distance=distance)
    n = int(nx) * int(ny) * int(nz)
    ptrs = []
    for name, t in (("static_sdf", static_sdf), ("combined_sdf", combined_sdf)):
        if t is not None:
            check_tensors(dev, torch.float32, **{name: t})
            if t.numel() != n:
                raise ValueError(f"{name} must hold nx*ny*nz = {n} elements")
        ptrs.append(t.data_ptr() if t is not None else None)
    if site_index.numel() != n or distance.numel() != n:
        raise ValueError(f"site_index / distance must hold nx*ny*nz = {n} elements")
    err = _lib.load().cb200_esdf_signed_distance(site_index.data_ptr(), ptrs[0], ptrs[1], distance.data_ptr(), int(nx), int(ny),
                                                 int(nz), float(voxel_size), float(adjacent_skip_steps), stream_ptr(dev))
    _lib.check(err, "esdf_signed_distance")
