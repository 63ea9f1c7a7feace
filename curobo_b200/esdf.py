"""Host mirror of the reference's exact Euclidean distance transform operator.

  ParallelBandingEDT  <- curobo/_src/perception/mapper/esdf/edt_parallel_banding.py:21-80 (same constructor and
                         `propagate(site_index)`; the class keeps the reference's name although the kernels underneath are a
                         three-pass in-place transform, not PBA+)
  seed_sites_from_occupancy / unsigned_distance: the steps either side of `propagate` in
                         BlockSparseESDFIntegrator._compute_esdf_impl (integrator_esdf.py:692-704) for a dense occupancy
                         grid -- the reference seeds from its block-sparse TSDF and signs the distance with it; both need
                         the TSDF hash, which is out of scope.
  DenseESDFBuilder    <- BlockSparseESDFIntegrator._compute_esdf_impl (integrator_esdf.py:640-704: seed -> propagate -> signed
                         distance) for a DENSE signed-distance source at the ESDF's resolution: the same three stages as
                         three launch groups, producing the fp16 [nx, ny, nz] grid the collision kernels read (VoxelData).
CUDA only.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .backends import pba as pba_cu
from .backends import tensor_checks as _tc

MAX_DIM = 1023  # 10-bit packed coordinates (perception/mapper/util/utils_quantization.py:33-34)


def validate_grid_size(grid_shape: Tuple[int, int, int], class_name: str) -> int:
    """esdf/kernel/wp_jfa.py:28-38, plus the 10-bit coordinate limit of the packing."""
    nx, ny, nz = (int(v) for v in grid_shape)
    n = nx * ny * nz
    if min(nx, ny, nz) < 1 or max(nx, ny, nz) > MAX_DIM:
        raise ValueError(f"{class_name}: grid dimensions must be in [1, {MAX_DIM}], got {grid_shape}")
    if n > 2 ** 31 - 1:
        raise ValueError(f"Grid too large for int32 site_index: {nx}x{ny}x{nz} = {n:,} voxels")
    return n


def seed_sites_from_occupancy(occupancy: torch.Tensor) -> torch.Tensor:
    """[nx, ny, nz] bool -> int32 site_index: packed own coordinates at occupied voxels, -1 elsewhere."""
    nx, ny, nz = occupancy.shape
    dev = occupancy.device
    x = torch.arange(nx, device=dev, dtype=torch.int32).view(nx, 1, 1)
    y = torch.arange(ny, device=dev, dtype=torch.int32).view(1, ny, 1)
    z = torch.arange(nz, device=dev, dtype=torch.int32).view(1, 1, nz)
    packed = (z << 20) | (y << 10) | x
    return torch.where(occupancy.bool(), packed, torch.full_like(packed, -1)).contiguous()


class ParallelBandingEDT:
    """Exact 3-D EDT: after `propagate`, every voxel of `site_index` holds the packed coordinates of a nearest site."""

    def __init__(self, grid_shape: Tuple[int, int, int], voxel_size: float, device: torch.device, m3: int = 2):
        device = torch.device(device)
        _tc.require_cuda(device, f"ParallelBandingEDT requires CUDA device, got {device}")
        self.grid_shape = tuple(int(v) for v in grid_shape)
        self.voxel_size = float(voxel_size)
        self.device = device
        self.m3 = m3
        self.n_voxels = validate_grid_size(self.grid_shape, "ParallelBandingEDT")
        # the reference's scratch buffer: part of its launcher signature; the kernels here work in place
        self._buffer = torch.empty(self.n_voxels, dtype=torch.int32, device=device)

    def propagate(self, site_index: torch.Tensor) -> None:
        nx, ny, nz = self.grid_shape
        pba_cu.launch_pba3d(site_index.view(-1), self._buffer, nx, ny, nz, m3=self.m3)

    def unsigned_distance(self, site_index: torch.Tensor, out: torch.Tensor = None, empty_value: float = 1e4) -> torch.Tensor:
        """fp16 [nx, ny, nz] distance field [m] from a propagated site_index."""
        nx, ny, nz = self.grid_shape
        if out is None:
            out = torch.empty(self.grid_shape, dtype=torch.float16, device=self.device)
        pba_cu.launch_edt_unsigned_distance(site_index.view(-1), out.view(-1), nx, ny, nz, self.voxel_size, empty_value)
        return out


class DenseESDFBuilder:
    """seed -> exact nearest-site transform -> signed fp16 distance, from dense SDF grids [nx, ny, nz] float32 (> 1e9 =
    unobserved): `combined_sdf` (seeding + sign fallback at the voxel) and `static_sdf` (sign next to the site).  All
    buffers are allocated once; `compute` is CUDA-graph capturable.  The result is the `features` layout of VoxelData
    (z fastest, fp16)."""

    def __init__(self, grid_shape: Tuple[int, int, int], voxel_size: float, truncation_distance: float, device,
                 adjacent_skip_steps: float = 1.0, seeding_method: str = "scatter", origin=(0.0, 0.0, 0.0)):
        """seeding_method: "scatter" (a voxel is a site when the seed rule holds for its own SDF: the reference's scatter kernel at
        equal resolution) or "gather" (the reference's default, mapper_cfg.py:103: the rule probed at the centre and half a voxel
        away along each axis -- a band up to one voxel thicker; needs the grid `origin` because the probes' voxels are decided by
        float rounding of world coordinates)."""
        if seeding_method not in ("scatter", "gather"):
            raise ValueError("seeding_method must be 'scatter' or 'gather'")
        self.seeding_method, self.origin = seeding_method, tuple(float(v) for v in origin)
        self.edt = ParallelBandingEDT(grid_shape, voxel_size, torch.device(device))
        self.truncation_distance = float(truncation_distance)
        self.adjacent_skip_steps = float(adjacent_skip_steps)
        self.site_index = torch.empty(self.edt.grid_shape, dtype=torch.int32, device=self.edt.device)
        self.dist_field = torch.empty(self.edt.grid_shape, dtype=torch.float16, device=self.edt.device)

    def to_voxel_data(self, origin=None, max_esdf_distance: float = 100.0):
        """The ESDF as the obstacle type the collision operators and the fused rollout read: a one-layer curobo_b200.scene.VoxelData
        whose `features` ALIAS `dist_field` (a later `compute` updates the world in place; call RolloutEngine.refresh_world() --
        or VoxelData.build_mip() -- afterwards so that the lower-bound level follows).  `origin` = grid centre (default: the
        builder's), identity rotation."""
        from .scene import VoxelData
        nx, ny, nz = self.edt.grid_shape
        o = self.origin if origin is None else tuple(float(v) for v in origin)
        dev = self.edt.device
        params = torch.tensor([[[nx, ny, nz, self.edt.voxel_size]]], dtype=torch.float32, device=dev)
        inv_pose = torch.tensor([[[-o[0], -o[1], -o[2], 1.0, 0.0, 0.0, 0.0]]], dtype=torch.float32, device=dev)
        return VoxelData(params, inv_pose, torch.ones((1, 1), dtype=torch.uint8, device=dev),
                         torch.ones(1, dtype=torch.int32, device=dev), self.dist_field.view(1, 1, -1), 1, 1, float(max_esdf_distance))

    def compute(self, combined_sdf: torch.Tensor, static_sdf: torch.Tensor = None) -> torch.Tensor:
        nx, ny, nz = self.edt.grid_shape
        if self.seeding_method == "gather":
            pba_cu.launch_esdf_seed_sites_gather(combined_sdf.view(-1), self.site_index.view(-1), nx, ny, nz, self.edt.voxel_size,
                                                 self.truncation_distance, self.origin)
        else:
            pba_cu.launch_esdf_seed_sites(combined_sdf.view(-1), self.site_index.view(-1), nx, ny, nz, self.edt.voxel_size,
                                          self.truncation_distance)
        self.edt.propagate(self.site_index)
        pba_cu.launch_esdf_signed_distance(self.site_index.view(-1), None if static_sdf is None else static_sdf.view(-1),
                                           combined_sdf.view(-1), self.dist_field.view(-1), nx, ny, nz, self.edt.voxel_size,
                                           self.adjacent_skip_steps)
        return self.dist_field


class DenseTSDF:
    """Depth images -> TSDF on a dense grid [nx, ny, nz] at the ESDF's resolution: the voxel-centric projective update of the
    reference's CameraProjectIntegrator (kernel/builder/builder_camera_integrate.py:399-489) without the block-sparse store in
    front of it.  `block_data` [nx, ny, nz, 2] float16 = (sum sdf * w, sum w); `combined_sdf(static_sdf)` is what
    DenseESDFBuilder.compute takes (sample_combined_sdf, kernel/wp_tsdf_sample.py:22-97).  Buffers are allocated once."""

    def __init__(self, grid_shape: Tuple[int, int, int], voxel_size: float, truncation_distance: float, device,
                 origin=(0.0, 0.0, 0.0), depth_min: float = 0.1, depth_max: float = 10.0, minimum_tsdf_weight: float = 0.1):
        self.grid_shape = tuple(int(v) for v in grid_shape)
        self.voxel_size, self.truncation_distance = float(voxel_size), float(truncation_distance)
        self.origin = tuple(float(v) for v in origin)
        self.depth_min, self.depth_max, self.minimum_tsdf_weight = float(depth_min), float(depth_max), float(minimum_tsdf_weight)
        self.device = torch.device(device)
        self.block_data = torch.zeros(self.grid_shape + (2,), dtype=torch.float16, device=self.device)
        self._combined = torch.empty(self.grid_shape, dtype=torch.float32, device=self.device)
        self.static_sdf = None                       # allocated by the first stamp_cuboids

    def reset(self) -> None:
        self.block_data.zero_()
        if self.static_sdf is not None:
            self.static_sdf.fill_(1e10)

    def stamp_cuboids(self, cuboids, env_idx: int = 0) -> torch.Tensor:
        """World cuboids (curobo_b200.scene.CuboidData) into the static channel (stamp_sdf_kernel, builder_stamp.py:263-315):
        returns the float32 [nx, ny, nz] static SDF (> 1e9 = nothing stamped), kept and min-combined across calls."""
        if self.static_sdf is None:
            self.static_sdf = torch.full(self.grid_shape, 1e10, dtype=torch.float32, device=self.device)
        nx, ny, nz = self.grid_shape
        pba_cu.launch_tsdf_stamp_cuboids(self.static_sdf.view(-1), nx, ny, nz, self.voxel_size, self.origin, self.truncation_distance,
                                         cuboids, env_idx)
        return self.static_sdf

    def integrate(self, depth_images: torch.Tensor, intrinsics: torch.Tensor, cam_positions: torch.Tensor,
                  cam_quaternions: torch.Tensor) -> None:
        nx, ny, nz = self.grid_shape
        pba_cu.launch_tsdf_integrate_depth(self.block_data.view(-1), nx, ny, nz, self.voxel_size, self.origin, intrinsics,
                                           cam_positions, cam_quaternions, depth_images, self.depth_min, self.depth_max,
                                           self.truncation_distance)

    def combined_sdf(self, static_sdf: torch.Tensor = None) -> torch.Tensor:
        pba_cu.launch_tsdf_combined_sdf(self.block_data.view(-1), None if static_sdf is None else static_sdf.view(-1),
                                        self._combined.view(-1), self.minimum_tsdf_weight)
        return self._combined
