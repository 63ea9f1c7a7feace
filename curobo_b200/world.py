"""Obstacle storage for the scene-collision part of the rollout (host side).

Tensor layouts are the reference's wire formats, so a `SceneData` holder can hand us its tensors as-is:
  - cuboids:  ``CuboidData``  curobo/_src/geom/data/data_cuboid.py:43-110
        dims[n_env,max_n,4] f32 (full extents x,y,z,pad), inv_pose[n_env,max_n,8] f32
        (x,y,z,qw,qx,qy,qz,pad  = world->obstacle transform), enable[n_env,max_n] u8, count[n_env] i32
  - ESDF voxel grids: ``VoxelData``  curobo/_src/geom/data/data_voxel.py:41-92
        params[n_env,max_n,4] f32 (nx,ny,nz,voxel_size), inv_pose[...,8], enable, count,
        features[n_env,max_n,nx*ny*nz] **fp16**, C-order (z fastest, data_voxel.py:728-742), max_dist.

The synthetic world builders follow the reference's own test/benchmark generators:
  - box ESDF: tests/_src/geom/sdf/test_voxel_collision.py:403-440 (`_make_box_esdf`), min over boxes
  - cuboid world: benchmark/cost_gradient_benchmark.py:486-499 (table + pillar), cache of 10
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


def _inv_pose_from_pose(pose: Sequence[float]) -> np.ndarray:
    """pose = [x,y,z,qw,qx,qy,qz] (obstacle in world) -> inverse pose, same layout + pad."""
    x, y, z, w, a, b, c = [float(v) for v in pose]
    n = (w * w + a * a + b * b + c * c) ** 0.5
    w, a, b, c = w / n, a / n, b / n, c / n
    R = np.array([[1 - 2 * (b * b + c * c), 2 * (a * b - w * c), 2 * (a * c + w * b)],
                  [2 * (a * b + w * c), 1 - 2 * (a * a + c * c), 2 * (b * c - w * a)],
                  [2 * (a * c - w * b), 2 * (b * c + w * a), 1 - 2 * (a * a + b * b)]])
    t = -R.T @ np.array([x, y, z])
    return np.array([t[0], t[1], t[2], w, -a, -b, -c, 0.0], dtype=np.float32)


@dataclass
class CuboidWorld:
    dims: np.ndarray       # [n_env,max_n,4] f32
    inv_pose: np.ndarray   # [n_env,max_n,8] f32
    enable: np.ndarray     # [n_env,max_n] u8
    count: np.ndarray      # [n_env] i32

    @property
    def max_n(self) -> int:
        return int(self.dims.shape[1])

    @property
    def num_envs(self) -> int:
        return int(self.dims.shape[0])

    @classmethod
    def create(cls, cuboids: List[dict], max_n: Optional[int] = None, num_envs: int = 1) -> "CuboidWorld":
        """cuboids: [{"dims":[x,y,z], "pose":[x,y,z,qw,qx,qy,qz]}, ...] loaded into every env."""
        max_n = max(max_n or len(cuboids), 1)
        dims = np.zeros((num_envs, max_n, 4), np.float32)
        inv = np.zeros((num_envs, max_n, 8), np.float32)
        inv[..., 3] = 1.0
        en = np.zeros((num_envs, max_n), np.uint8)
        cnt = np.full((num_envs,), len(cuboids), np.int32)
        for e in range(num_envs):
            for i, c in enumerate(cuboids):
                dims[e, i, :3] = c["dims"]
                inv[e, i] = _inv_pose_from_pose(c["pose"])
                en[e, i] = 1
        return cls(dims, inv, en, cnt)


@dataclass
class VoxelWorld:
    params: np.ndarray     # [n_env,max_n,4] f32 (nx,ny,nz,voxel)
    inv_pose: np.ndarray   # [n_env,max_n,8] f32
    enable: np.ndarray     # [n_env,max_n] u8
    count: np.ndarray      # [n_env] i32
    features: np.ndarray   # [n_env,max_n,n_vox] f16
    max_dist: float = 100.0

    @property
    def max_n(self) -> int:
        return int(self.params.shape[1])

    @property
    def num_envs(self) -> int:
        return int(self.params.shape[0])

    @classmethod
    def from_grid(cls, sdf: np.ndarray, voxel_size: float,
                  pose: Sequence[float] = (0, 0, 0, 1, 0, 0, 0), max_dist: float = 100.0) -> "VoxelWorld":
        """One grid, one env. sdf: [nx,ny,nz] (any float dtype; stored fp16)."""
        nx, ny, nz = sdf.shape
        params = np.array([[[nx, ny, nz, voxel_size]]], np.float32)
        inv = _inv_pose_from_pose(pose).reshape(1, 1, 8)
        return cls(params, inv, np.ones((1, 1), np.uint8), np.ones((1,), np.int32),
                   np.ascontiguousarray(sdf, dtype=np.float16).reshape(1, 1, -1), float(max_dist))


# ------------------------------------------------------------------------------------------
# synthetic worlds (seeded) used by tests and bench.py
# ------------------------------------------------------------------------------------------

def make_box_esdf(n: int = 256, voxel_size: float = 0.01, num_boxes: int = 12, seed: int = 0,
                  ground_z: Optional[float] = -0.05, grid_center=(0.0, 0.0, 0.0),
                  xp=np) -> np.ndarray:
    """Analytic SDF of seeded random boxes (+ ground half-space), sampled at voxel centres.

    Voxel (i,j,k) centre = grid_center + (idx - (n-1)/2) * voxel  (test_voxel_collision.py:417-423).
    `xp` may be numpy or torch (same code path builds the 256^3 grid on the GPU in bench.py)."""
    rng = np.random.default_rng(seed)
    half_extent = 0.5 * n * voxel_size
    centers = rng.uniform(-0.7 * half_extent, 0.7 * half_extent, size=(num_boxes, 3))
    centers[:, 2] = rng.uniform(0.1, 0.8 * half_extent, size=num_boxes)
    # keep boxes away from the robot base column so that not every configuration collides
    for c in centers:
        if abs(c[0]) < 0.25 and abs(c[1]) < 0.25:
            c[0] += 0.45 if c[0] >= 0 else -0.45
    halves = rng.uniform(0.04, 0.15, size=(num_boxes, 3))
    if xp is np:
        ax = (np.arange(n, dtype=np.float32) - (n - 1) / 2.0) * voxel_size
        gx, gy, gz = np.meshgrid(ax + grid_center[0], ax + grid_center[1], ax + grid_center[2], indexing="ij")
        sdf = np.full((n, n, n), 1e3, np.float32)
        for c, h in zip(centers, halves):
            dx, dy, dz = np.abs(gx - c[0]) - h[0], np.abs(gy - c[1]) - h[1], np.abs(gz - c[2]) - h[2]
            out = np.sqrt(np.maximum(dx, 0) ** 2 + np.maximum(dy, 0) ** 2 + np.maximum(dz, 0) ** 2)
            ins = np.minimum(np.maximum(np.maximum(dx, dy), dz), 0)
            sdf = np.minimum(sdf, (out + ins).astype(np.float32))
        if ground_z is not None:
            sdf = np.minimum(sdf, (gz - ground_z).astype(np.float32))
        return sdf.astype(np.float16)
    torch = xp
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ax = (torch.arange(n, dtype=torch.float32, device=dev) - (n - 1) / 2.0) * voxel_size
    gx, gy, gz = torch.meshgrid(ax + grid_center[0], ax + grid_center[1], ax + grid_center[2], indexing="ij")
    sdf = torch.full((n, n, n), 1e3, dtype=torch.float32, device=dev)
    for c, h in zip(centers, halves):
        dx, dy, dz = (gx - float(c[0])).abs() - float(h[0]), (gy - float(c[1])).abs() - float(h[1]), (gz - float(c[2])).abs() - float(h[2])
        out = torch.sqrt(dx.clamp(min=0) ** 2 + dy.clamp(min=0) ** 2 + dz.clamp(min=0) ** 2)
        ins = torch.maximum(torch.maximum(dx, dy), dz).clamp(max=0)
        sdf = torch.minimum(sdf, out + ins)
    if ground_z is not None:
        sdf = torch.minimum(sdf, gz - ground_z)
    return sdf.to(torch.float16)


def make_benchmark_cuboid_world(max_n: int = 10) -> CuboidWorld:
    """Table + pillar of benchmark/cost_gradient_benchmark.py:486-499, padded to a cache of 10."""
    return CuboidWorld.create([
        {"dims": [2.2, 2.2, 0.2], "pose": [0.0, 0.0, -0.1, 1, 0, 0, 0]},
        {"dims": [0.1, 0.1, 1.5], "pose": [0.45, 0.0, 0.3, 1, 0, 0, 0]},
    ], max_n=max_n)


def make_single_box_esdf(grid_dims=(0.5, 0.5, 0.5), voxel_size=0.02, grid_center=(0.0, 0.0, 0.0),
                         box_center=(0.0, 0.0, 0.0), box_half=(0.05, 0.05, 0.05), max_dist: float = 100.0) -> VoxelWorld:
    """The reference's own test ESDF (tests/_src/geom/sdf/test_voxel_collision.py:403-440)."""
    n = [int(round(d / voxel_size)) for d in grid_dims]
    ax = [grid_center[a] + (np.arange(n[a], dtype=np.float32) - (n[a] - 1) / 2.0) * voxel_size for a in range(3)]
    gx, gy, gz = np.meshgrid(*ax, indexing="ij")
    d = [np.abs(g - box_center[a]) - box_half[a] for a, g in enumerate((gx, gy, gz))]
    out = np.sqrt(sum(np.maximum(x, 0) ** 2 for x in d))
    ins = np.minimum(np.maximum(np.maximum(d[0], d[1]), d[2]), 0)
    return VoxelWorld.from_grid((out + ins).astype(np.float16), voxel_size,
                                pose=[grid_center[0], grid_center[1], grid_center[2], 1, 0, 0, 0], max_dist=max_dist)


def make_empty_esdf(dims=(0.5, 0.5, 0.5), voxel_size=0.02, center=(0.0, 0.0, 0.0), fill_value=1.0,
                    max_dist: float = 100.0) -> VoxelWorld:
    """All-free-space grid (tests/_src/geom/sdf/test_voxel_collision.py:381-400)."""
    n = [int(round(d / voxel_size)) for d in dims]
    return VoxelWorld.from_grid(np.full(n, fill_value, np.float16), voxel_size,
                                pose=[center[0], center[1], center[2], 1, 0, 0, 0], max_dist=max_dist)


def _look_at_quat(eye, target):
    """camera -> world quaternion (wxyz) of a pinhole camera at `eye` whose +z axis points at `target` (x right, y down)."""
    zc = np.asarray(target, np.float64) - np.asarray(eye, np.float64)
    zc /= np.linalg.norm(zc)
    up = np.array([0.0, 0.0, 1.0]) if abs(zc[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    xc = np.cross(zc, up)
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    R = np.stack([xc, yc, zc], 1)                                             # columns = camera axes in the world
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    if w > 1e-6:
        q = np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    else:
        q = np.array([0.0, 1.0, 0.0, 0.0])
    return (q / np.linalg.norm(q)).astype(np.float32), R


def depth_scene(shape, voxel, n_cam=2, hw=(48, 64), seed=0):
    """Synthetic input of the depth -> TSDF stage (tests, bench): n_cam pinhole cameras around a ball of radius 0.3 * extent at the
    grid centre; rendered z-depth (ray / sphere, a far wall behind), 3 % invalid (zero) pixels.  Returns intrinsics [C,3,3],
    positions [C,3], quaternions [C,4] (wxyz, camera -> world), depth [C,H,W] (float32) and the ball radius."""
    rng = np.random.default_rng(seed)
    ext = voxel * min(shape)
    radius = 0.3 * ext
    H, W = hw
    K = np.zeros((n_cam, 3, 3), np.float32)
    pos = np.zeros((n_cam, 3), np.float32)
    quat = np.zeros((n_cam, 4), np.float32)
    depth = np.zeros((n_cam, H, W), np.float32)
    for c in range(n_cam):
        ang = 2 * np.pi * c / n_cam + 0.3
        eye = np.array([np.cos(ang), np.sin(ang), 0.35]) * 1.6 * ext
        q, R = _look_at_quat(eye, (0.0, 0.0, 0.0))
        f = 0.9 * W
        K[c] = [[f, 0, W / 2 - 0.5], [0, f, H / 2 - 0.5], [0, 0, 1]]
        pos[c], quat[c] = eye, q
        v, u = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
        d = np.stack([(u - K[c, 0, 2]) / f, (v - K[c, 1, 2]) / f, np.ones_like(u)], -1)       # camera-frame ray, z = 1
        dw = d @ R.T
        a = (dw * dw).sum(-1)
        b = 2 * (dw @ eye)
        cc = eye @ eye - radius * radius
        disc = b * b - 4 * a * cc
        t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 3.0 * ext)     # z-depth of the hit (d.z = 1)
        depth[c] = t.astype(np.float32)
        depth[c][rng.random((H, W)) < 0.03] = 0.0
    return K, pos, quat, depth, radius
