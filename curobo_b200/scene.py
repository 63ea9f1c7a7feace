"""Scene-collision operators on device tensors (drop-in at the autograd level).

The reference has no backend switch here (Warp kernels are launched from the autograd Functions,
curobo/_src/geom/collision/wp_autograd.py:37-247), so the drop-in is the Function pair below with the
reference's own signatures.  Obstacle holders are duck-typed on the reference's attribute names
(`CuboidData`: dims/inv_pose/enable/count/max_n/num_envs, geom/data/data_cuboid.py:43-110;
 `VoxelData`:  params/inv_pose/enable/count/features/max_n/num_envs/max_esdf_distance,
 geom/data/data_voxel.py:41-92) so a reference `SceneData` can be passed directly.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import lib as _lib
from .backends.tensor_checks import check_tensors, stream_ptr
from .world import CuboidWorld, VoxelWorld


@dataclass
class CuboidData:
    dims: torch.Tensor
    inv_pose: torch.Tensor
    enable: torch.Tensor
    count: torch.Tensor
    max_n: int
    num_envs: int

    @classmethod
    def from_world(cls, w: CuboidWorld, device) -> "CuboidData":
        t = lambda a: torch.as_tensor(a).to(device).contiguous()  # noqa: E731
        return cls(t(w.dims), t(w.inv_pose), t(w.enable), t(w.count), w.max_n, w.num_envs)


@dataclass
class VoxelData:
    params: torch.Tensor
    inv_pose: torch.Tensor
    enable: torch.Tensor
    count: torch.Tensor
    features: torch.Tensor        # fp16 [n_env, max_n, n_vox(,1)]
    max_n: int
    num_envs: int
    max_esdf_distance: float

    @classmethod
    def from_world(cls, w: VoxelWorld, device) -> "VoxelData":
        t = lambda a: torch.as_tensor(a).to(device).contiguous()  # noqa: E731
        return cls(t(w.params), t(w.inv_pose), t(w.enable), t(w.count), t(w.features), w.max_n, w.num_envs,
                   float(w.max_dist))

    def build_mip(self) -> torch.Tensor:
        """(Re)build the optional lower-bound pyramid level of the ESDF (include/curobo_b200.h: cb200_voxel_set.mip)
        and attach it to this object; call again after the ESDF values change.  See build_voxel_mip."""
        return build_voxel_mip(self)


def build_voxel_mip(d) -> torch.Tensor:
    """One fp16 lower bound per block of cb200_voxel_mip_block()^3 trilinear base corners, stored on `d` as `d.cb200_mip`
    [num_envs * max_n, stride] so that c_voxel_set picks it up.  `d` is a VoxelData (ours or the reference's: only the
    attributes params / features / ... are read).  Results of every operator are identical with or without it."""
    dev = d.features.device
    L = _lib.load()
    params_host = d.params.detach().reshape(-1, 4).cpu().contiguous()
    n_layers = int(params_host.shape[0])
    stride = int(L.cb200_voxel_mip_stride(params_host.data_ptr(), n_layers))
    if stride < 1:
        raise ValueError("cannot size the ESDF pyramid level")
    mip = torch.empty((n_layers, stride), dtype=torch.int16, device=dev)
    try:
        d.cb200_mip = None
    except Exception:  # noqa: BLE001  (frozen holders)
        pass
    vs = c_voxel_set(d, dev)
    vs.mip, vs.mip_stride = mip.data_ptr(), stride
    _lib.check(L.cb200_voxel_build_mip(C.byref(vs), stream_ptr(dev)), "voxel_build_mip")
    d.cb200_mip = mip
    d.cb200_mip_stamp = _features_stamp(d)
    return mip


def _features_stamp(d):
    """Identity of the ESDF values a pyramid level was built from: storage pointer, shape and torch's in-place version counter.
    Any torch-side update of `features` (copy_, index writes, replacement of the tensor) changes it.  Writes through raw
    pointers by foreign kernels do not -- which is why the level is opt-in and owners of such grids call build_mip() /
    RolloutEngine.refresh_world() after every update."""
    f = d.features
    return (int(f.data_ptr()), tuple(f.shape), int(getattr(f, "_version", 0)))


def voxel_mip_is_fresh(d) -> bool:
    return getattr(d, "cb200_mip", None) is not None and getattr(d, "cb200_mip_stamp", None) == _features_stamp(d)


@dataclass
class SceneData:
    cuboid: Optional[CuboidData] = None
    voxel: Optional[VoxelData] = None
    mesh: Optional[object] = None          # curobo_b200.mesh.MeshData

    def get_valid_data(self) -> List[object]:
        return [d for d in (self.cuboid, self.voxel, self.mesh) if d is not None]


@dataclass
class CollisionBuffer:
    """distance [B,H,S] + gradient [B,H,S,4] (curobo/_src/geom/collision/buffer_collision.py:24-98)."""
    distance: torch.Tensor
    gradient: torch.Tensor

    @classmethod
    def from_shape(cls, shape, device) -> "CollisionBuffer":
        b, h, n, _ = shape
        return cls(torch.zeros((b, h, n), dtype=torch.float32, device=device),
                   torch.zeros((b, h, n, 4), dtype=torch.float32, device=device))


def _split_scene(scene):
    cub = vox = mesh = None
    for d in scene.get_valid_data():
        if hasattr(d, "features"):
            vox = d
        elif hasattr(d, "triangles"):
            mesh = d
        elif hasattr(d, "dims"):
            cub = d
        else:
            raise ValueError("b200 scene collision supports cuboid, voxel (ESDF) and mesh obstacles")
    return cub, vox, mesh


def c_cuboid_set(d: Optional[object], dev=None) -> Optional[_lib.CuboidSet]:
    if d is None:
        return None
    if dev is not None:
        check_tensors(dev, torch.float32, cuboid_dims=d.dims, cuboid_inv_pose=d.inv_pose)
        check_tensors(dev, torch.uint8, cuboid_enable=d.enable)
        check_tensors(dev, torch.int32, cuboid_count=d.count)
    return _lib.CuboidSet(d.dims.data_ptr(), d.inv_pose.data_ptr(), d.enable.data_ptr(), d.count.data_ptr(),
                          int(d.max_n), int(d.num_envs))


def c_voxel_set(d: Optional[object], dev=None) -> Optional[_lib.VoxelSet]:
    if d is None:
        return None
    if dev is not None:
        check_tensors(dev, torch.float32, voxel_params=d.params, voxel_inv_pose=d.inv_pose)
        check_tensors(dev, torch.float16, voxel_features=d.features)
        check_tensors(dev, torch.uint8, voxel_enable=d.enable)
        check_tensors(dev, torch.int32, voxel_count=d.count)
    n_vox = int(d.features.shape[2])
    # an attached lower-bound level is handed to the kernels only while it provably matches the grid it was built from
    mip = getattr(d, "cb200_mip", None) if voxel_mip_is_fresh(d) else None
    return _lib.VoxelSet(d.params.data_ptr(), d.inv_pose.data_ptr(), d.enable.data_ptr(), d.count.data_ptr(),
                         d.features.data_ptr(), n_vox, int(d.max_n), int(d.num_envs), float(d.max_esdf_distance),
                         mip.data_ptr() if mip is not None else None, int(mip.shape[1]) if mip is not None else 0)


def _launch(sweep, query_spheres, buffer, scene, weight, activation_distance, speed_dt, enable_speed_metric,
            env_query_idx, use_multi_env):
    dev = query_spheres.device
    b, h, n, _ = query_spheres.shape
    check_tensors(dev, torch.float32, query_spheres=query_spheres, distance=buffer.distance,
                  gradient=buffer.gradient, weight=weight, activation_distance=activation_distance)
    cub, vox, mesh = _split_scene(scene)
    cs, vs = c_cuboid_set(cub, dev), c_voxel_set(vox, dev)
    cp = C.byref(cs) if cs is not None else None
    vp = C.byref(vs) if vs is not None else None
    eq = None
    if env_query_idx is not None:
        check_tensors(dev, torch.int32, env_query_idx=env_query_idx)
        eq = env_query_idx.data_ptr()
    L = _lib.load()
    if not sweep:
        err = L.cb200_sphere_obstacle_collision(
            buffer.distance.data_ptr(), buffer.gradient.data_ptr(), query_spheres.data_ptr(), cp, vp,
            weight.data_ptr(), activation_distance.data_ptr(), eq, b, h, n, int(bool(use_multi_env)), stream_ptr(dev))
    else:
        if enable_speed_metric:
            check_tensors(dev, torch.float32, speed_dt=speed_dt)
        err = L.cb200_swept_sphere_obstacle_collision(
            buffer.distance.data_ptr(), buffer.gradient.data_ptr(), query_spheres.data_ptr(), cp, vp,
            weight.data_ptr(), activation_distance.data_ptr(), speed_dt.data_ptr() if speed_dt is not None else None,
            int(bool(enable_speed_metric)), eq, b, h, n, int(bool(use_multi_env)), stream_ptr(dev))
    _lib.check(err, "sphere_obstacle_collision")
    if mesh is not None:
        # one more launch for the mesh obstacle type, adding to the buffers the launch above wrote (it zero-fills them when
        # there are no cuboids / grids) -- the reference launches its generic kernel once per obstacle type too
        from .mesh import c_mesh_set
        ms = c_mesh_set(mesh, dev)
        err = L.cb200_sphere_mesh_collision(
            buffer.distance.data_ptr(), buffer.gradient.data_ptr(), query_spheres.data_ptr(), C.byref(ms), weight.data_ptr(),
            activation_distance.data_ptr(), speed_dt.data_ptr() if (sweep and speed_dt is not None) else None,
            int(bool(sweep and enable_speed_metric)), eq, b, h, n, int(bool(use_multi_env)), int(bool(sweep)), 1, stream_ptr(dev))
        _lib.check(err, "sphere_mesh_collision")


class SphereObstacleCollision(torch.autograd.Function):
    """Same call signature as curobo/_src/geom/collision/wp_autograd.py:37-121."""

    @staticmethod
    def forward(ctx, query_spheres, buffer, scene, weight, activation_distance, max_distance, env_query_idx,
                use_multi_env, return_loss=False):
        _launch(False, query_spheres.detach(), buffer, scene, weight, activation_distance, None, False,
                env_query_idx, use_multi_env)
        ctx.return_loss = return_loss
        ctx.save_for_backward(buffer.gradient)
        return buffer.distance

    @staticmethod
    def backward(ctx, grad_output):
        grad_sph = None
        if ctx.needs_input_grad[0]:
            (g,) = ctx.saved_tensors
            grad_sph = g * grad_output.unsqueeze(-1) if ctx.return_loss else g
        return grad_sph, None, None, None, None, None, None, None, None


class SweptSphereObstacleCollision(torch.autograd.Function):
    """Same call signature as curobo/_src/geom/collision/wp_autograd.py:124-247."""

    @staticmethod
    def forward(ctx, query_spheres, buffer, scene, weight, activation_distance, max_distance, speed_dt,
                enable_speed_metric, env_query_idx, use_multi_env, return_loss=False):
        _launch(True, query_spheres.detach(), buffer, scene, weight, activation_distance, speed_dt,
                enable_speed_metric, env_query_idx, use_multi_env)
        ctx.return_loss = return_loss
        ctx.save_for_backward(buffer.gradient)
        return buffer.distance

    @staticmethod
    def backward(ctx, grad_output):
        grad_sph = None
        if ctx.needs_input_grad[0]:
            (g,) = ctx.saved_tensors
            grad_sph = g * grad_output.unsqueeze(-1) if ctx.return_loss else g
        return grad_sph, None, None, None, None, None, None, None, None, None, None
