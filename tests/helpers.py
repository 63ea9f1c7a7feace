"""Shared test helpers: seeded synthetic inputs of the shapes in BASELINE.md section 3."""
import ctypes as C
import os

import numpy as np

from curobo_b200.robot_model import load_robot
from curobo_b200.world import CuboidWorld, VoxelWorld, make_benchmark_cuboid_world, make_box_esdf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_q(rm, n, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    lo, hi = rm.position_limits[0], rm.position_limits[1]
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * scale
    return rng.uniform(mid - half, mid + half, size=(n, rm.num_dof)).astype(np.float32)


def humanoid_q(rm, n, seed=0, scale=0.6):
    """Floating-base humanoid standing inside a 2.56 m world: virtual base joints (base_j_*) near the origin,
    pelvis ~0.8 m above the ground plane, small base rotations; other joints uniform in `scale` of their range."""
    q = random_q(rm, n, seed=seed, scale=scale)
    rng = np.random.default_rng(seed + 1000)
    for d, name in enumerate(rm.joint_names):
        if name in ("base_j_x", "base_j_y"):
            q[:, d] = rng.uniform(-0.3, 0.3, size=n)
        elif name == "base_j_z":
            q[:, d] = rng.uniform(0.75, 0.85, size=n)
        elif name.startswith("base_j_"):
            q[:, d] = rng.uniform(-0.2, 0.2, size=n)
    return q.astype(np.float32)


def random_walk_q(rm, b, h, seed=0, sigma=0.05):
    rng = np.random.default_rng(seed)
    q0 = random_q(rm, b, seed=seed + 1, scale=0.8)
    steps = rng.normal(0, sigma, size=(b, h, rm.num_dof)).astype(np.float32)
    steps[:, 0] = 0
    q = q0[:, None, :] + np.cumsum(steps, axis=1)
    return np.clip(q, rm.position_limits[0], rm.position_limits[1]).astype(np.float32)


def small_voxel_world(n=64, voxel=0.04, seed=3, num_boxes=10):
    sdf = make_box_esdf(n=n, voxel_size=voxel, num_boxes=num_boxes, seed=seed, ground_z=-0.05)
    return VoxelWorld.from_grid(sdf.reshape(n, n, n), voxel)


def hostmath():
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostmath", "libcb200_hostmath.so"))
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hm_scene(lib, spheres, weight, eta, sweep, speed, dt, cub: CuboidWorld = None, vox: VoxelWorld = None, mip=None):
    sp = np.ascontiguousarray(spheres, np.float32)
    B, H, S, _ = sp.shape
    cost = np.zeros((B, H, S), np.float32)
    grad = np.zeros((B, H, S, 4), np.float32)
    args = [ptr(sp), C.c_int(B), C.c_int(H), C.c_int(S), C.c_float(weight), C.c_float(eta), C.c_int(int(sweep)),
            C.c_int(int(speed)), C.c_float(dt)]
    if cub is not None:
        args += [ptr(cub.dims), ptr(cub.inv_pose), ptr(cub.enable), ptr(cub.count), C.c_int(cub.max_n)]
    else:
        args += [None, None, None, None, C.c_int(0)]
    if vox is not None:
        feat = np.ascontiguousarray(vox.features).view(np.uint16)
        args += [ptr(vox.params), ptr(vox.inv_pose), ptr(vox.enable), ptr(vox.count), ptr(feat),
                 C.c_int(vox.features.shape[2]), C.c_int(vox.max_n), C.c_float(vox.max_dist)]
    else:
        args += [None, None, None, None, None, C.c_int(0), C.c_int(0), C.c_float(0)]
    args += [ptr(cost), ptr(grad)]
    args += [ptr(mip) if mip is not None else None, C.c_int(int(mip.shape[-1]) if mip is not None else 0)]
    lib.hm_scene(*args)
    return cost, grad


def mip_block() -> int:
    from curobo_b200 import lib as _l
    return int(_l.load().cb200_voxel_mip_block())


def numpy_voxel_mip(vox: VoxelWorld):
    """Reference construction of the ESDF lower-bound level (cb200_voxel_build_mip): for every block of B^3 base
    corners [Bc, Bc+B-1] the minimum over fine voxels [Bc, Bc+B] per axis; returns uint16 [layers, stride]."""
    Bk = mip_block()
    layers = vox.params.reshape(-1, 4)
    feats = vox.features.reshape(layers.shape[0], -1)
    dims = [(int(p[0]), int(p[1]), int(p[2])) for p in layers]
    cd = lambda n: (n + Bk - 1) // Bk  # noqa: E731
    stride = max(cd(nx) * cd(ny) * cd(nz) for nx, ny, nz in dims)
    out = np.full((len(dims), stride), 0x7bff, np.uint16)
    for k, (nx, ny, nz) in enumerate(dims):
        g = feats[k, : nx * ny * nz].astype(np.float32).reshape(nx, ny, nz)
        mx, my, mz = cd(nx), cd(ny), cd(nz)
        m = np.zeros((mx, my, mz), np.float16)
        for cx in range(mx):
            for cy in range(my):
                for cz in range(mz):
                    m[cx, cy, cz] = g[Bk * cx: Bk * cx + Bk + 1, Bk * cy: Bk * cy + Bk + 1, Bk * cz: Bk * cz + Bk + 1].min()
        out[k, : mx * my * mz] = m.reshape(-1).view(np.uint16)
    return out
