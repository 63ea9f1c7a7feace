"""Generate tests/golden/tsdf_reference_golden.npz: the REFERENCE's own `integrate_voxels_kernel` (perception/mapper/kernel/builder/
builder_camera_integrate.py:399-489, built by make_camera_integrate_kernels with the coordinate functions of builder_coord.py)
executed on the CPU thread by thread under the pure-Python Warp stand-in (oracle/warp_shim), with EVERY block of a small grid
marked visible, so that its block pool is a dense grid in block order.  The fixture stores inputs and the resulting
(sum sdf * w, sum w) per voxel re-ordered to [nx, ny, nz]; tests/test_edt_cpu.py replays the inputs through
oracle/edt_oracle.tsdf_integrate_depth and compares.  Needs /root/reference (authoring container only):

    python tests/golden/make_tsdf_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import _reference_under_shim as R  # noqa: E402

R.prepare()
import warp as wp  # noqa: E402  (the stand-in)

from curobo_b200.world import depth_scene  # noqa: E402

coord = R.ref("curobo._src.perception.mapper.kernel.builder.builder_coord")
cam = R.ref("curobo._src.perception.mapper.kernel.builder.builder_camera_integrate")


def run_case(shape, bs, voxel, origin, n_cam, hw, seed, n_frames, depth_min, depth_max, trunc):
    nx, ny, nz = shape
    ck = coord.make_coord_kernels(bs, grid_shape=(nz, ny, nx), origin_xyz=origin, voxel_size=voxel)   # (D, H, W) = (z, y, x)
    K, pos, quat, depth, _ = depth_scene(shape, voxel, n_cam=n_cam, hw=hw, seed=seed)
    pos = (pos + np.asarray(origin, np.float32)).astype(np.float32)                                    # cameras follow the grid
    none = None
    kernels = cam.make_camera_integrate_kernels(
        bs, feature_dim=1, num_cameras=n_cam, image_height=hw[0], image_width=hw[1], num_samples=1, grid_shape=(nz, ny, nx),
        origin_xyz=origin, voxel_size=voxel, truncation_distance=trunc, feature_grid_shape=None, feature_channels_per_thread=1,
        max_feature_tile_channels=1, max_support_pixels_per_block_camera=1, pack_key_only=none, unpack_block_key=none,
        find_or_insert_block=none, hash_lookup=none, voxel_to_world=ck["voxel_to_world"],
        voxel_to_world_corner=ck["voxel_to_world_corner"], world_to_continuous_voxel=ck["world_to_continuous_voxel"],
        block_local_to_world=ck["block_local_to_world"], block_grid_to_key_coords=ck["block_grid_to_key_coords"],
        block_key_to_grid_coords=ck["block_key_to_grid_coords"])
    kern = kernels["integrate_voxels_kernel"]
    gb = (nx // bs, ny // bs, nz // bs)
    grid_blocks = [(gx, gy, gz) for gx in range(gb[0]) for gy in range(gb[1]) for gz in range(gb[2])]
    keys = []
    for g in grid_blocks:
        k = ck["block_grid_to_key_coords"](wp.int32(g[0]), wp.int32(g[1]), wp.int32(g[2]))
        keys.append([int(k[0]), int(k[1]), int(k[2])])
    nb = len(keys)
    block_coords = wp.from_numpy(np.asarray(keys, np.int32).reshape(-1), dtype=wp.int32)
    block_data = wp.from_numpy(np.zeros((nb, bs ** 3, 2), np.float16), dtype=wp.float16, ndim=3)
    vis = wp.from_numpy(np.arange(nb, dtype=np.int32), dtype=wp.int32)
    frames = []
    for _ in range(n_frames):
        wp.launch(kern, dim=(nb, bs ** 3), inputs=[
            vis, nb, wp.from_numpy(K, dtype=wp.float32, ndim=3), wp.from_numpy(pos, dtype=wp.float32, ndim=2),
            wp.from_numpy(quat, dtype=wp.float32, ndim=2), wp.from_numpy(depth, dtype=wp.float32, ndim=3), float(depth_min),
            float(depth_max), block_coords, block_data])
        bd = np.asarray(block_data.numpy(), np.float16).reshape(nb, bs ** 3, 2)
        dense = np.zeros((nx, ny, nz, 2), np.float16)
        for b, g in enumerate(grid_blocks):                      # local index = lz * bs^2 + ly * bs + lx (builder_coord.py:163-176)
            blk = bd[b].reshape(bs, bs, bs, 2)                   # [lz, ly, lx]
            dense[g[0] * bs:(g[0] + 1) * bs, g[1] * bs:(g[1] + 1) * bs, g[2] * bs:(g[2] + 1) * bs] = blk.transpose(2, 1, 0, 3)
        frames.append(dense)
    return dict(shape=np.asarray(shape), voxel=np.float32(voxel), origin=np.asarray(origin, np.float32), K=K, pos=pos, quat=quat,
                depth=depth, depth_min=np.float32(depth_min), depth_max=np.float32(depth_max), trunc=np.float32(trunc),
                block_data=np.stack(frames))


def main():
    out = {}
    cases = {"a": dict(shape=(12, 8, 16), bs=4, voxel=0.05, origin=(0.1, -0.05, 0.2), n_cam=2, hw=(24, 32), seed=5, n_frames=2,
                       depth_min=0.05, depth_max=5.0, trunc=0.15),
             "b": dict(shape=(8, 8, 8), bs=8, voxel=0.04, origin=(0.0, 0.0, 0.0), n_cam=3, hw=(16, 16), seed=6, n_frames=1,
                       depth_min=0.3, depth_max=0.9, trunc=0.08)}
    for name, kw in cases.items():
        for k, v in run_case(**kw).items():
            out[f"{name}/{k}"] = v
        print(name, "observed voxels", int((out[f"{name}/block_data"][-1][..., 1] > 0).sum()), "of", int(np.prod(kw["shape"])))
    np.savez_compressed(os.path.join(HERE, "tsdf_reference_golden.npz"), **out)


if __name__ == "__main__":
    main()
