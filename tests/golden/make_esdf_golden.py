"""Generate tests/golden/esdf_reference_golden.npz: the REFERENCE's own ESDF kernel sources -- scatter seeding
(seed_esdf_sites_from_block_sparse_kernel, perception/mapper/kernel/builder/builder_esdf.py:192-266) and the signed distance step
(compute_esdf_from_min_tsdf_kernel, :412-499, with sample_combined_sdf / sample_static_sdf of kernel/wp_tsdf_sample.py) -- executed
on the CPU thread by thread under the pure-Python Warp stand-in (oracle/warp_shim).  The block-sparse TSDF is a small grid with EVERY
block allocated; the hash table is replaced by a Python dict behind the `hash_lookup` closure parameter the builder takes (the hash
itself is out of scope).  Dynamic channel = the depth integration of make_tsdf_golden.py's case "a"; static channel = a slab.
The nearest-site propagation between the two kernels is the oracle's own exact transform (its parity is pinned elsewhere: scipy and
the reference's compiled PBA+ kernels).  tests/test_edt_cpu.py replays the inputs through oracle/edt_oracle.py.  Needs /root/reference:

    python tests/golden/make_esdf_golden.py
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

from oracle import edt_oracle as E  # noqa: E402

from curobo_b200.world import CuboidWorld  # noqa: E402

coord = R.ref("curobo._src.perception.mapper.kernel.builder.builder_coord")
stamp = R.ref("curobo._src.perception.mapper.kernel.builder.builder_stamp")
data_cuboid = R.ref("curobo._src.geom.data.data_cuboid")
esdf = R.ref("curobo._src.perception.mapper.kernel.builder.builder_esdf")
types_ = R.ref("curobo._src.perception.mapper.kernel.warp_types")


def to_blocks(dense, bs, grid_blocks):
    """[nx, ny, nz, ...] -> [n_blocks, bs^3, ...] with local index lz * bs^2 + ly * bs + lx."""
    out = []
    for g in grid_blocks:
        blk = dense[g[0] * bs:(g[0] + 1) * bs, g[1] * bs:(g[1] + 1) * bs, g[2] * bs:(g[2] + 1) * bs]
        out.append(np.moveaxis(blk, (0, 1, 2), (2, 1, 0)).reshape((bs ** 3,) + dense.shape[3:]))
    return np.stack(out)


def run_stamp(ck, shape, bs, voxel, origin, trunc):
    """stamp_sdf_kernel (builder_stamp.py:263-315) over every block of the grid, two environments of cuboids (one disabled slot),
    two stamping calls into the same static channel (min-combine)."""
    nx, ny, nz = shape
    grid_blocks = [(gx, gy, gz) for gx in range(nx // bs) for gy in range(ny // bs) for gz in range(nz // bs)]
    keys = []
    for g in grid_blocks:
        k = ck["block_grid_to_key_coords"](wp.int32(g[0]), wp.int32(g[1]), wp.int32(g[2]))
        keys.append((int(k[0]), int(k[1]), int(k[2])))

    def unpack_block_key(key):
        k = keys[int(key)]
        return wp.vec3i(k[0], k[1], k[2])
    sk = stamp.make_stamp_kernels(bs, grid_shape=(nz, ny, nx), origin_xyz=origin, voxel_size=voxel, truncation_distance=trunc,
                                  pack_key_only=None, unpack_block_key=unpack_block_key,
                                  block_local_to_world=ck["block_local_to_world"], hash_lookup=None,
                                  hash_table_insert_with_pool_idx=None, free_list_pop=None)
    o = np.asarray(origin)
    from curobo_b200.world import _inv_pose_from_pose
    cw = CuboidWorld.create([{"dims": [0.3, 0.2, 0.25], "pose": list(o + [0.05, 0.0, 0.1]) + [0.9, 0.1, 0.3, -0.2]},
                             {"dims": [0.8, 0.6, 0.1], "pose": list(o + [0.0, 0.0, -0.3]) + [1, 0, 0, 0]},
                             {"dims": [0.2, 0.2, 0.2], "pose": list(o + [-0.2, 0.1, 0.2]) + [1, 0, 0, 0]}], max_n=3, num_envs=2)
    cw.dims[1], cw.enable[1], cw.count[1] = 0.0, 0, 1                          # environment 1: a single rotated slab
    cw.inv_pose[1] = 0.0
    cw.inv_pose[1, :, 3] = 1.0
    cw.dims[1, 0, :3] = [0.12, 0.5, 0.5]
    cw.inv_pose[1, 0] = _inv_pose_from_pose(list(o + [0.2, 0.0, 0.0]) + [0.7071068, 0, 0, 0.7071068])
    cw.enable[1, 0] = 1
    cw.enable[0, 2] = 0                                                          # a disabled slot
    st = data_cuboid.CuboidDataWarp()
    st.dims = wp.from_numpy(cw.dims.reshape(-1, 4), dtype=wp.float32, ndim=2)
    st.inv_pose = wp.from_numpy(cw.inv_pose.reshape(-1, 8), dtype=wp.float32, ndim=2)
    st.enable = wp.from_numpy(cw.enable.reshape(-1), dtype=wp.uint8)
    st.n_per_env = wp.from_numpy(cw.count.reshape(-1), dtype=wp.int32)
    st.max_n, st.num_envs = wp.int32(cw.max_n), wp.int32(cw.num_envs)
    nb = len(keys)
    static = wp.from_numpy(np.full((nb, bs ** 3), np.inf, np.float16), dtype=wp.float16, ndim=2)
    sums = wp.from_numpy(np.zeros(nb, np.int32), dtype=wp.int32)
    outs = []
    for env in (0, 1):                                                           # env 1 is stamped on top of env 0
        wp.launch(sk["stamp_sdf_kernel"], dim=(nb, bs ** 3),
                  inputs=[wp.from_numpy(np.arange(nb, dtype=np.int64), dtype=wp.int64),
                          wp.from_numpy(np.arange(nb, dtype=np.int32), dtype=wp.int32), nb, st, env, static, sums])
        blocks = np.asarray(static.numpy(), np.float16).reshape(nb, bs, bs, bs)   # [lz, ly, lx]
        dense = np.zeros(shape, np.float16)
        for b, g in enumerate(grid_blocks):
            dense[g[0] * bs:(g[0] + 1) * bs, g[1] * bs:(g[1] + 1) * bs, g[2] * bs:(g[2] + 1) * bs] = blocks[b].transpose(2, 1, 0)
        outs.append(dense.copy())
    return np.stack(outs), cw


def main():
    t = np.load(os.path.join(HERE, "tsdf_reference_golden.npz"))
    shape = tuple(int(v) for v in t["a/shape"])
    nx, ny, nz = shape
    bs, voxel, origin, trunc = 4, float(t["a/voxel"]), tuple(float(v) for v in t["a/origin"]), float(t["a/trunc"])
    min_weight, skip = 0.5, 1.0
    bd = t["a/block_data"][-1]                                                   # [nx, ny, nz, 2] float16
    static = np.full(shape, np.inf, np.float16)
    static[:, :, :2] = np.float16(-0.03)                                         # a slab at low z ...
    static[:, :, 2] = np.float16(0.02)                                           # ... and its surface layer
    static[2:5, 1:4, 9:12] = np.float16(-0.2)                                    # an interior (beyond the truncation edge) box
    ck = coord.make_coord_kernels(bs, grid_shape=(nz, ny, nx), origin_xyz=origin, voxel_size=voxel)
    stamped, cw = run_stamp(ck, shape, bs, voxel, origin, trunc)
    table = {}

    def hash_lookup(hash_table, kx, ky, kz, capacity):
        return wp.int32(hash_table.get((int(kx), int(ky), int(kz)), -1))
    ek = esdf.make_esdf_kernels(bs, grid_shape=(nz, ny, nx), esdf_grid_shape=(nx, ny, nz), origin_xyz=origin, voxel_size=voxel,
                                truncation_distance=trunc, hash_lookup=hash_lookup,
                                block_grid_to_key_coords=ck["block_grid_to_key_coords"],
                                block_key_to_voxel_base=ck["block_key_to_voxel_base"])
    grid_blocks = [(gx, gy, gz) for gx in range(nx // bs) for gy in range(ny // bs) for gz in range(nz // bs)]
    keys = []
    for b, g in enumerate(grid_blocks):
        k = ck["block_grid_to_key_coords"](wp.int32(g[0]), wp.int32(g[1]), wp.int32(g[2]))
        keys.append([int(k[0]), int(k[1]), int(k[2])])
        table[tuple(keys[-1])] = b
    nb = len(keys)
    i32 = lambda a: wp.from_numpy(np.asarray(a, np.int32).reshape(-1), dtype=wp.int32)  # noqa: E731
    tsdf = types_.BlockSparseTSDFWarp(
        hash_table=table, hash_capacity=0,
        block_data=wp.from_numpy(to_blocks(bd, bs, grid_blocks), dtype=wp.float16, ndim=3),
        static_block_data=wp.from_numpy(to_blocks(static, bs, grid_blocks), dtype=wp.float16, ndim=2),
        has_dynamic=True, has_static=True, has_features=False, feature_dim=0, block_coords=i32(keys),
        block_to_hash_slot=i32(np.zeros(nb)), num_allocated=i32([nb]))
    tsdf.hash_capacity = 0
    n = nx * ny * nz
    site = wp.from_numpy(np.full(n, -1, np.int32), dtype=wp.int32)
    org = wp.from_numpy(np.asarray(origin, np.float32), dtype=wp.float32)
    vs = wp.from_numpy(np.asarray([voxel], np.float32), dtype=wp.float32)
    wp.launch(ek["seed_esdf_sites_from_block_sparse_kernel"], dim=(nb, bs ** 3), inputs=[tsdf, site, org, vs, float(min_weight)])
    seeds = np.asarray(site.numpy(), np.int32).reshape(shape).copy()
    site_g = wp.from_numpy(np.full(n, -1, np.int32), dtype=wp.int32)            # the default method: gather (pre-cleared to -1)
    wp.launch(ek["seed_esdf_sites_gather_kernel"], dim=n, inputs=[tsdf, site_g, org, vs, float(min_weight)])
    seeds_gather = np.asarray(site_g.numpy(), np.int32).reshape(shape).copy()
    prop = E.pba3d(seeds, "zyx")                                                 # exact nearest sites (oracle)
    site2 = wp.from_numpy(prop.reshape(-1).astype(np.int32), dtype=wp.int32)
    dist = wp.from_numpy(np.zeros(n, np.float16), dtype=wp.float16)
    wp.launch(ek["compute_esdf_from_min_tsdf_kernel"], dim=n, inputs=[site2, vs, dist, tsdf, float(min_weight), org, float(skip)])
    out = dict(shape=np.asarray(shape), voxel=np.float32(voxel), origin=np.asarray(origin, np.float32), trunc=np.float32(trunc),
               min_weight=np.float32(min_weight), skip=np.float32(skip), block_data=bd, static=static, seeds=seeds,
               seeds_gather=seeds_gather, propagated=prop, stamped=stamped, cub_dims=cw.dims, cub_inv_pose=cw.inv_pose,
               cub_enable=cw.enable, cub_count=cw.count, cub_max_n=np.int32(cw.max_n),
               dist_field=np.asarray(dist.numpy(), np.float16).reshape(shape))
    np.savez_compressed(os.path.join(HERE, "esdf_reference_golden.npz"), **out)
    d = out["dist_field"].astype(np.float32)
    print("gather seeds", int((seeds_gather >= 0).sum()), "| stamped voxels", [int(np.isfinite(x).sum()) for x in stamped])
    print("seeds", int((seeds >= 0).sum()), "of", n, "| negative distances", int((d < 0).sum()), "| unsigned-empty", int((d > 1e3).sum()))


if __name__ == "__main__":
    main()
