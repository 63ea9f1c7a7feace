"""CPU pins of the exact nearest-site transform: the oracle (oracle/edt_oracle.py, a restatement of the
reference's PBA+ phases) against brute force and scipy's exact EDT; the product's column routines compiled for the host
(tests/hostmath) against the oracle and scipy.  Parity is on the squared distance to the reported site (integer, bit exact) and
on "the reported site is a site": which of several equidistant sites is reported depends on the sweep order, in the reference
too."""
import ctypes as C

import numpy as np
import pytest
from scipy import ndimage

from edt_cases import MEDIUM, SMALL, occupancy
from helpers import hostmath
from oracle import edt_oracle as E


def hm_pba3d(sites):
    g = np.ascontiguousarray(sites, np.int32).copy()
    nx, ny, nz = g.shape
    hostmath().hm_pba3d(g.ctypes.data_as(C.c_void_p), C.c_int(nx), C.c_int(ny), C.c_int(nz))
    return g


def hm_pba3d_tiles(sites, n_ctas=3):
    g = np.ascontiguousarray(sites, np.int32).copy()
    nx, ny, nz = g.shape
    hostmath().hm_pba3d_tiles(g.ctypes.data_as(C.c_void_p), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(n_ctas))
    return g


def check_result(res, occ):
    d2 = E.squared_distance(res)
    if not occ.any():
        assert (res == E.EMPTY).all()
        return
    assert (res >= 0).all()
    sx, sy, sz = E.unpack(res)
    assert (sx < occ.shape[0]).all() and (sy < occ.shape[1]).all() and (sz < occ.shape[2]).all()
    assert occ[sx, sy, sz].all(), "a reported site is not a site"
    want = ndimage.distance_transform_edt(~occ) ** 2
    assert np.array_equal(d2, np.rint(want).astype(np.int64)), "squared distances differ from the exact EDT"
    assert (d2[occ] == 0).all()


@pytest.mark.parametrize("kind,shape,p", SMALL)
@pytest.mark.parametrize("order", ["xyz", "zyx"])
def test_oracle_is_exact(kind, shape, p, order):
    occ = occupancy(kind, shape, seed=3, p=p)
    res = E.pba3d(E.seed_grid(occ), order)
    check_result(res, occ)
    assert np.array_equal(E.squared_distance(res), E.brute_force_squared_distance(occ))


@pytest.mark.parametrize("kind,shape,p", SMALL + MEDIUM)
def test_host_compiled_column_routines_are_exact(kind, shape, p):
    occ = occupancy(kind, shape, seed=5, p=p)
    sites = E.seed_grid(occ)
    res = hm_pba3d(sites)
    check_result(res, occ)
    if np.prod(shape) <= 2000:
        assert np.array_equal(E.squared_distance(res), E.squared_distance(E.pba3d(sites, "zyx")))


@pytest.mark.parametrize("kind,shape,p", SMALL + MEDIUM)
def test_kernel_tile_schedule_emulated_on_the_host(kind, shape, p):
    """The per-lane functions the CUDA kernels execute (FloodZ / Envelope<AXIS>::run, cb200_edt.cuh), driven by a host loop
    that strides CTAs over tiles and runs the 32 lanes between barriers: covers the kernels' index arithmetic (tile -> rows /
    columns, partial tiles, the transposed z tile) without a GPU.  Same result as the plain column-by-column driver."""
    occ = occupancy(kind, shape, seed=9, p=p)
    sites = E.seed_grid(occ)
    res = hm_pba3d_tiles(sites, n_ctas=3)
    check_result(res, occ)
    assert np.array_equal(res, hm_pba3d(sites)), "tile schedule and column driver disagree"
    assert np.array_equal(res, hm_pba3d_tiles(sites, n_ctas=1000)), "result depends on the number of CTAs"


def test_non_sites_may_be_any_negative_value_and_input_is_not_required_to_be_minus_one():
    occ = occupancy("random", (9, 8, 7), seed=1, p=0.05)
    sites = E.seed_grid(occ)
    rng = np.random.default_rng(0)
    noisy = np.where(sites < 0, -rng.integers(1, 2 ** 31, sites.shape), sites).astype(np.int32)
    assert np.array_equal(hm_pba3d(noisy), hm_pba3d(sites))
    assert np.array_equal(E.pba3d(noisy), E.pba3d(sites))


def test_maximum_coordinate_range_and_64_bit_dominance_products():
    """A 1023-long axis: coordinates use all 10 bits and the dominance products exceed 32 bits."""
    occ = np.zeros((1023, 2, 3), bool)
    occ[0, 0, 0] = occ[1022, 1, 2] = occ[511, 0, 1] = True
    check_result(hm_pba3d(E.seed_grid(occ)), occ)
    occ = np.zeros((2, 1023, 2), bool)
    occ[1, 1022, 1] = occ[0, 3, 0] = True
    check_result(hm_pba3d(E.seed_grid(occ)), occ)
    occ = np.zeros((2, 2, 1023), bool)
    occ[1, 1, 1000] = occ[0, 0, 17] = True
    check_result(hm_pba3d(E.seed_grid(occ)), occ)


def test_unsigned_distance_step():
    occ = occupancy("shells", (12, 10, 9), seed=2)
    res = E.pba3d(E.seed_grid(occ), "zyx")
    d = E.unsigned_distance_fp16(res, 0.02)
    assert d.dtype == np.float16 and (d[occ] == 0).all()
    want = (ndimage.distance_transform_edt(~occ) * 0.02).astype(np.float16)
    assert np.array_equal(d, want)
    assert (E.unsigned_distance_fp16(E.pba3d(E.seed_grid(np.zeros((3, 3, 3), bool))), 0.02) == np.float16(1e4)).all()


def test_host_mirror_seeding_and_grid_validation():
    """curobo_b200.esdf: the seeding helper (plain torch, device-agnostic) packs like the oracle; grid validation follows
    esdf/kernel/wp_jfa.py:28-38 plus the 10-bit coordinate limit; the operator refuses a CPU device like the reference."""
    import torch
    from curobo_b200 import esdf
    occ = occupancy("random", (5, 9, 7), seed=4, p=0.2)
    got = esdf.seed_sites_from_occupancy(torch.as_tensor(occ))
    assert got.dtype == torch.int32 and got.is_contiguous()
    assert np.array_equal(got.numpy(), E.seed_grid(occ))
    assert esdf.validate_grid_size((1023, 2, 3), "t") == 1023 * 6
    for bad in ((1024, 2, 2), (0, 4, 4), (4, 4, -1)):
        with pytest.raises(ValueError):
            esdf.validate_grid_size(bad, "t")
    with pytest.raises(ValueError):
        esdf.ParallelBandingEDT((8, 8, 8), 0.02, torch.device("cpu"))


def test_random_shapes_and_densities_property():
    """Property run over random grid shapes (1..70 per axis, so partial 32-column tiles in every pass) and site densities from
    a single site to nearly full: the emulated kernel schedule is exact against scipy."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(st.integers(1, 70), st.integers(1, 70), st.integers(1, 70), st.floats(0.0, 1.0), st.integers(0, 2 ** 16))
    def run(nx, ny, nz, u, seed):
        rng = np.random.default_rng(seed)
        p = u ** 4                                            # mostly sparse, sometimes dense
        occ = rng.random((nx, ny, nz)) < p
        if not occ.any():
            occ[tuple(int(rng.integers(0, n)) for n in occ.shape)] = True
        check_result(hm_pba3d_tiles(E.seed_grid(occ), n_ctas=int(rng.integers(1, 9))), occ)

    run()


def test_signed_distance_oracle_properties():
    """The oracle's restatement of the seeding rule and of compute_esdf_from_min_tsdf_kernel on an analytic ball: seeds lie on the
    surface shell, the signed field is negative exactly inside the ball (away from the one-voxel shell where the reference
    looks at the voxel itself), and |field| is the exact distance to the nearest seed."""
    n, voxel, r = 24, 0.05, 7.0
    g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float32)
    d = (np.linalg.norm(g - (n - 1) / 2.0, axis=-1) - r).astype(np.float32)
    trunc = 4 * voxel
    sdf = np.clip(d * voxel, -trunc, trunc).astype(np.float32)
    seeds = E.seed_sites_from_sdf(sdf, voxel, trunc)
    shell = np.abs(sdf) <= np.float32(voxel) * np.float32(0.9)
    deep = sdf < -(np.float32(trunc) - np.float32(voxel) * np.float32(1.1))
    assert np.array_equal(seeds >= 0, shell | deep) and shell.sum() > 0 and deep.sum() > 0
    res = E.pba3d(seeds, "zyx") if n ** 3 <= 4000 else hm_pba3d(seeds)
    f = E.signed_distance_fp16(res, sdf, sdf, voxel, 1.0).astype(np.float32)
    want_abs = (ndimage.distance_transform_edt(seeds < 0) * voxel).astype(np.float16).astype(np.float32)
    assert np.array_equal(np.abs(f), want_abs)
    clear = np.abs(d) > 1.5
    assert ((f < 0) == (d < 0))[clear & (seeds < 0)].all()
    unknown = np.full_like(sdf, 1e10)
    assert (E.signed_distance_fp16(res, unknown, unknown, voxel, 1.0).astype(np.float32) >= 0).all()
    assert (E.signed_distance_fp16(np.full((3, 3, 3), E.EMPTY, np.int32), None, None, voxel) == np.float16(1e4)).all()


def test_tsdf_depth_integration_oracle_on_a_plane():
    """Oracle of the depth -> TSDF stage (restatement of integrate_voxels_kernel, builder_camera_integrate.py:399-489) on a scene
    with a closed form: a camera at (0, 0, -1) looking along +z (identity rotation) sees a fronto-parallel plane at depth 1, i.e.
    the world plane z = 0.  Inside the frustum sdf = depth - z_cam = -z_world, kept for z_world <= truncation (behind the surface
    only down to -truncation is cut: sdf >= -truncation), clamped to +truncation in front; weights = max(pixel coverage, 1) per
    integration; unobserved elsewhere.  Then sample_combined_sdf: weight threshold and the min with a static channel."""
    from oracle import edt_oracle as E
    shape, voxel, trunc = (16, 16, 24), 0.05, 0.2
    H, W, f = 64, 64, 30.0
    K = np.array([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]], np.float32)
    pos = np.array([[0.0, 0.0, -1.0]], np.float32)
    quat = np.array([[1.0, 0.0, 0.0, 0.0]], np.float32)
    depth = np.ones((1, H, W), np.float32)
    bd = E.tsdf_integrate_depth(np.zeros(shape + (2,), np.float16), voxel, (0, 0, 0), K, pos, quat, depth, 0.1, 5.0, trunc)
    wz = (np.arange(shape[2]) + 0.5 - shape[2] / 2) * voxel
    w = bd[..., 1].astype(np.float32)
    seen = w > 0
    assert seen[:, :, wz <= trunc + 1e-6].all(), "the frustum covers the whole grid in front of and just behind the plane"
    assert not seen[:, :, wz > trunc + 1e-6].any(), "more than the truncation distance behind the surface is not updated"
    sdf = E.tsdf_combined_sdf(bd, None, 0.5)
    front = np.broadcast_to(wz < -trunc, shape) & seen
    band = np.broadcast_to(np.abs(wz) <= trunc, shape) & seen
    assert np.allclose(sdf[front], trunc, atol=2e-3)
    assert np.allclose(sdf[band], np.broadcast_to(-wz, shape)[band], atol=2e-3)
    assert (sdf[~seen] > 1e9).all()
    zc = wz + 1.0
    cov = (f * voxel / zc) ** 2
    assert np.allclose(w[8, 8, wz <= trunc], np.maximum(cov, 1.0)[wz <= trunc], rtol=2e-3)
    bd2 = E.tsdf_integrate_depth(bd, voxel, (0, 0, 0), K, pos, quat, depth, 0.1, 5.0, trunc)          # a second frame accumulates
    assert np.allclose(bd2[..., 1].astype(np.float32)[seen], 2 * w[seen], rtol=2e-3)
    assert np.allclose(E.tsdf_combined_sdf(bd2, None, 0.5)[band], sdf[band], atol=3e-3)
    assert (E.tsdf_combined_sdf(bd, None, 1e6) > 1e9).all()                                           # weight threshold
    static = np.full(shape, 1e10, np.float32)
    static[:, :, :3] = -0.3
    comb = E.tsdf_combined_sdf(bd, static, 0.5)
    assert (comb[:, :, :3] == -0.3).all() and np.array_equal(comb[:, :, 3:], sdf[:, :, 3:])
    depth0 = depth.copy()
    depth0[0, :, : W // 2] = 0.0                                                                      # invalid pixels: no update
    bd3 = E.tsdf_integrate_depth(np.zeros(shape + (2,), np.float16), voxel, (0, 0, 0), K, pos, quat, depth0, 0.1, 5.0, trunc)
    assert (bd3[: shape[0] // 2 - 1, :, :, 1] == 0).all() and (bd3[shape[0] // 2 + 1:, :, wz <= trunc, 1] > 0).all()


def test_tsdf_oracle_vs_the_references_own_kernel_source():
    """tests/golden/tsdf_reference_golden.npz holds the output of the REFERENCE's `integrate_voxels_kernel` source
    (builder_camera_integrate.py:399-489 + the coordinate functions of builder_coord.py) executed under the pure-Python Warp stand-in
    with every block of a small grid visible (tests/golden/make_tsdf_golden.py).  The oracle's dense restatement must reproduce it:
    same voxels updated, (sum sdf * w, sum w) equal to fp16 rounding -- the stand-in evaluates scalar float32 expressions in the
    kernel's order, numpy evaluates the same expressions vectorised, so at most a voxel on a pixel edge may differ."""
    import os
    from oracle import edt_oracle as E
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tsdf_reference_golden.npz"))
    for case in ("a", "b"):
        shape = tuple(int(v) for v in g[f"{case}/shape"])
        want_frames = g[f"{case}/block_data"]
        bd = np.zeros(shape + (2,), np.float16)
        for want in want_frames:
            bd = E.tsdf_integrate_depth(bd, float(g[f"{case}/voxel"]), g[f"{case}/origin"], g[f"{case}/K"], g[f"{case}/pos"],
                                        g[f"{case}/quat"], g[f"{case}/depth"], float(g[f"{case}/depth_min"]),
                                        float(g[f"{case}/depth_max"]), float(g[f"{case}/trunc"]))
            seen_w, seen_g = want[..., 1] > 0, bd[..., 1] > 0
            differ = (seen_w != seen_g) | ~np.isclose(bd.astype(np.float32), want.astype(np.float32), rtol=2e-3, atol=2e-3).all(-1)
            assert differ.sum() <= max(1, int(0.003 * differ.size)), f"case {case}: {int(differ.sum())} of {differ.size} voxels differ"
            assert seen_w.sum() > 100


def test_esdf_seeding_and_sign_oracle_vs_the_references_own_kernel_sources():
    """tests/golden/esdf_reference_golden.npz: the reference's scatter seeding kernel and compute_esdf_from_min_tsdf_kernel
    (builder_esdf.py:192-266, :412-499, with sample_combined_sdf / sample_static_sdf) executed under the Warp stand-in on a small
    all-blocks-allocated TSDF (tests/golden/make_esdf_golden.py).  The oracle's dense restatement (what the CUDA kernels are held to)
    must give the same seed sites and the same signed fp16 distances."""
    import os
    from oracle import edt_oracle as E
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "esdf_reference_golden.npz"))
    voxel, trunc, minw, skip = float(g["voxel"]), float(g["trunc"]), float(g["min_weight"]), float(g["skip"])
    static = g["static"].astype(np.float32)
    combined = E.tsdf_combined_sdf(g["block_data"], static, minw)
    seeds = E.seed_sites_from_sdf(combined, voxel, trunc)
    assert np.array_equal(seeds, g["seeds"]), f"{int((seeds != g['seeds']).sum())} seed voxels differ"
    assert (seeds >= 0).sum() > 100
    gather = E.seed_sites_gather_from_sdf(combined, voxel, trunc, g["origin"])          # the reference's default method
    assert np.array_equal(gather, g["seeds_gather"]), f"{int((gather != g['seeds_gather']).sum())} gather seed voxels differ"
    assert (gather >= 0).sum() > (seeds >= 0).sum() and ((seeds >= 0) <= (gather >= 0)).all(), "gather dilates the scatter band"
    static_in = np.where(np.isfinite(static), static, np.float32(1e10)).astype(np.float32)
    want = g["dist_field"].astype(np.float32)
    got = E.signed_distance_fp16(g["propagated"], static_in, combined, voxel, skip).astype(np.float32)
    assert np.array_equal(np.sign(got), np.sign(want)), f"{int((np.sign(got) != np.sign(want)).sum())} signs differ"
    assert np.abs(got - want).max() <= 1e-3 and (want < 0).sum() > 10 and (want > 0).sum() > 100


def test_stamp_cuboids_oracle_vs_the_references_own_kernel_source():
    """Static channel: the oracle's dense restatement of stamp_sdf_kernel (builder_stamp.py:263-315, cuboid overloads of
    data_cuboid.py:461-545) against the reference's kernel source under the Warp stand-in: two environments stamped one after the
    other into the same channel (min-combine), a rotated cuboid, a disabled slot.  Same voxels stamped, fp16-equal values."""
    import os
    from oracle import edt_oracle as E
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "esdf_reference_golden.npz"))
    shape = tuple(int(v) for v in g["shape"])
    st = np.full(shape, 1e10, np.float32)
    for env in (0, 1):
        st = E.tsdf_stamp_cuboids(st, float(g["voxel"]), g["origin"], float(g["trunc"]), g["cub_dims"], g["cub_inv_pose"],
                                  g["cub_enable"], g["cub_count"], int(g["cub_max_n"]), env)
        want = g["stamped"][env].astype(np.float32)
        assert np.array_equal(st < 1e9, np.isfinite(want)), f"env {env}: {int(((st < 1e9) != np.isfinite(want)).sum())} voxels differ"
        m = np.isfinite(want)
        assert m.sum() > 1000 and np.abs(st[m] - want[m]).max() <= 2e-4       # one fp16 ulp at 0.15
    assert (st[st < 1e9] < 0).sum() > 50
