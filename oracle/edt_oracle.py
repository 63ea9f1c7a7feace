"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's exact 3-D nearest-site transform (PBA+ 3-D EDT).

Follows curobo/_src/curobolib/kernels/parallel_banding/pba3d_kernel.cuh phase by phase:
    flood_axis      <- kernel_flood_z           :67-104   (bidirectional 1-D sweep along the first axis)
    maurer_axis     <- kernel_maurer_axis       :160-222  (stack of Voronoi-dominant sites, is_voronoi_dominated :119-142)
    color_axis      <- kernel_color_axis        :235-350  (walk the stack from the far end, nearest site per row)
    pba3d           <- launch_pba3d, backends/cuda_core_backend/pba.py:60-124 (FloodZ, Maurer+Color on Y, Maurer+Color on X)
Packing: site = (z << 20) | (y << 10) | x in cuRobo indices (perception/mapper/util/utils_quantization.py:40-54,
site_encoding.cuh:13-19 with the kernels' (sx, sy, sz) = (nz, ny, nx)); non-sites are negative; the transform leaves
0x80000000 where no site exists at all.  Grid layout [nx, ny, nz], x slowest.

The reference sweeps cuRobo-x first, then y, then z (through two transposes); the transform is separable, so the ORDER of the
axes only changes which of several equidistant sites is reported -- never the distance.  `pba3d(order=...)` takes the order
so the tests can restate both the reference's order ("xyz") and the product kernels' order ("zyx", which needs no transpose
because z is the contiguous axis).  Parity is therefore defined on the squared distance to the reported site (integer, exact)
plus "the reported site is a site"; pinned against scipy.ndimage.distance_transform_edt (an independent exact EDT) and brute
force in tests/test_edt_cpu.py.  Pure-Python loops: for small grids only.
"""
import numpy as np

EMPTY = -(2 ** 31)
MASK = 0x3FF
SHIFT = {"x": 0, "y": 10, "z": 20}


def pack(x, y, z):
    return (np.asarray(z, np.int64) << 20 | np.asarray(y, np.int64) << 10 | np.asarray(x, np.int64)).astype(np.int32)


def unpack(v):
    v = np.asarray(v, np.int64)
    return v & MASK, (v >> 10) & MASK, (v >> 20) & MASK


def seed_grid(occupancy):
    """occupancy [nx,ny,nz] bool -> site_index int32: own packed coordinates at sites, -1 elsewhere
    (integrator_esdf.py:526 fill_(-1) + the seeding kernels)."""
    occ = np.asarray(occupancy, bool)
    x, y, z = np.meshgrid(*[np.arange(n) for n in occ.shape], indexing="ij")
    return np.where(occ, pack(x, y, z), np.int32(-1)).astype(np.int32)


def _coord(v, axis):
    return (int(v) >> SHIFT[axis]) & MASK


def flood_axis(col, axis):
    """kernel_flood_z on one column (list of packed ints along `axis`)."""
    n = len(col)
    out = [EMPTY] * n
    carry = EMPTY
    for i in range(n):                                   # forward :79-84
        if col[i] >= 0:
            carry = int(col[i])
        out[i] = carry
    big = 2 ** 31 - 1
    for i in range(n - 2, -1, -1):                       # backward :91-103 (ties keep the backward site)
        db = abs((_coord(carry, axis) if carry >= 0 else big) - i)
        f = out[i]
        df = abs((_coord(f, axis) if f >= 0 else big) - i)
        if df < db:
            carry = f
        out[i] = carry
    return out


def _h(v, axis, q):
    """squared distance from site v to the column through q, measured off-axis."""
    sx, sy, sz = int(v) & MASK, (int(v) >> 10) & MASK, (int(v) >> 20) & MASK
    d = 0
    if axis != "x":
        d += (sx - q[0]) ** 2
    if axis != "y":
        d += (sy - q[1]) ** 2
    if axis != "z":
        d += (sz - q[2]) ** 2
    return d


def maurer_color_axis(col, axis, q):
    """kernel_maurer_axis + kernel_color_axis on one column.  col[r] = nearest site of row r within the axes already done (or
    negative).  Stack entries are (row, site); a new site pops every top entry B for which the parabolas of the entry below (A)
    and the new site (C) meet before B's row range (is_voronoi_dominated, written here with g = row^2 + h:
    (gB - gA)(rC - rB) > (gC - gB)(rB - rA), identical to the reference's lhs > rhs after expanding its sums of coordinates)."""
    stack = []
    for r, v in enumerate(col):
        if v < 0:
            continue
        gc = r * r + _h(v, axis, q)
        while len(stack) >= 2:
            (ra, va), (rb, vb) = stack[-2], stack[-1]
            ga, gb = ra * ra + _h(va, axis, q), rb * rb + _h(vb, axis, q)
            if (gb - ga) * (r - rb) > (gc - gb) * (rb - ra):
                stack.pop()
            else:
                break
        stack.append((r, int(v)))
    n = len(col)
    out = [EMPTY] * n
    if not stack:
        return out
    k = len(stack) - 1                                    # color: from the last row down, pop while the entry below is
    for t in range(n - 1, -1, -1):                        # at least as close (:303-322, `cand_sq > min_dist_sq -> break`)
        best = (stack[k][0] - t) ** 2 + _h(stack[k][1], axis, q)
        while k > 0:
            cand = (stack[k - 1][0] - t) ** 2 + _h(stack[k - 1][1], axis, q)
            if cand > best:
                break
            best = cand
            k -= 1
        out[t] = stack[k][1]
    return out


def pba3d(site_index, order="xyz"):
    """Nearest-site transform of site_index [nx,ny,nz] int32 (negative = no site).  order[0] is flooded, order[1:] use
    Maurer stacks; "xyz" is the reference's schedule (pba.py:86-124)."""
    g = np.array(site_index, dtype=np.int64).copy()
    assert g.ndim == 3 and max(g.shape) <= 1023
    g[g < 0] = EMPTY
    ax_id = {"x": 0, "y": 1, "z": 2}
    for step, axis in enumerate(order):
        a = ax_id[axis]
        others = [i for i in range(3) if i != a]
        for i in range(g.shape[others[0]]):
            for j in range(g.shape[others[1]]):
                idx = [0, 0, 0]
                idx[others[0]], idx[others[1]] = i, j
                sl = [i, j]
                sl.insert(a, slice(None))
                col = [int(v) for v in g[tuple(sl)]]
                res = flood_axis(col, axis) if step == 0 else maurer_color_axis(col, axis, idx)
                g[tuple(sl)] = res
    return g.astype(np.int32)


def squared_distance(result, shape=None):
    """d^2 (voxel units) from every voxel to its reported site; -1 where the transform reports no site."""
    r = np.asarray(result)
    x, y, z = np.meshgrid(*[np.arange(n) for n in r.shape], indexing="ij")
    sx, sy, sz = unpack(r)
    d2 = (sx - x) ** 2 + (sy - y) ** 2 + (sz - z) ** 2
    return np.where(r < 0, -1, d2).astype(np.int64)


def brute_force_squared_distance(occupancy):
    occ = np.asarray(occupancy, bool)
    pts = np.argwhere(occ)
    if len(pts) == 0:
        return np.full(occ.shape, -1, np.int64)
    x, y, z = np.meshgrid(*[np.arange(n) for n in occ.shape], indexing="ij")
    q = np.stack([x, y, z], -1).reshape(-1, 1, 3)
    best = np.full(q.shape[0], np.iinfo(np.int64).max, np.int64)
    for s in range(0, len(pts), 512):
        d = ((q - pts[None, s:s + 512]) ** 2).sum(-1)
        best = np.minimum(best, d.min(1))
    return best.reshape(occ.shape)


def unsigned_distance_fp16(result, voxel_size, empty_value=1e4):
    """The distance step of the ESDF builder without the TSDF sign (builder_esdf.py:434-446): fp16(|voxel - site| * voxel_size),
    fp16(1e4) where no site exists."""
    d2 = squared_distance(result)
    d = np.sqrt(np.maximum(d2, 0).astype(np.float32)) * np.float32(voxel_size)
    return np.where(d2 < 0, np.float32(empty_value), d).astype(np.float16)


def seed_sites_from_sdf(sdf, voxel_size, truncation):
    """Seed rule of the ESDF builder for a dense SDF at the ESDF's resolution (builder_esdf.py:255-261, 286-300): a voxel is a
    site when it is observed (sdf <= 1e9) and |sdf| <= 0.9 voxel (surface) or sdf < -(truncation - 1.1 voxel) (truncation
    boundary); sites hold their own packed coordinates, everything else -1.  = the SCATTER seeding kernel at equal resolution
    (builder_esdf.py:192-266; the gather variant dilates the band and is not restated).  Pinned, together with
    signed_distance_fp16 below, on the reference's kernel sources executed under the Warp stand-in
    (tests/golden/make_esdf_golden.py -> esdf_reference_golden.npz)."""
    sdf = np.asarray(sdf, np.float32)
    vs, tr = np.float32(voxel_size), np.float32(truncation)
    seed = ~(sdf > np.float32(1e9)) & ((np.abs(sdf) <= vs * np.float32(0.9)) | (sdf < -(tr - vs * np.float32(1.1))))
    return seed_grid(seed)


def seed_sites_gather_from_sdf(sdf, voxel_size, truncation, origin):
    """The reference's DEFAULT seeding (mapper_cfg.py:103; seed_esdf_sites_gather_kernel builder_esdf.py:308-404 with
    _check_seed_at_world_pos :267-306) for a dense SDF on the ESDF's own grid: 7 probes per voxel (centre, +- half a voxel along
    each axis), world -> voxel by int((w - origin) / voxel + n / 2) in float32 in the kernel's order.  Pinned on the reference's
    kernel source under the Warp stand-in (tests/golden/make_esdf_golden.py)."""
    f = np.float32
    sdf = np.asarray(sdf, f)
    nx, ny, nz = sdf.shape
    vs, tr = f(voxel_size), f(truncation)
    o = np.asarray(origin, f)
    surface, edge = vs * f(0.9), -(tr - vs * f(1.1))
    half = vs * f(0.5)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cx = o[0] + (ix.astype(f) + f(0.5) - f(nx) * f(0.5)) * vs
    cy = o[1] + (iy.astype(f) + f(0.5) - f(ny) * f(0.5)) * vs
    cz = o[2] + (iz.astype(f) + f(0.5) - f(nz) * f(0.5)) * vs

    def rule(wx, wy, wz):
        gx = np.trunc((wx - o[0]) / vs + f(nx) * f(0.5)).astype(np.int64)
        gy = np.trunc((wy - o[1]) / vs + f(ny) * f(0.5)).astype(np.int64)
        gz = np.trunc((wz - o[2]) / vs + f(nz) * f(0.5)).astype(np.int64)
        ok = (gx >= 0) & (gx < nx) & (gy >= 0) & (gy < ny) & (gz >= 0) & (gz < nz)
        d = sdf[np.clip(gx, 0, nx - 1), np.clip(gy, 0, ny - 1), np.clip(gz, 0, nz - 1)]
        return ok & ~(d > f(1e9)) & ((np.abs(d) <= surface) | (d < edge))
    hit = rule(cx, cy, cz)
    for dx, dy, dz in ((half, 0, 0), (-half, 0, 0), (0, half, 0), (0, -half, 0), (0, 0, half), (0, 0, -half)):
        hit |= rule((cx + f(dx)).astype(f), (cy + f(dy)).astype(f), (cz + f(dz)).astype(f))
    return seed_grid(hit)


def _round_half_away(v):
    return np.where(v < 0, -np.floor(np.float32(0.5) - v), np.floor(v + np.float32(0.5)))


def signed_distance_fp16(result, static_sdf, combined_sdf, voxel_size, skip_steps=1.0):
    """compute_esdf_from_min_tsdf_kernel (builder_esdf.py:410-503) with the TSDF hash look-ups replaced by dense arrays at the
    ESDF's resolution: distance to the reported site, sign from the static SDF one skip step from the site towards the voxel when
    the voxel is more than one voxel away, else / if unobserved from the combined SDF at the voxel, unsigned if that is
    unobserved too; fp16(1e4) where no site exists.  wp.round = round half away from zero."""
    r = np.asarray(result)
    nx, ny, nz = r.shape
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    sx, sy, sz = unpack(r)
    dx, dy, dz = (x - sx).astype(np.float32), (y - sy).astype(np.float32), (z - sz).astype(np.float32)
    dist = np.sqrt(dx * dx + dy * dy + dz * dz).astype(np.float32)
    edt = (dist * np.float32(voxel_size)).astype(np.float32)
    src = np.full(r.shape, np.float32(1e10), np.float32)
    if static_sdf is not None and skip_steps > 0.0:
        far = (r >= 0) & (dist > 1.0)
        inv = np.where(far, np.float32(1.0) / np.where(far, dist, 1).astype(np.float32), 0).astype(np.float32)
        k = np.float32(skip_steps)
        ax = sx + _round_half_away(dx * inv * k).astype(np.int64)
        ay = sy + _round_half_away(dy * inv * k).astype(np.int64)
        az = sz + _round_half_away(dz * inv * k).astype(np.int64)
        ok = far & (ax >= 0) & (ax < nx) & (ay >= 0) & (ay < ny) & (az >= 0) & (az < nz)
        st = np.asarray(static_sdf, np.float32)
        src = np.where(ok, st[np.clip(ax, 0, nx - 1), np.clip(ay, 0, ny - 1), np.clip(az, 0, nz - 1)], src)
    if combined_sdf is not None:
        src = np.where(src > 1e9, np.asarray(combined_sdf, np.float32), src)
    signed = np.where(~(src > 1e9) & (src < 0), -edt, edt)
    return np.where(r < 0, np.float32(1e4), signed).astype(np.float16)


def tsdf_integrate_depth(block_data, voxel_size, origin, intrinsics, cam_positions, cam_quaternions, depth_images, depth_min,
                         depth_max, truncation):
    """Dense restatement of integrate_voxels_kernel (perception/mapper/kernel/builder/builder_camera_integrate.py:399-489) --
    TEST INFRASTRUCTURE.  block_data [nx, ny, nz, 2] float16 = (sum sdf * w, sum w), returned updated (a copy).
    Voxel centre (idx + 0.5 - n / 2) voxel + origin (builder_coord.py:57-66); camera frame by wp.quat_rotate(quat_inverse(q), .)
    = v (2 w^2 - 1) + 2 w (qv x v) + 2 qv (qv . v); pixel index by truncation towards zero; sdf = depth - z_cam kept when
    >= -truncation, clamped to +truncation; weight = max((fx voxel / z)(fy voxel / z), 1) (compute_tsdf_weight == 1,
    wp_integrate_common.py:57-105); fp32 accumulation over the cameras, one fp16 rounding per call.  All arithmetic in float32 in
    the kernel's order.  Pinned on the reference's own kernel source executed under the Warp stand-in
    (tests/golden/make_tsdf_golden.py -> tsdf_reference_golden.npz) and on a closed-form scene (tests/test_edt_cpu.py)."""
    f = np.float32
    bd = np.array(block_data, dtype=np.float16, copy=True)
    nx, ny, nz = bd.shape[:3]
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    wx = (ix.astype(f) + f(0.5) - f(nx) * f(0.5)) * f(voxel_size) + f(origin[0])
    wy = (iy.astype(f) + f(0.5) - f(ny) * f(0.5)) * f(voxel_size) + f(origin[1])
    wz = (iz.astype(f) + f(0.5) - f(nz) * f(0.5)) * f(voxel_size) + f(origin[2])
    tot_sw = np.zeros((nx, ny, nz), f)
    tot_w = np.zeros((nx, ny, nz), f)
    C, H, W = depth_images.shape
    for c in range(C):
        K = np.asarray(intrinsics[c], f)
        cp = np.asarray(cam_positions[c], f)
        q = np.asarray(cam_quaternions[c], f)                                 # wxyz
        qx, qy, qz, qw = -q[1], -q[2], -q[3], q[0]
        vx, vy, vz = wx - cp[0], wy - cp[1], wz - cp[2]
        cc = f(2.0) * qw * qw - f(1.0)
        crx, cry, crz = qy * vz - qz * vy, qz * vx - qx * vz, qx * vy - qy * vx
        d = qx * vx + qy * vy + qz * vz
        xc = vx * cc + crx * qw * f(2.0) + qx * d * f(2.0)
        yc = vy * cc + cry * qw * f(2.0) + qy * d * f(2.0)
        zc = vz * cc + crz * qw * f(2.0) + qz * d * f(2.0)
        ok = zc > f(depth_min)
        zs = np.where(ok, zc, f(1.0))
        fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        u = fx * xc / zs + cx
        v = fy * yc / zs + cy
        with np.errstate(invalid="ignore"):
            px = np.trunc(np.clip(u, -1e9, 1e9)).astype(np.int64)
            py = np.trunc(np.clip(v, -1e9, 1e9)).astype(np.int64)
        ok &= (px >= 0) & (px < W) & (py >= 0) & (py < H)
        depth = np.asarray(depth_images[c], f)[np.clip(py, 0, H - 1), np.clip(px, 0, W - 1)]
        ok &= (depth >= f(depth_min)) & (depth <= f(depth_max))
        sdf = depth - zc
        ok &= sdf >= -f(truncation)
        sdf_c = np.minimum(sdf, f(truncation))
        cov = (fx * f(voxel_size) / zs) * (fy * f(voxel_size) / zs)
        w = np.maximum(cov, f(1.0))
        tot_sw += np.where(ok, sdf_c * w, f(0.0)).astype(f)
        tot_w += np.where(ok, w, f(0.0)).astype(f)
    upd = tot_w > 0
    bd[..., 0] = np.where(upd, (bd[..., 0].astype(f) + tot_sw).astype(np.float16), bd[..., 0])
    bd[..., 1] = np.where(upd, (bd[..., 1].astype(f) + tot_w).astype(np.float16), bd[..., 1])
    return bd


def tsdf_combined_sdf(block_data, static_sdf, min_weight):
    """sample_combined_sdf (perception/mapper/kernel/wp_tsdf_sample.py:22-97) for the dense grid: float32 [nx, ny, nz]."""
    f = np.float32
    sw, w = block_data[..., 0].astype(f), block_data[..., 1].astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.where(w > f(min_weight), sw / np.where(w > f(min_weight), w, f(1.0)), f(1e10)).astype(f)
    if static_sdf is not None:
        d = np.minimum(d, np.asarray(static_sdf, f))
    return d


def tsdf_stamp_cuboids(static_sdf, voxel_size, origin, truncation, dims, inv_pose, enable, count, max_n, env=0):
    """Dense restatement of stamp_sdf_kernel (perception/mapper/kernel/builder/builder_stamp.py:263-315) with the cuboid overloads
    of geom/data/data_cuboid.py:461-545 -- TEST INFRASTRUCTURE.  static_sdf float32 [nx, ny, nz] (> 1e9 = nothing stamped), returned
    updated (a copy): where |min over enabled cuboids of the box SDF at the voxel centre| <= truncation the voxel takes
    clamp(min(existing, min), +-truncation) rounded through fp16.  dims [n_env * max_n, 4], inv_pose [n_env * max_n, 8]
    (x y z qw qx qy qz), enable, count: the CuboidWorld arrays.  Pinned on the reference's kernel source under the Warp stand-in
    (tests/golden/make_esdf_golden.py)."""
    f = np.float32
    out = np.array(static_sdf, dtype=f, copy=True)
    nx, ny, nz = out.shape
    vs, tr = f(voxel_size), f(truncation)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    c = np.stack([(ix.astype(f) + f(0.5) - f(nx) * f(0.5)) * vs + f(origin[0]), (iy.astype(f) + f(0.5) - f(ny) * f(0.5)) * vs + f(origin[1]),
                  (iz.astype(f) + f(0.5) - f(nz) * f(0.5)) * vs + f(origin[2])], -1).astype(f)
    dims = np.asarray(dims, f).reshape(-1, 4)
    inv = np.asarray(inv_pose, f).reshape(-1, 8)
    en = np.asarray(enable).reshape(-1)
    mn = np.full((nx, ny, nz), f(1e10), f)
    for k in range(int(np.asarray(count).reshape(-1)[env])):
        kk = env * int(max_n) + k
        if en[kk] != 1:
            continue
        p, q = inv[kk, :3], inv[kk, 3:7]                                       # wp.transform_point(t, x) = quat_rotate(q, x) + p
        qv, qw = np.array([q[1], q[2], q[3]], f), q[0]
        cc = f(2.0) * qw * qw - f(1.0)
        cr = np.cross(np.broadcast_to(qv, c.shape), c).astype(f)
        d = (c @ qv).astype(f)[..., None]
        lp = (c * cc + cr * qw * f(2.0) + qv * d * f(2.0)).astype(f) + p
        h = dims[kk, :3] * f(0.5)
        qd = np.abs(lp) - h
        od = np.sqrt((np.maximum(qd, f(0.0)) ** 2).sum(-1)).astype(f)
        sdf = (od + np.minimum(qd.max(-1), f(0.0))).astype(f)
        mn = np.minimum(mn, sdf)
    upd = np.abs(mn) <= tr
    fin = np.clip(np.minimum(out, mn), -tr, tr).astype(np.float16).astype(f)
    out[upd] = fin[upd]
    return out
