"""Generate tests/golden/warp_reference_golden.npz: outputs of the REFERENCE's own Warp kernel sources (scene collision, swept
collision, speed metric, tool pose, c-space STATE / POSITION costs), executed on the CPU thread by thread under the pure-Python
Warp stand-in (oracle/warp_shim), on seeded inputs.  The fixture stores inputs and outputs; tests/test_warp_reference_golden_cpu.py
replays the inputs through oracle/rollout_oracle.py and compares.  Needs /root/reference (authoring container only):

    python tests/golden/make_warp_golden.py
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

from curobo_b200.robot_model import load_robot  # noqa: E402
from curobo_b200.world import CuboidWorld, VoxelWorld, make_benchmark_cuboid_world, make_single_box_esdf  # noqa: E402

OUT = {}


def put(case, **arrays):
    for k, v in arrays.items():
        OUT[f"{case}/{k}"] = np.asarray(v)


def A(x, dtype=None):
    x = np.ascontiguousarray(x)
    if dtype is None:
        dtype = {np.dtype(np.float32): wp.float32, np.dtype(np.int32): wp.int32, np.dtype(np.uint8): wp.uint8,
                 np.dtype(np.float16): wp.float16}[x.dtype]
    if dtype in (wp.vec3, wp.vec4):
        return wp.from_numpy(x.reshape(-1, dtype.n), dtype=dtype)
    return wp.from_numpy(x.reshape(-1), dtype=dtype)


# ------------------------------------------------------------------------------------------------ c-space costs
def gen_cspace_state():
    m = R.ref("curobo._src.cost.wp_cspace_state")
    rm = load_robot("franka")
    rng = np.random.default_rng(3)
    B, H, D = 4, 5, 7
    q = rng.uniform(rm.position_limits[0] - 0.15, rm.position_limits[1] + 0.15, size=(B, H, D)).astype(np.float32)
    v, a, j = [rng.normal(0, s, size=(B, H, D)).astype(np.float32) for s in (2.0, 12.0, 400.0)]
    tau = rng.normal(0, 60.0, size=(B, H, D)).astype(np.float32)
    dt = rng.uniform(0.02, 0.2, size=B).astype(np.float32)
    lim = [np.asarray(x, np.float32) for x in (rm.position_limits, rm.velocity_limits, rm.acceleration_limits, rm.jerk_limits,
                                               rm.effort_limits)]
    w = np.array([10000.0, 10000.0, 100.0, 50.0, 30.0], np.float32)
    act = np.array([0.01, 0.01, 0.01, 0.01, 0.05], np.float32)
    reg = np.array([1000.0, 10000.0, 5.0, 0.02, 0.4], np.float32)
    tgt = rng.uniform(-1, 1, size=(2, D)).astype(np.float32)
    it = (np.arange(B) % 2).astype(np.int32)
    tw, ntf, dofw = np.array([3.0], np.float32), np.array([0.5], np.float32), rng.uniform(0.5, 1.5, D).astype(np.float32)
    for retime in (True, False):
        outs = [A(np.zeros(B * H * D, np.float32)) for _ in range(6)]
        args = [A(q), A(v), A(a), A(j), A(tau), A(dt), A(tgt), A(it)] + [A(x) for x in lim] + [A(w), A(act), A(reg), A(tw), A(ntf),
                                                                                             A(dofw)] + outs + [1, B, H, D, retime, retime]
        wp.launch(m.forward_cspace_state_warp, dim=B * H * D, inputs=args)
        put(f"cspace_state_retime{int(retime)}", q=q, v=v, a=a, j=j, tau=tau, dt=dt, lim_p=lim[0], lim_v=lim[1], lim_a=lim[2],
            lim_j=lim[3], lim_tau=lim[4], weight=w, act=act, reg=reg, target=tgt, idxs_target=it, target_weight=tw, ntf=ntf,
            dof_weight=dofw, cost=outs[0].data.reshape(B, H, D), **{f"grad_{n}": o.data.reshape(B, H, D) for n, o in
                                                                   zip("pvajt", outs[1:])})


def gen_cspace_position():
    m = R.ref("curobo._src.cost.wp_cspace_position")
    rm = load_robot("franka")
    rng = np.random.default_rng(4)
    B, H, D = 6, 1, 7
    q = rng.uniform(rm.position_limits[0] - 0.2, rm.position_limits[1] + 0.2, size=(B, H, D)).astype(np.float32)
    lim_p, lim_tau, lim_v = (np.asarray(x, np.float32) for x in (rm.position_limits, rm.effort_limits, rm.velocity_limits))
    w, act = np.array([5000.0, 0.0], np.float32), np.array([0.01, 0.01], np.float32)
    tgt = rng.uniform(-1, 1, size=(2, D)).astype(np.float32)
    it = (np.arange(B) % 2).astype(np.int32)
    tw, dofw = np.array([2.0], np.float32), rng.uniform(0.0, 1.5, D).astype(np.float32)
    z = np.zeros((B, H, D), np.float32)
    outs = [A(np.zeros(B * H * D, np.float32)) for _ in range(3)]
    args = [A(q), A(z), A(tgt), A(it), A(lim_p), A(lim_tau), A(w), A(act), A(tw), A(dofw), A(np.zeros(2, np.float32)),
            A(np.zeros((1, D), np.float32)), A(np.zeros((1, D), np.float32)), A(np.zeros(B, np.int32)), A(lim_v),
            A(np.zeros(B, np.float32))] + outs + [1, B, H, D]
    wp.launch(m.forward_cspace_position_warp, dim=B * H * D, inputs=args)
    put("cspace_position", q=q, lim_p=lim_p, weight=w, act=act, target=tgt, idxs_target=it, target_weight=tw, dof_weight=dofw,
        cost=outs[0].data.reshape(B, H, D), grad_p=outs[1].data.reshape(B, H, D))


# ------------------------------------------------------------------------------------------------ tool pose
def gen_tool_pose():
    m = R.ref("curobo._src.cost.wp_tool_pose")
    rng = np.random.default_rng(9)
    B, H, L, G, NG = 4, 3, 2, 3, 3
    pos = rng.normal(size=(B, H, L, 3)).astype(np.float32)
    quat = rng.normal(size=(B, H, L, 4)).astype(np.float32)
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    gpos = rng.normal(size=(G, L, NG, 3)).astype(np.float32)
    gquat = rng.normal(size=(G, L, NG, 4)).astype(np.float32)
    gquat /= np.linalg.norm(gquat, axis=-1, keepdims=True)
    # a few near-goal poses so that the convergence-tolerance branch is exercised
    pos[0, -1, 0] = gpos[1, 0, 2] + 1e-4
    quat[0, -1, 0] = gquat[1, 0, 2]
    idx = rng.integers(0, G, size=B).astype(np.int32)
    idx[0] = 1
    w = np.array([1000.0, 30.0], np.float32)
    at = rng.uniform(0.2, 1.5, size=(L, 6)).astype(np.float32)
    ant = rng.uniform(0.0, 1.0, size=(L, 6)).astype(np.float32)
    tt, tnt = np.full((L, 2), 1e-3, np.float32), np.full((L, 2), 1e-2, np.float32)
    N = B * H * L
    for method in (0, 1):
        kern = m.create_goalset_pose_distance_kernel_with_constants(NG, method)
        od, opd, ord_ = A(np.zeros(2 * N, np.float32)), A(np.zeros(N, np.float32)), A(np.zeros(N, np.float32))
        opg, org, ogi = A(np.zeros((N, 3), np.float32), wp.vec3), A(np.zeros((N, 4), np.float32), wp.vec4), A(np.zeros(N, np.int32))
        args = [A(pos, wp.vec3), A(quat, wp.vec4), A(gpos, wp.vec3), A(gquat, wp.vec4), A(idx), A(w), A(at), A(ant), A(tt), A(tnt),
                A(np.zeros(L, np.uint8)), od, opd, ord_, opg, org, ogi, B, H, L]
        wp.launch(kern, dim=N, inputs=args)
        put(f"tool_pose_method{method}", pos=pos, quat=quat, goal_pos=gpos, goal_quat=gquat, idxs_goal=idx, weight=w, axes_t=at,
            axes_nt=ant, tol_t=tt, tol_nt=tnt, distance=od.data.reshape(B, H, 2 * L), pos_dist=opd.data.reshape(B, H, L),
            rot_dist=ord_.data.reshape(B, H, L), grad_pos=opg.data.reshape(B, H, L, 3), grad_quat=org.data.reshape(B, H, L, 4),
            goalset_idx=ogi.data.reshape(B, H, L))


# ------------------------------------------------------------------------------------------------ scene collision
def cuboid_struct(cw):
    dc = R.ref("curobo._src.geom.data.data_cuboid")
    s = dc.CuboidDataWarp()
    s.dims = wp.from_numpy(cw.dims.reshape(-1, 4), dtype=wp.float32, ndim=2)
    s.inv_pose = wp.from_numpy(cw.inv_pose.reshape(-1, 8), dtype=wp.float32, ndim=2)
    s.enable = wp.from_numpy(cw.enable.reshape(-1), dtype=wp.uint8)
    s.n_per_env = wp.from_numpy(cw.count.reshape(-1), dtype=wp.int32)
    s.max_n, s.num_envs = wp.int32(cw.max_n), wp.int32(cw.num_envs)
    return s


def voxel_struct(vw):
    dv = R.ref("curobo._src.geom.data.data_voxel")
    s = dv.VoxelDataWarp()
    p = vw.params.reshape(-1, 4)
    dims = np.zeros_like(p)
    dims[:, :3] = p[:, :3] * p[:, 3:4]
    s.params = wp.from_numpy(p, dtype=wp.float32, ndim=2)
    s.dims = wp.from_numpy(dims, dtype=wp.float32, ndim=2)
    s.inv_pose = wp.from_numpy(vw.inv_pose.reshape(-1, 8), dtype=wp.float32, ndim=2)
    s.enable = wp.from_numpy(vw.enable.reshape(-1), dtype=wp.uint8)
    s.features = wp.from_numpy(vw.features.reshape(-1), dtype=wp.float16)
    s.n_per_env = wp.from_numpy(vw.count.reshape(-1), dtype=wp.int32)
    s.n_voxels_per_layer, s.max_n, s.num_envs = wp.int32(vw.features.shape[2]), wp.int32(vw.max_n), wp.int32(vw.num_envs)
    s.max_dist = wp.float32(vw.max_dist)
    return s


def run_collision(sph, weight, eta, structs, env_idx, sweep, speed_dt):
    ck = R.ref("curobo._src.geom.collision.wp_collision_kernel")
    sk = R.ref("curobo._src.geom.collision.wp_sweep_collision_kernel")
    sm = R.ref("curobo._src.geom.collision.wp_speed_metric")
    B, H, S, _ = sph.shape
    dist, grad = A(np.zeros(B * H * S, np.float32)), A(np.zeros(B * H * S * 4, np.float32))
    sp = A(sph, wp.vec4)
    multi = wp.uint8(1 if env_idx is not None else 0)
    env = A(np.zeros(B, np.int32) if env_idx is None else np.asarray(env_idx, np.int32))
    kern = sk.swept_sphere_obstacle_collision_kernel if sweep else ck.sphere_obstacle_collision_kernel
    for st, max_n in structs:                              # the reference launches once per obstacle type, accumulating
        wp.launch(kern, dim=B * H * S * max_n, inputs=[st, sp, A(np.array([weight], np.float32)), A(np.array([eta], np.float32)),
                                                       env, dist, grad, B, H, S, max_n, multi])
    if speed_dt is not None:
        wp.launch(sm.apply_speed_metric, dim=B * H * S, inputs=[sp, dist, grad, A(np.array([speed_dt], np.float32)), B, H, S])
    return dist.data.reshape(B, H, S).copy(), grad.data.reshape(B, H, S, 4).copy()


def save_world(case, cub, vox):
    if cub is not None:
        put(case, cub_dims=cub.dims, cub_inv_pose=cub.inv_pose, cub_enable=cub.enable, cub_count=cub.count)
    if vox is not None:
        put(case, vox_params=vox.params, vox_inv_pose=vox.inv_pose, vox_enable=vox.enable, vox_count=vox.count,
            vox_features=vox.features, vox_max_dist=np.float32(vox.max_dist))


def trajectory_spheres(rng, B, H, S, lo, hi, step):
    sph = np.zeros((B, H, S, 4), np.float32)
    sph[:, 0, :, :3] = rng.uniform(lo, hi, size=(B, S, 3))
    for h in range(1, H):
        sph[:, h, :, :3] = sph[:, h - 1, :, :3] + rng.normal(0, step, size=(B, S, 3))
    sph[..., 3] = rng.uniform(0.02, 0.07, size=(B, 1, S))
    sph[0, :, 1, 3] = -1.0                                  # a disabled sphere
    return sph


def gen_collision():
    rng = np.random.default_rng(0)
    w, eta = 100.0, 0.02
    cub = CuboidWorld.create([{"dims": [2.2, 2.2, 0.2], "pose": [0, 0, -0.1, 1, 0, 0, 0]},
                              {"dims": [0.1, 0.1, 1.5], "pose": [0.45, 0, 0.3, 1, 0, 0, 0]},
                              {"dims": [0.3, 0.2, 0.25], "pose": [0.2, -0.3, 0.5, 0.8, 0.2, -0.4, 0.4]}], max_n=5)
    vox = make_single_box_esdf(grid_center=(0.35, 0.25, 0.45))
    # 1. discrete, cuboids + ESDF together
    sph = trajectory_spheres(rng, 2, 1, 30, [-0.1, -0.5, -0.2], [0.8, 0.5, 1.0], 0.0)
    d, g = run_collision(sph, w, eta, [(cuboid_struct(cub), cub.max_n), (voxel_struct(vox), vox.max_n)], None, False, None)
    put("collision_discrete", spheres=sph, weight=np.float32(w), eta=np.float32(eta), cost=d, grad=g)
    save_world("collision_discrete", cub, vox)
    # 2. two environments: env 1 has a wall and a disabled cuboid
    c2 = CuboidWorld.create([{"dims": [0.1, 1.0, 1.0], "pose": [0.3, 0, 0.5, 1, 0, 0, 0]},
                             {"dims": [5.0, 5.0, 5.0], "pose": [0, 0, 0, 1, 0, 0, 0]}], max_n=5, num_envs=2)
    c2.dims[0], c2.inv_pose[0], c2.enable[0], c2.count[0] = cub.dims[0], cub.inv_pose[0], cub.enable[0], cub.count[0]
    c2.enable[1, 1] = 0
    sph = trajectory_spheres(rng, 3, 1, 20, [-0.1, -0.5, -0.2], [0.8, 0.5, 1.0], 0.0)
    env = np.array([1, 0, 1], np.int32)
    d, g = run_collision(sph, w, eta, [(cuboid_struct(c2), c2.max_n)], env, False, None)
    put("collision_multi_env", spheres=sph, weight=np.float32(w), eta=np.float32(eta), env_query_idx=env, cost=d, grad=g)
    save_world("collision_multi_env", c2, None)
    # 3. swept (cuboids + ESDF), then 4. the same with the speed metric
    sph = trajectory_spheres(rng, 2, 6, 12, [0.0, -0.4, 0.0], [0.7, 0.4, 0.9], 0.05)
    for case, sdt in (("collision_swept", None), ("collision_swept_speed", 0.05)):
        d, g = run_collision(sph, w, eta, [(cuboid_struct(cub), cub.max_n), (voxel_struct(vox), vox.max_n)], None, True, sdt)
        put(case, spheres=sph, weight=np.float32(w), eta=np.float32(eta), cost=d, grad=g,
            **({"speed_dt": np.float32(sdt)} if sdt is not None else {}))
        save_world(case, cub, vox)


def gen_collision_edges():
    """Branch coverage: sphere centres inside cuboids (inside-gradient branch), on faces / edges, far away; an ESDF grid with a
    rotated, offset pose, spheres outside the grid, on its border voxels and deep inside the box; large radii (both activation
    regimes: pen < eta quadratic, pen >= eta linear); stationary and fast-moving trajectories for the sweep."""
    rng = np.random.default_rng(5)
    w, eta = 250.0, 0.04
    cub = CuboidWorld.create([{"dims": [0.4, 0.3, 0.2], "pose": [0.1, 0.0, 0.1, 0.9, 0.1, 0.3, -0.2]},
                              {"dims": [0.2, 0.6, 0.3], "pose": [-0.2, 0.2, 0.0, 0.6, -0.5, 0.4, 0.3]}], max_n=3)
    vox = make_single_box_esdf(grid_dims=(0.6, 0.5, 0.4), voxel_size=0.025, grid_center=(0.05, -0.1, 0.15),
                               box_half=(0.12, 0.1, 0.08), pose_quat=(0.85, 0.2, -0.3, 0.35)) if "pose_quat" in \
        make_single_box_esdf.__code__.co_varnames else None
    if vox is None:                                          # build the rotated grid by hand from an axis-aligned one
        base = make_single_box_esdf(grid_dims=(0.6, 0.5, 0.4), voxel_size=0.025)
        from curobo_b200.world import _inv_pose_from_pose
        vox = VoxelWorld(base.params, _inv_pose_from_pose([0.05, -0.1, 0.15, 0.85, 0.2, -0.3, 0.35]).reshape(1, 1, 8), base.enable,
                         base.count, base.features, base.max_dist)
    B, H, S = 2, 5, 26
    sph = np.zeros((B, H, S, 4), np.float32)
    centres = np.concatenate([rng.uniform(-0.45, 0.45, size=(S - 8, 3)),
                              np.array([[0.1, 0.0, 0.1], [0.12, 0.02, 0.11], [-0.2, 0.2, 0.0], [0.05, -0.1, 0.15],
                                        [0.06, -0.09, 0.16], [2.0, 2.0, 2.0], [0.3, 0.0, 0.1], [0.05, -0.1, 0.33]])]).astype(np.float32)
    sph[0, :, :, :3] = centres[None]                         # batch 0: stationary (swept == discrete there)
    sph[1, 0, :, :3] = centres
    for h in range(1, H):
        sph[1, h, :, :3] = sph[1, h - 1, :, :3] + rng.normal(0, 0.12, size=(S, 3))   # batch 1: long jumps between waypoints
    sph[..., 3] = rng.uniform(0.01, 0.15, size=(1, 1, S))
    structs = [(cuboid_struct(cub), cub.max_n), (voxel_struct(vox), vox.max_n)]
    for case, sweep, sdt in (("collision_edge_discrete", False, None), ("collision_edge_swept", True, None),
                             ("collision_edge_swept_speed", True, 0.02)):
        d, g = run_collision(sph, w, eta, structs, None, sweep, sdt)
        put(case, spheres=sph, weight=np.float32(w), eta=np.float32(eta), cost=d, grad=g,
            **({"speed_dt": np.float32(sdt)} if sdt is not None else {}))
        save_world(case, cub, vox)


if __name__ == "__main__":
    import time
    for fn in (gen_cspace_state, gen_cspace_position, gen_tool_pose, gen_collision, gen_collision_edges):
        t = time.time()
        fn()
        print(f"{fn.__name__}: {time.time() - t:.1f} s")
    path = os.path.join(HERE, "warp_reference_golden.npz")
    np.savez_compressed(path, **OUT)
    print(f"wrote {path}: {len(OUT)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
