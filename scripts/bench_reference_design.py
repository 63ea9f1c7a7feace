#!/usr/bin/env python
"""Side measurement (NOT bench.py, not the product): the REFERENCE DESIGN on the same B200.

BASELINE.md section 2b / SURVEY.md 8d "second, stronger baseline": the reference's own CUDA kernels compiled for
sm_100a (oracle/_ref: kinematics_forward_spheres_kernel, self_collision_max_distance_kernel,
kinematics_backward_kernel) chained UNFUSED with this repository's stand-alone restatements of the Warp kernels that
cannot be compiled here (scene collision, tool pose, c-space) plus the torch glue the reference's autograd runs
(gradient accumulation, cost sum) -- the same data flow through HBM as RobotRollout.evaluate_action + backward,
minus the reference's memsets / extra streams (so it flatters the reference design).  The chain is captured in a CUDA
graph, like the reference runs it.  Next to it: the fused kernel on the same inputs, and the max difference.

    python scripts/bench_reference_design.py            # prints one JSON line per workload
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import bench  # noqa: E402
import ref_kernels  # noqa: E402
from curobo_b200 import cost as cb_cost  # noqa: E402
from curobo_b200.kinematics import KinematicsParams  # noqa: E402
from curobo_b200.scene import CollisionBuffer, SceneData, _launch as scene_launch  # noqa: E402


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def run(workload, steps=200, warmup=10, device="cuda:0", quiet=False):
    dev = torch.device(device)
    wl = bench.make_workload(workload)
    rm, cfg = wl["robot"], wl["cfg"]
    eng = bench.build_engine(wl, dev)
    B, H = wl["B"], wl["H"]
    N, D, S, L, nl = B * H, rm.num_dof, rm.num_spheres, rm.num_tool_frames, rm.num_links
    q = torch.as_tensor(wl["q"]).to(dev).reshape(N, D).contiguous()
    kp = KinematicsParams.from_robot_model(rm, dev)
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
    lib = ref_kernels.lib()
    # buffers (allocated once, like the reference's create_buffers)
    pos, quat, sph, com, cum = z(N, L, 3), z(N, L, 4), z(N, S, 4), z(N, 4), z(N, nl, 3, 4)
    eq = z(1, dt=torch.int32)
    sc_dist, sc_vec, sc_sparse = z(B, H), z(B, H, S, 4), z(B, H, S, dt=torch.uint8)
    nb = rm.num_blocks_per_batch
    sc_pd, sc_bv, sc_bi = z(1), z(B, H, nb), z(B, H, nb, 2, dt=torch.int16)
    w_self = torch.tensor([cfg.self_weight], device=dev)
    padding = torch.as_tensor(rm.sphere_padding).to(dev)
    pairs = torch.as_tensor(rm.collision_pairs).to(dev)
    cbuf = CollisionBuffer.from_shape((B, H, S, 4), dev)
    scene = SceneData(eng.cuboid, eng.voxel)
    w_scene = torch.tensor([cfg.scene_weight], device=dev)
    eta = torch.tensor([cfg.scene_activation], device=dev)
    env0 = z(B, dt=torch.int32)
    has_pose = wl["goal"] is not None and cfg.pose_weight is not None
    if has_pose:
        gp, gq, gidx = (torch.as_tensor(x).to(dev) for x in wl["goal"])
        gidx = gidx.view(B, 1).contiguous()
        pw = torch.tensor(list(cfg.pose_weight), device=dev)
        ones6, zeros2 = torch.ones((L, 6), device=dev), z(L, 2)
        proj = z(L, 1, dt=torch.uint8)
        p_cost, p_pd, p_rd = z(B, H, 2 * L), z(B, H, L), z(B, H, L)
        p_gp, p_gq, p_gi = z(B, H, L, 3), z(B, H, L, 4), z(B, H, L, dt=torch.int32)
    g_pos0, g_quat0 = z(N, L, 3), z(N, L, 4)
    lim = torch.as_tensor(rm.position_limits).to(dev).contiguous()
    cs_w = torch.tensor(list(cfg.cspace_weight[:2]), device=dev)
    cs_a = torch.tensor(list(cfg.cspace_activation[:2]), device=dev)
    cs_cost, cs_gp, cs_gt = z(B, H, D), z(B, H, D), z(B, H, D)
    zeros_bhd, zi, zd = z(B, H, D), z(B, dt=torch.int32), z(1, D)
    efl = torch.as_tensor(rm.effort_limits).to(dev).contiguous()
    vlim = torch.as_tensor(rm.velocity_limits).to(dev).contiguous()
    tw1, tdw, z2, z1 = z(1), torch.ones(D, device=dev), z(2), z(1)
    g_sph, g_q, g_com = z(N, S, 4), z(N, D), z(N, 4)
    total_cost, grad_q = z(B, H), z(B, H, D)
    q3 = q.view(B, H, D)

    def unfused():
        st = _stream(dev)
        err = lib.ref_kinematics_forward_spheres(
            _p(pos), _p(quat), _p(sph), _p(com), _p(cum), _p(q), _p(kp.fixed_transforms), _p(kp.link_spheres),
            _p(kp.link_masses_com), _p(kp.joint_map_type), _p(kp.joint_map), _p(kp.link_map), _p(kp.tool_frame_map),
            _p(kp.link_sphere_idx_map), _p(kp.joint_offset_map), _p(eq), kp.num_envs, N, 1, kp.num_dof, kp.num_spheres,
            kp.num_links, kp.num_pose_links, st)
        assert err == 0
        sph4 = sph.view(B, H, S, 4)
        err = lib.ref_self_collision_distance(_p(sc_dist), _p(sc_vec), _p(sc_pd), _p(sc_sparse), _p(sph4), _p(padding),
                                              _p(w_self), _p(pairs), _p(sc_bv), _p(sc_bi), nb, rm.max_threads_per_block, B, H,
                                              S, pairs.shape[0], 1, st)
        assert err == 0
        scene_launch(cfg.use_sweep, sph4, cbuf, scene, w_scene, eta, None, False, env0, False)
        if has_pose:
            cb_cost.tool_pose_distance(pos.view(B, H, L, 3), quat.view(B, H, L, 4), gp, gq, gidx, pw, ones6, ones6, zeros2,
                                       zeros2, proj, p_cost, p_pd, p_rd, p_gp, p_gq, p_gi, cfg.pose_lie)
        cb_cost.cspace_position_cost(q3, zeros_bhd, zd, zi, lim, efl, cs_w, cs_a, tw1, tdw, z2, zd, zd, zi, vlim, z1, cs_cost,
                                     cs_gp, cs_gt)
        torch.add(sc_vec.view(N, S, 4), cbuf.gradient.view(N, S, 4), out=g_sph)          # autograd accumulation
        err = lib.ref_kinematics_backward(
            _p(g_q), _p(p_gp.view(N, L, 3) if has_pose else g_pos0), _p(p_gq.view(N, L, 4) if has_pose else g_quat0),
            _p(g_sph), _p(g_com), _p(g_com), None, _p(cum), _p(kp.link_spheres), _p(kp.link_masses_com), _p(kp.link_map),
            _p(kp.joint_map), _p(kp.joint_map_type), _p(kp.tool_frame_map), _p(kp.link_sphere_idx_map),
            _p(kp.link_chain_data), _p(kp.link_chain_offsets), _p(kp.joint_links_data), _p(kp.joint_links_offsets),
            _p(kp.joint_affects_endeffector), _p(kp.joint_offset_map), _p(eq), kp.num_envs, N, 1, kp.num_dof,
            kp.num_spheres, kp.num_links, kp.num_pose_links, st)
        assert err == 0
        torch.add(g_q.view(B, H, D), cs_gp, out=grad_q)
        c = sc_dist + cbuf.distance.sum(-1) + cs_cost.sum(-1)
        if has_pose:
            c = c + p_cost.sum(-1)
        total_cost.copy_(c)

    def time_graph(fn):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        for _ in range(warmup):
            g.replay()
        torch.cuda.synchronize()
        st_, en_ = [torch.cuda.Event(enable_timing=True) for _ in range(steps)], [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush.fill_(i & 0xFF)
            st_[i].record()
            g.replay()
            en_[i].record()
        torch.cuda.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in zip(st_, en_)]))

    if os.environ.get("CB200_STAGE_TIMES"):
        import time
        stages = {}
        def t_stage(name, fn, reps=200):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            stages[name] = a.elapsed_time(b) / reps * 1e3
        st = lambda: _stream(dev)  # noqa: E731
        sph4 = sph.view(B, H, S, 4)
        t_stage("ref_fk_forward", lambda: lib.ref_kinematics_forward_spheres(
            _p(pos), _p(quat), _p(sph), _p(com), _p(cum), _p(q), _p(kp.fixed_transforms), _p(kp.link_spheres),
            _p(kp.link_masses_com), _p(kp.joint_map_type), _p(kp.joint_map), _p(kp.link_map), _p(kp.tool_frame_map),
            _p(kp.link_sphere_idx_map), _p(kp.joint_offset_map), _p(eq), kp.num_envs, N, 1, kp.num_dof, kp.num_spheres,
            kp.num_links, kp.num_pose_links, st()))
        t_stage("ref_self_collision", lambda: lib.ref_self_collision_distance(
            _p(sc_dist), _p(sc_vec), _p(sc_pd), _p(sc_sparse), _p(sph4), _p(padding), _p(w_self), _p(pairs), _p(sc_bv),
            _p(sc_bi), nb, rm.max_threads_per_block, B, H, S, pairs.shape[0], 1, st()))
        t_stage("our_scene_per_op", lambda: scene_launch(cfg.use_sweep, sph4, cbuf, scene, w_scene, eta, None, False, env0, False))
        if has_pose:
            t_stage("our_tool_pose_per_op", lambda: cb_cost.tool_pose_distance(
                pos.view(B, H, L, 3), quat.view(B, H, L, 4), gp, gq, gidx, pw, ones6, ones6, zeros2, zeros2, proj, p_cost, p_pd,
                p_rd, p_gp, p_gq, p_gi, cfg.pose_lie))
        t_stage("our_cspace_per_op", lambda: cb_cost.cspace_position_cost(
            q3, zeros_bhd, zd, zi, lim, efl, cs_w, cs_a, tw1, tdw, z2, zd, zd, zi, vlim, z1, cs_cost, cs_gp, cs_gt))
        t_stage("torch_grad_add", lambda: torch.add(sc_vec.view(N, S, 4), cbuf.gradient.view(N, S, 4), out=g_sph))
        t_stage("ref_fk_backward", lambda: lib.ref_kinematics_backward(
            _p(g_q), _p(p_gp.view(N, L, 3) if has_pose else g_pos0), _p(p_gq.view(N, L, 4) if has_pose else g_quat0),
            _p(g_sph), _p(g_com), _p(g_com), None, _p(cum), _p(kp.link_spheres), _p(kp.link_masses_com), _p(kp.link_map),
            _p(kp.joint_map), _p(kp.joint_map_type), _p(kp.tool_frame_map), _p(kp.link_sphere_idx_map),
            _p(kp.link_chain_data), _p(kp.link_chain_offsets), _p(kp.joint_links_data), _p(kp.joint_links_offsets),
            _p(kp.joint_affects_endeffector), _p(kp.joint_offset_map), _p(eq), kp.num_envs, N, 1, kp.num_dof,
            kp.num_spheres, kp.num_links, kp.num_pose_links, st()))
        t_stage("our_fk_forward_per_op", lambda: __import__("curobo_b200.backends.kinematics", fromlist=["x"]).launch_kinematics_forward_spheres(
            pos.view(B, H, L, 3), quat.view(B, H, L, 4), sph4, None, cum.view(B, H, nl, 3, 4), q, kp.fixed_transforms, kp.link_spheres,
            None, kp.joint_map_type, kp.joint_map, kp.link_map, kp.tool_frame_map, kp.link_sphere_idx_map, kp.joint_offset_map,
            eq, 1, N, 1, D, S))
        print(json.dumps({"workload": workload, "stage_us_warm_L2": {k: round(v, 2) for k, v in stages.items()}}), flush=True)
    ms_unfused = time_graph(unfused)
    uc, ug = total_cost.clone(), grad_q.clone()
    ms_fused = time_graph(lambda: eng.evaluate_action(q3))
    fc, fg = eng.out.cost, eng.out.grad_q
    dc = float((uc - fc).abs().max() / fc.abs().max())
    dg = float((ug - fg).abs().max() / fg.abs().max())
    n_launch = 8 + (1 if has_pose else 0) + 3 + (1 if has_pose else 0)
    res = {"workload": workload, "rows": N, "reference_design_unfused_ms": ms_unfused,
           "reference_design_evals_per_s": N / (ms_unfused * 1e-3), "fused_ms": ms_fused,
           "fused_evals_per_s": N / (ms_fused * 1e-3), "speedup": ms_unfused / ms_fused,
           "launches_unfused_approx": n_launch, "launches_fused": 1, "max_rel_diff_cost": dc,
           "max_rel_diff_grad": dg, "timer": "cuda events around a CUDA-graph replay, L2 flushed between steps"}
    if not quiet:
        print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    if not ref_kernels.available():
        raise SystemExit("oracle/_ref/libcurobo_ref.so is missing (build it where /root/reference exists)")
    as_json = "--json" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["franka_ik_512x32_cuboid", "franka_16384_esdf", "g1_29_8192_esdf"]
    out = {}
    for w in names:
        r = run(w, quiet=as_json)
        out[w] = {"reference_design_ms": r["reference_design_unfused_ms"], "fused_ms": r["fused_ms"], "speedup": r["speedup"],
                  "max_rel_diff_cost": r["max_rel_diff_cost"], "max_rel_diff_grad": r["max_rel_diff_grad"]}
    if as_json:    # one line for bench.py's reference_design_gpu leg
        print(json.dumps({"what": "the reference's own CUDA kernels (sm_100a, oracle/_ref) chained unfused from a CUDA graph vs the "
                                  "fused kernel, same inputs, L2 flushed between steps", "workloads": out}), flush=True)
