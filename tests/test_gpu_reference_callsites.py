"""The drop-in, proven by EXECUTION: the reference's own autograd Functions -- curobo/_src/curobolib/cuda_ops/
{kinematics,geometry,trajectory,optimization,dynamics}.py, i.e. the call sites of the kernel-backend modules -- run their
forward / backward on top of `curobo_b200.backends` (the INTEGRATION.md section 1 overlay) and must give what the reference's
own compiled CUDA kernels (oracle/_ref/libcurobo_ref.so) and the oracle give.

The reference's Python is imported from oracle/_ref/pyref: byte code compiled from the sources where they lie under
/root/reference by the committed recipe oracle/build_pyref.py (a git-ignored build product like the compiled kernels; the GPU box
has no /root/reference).  Third-party packages the reference imports but this image lacks (warp, trimesh, ...) are stubbed by
tests/golden/_reference_under_shim.py; none of them is on these call paths.  Skipped when the byte code was not built."""
import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(PYREF, "MANIFEST.json")),
                                 reason="oracle/_ref/pyref not built (python oracle/build_pyref.py where /root/reference exists)")]

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_kernels  # noqa: E402
from bspline_cases import make_case  # noqa: E402
from helpers import random_q  # noqa: E402
from optim_cases import lbfgs_case  # noqa: E402
from oracle import bspline_oracle as bo  # noqa: E402
from oracle import optim_oracle as oo  # noqa: E402
from oracle import rollout_oracle as O  # noqa: E402

DEV = "cuda:0"


def T(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a)
    return (t if dt is None else t.to(dt)).to(DEV).clone()


T._copies = True


_MODS = None


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def load_reference():
    """The reference's modules with the b200 backend overlaid."""
    global _MODS
    if _MODS is not None:
        return _MODS
    import _reference_under_shim as shim
    shim.prepare(root=PYREF)
    manifest = json.load(open(os.path.join(PYREF, "MANIFEST.json")))
    assert manifest["python"].split(".")[:2] == sys.version.split()[0].split(".")[:2], "byte code of another Python"
    cb = importlib.import_module("curobo._src.curobolib.backends")
    assert os.path.abspath(cb.__file__).startswith(PYREF), "curobo must come from the byte-code build of the reference"
    from curobo_b200.backends import dynamics, geometry, kinematics, optimization, pba, trajectory
    cb._backend_modules = {"kinematics": kinematics, "geometry": geometry, "trajectory": trajectory,
                           "optimization": optimization, "dynamics": dynamics, "pba": pba}
    cb._backend_name = "b200"

    class Mods:
        kinematics = importlib.import_module("curobo._src.curobolib.cuda_ops.kinematics")
        geometry = importlib.import_module("curobo._src.curobolib.cuda_ops.geometry")
        trajectory = importlib.import_module("curobo._src.curobolib.cuda_ops.trajectory")
        optimization = importlib.import_module("curobo._src.curobolib.cuda_ops.optimization")
        lbfgs = importlib.import_module("curobo._src.optim.gradient.lbfgs")
        params = importlib.import_module("curobo._src.robot.types.kinematics_params")
        device_cfg = importlib.import_module("curobo._src.types.device_cfg")
    assert cb.get_backend_name() == "b200"
    # the proxies the call sites hold resolve to our modules
    assert Mods.kinematics.kinematics_cu.launch_kinematics_forward_spheres is kinematics.launch_kinematics_forward_spheres
    assert Mods.geometry.geometry_cu.self_collision_distance is geometry.self_collision_distance
    _MODS = Mods
    return Mods


def _reference_kinematics_params(ref, rm):
    """The reference's own KinematicsParams dataclass filled from our loader's tensors (robot/types/kinematics_params.py:23-158)."""
    ls = rm.link_spheres if rm.link_spheres.ndim == 3 else rm.link_spheres[None]
    dc = ref.device_cfg.DeviceCfg(device=torch.device(DEV))
    return ref.params.KinematicsParams(
        fixed_transforms=T(rm.fixed_transforms), link_map=T(rm.link_map), joint_map=T(rm.joint_map),
        joint_map_type=T(rm.joint_map_type), joint_offset_map=T(rm.joint_offset_map.reshape(-1)),
        tool_frame_map=T(rm.tool_frame_map), link_chain_data=T(rm.link_chain_data), link_chain_offsets=T(rm.link_chain_offsets),
        joint_links_data=T(rm.joint_links_data), joint_links_offsets=T(rm.joint_links_offsets),
        joint_affects_endeffector=T(rm.joint_affects_endeffector.astype(bool)),
        tool_frames=[f"tool_{i}" for i in range(rm.num_tool_frames)], joint_limits=None, non_fixed_joint_names=[],
        num_dof=rm.num_dof, link_spheres=T(ls), link_sphere_idx_map=T(rm.link_sphere_idx_map), total_spheres=rm.num_spheres,
        link_masses_com=T(rm.link_masses_com), device_cfg=dc)


@pytest.mark.parametrize("robot,n", [("franka", 64), ("g1_29", 24)])
def test_reference_kinematics_function_over_b200_backend(ref, robot, n):
    """KinematicsFusedFunction.apply (cuda_ops/kinematics.py:93-356): forward through launch_kinematics_forward_spheres, then
    loss.backward() through launch_kinematics_backward -- both resolved to curobo_b200.backends.kinematics."""
    from curobo_b200.robot_model import load_robot
    rm = load_robot(robot)
    kp = _reference_kinematics_params(ref, rm)
    F = ref.kinematics.KinematicsFusedFunction
    dc = ref.device_cfg.DeviceCfg(device=torch.device(DEV))
    buf = F.create_buffers(n, 1, kp, dc)
    qn = random_q(rm, n, seed=5)
    q = T(qn).view(n, 1, -1).requires_grad_(True)
    env = torch.zeros(1, dtype=torch.int32, device=DEV)
    pos, quat, sph, com, jac = F.apply(
        q, buf["batch_link_position"], buf["batch_link_quaternion"], buf["batch_robot_spheres"], buf["batch_com"],
        buf["batch_jacobian"], buf["batch_cumul_mat"], kp, buf["grad_out_q"], buf["grad_out_q_jacobian"],
        buf["grad_in_link_pos"], buf["grad_in_link_quat"], buf["grad_in_robot_spheres"], buf["grad_in_com"],
        False, True, False, env, 1)
    cum, want_sph, want_pos, want_quat = O.fk_forward(rm, qn)
    assert np.allclose(pos.detach().cpu().numpy().reshape(want_pos.shape), want_pos, atol=1e-5)
    assert np.allclose(sph.detach().cpu().numpy().reshape(want_sph.shape), want_sph, atol=1e-5)
    assert np.allclose(buf["batch_cumul_mat"].cpu().numpy().reshape(cum.shape), cum, atol=1e-5)
    qd = np.abs(np.sum(quat.detach().cpu().numpy().reshape(want_quat.shape) * want_quat, axis=-1))
    assert np.allclose(qd, 1.0, atol=1e-5)
    rng = np.random.default_rng(1)
    gs = rng.normal(size=want_sph.shape).astype(np.float32)
    gp = rng.normal(size=want_pos.shape).astype(np.float32)
    gq = rng.normal(size=want_quat.shape).astype(np.float32)
    loss = (sph.view(want_sph.shape) * T(gs)).sum() + (pos.view(want_pos.shape) * T(gp)).sum() + (quat.view(want_quat.shape) * T(gq)).sum()
    loss.backward()
    want_g = O.fk_backward(rm, cum, gs, gp, gq)
    g = q.grad.cpu().numpy().reshape(want_g.shape)
    assert np.allclose(g, want_g, rtol=1e-3, atol=1e-5 * np.abs(want_g).max())
    if ref_kernels.available():   # and against the reference's own compiled kernels on the same inputs
        from curobo_b200.kinematics import KinematicsParams
        okp = KinematicsParams.from_robot_model(rm, DEV)
        rp, rq, rs, rc = ref_kernels.fk_forward(okp, T(qn))
        assert np.allclose(sph.detach().cpu().numpy().reshape(want_sph.shape), rs.cpu().numpy().reshape(want_sph.shape), atol=1e-5)
        rg = ref_kernels.fk_backward(okp, T(cum), T(gp), T(gq), T(gs)).cpu().numpy().reshape(want_g.shape)
        assert np.allclose(g, rg, rtol=1e-3, atol=1e-5 * np.abs(want_g).max())


@pytest.mark.parametrize("robot,n", [("franka", 48), ("g1_43", 6)])
def test_reference_self_collision_function_over_b200_backend(ref, robot, n):
    """SelfCollisionDistance.apply (cuda_ops/geometry.py:18-104) -> geometry_cu.self_collision_distance = ours; single-block
    (Franka) and map-reduce sized (G1-43: num_blocks_per_batch = 2) scratch, exactly as SelfCollisionCost allocates them
    (cost/cost_self_collision.py:31-89)."""
    from curobo_b200.robot_model import load_robot
    rm = load_robot(robot)
    pool = random_q(rm, 256 if robot == "franka" else 4 * n, seed=23)
    _, sph_pool, _, _ = O.fk_forward(rm, pool)
    cost_pool, _, _ = O.self_collision(sph_pool, rm.sphere_padding, rm.collision_pairs, 5000.0)
    keep = np.argsort(-cost_pool)[:n]                      # the colliding configurations first, then free ones
    sph_np = np.ascontiguousarray(sph_pool[keep])
    S, nb = rm.num_spheres, rm.num_blocks_per_batch
    sph = T(sph_np.reshape(n, 1, S, 4).astype(np.float32)).requires_grad_(True)
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=DEV)  # noqa: E731
    out_d, out_v, pd, sparse = z(n, 1), z(n, 1, S, 4), z(1), z(n, 1, S, dt=torch.uint8)
    bv, bi = z(n, 1, nb), z(n, 1, nb, 2, dt=torch.int16)
    w = torch.tensor([5000.0], device=DEV)
    d = ref.geometry.SelfCollisionDistance.apply(sph, out_d, out_v, pd, sparse, w, T(rm.sphere_padding), T(rm.collision_pairs),
                                                 bv, bi, nb, rm.max_threads_per_block, False, False)
    d.sum().backward()
    want_c, want_g, _ = O.self_collision(sph_np, rm.sphere_padding, rm.collision_pairs, 5000.0)
    assert np.allclose(d.detach().cpu().numpy().reshape(n), want_c, rtol=1e-4, atol=1e-6 * max(1.0, want_c.max()))
    assert np.allclose(sph.grad.cpu().numpy().reshape(n, S, 4), want_g.reshape(n, S, 4), rtol=1e-3,
                       atol=1e-5 * max(1.0, np.abs(want_g).max()))
    assert float(want_c.max()) > 0.0, "the case must contain colliding configurations"
    if ref_kernels.available():
        rd, rv = ref_kernels.self_collision(rm, T(sph_np.reshape(n, 1, S, 4).astype(np.float32)), T(rm.sphere_padding),
                                            T(rm.collision_pairs), 5000.0)
        assert np.allclose(d.detach().cpu().numpy().reshape(n), rd.cpu().numpy().reshape(n), rtol=1e-4, atol=1e-6 * max(1.0, want_c.max()))
        assert np.allclose(sph.grad.cpu().numpy().reshape(n, S, 4), rv.cpu().numpy().reshape(n, S, 4), rtol=1e-4,
                           atol=1e-5 * max(1.0, np.abs(want_g).max()))


@pytest.mark.parametrize("implicit", [False, True])
def test_reference_bspline_function_over_b200_backend(ref, implicit):
    """BSplineIdxKernel.apply (cuda_ops/trajectory.py:299-430): forward + backward through trajectory_cu = ours."""
    c = make_case(seed=17, B=6, nk=8, D=7, steps=4, degree=4, implicit=implicit)
    B, Tn, D, nk = c["B"], c["T"], c["D"], c["nk"]
    u = T(c["knots"]).requires_grad_(True)
    outs = [torch.zeros((B, Tn, D), device=DEV) for _ in range(4)]
    out_dt = torch.zeros(B, device=DEV)
    gk = torch.zeros((B, nk, D), device=DEV)
    p, v, a, j = ref.trajectory.BSplineIdxKernel.apply(
        u, *[T(x) for x in c["start"]], *[T(x) for x in c["goal"]], T(c["start_idx"]), T(c["goal_idx"]), *outs, out_dt,
        T(c["traj_dt"]), T(c["implicit"]), gk, 4)
    want = bo.bspline_forward(c["knots"], c["start"], c["goal"], c["start_idx"], c["goal_idx"], c["traj_dt"], c["implicit"], Tn, 4)
    for got, w in zip((p, v, a, j), want[:4]):
        assert np.allclose(got.detach().cpu().numpy(), w, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(w).max()))
    g = [T(x) for x in c["grads"]]
    ((p * g[0]).sum() + (v * g[1]).sum() + (a * g[2]).sum() + (j * g[3]).sum()).backward()
    wk = bo.bspline_backward(*c["grads"], c["traj_dt"], c["goal_idx"], c["implicit"], nk, 4)
    assert np.allclose(u.grad.cpu().numpy(), wk, rtol=1e-4, atol=1e-5 * np.abs(wk).max())


def test_reference_lbfgs_function_over_b200_backend(ref):
    """LBFGScu.apply (cuda_ops/optimization.py:192-252) -> optimization_cu.launch_lbfgs_step = ours; buffers shaped as
    QuasiNewtonBuffers allocates them ([m,B,V,1], [m,B,1,1], [B,V,1]; optim/components/quasi_newton_buffers.py:20-130)."""
    c = lbfgs_case(seed=3, B=12, V=28, m=7)
    m, B, V = c["Y"].shape
    step = torch.zeros((B, V), device=DEV)
    rho, y, s_ = T(c["rho"]).view(m, B, 1, 1), T(c["Y"]).view(m, B, V, 1), T(c["S"]).view(m, B, V, 1)
    x0, g0 = T(c["x_0"]).view(B, V, 1), T(c["grad_0"]).view(B, V, 1)
    out = ref.optimization.LBFGScu.apply(step, rho, y, s_, T(c["q"]), T(c["grad_q"]), x0, g0, 0.01, True, True)
    w_step, w_rho, w_Y, w_S, w_x0, w_g0 = oo.lbfgs_step(c["rho"], c["Y"], c["S"], c["q"], c["grad_q"], c["x_0"], c["grad_0"], 0.01, True)
    assert np.allclose(out.cpu().numpy().reshape(B, V), w_step, rtol=1e-4, atol=1e-6 * np.abs(w_step).max())
    assert np.array_equal(y.cpu().numpy().reshape(m, B, V), w_Y) and np.array_equal(s_.cpu().numpy().reshape(m, B, V), w_S)
    assert np.array_equal(x0.cpu().numpy().reshape(B, V), w_x0)
    if ref_kernels.available():
        r = {k: T(v) for k, v in c.items()}
        rstep = torch.zeros_like(step)
        ref_kernels.lbfgs_step(rstep, r["rho"], r["Y"], r["S"], r["q"], r["x_0"], r["grad_0"], r["grad_q"], 0.01, True, True)
        assert np.allclose(out.cpu().numpy().reshape(B, V), rstep.cpu().numpy(), rtol=1e-6, atol=1e-7 * np.abs(w_step).max())


def test_reference_lbfgs_optimizer_drives_the_b200_rollout(ref, iters=100, problems=24):
    """One level up: the reference's own optimizer -- LBFGSOpt + GradientOptCore (optim/gradient/lbfgs.py:157-400,
    optim/components/gradient_opt_core.py:255-480: line-search strategy, best tracker, `_compute_cost_constraint_and_gradient`) --
    is constructed on two `B200RobotRollout` instances (the `Rollout` protocol, rollout/rollout_protocol.py:35-176) and solves a batch
    of IK problems.  Everything below its Python is this repository: the fused rollout kernel behind `evaluate_action`, and
    `launch_lbfgs_step` / `launch_line_search` behind its LBFGScu / line-search calls (backend overlay).  It must solve what
    this repository's own LBFGSOpt solves with the same settings."""
    from curobo_b200.kinematics import Kinematics
    from curobo_b200.optim import LBFGSOpt as OurLBFGS, LBFGSOptCfg as OurCfg
    from curobo_b200.robot_model import load_robot
    from curobo_b200.rollout import RolloutConfig
    from curobo_b200.rollout_protocol import B200RobotRollout
    from curobo_b200.scene import CuboidData
    from curobo_b200.world import make_benchmark_cuboid_world
    rm = load_robot("franka")
    P, n, D = problems, 4, rm.num_dof
    scales = [0.0, 0.1, 0.5, 1.0]
    cub = make_benchmark_cuboid_world()
    rollouts = [B200RobotRollout(rm, RolloutConfig.ik(), DEV, cuboid=CuboidData.from_world(cub, DEV), horizon=1) for _ in range(2)]
    dc = ref.device_cfg.DeviceCfg(device=torch.device(DEV))
    cfg = ref.lbfgs.LBFGSOptCfg(num_iters=iters, inner_iters=25, num_problems=P, device_cfg=dc, line_search_scale=scales,
                                step_scale=0.98, history=7, epsilon=0.01, initial_step_scale=0.001)
    opt = ref.lbfgs.LBFGSOpt(cfg, rollouts, use_cuda_graph=False)
    _, _, gp, gq = O.fk_forward(rm, random_q(rm, P, seed=11) * 0.8)
    idx = np.repeat(np.arange(P), n).astype(np.int32)                       # rows: problem-major, particle-minor
    for ro in rollouts:
        ro.update_params(goal_position=T(gp[:, :, None, :]), goal_quat=T(gq[:, :, None, :]), idxs_goal=T(idx))
    x0 = T(random_q(rm, P, seed=12)).view(P, 1, D)
    # the CUDA event timer around optimize() needs a device; the emulated run calls the loop underneath it
    q_ref = (opt.optimize(x0) if DEV != "cpu" else opt._core._optimize_impl(x0)).reshape(P, D).clone()

    def pos_err(q):
        st = Kinematics(rm, DEV).compute_kinematics(q.view(P, 1, D))
        return np.linalg.norm(st.tool_pose_position.reshape(P, -1, 3)[:, 0].detach().cpu().numpy() - gp[:, 0], axis=-1)

    e0, e_ref = pos_err(x0.reshape(P, D)), pos_err(q_ref)
    assert np.median(e_ref) < (0.05 if iters >= 100 else 0.3) * np.median(e0), (np.median(e0), np.median(e_ref))
    # this repository's optimizer, same settings, same rollout engine
    eng = rollouts[0].engine

    def cost_grad(x):
        out = eng.evaluate_action(x.view(P * n, 1, D))
        return out.cost.view(-1), out.grad_q.view(P * n, D)

    ours = OurLBFGS(OurCfg(num_iters=iters, line_search_scale=scales, initial_step_scale=0.001), P, 1, D,
                    T(rm.position_limits[0]), T(rm.position_limits[1]), cost_grad, DEV)
    e_ours = pos_err(ours.optimize(x0.reshape(P, D)).reshape(P, D))
    assert np.median(e_ref) <= 2.0 * np.median(e_ours) + 2e-3, (np.median(e_ref), np.median(e_ours))
    if iters >= 100:
        # one seed per problem: local minima are expected; both optimizers must solve the same share of the problems
        assert (e_ref < 5e-3).mean() >= 0.4 and abs((e_ref < 5e-3).mean() - (e_ours < 5e-3).mean()) <= 0.2, (e_ref, e_ours)
