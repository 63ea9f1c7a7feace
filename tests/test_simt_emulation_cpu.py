"""Whole CUDA kernels executed on the CPU: tests/simt compiles a kernel translation unit of curobo_b200/csrc as ordinary C++
(kernels become functions, threadIdx & co. are thread-local, __syncthreads() is a real barrier between the std::threads that
play a CTA's threads, shared memory is a static buffer, atomicAdd is a real atomic) and runs CTAs one after another.  This
value-checks kernel instantiations and schedules that the GPU tests of the last GPU session did not reach -- here every
rows-per-CTA variant of the RNEA CTA kernels (the launcher resolves small test batches to 8 rows per CTA; the bench sizes use
16 and 32) -- and exercises the barriers with genuinely concurrent threads."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dynamics_cases import RANDOM_TREES, make_case, model_args, pack_cache, random_tree_case
from helpers import ptr
from oracle import dynamics_oracle as do

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")


def build(name, src_deps, extra_units=()):
    so = os.path.join(SIMT, f"libsimt_{name}.so")
    units = [os.path.join(SIMT, f"simt_{name}.cpp")] + [os.path.join(SIMT, u) for u in extra_units]
    deps = units + [os.path.join(SIMT, "cuda_runtime.h")] + src_deps
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-w", "-I", SIMT, *units, "-o", so], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="module")
def emu():
    csrc = os.path.join(ROOT, "curobo_b200", "csrc")
    return build("dynamics", [os.path.join(csrc, "cb200_dynamics.cu"), os.path.join(csrc, "cb200_dynamics.cuh")])


def run(emu, c, R, grid):
    B, nl, D, nlev = c["B"], c["nl"], c["D"], c["n_levels"]
    m = [np.ascontiguousarray(x) for x in model_args(c)]
    model = [ptr(x) for x in m] + [ptr(c["starts"]), ptr(c["order"])]
    tau = np.full((B, D), np.nan, np.float32)
    cache = np.zeros((B, nl * 20), np.float32)
    assert emu.em_rnea_forward(R, grid, ptr(tau), ptr(c["q"]), ptr(c["qd"]), ptr(c["qdd"]), *model, ptr(cache), B, nl, D, nlev,
                               None) == 0
    g = [np.full((B, D), np.nan, np.float32) for _ in range(3)]
    assert emu.em_rnea_backward(R, grid, *[ptr(x) for x in g], ptr(c["grad_tau"]), ptr(c["q"]), ptr(c["qd"]), *model, ptr(cache), B,
                                nl, D, nlev, None) == 0
    return tau, cache, g


def check(c, tau, cache, g):
    m = model_args(c)
    tau_w, cache_w = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    want = do.rnea_backward(c["grad_tau"], c["q"], c["qd"], cache_w, *m)
    assert np.isfinite(tau).all() and all(np.isfinite(x).all() for x in g)
    assert np.allclose(tau, tau_w, rtol=2e-4, atol=3e-5 * np.abs(tau_w).max())
    cw = pack_cache(cache_w, c["nl"]).reshape(c["B"], c["nl"], 20)[:, :, :18]
    assert np.allclose(cache.reshape(c["B"], c["nl"], 20)[:, :, :18], cw, rtol=2e-4, atol=3e-5 * np.abs(cw).max())
    for got, w in zip(g, want):
        assert np.allclose(got, w, rtol=5e-4, atol=1e-4 * max(float(np.abs(w).max()), 1e-6))


@pytest.mark.parametrize("robot,B", [("franka", 70), ("g1_29", 37)])
@pytest.mark.parametrize("R", [8, 16, 32])
def test_rnea_cta_kernels_every_rows_per_cta_variant(emu, robot, B, R):
    c = make_case(robot, B, 17)
    grid = max(1, (B + R - 1) // R - 1)          # one CTA fewer than tiles: the grid-stride loop over row blocks is exercised
    check(c, *run(emu, c, R, grid))


@pytest.mark.parametrize("nl,B,seed,mimic", RANDOM_TREES)
def test_rnea_cta_kernels_on_synthetic_trees(emu, nl, B, seed, mimic):
    c = random_tree_case(nl, B + 9, seed, mimic)
    check(c, *run(emu, c, 16, 1))


# ------------------------------------------------------------------------------------------------ nearest-site transform
@pytest.fixture(scope="module")
def emu_edt():
    csrc = os.path.join(ROOT, "curobo_b200", "csrc")
    return build("edt", [os.path.join(csrc, "cb200_edt.cu"), os.path.join(csrc, "cb200_edt.cuh")])


def test_edt_kernels_executed_by_threads(emu_edt):
    """edt_flood_z_kernel / edt_envelope_kernel<1> / <0> themselves (not a restatement of their schedule): 32 std::threads per
    one-warp CTA, __syncwarp() as a real barrier, the grid-stride loop over tiles with fewer CTAs than tiles."""
    from scipy import ndimage
    from edt_cases import MEDIUM, SMALL, occupancy
    from oracle import edt_oracle as E
    for kind, shape, p in SMALL + MEDIUM[:3]:
        occ = occupancy(kind, shape, seed=13, p=p)
        g = np.ascontiguousarray(E.seed_grid(occ))
        assert emu_edt.em_pba3d(g.ctypes.data_as(C.c_void_p), *[int(v) for v in shape], 3) == 0
        d2 = E.squared_distance(g)
        if not occ.any():
            assert (g == E.EMPTY).all()
            continue
        sx, sy, sz = E.unpack(g)
        assert (g >= 0).all() and occ[sx, sy, sz].all()
        assert np.array_equal(d2, np.rint(ndimage.distance_transform_edt(~occ) ** 2).astype(np.int64)), (kind, shape)


# ------------------------------------------------------------------------------------------------ B-spline kernels
@pytest.fixture(scope="module")
def emu_traj():
    csrc = os.path.join(ROOT, "curobo_b200", "csrc")
    return build("trajectory", [os.path.join(csrc, "cb200_trajectory.cu"), os.path.join(csrc, "cb200_bspline.cuh")])


def test_bspline_kernels_executed_by_threads(emu_traj):
    """bspline_forward_kernel<3|4|5> and bspline_backward_kernel<3|4|5> themselves against the B-spline oracle (all boundary
    modes, mixed implicit goals, non power-of-two steps), with a grid smaller than the work so the grid-stride loops run."""
    from bspline_cases import CASES, make_case
    from oracle import bspline_oracle as bo
    for kw in CASES:
        c = make_case(**kw)
        B, T, D, nk, deg = c["B"], c["T"], c["D"], c["nk"], c["degree"]
        outs = [np.full((B, T, D), np.nan, np.float32) for _ in range(4)]
        odt = np.zeros(B, np.float32)
        assert emu_traj.em_bspline_forward(2, *[ptr(o) for o in outs], ptr(odt), ptr(c["knots"]), *[ptr(x) for x in c["start"]],
                                           *[ptr(x) for x in c["goal"]], ptr(c["start_idx"]), ptr(c["goal_idx"]), ptr(c["traj_dt"]),
                                           ptr(c["implicit"]), None, B, T, D, nk, deg) == 0
        want = bo.bspline_forward(c["knots"], c["start"], c["goal"], c["start_idx"], c["goal_idx"], c["traj_dt"], c["implicit"], T, deg)
        for got, w in zip(outs, want[:4]):
            assert np.allclose(got, w, rtol=2e-5, atol=2e-5 * max(1.0, float(np.abs(w).max()))), kw
        assert np.allclose(odt, c["traj_dt"][c["goal_idx"]])
        gk = np.full((B, nk, D), np.nan, np.float32)
        assert emu_traj.em_bspline_backward(2, ptr(gk), *[ptr(np.ascontiguousarray(g)) for g in c["grads"]], ptr(c["traj_dt"]),
                                            ptr(c["goal_idx"]), ptr(c["implicit"]), B, T, D, nk, deg) == 0
        wk = bo.bspline_backward(*c["grads"], c["traj_dt"], c["goal_idx"], c["implicit"], nk, deg)
        assert np.allclose(gk, wk, rtol=1e-4, atol=1e-5 * float(np.abs(wk).max())), kw


# ------------------------------------------------------------------------------------------------ race detection
def test_rnea_cta_kernels_are_race_free_under_thread_sanitizer(tmp_path):
    """The same kernels, instrumented with ThreadSanitizer: with CTA threads as real threads and __syncthreads() as a real barrier,
    a missing barrier in the kernel source (two phases touching the same shared-memory slot without a barrier between them)
    is a data race TSan reports.  All three rows-per-CTA variants, forward and adjoint, on a synthetic tree with mimic joints."""
    exe = os.path.join(SIMT, "tsan_dynamics")
    src = os.path.join(SIMT, "tsan_dynamics_main.cpp")
    csrc = os.path.join(ROOT, "curobo_b200", "csrc")
    deps = [src, os.path.join(SIMT, "simt_dynamics.cpp"), os.path.join(SIMT, "cuda_runtime.h"), os.path.join(csrc, "cb200_dynamics.cu"),
            os.path.join(csrc, "cb200_dynamics.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-w", "-I", SIMT, src, "-o", exe],
                           capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("ThreadSanitizer build not available: " + r.stderr[-300:])
    c = random_tree_case(17, 21, 32, True)
    m = [np.ascontiguousarray(x) for x in model_args(c)]

    def pad(a):
        b = np.ascontiguousarray(a).tobytes()
        return b + b"\0" * ((16 - len(b) % 16) % 16)

    blob = b"".join(pad(x) for x in [np.array([c["B"], c["nl"], c["D"], c["n_levels"]], np.int32), c["q"], c["qd"], c["qdd"],
                                     c["grad_tau"], m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], c["starts"], c["order"]])
    path = tmp_path / "case.bin"
    path.write_bytes(blob)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66 report_signal_unsafe=0")
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, env=env, timeout=600)
    if r.returncode != 0 and "FATAL: ThreadSanitizer" in r.stderr and "data race" not in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + r.stderr[-200:])
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout[-200:], r.stderr[-1500:])


def _tsan_build(src_name, exe_name, extra=()):
    exe, src = os.path.join(SIMT, exe_name), os.path.join(SIMT, src_name)
    extra = [os.path.join(ROOT, e) if e.endswith(".cpp") else e for e in extra]
    csrc = os.path.join(ROOT, "curobo_b200", "csrc")
    deps = [src] + [e for e in extra if e.endswith(".cpp")] + [os.path.join(SIMT, f) for f in os.listdir(SIMT) if f.endswith((".h", ".cpp"))] + \
        [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if os.path.exists(exe) and all(os.path.getmtime(d) <= os.path.getmtime(exe) for d in deps):
        return exe
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-w", *extra, "-I", SIMT, src, "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-300:])
    return exe


def _tsan_run(cmd):
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66 report_signal_unsafe=0")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    if r.returncode != 0 and "FATAL: ThreadSanitizer" in r.stderr and "data race" not in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + r.stderr[-200:])
    return r


def test_edt_kernels_are_race_free_under_thread_sanitizer():
    r = _tsan_run([_tsan_build("tsan_edt_main.cpp", "tsan_edt")])
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout[-200:], r.stderr[-1500:])


def test_race_detector_has_teeth():
    """Mutation: the same EDT kernels with every barrier compiled out (-DCB200_SIMT_DROP_BARRIERS): the flood pass then reads
    tile rows other lanes are still writing, and ThreadSanitizer must say so (exit code 66)."""
    r = _tsan_run([_tsan_build("tsan_edt_main.cpp", "tsan_edt_nobarrier", ("-DCB200_SIMT_DROP_BARRIERS",))])
    assert r.returncode == 66 and "data race" in r.stderr, (r.returncode, r.stderr[-500:])


# ------------------------------------------------------------------------------------------------ the fused rollout kernels
@pytest.fixture(scope="module")
def emu_main():
    """The product's main translation unit (fused kernels + per-operator kernels + blob packer + the whole C ABI with its launch
    logic) as a host library: same cb200_* symbols and signatures as libcurobo_b200.so, host pointers instead of device pointers."""
    from curobo_b200 import lib as cblib
    csrc = os.path.join(ROOT, "curobo_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in ("cb200_kernels.cu", "cb200_trajectory.cu", "cb200_warp.cuh", "cb200_math.cuh",
                                            "cb200_blob.h", "cb200_bspline.cuh", "cb200_launch.h")] + [os.path.join(SIMT, "cuda_fp16.h")]
    L = build("kernels", deps, extra_units=("simt_trajectory_abi.cpp",))
    for name, (args, res) in cblib._SIGS.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
    assert L.cb200_abi_version() == 6
    return L


def host_engine(emu_main, monkeypatch, rm, cfg, cuboid=None, voxel=None):
    """A RolloutEngine whose native library is the emulated one and whose tensors live on the host: the product refuses non-CUDA
    devices in exactly one function (backends.tensor_checks.require_cuda); it and the library loader are swapped here, only here."""
    from curobo_b200 import lib as cblib
    from curobo_b200 import rollout as R
    from curobo_b200.backends import tensor_checks as tc
    monkeypatch.setattr(cblib, "_LIB", emu_main)
    monkeypatch.setattr(cblib, "load", lambda: emu_main)
    monkeypatch.setattr(tc, "require_cuda", lambda device, message: None)
    monkeypatch.setattr(tc, "_stream_of", lambda device: 0)
    return R.RolloutEngine(rm, cfg, "cpu", cuboid, voxel, use_voxel_mip=False)


def test_fused_ik_rollout_kernel_executed_by_threads(emu_main, monkeypatch):
    """rollout_fused_kernel (warp per evaluation: FK, spheres, self-collision, cuboid collision, tool pose, c-space, J^T backward,
    with its shuffles / ballots / warp reductions and the blob staging) and the launcher around it, on the smoke() problem."""
    import torch
    from curobo_b200.robot_model import load_robot
    from curobo_b200.rollout import RolloutConfig
    from curobo_b200.scene import CuboidData
    from curobo_b200.world import make_benchmark_cuboid_world
    from helpers import random_q
    from oracle import rollout_oracle as O
    rm = load_robot("franka")
    B = 40
    q = random_q(rm, B, seed=3)[:, None, :]
    _, _, gp, gq = O.fk_forward(rm, random_q(rm, 4, seed=4))
    goal = (gp[:, :, None, :].copy(), gq[:, :, None, :].copy())
    idx = (np.arange(B) % 4).astype(np.int32)
    cub = make_benchmark_cuboid_world()
    cfg = RolloutConfig.ik()
    eng = host_engine(emu_main, monkeypatch, rm, cfg, cuboid=CuboidData.from_world(cub, "cpu"))
    eng.update_goal(torch.as_tensor(goal[0]), torch.as_tensor(goal[1]), torch.as_tensor(idx))
    out = eng.evaluate_action(torch.as_tensor(q))
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), world_cuboid=cub, goal_pos=goal[0], goal_quat=goal[1], idxs_goal=idx)
    np.testing.assert_allclose(out.cost.numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    g = want["grad_q"]
    np.testing.assert_allclose(out.grad_q.numpy(), g, rtol=2e-3, atol=2e-5 * np.abs(g).max())


@pytest.mark.parametrize("B,H,speed", [(2, 9, True), (2, 5, False)])
def test_trajectory_rollout_kernel_executed_by_threads(emu_main, monkeypatch, B, H, speed):
    """rollout_traj_kernel (CTA per trajectory chunk, neighbour waypoints through shared memory, swept collision over cuboids and
    an fp16 ESDF, speed metric, STATE c-space with retimed weights, terminal-only pose) + the launcher's trajectory path."""
    import torch
    from curobo_b200.robot_model import load_robot
    from curobo_b200.rollout import RolloutConfig
    from curobo_b200.scene import CuboidData, VoxelData
    from curobo_b200.world import make_benchmark_cuboid_world
    from helpers import random_q, random_walk_q, small_voxel_world
    from oracle import rollout_oracle as O
    rm = load_robot("franka")
    q = random_walk_q(rm, B, H, seed=60 + H)
    rng = np.random.default_rng(H)
    dt = np.full(B, 0.05, np.float32)
    v = np.gradient(q, axis=1).astype(np.float32) / 0.05
    a_ = rng.normal(0, 5.0, size=q.shape).astype(np.float32)
    j_ = rng.normal(0, 200.0, size=q.shape).astype(np.float32)
    cfg = RolloutConfig.trajopt()
    cfg.use_speed_metric = speed
    cub, vox = make_benchmark_cuboid_world(), small_voxel_world()
    _, _, p, qt = O.fk_forward(rm, random_q(rm, B, seed=61))
    gp, gq = p[:, :, None, :].copy(), qt[:, :, None, :].copy()
    idx = np.arange(B, dtype=np.int32)
    eng = host_engine(emu_main, monkeypatch, rm, cfg, cuboid=CuboidData.from_world(cub, "cpu"), voxel=VoxelData.from_world(vox, "cpu"))
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x))  # noqa: E731
    eng.update_goal(t(gp), t(gq), t(idx), non_terminal_axes=torch.zeros((1, 6), dtype=torch.float32))
    out = eng.evaluate_action(t(q), vel=t(v), acc=t(a_), jerk=t(j_), dt=t(dt))
    ocfg = cfg.to_oracle_cfg(1)
    ocfg["pose_non_terminal_axes"] = np.zeros((1, 6), np.float32)
    want = O.rollout_cost_grad(rm, q, ocfg, world_cuboid=cub, world_voxel=vox, goal_pos=gp, goal_quat=gq, idxs_goal=idx, vel=v,
                               acc=a_, jerk=j_, dt=dt)
    np.testing.assert_allclose(out.scene_cost.numpy(), want["scene_cost"], rtol=2e-4, atol=1e-5 * max(want["scene_cost"].max(), 1e-6))
    np.testing.assert_allclose(out.cost.numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    g = want["grad_q"]
    np.testing.assert_allclose(out.grad_q.numpy(), g, rtol=2e-3, atol=2e-5 * np.abs(g).max())


def test_humanoid_rollout_kernel_executed_by_threads(emu_main, monkeypatch):
    """G1-29 (400 spheres, 55,414 collision pairs, the link-level broad phase and the second-level cull with ballot compaction,
    ESDF collision, tool pose) through the same kernel family."""
    import torch
    from curobo_b200.robot_model import load_robot
    from curobo_b200.rollout import RolloutConfig
    from curobo_b200.scene import VoxelData
    from helpers import humanoid_q, small_voxel_world
    from oracle import rollout_oracle as O
    rm = load_robot("g1_29")
    q = humanoid_q(rm, 5, seed=45)[:, None, :]
    _, _, p, qt = O.fk_forward(rm, humanoid_q(rm, 3, seed=46, scale=0.5))
    gp, gq = p[:, :, None, :].copy(), qt[:, :, None, :].copy()
    idx = (np.arange(q.shape[0]) % 3).astype(np.int32)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0), pose_weight=(2000.0, 100.0))
    vox = small_voxel_world()
    eng = host_engine(emu_main, monkeypatch, rm, cfg, voxel=VoxelData.from_world(vox, "cpu"))
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x))  # noqa: E731
    eng.update_goal(t(gp), t(gq), t(idx))
    out = eng.evaluate_action(t(q))
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(rm.num_tool_frames), world_voxel=vox, goal_pos=gp, goal_quat=gq, idxs_goal=idx)
    np.testing.assert_allclose(out.self_cost.numpy(), want["self_cost"], rtol=1e-4, atol=1e-6 * max(want["self_cost"].max(), 1e-6))
    np.testing.assert_allclose(out.cost.numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    g = want["grad_q"]
    np.testing.assert_allclose(out.grad_q.numpy(), g, rtol=2e-3, atol=2e-5 * np.abs(g).max())


def _write_sections(path, sections):
    with open(path, "wb") as f:
        f.write(np.int32(len(sections)).tobytes())
        for name, data in sections.items():
            b = data if isinstance(data, bytes) else np.ascontiguousarray(data).tobytes()
            f.write(name.encode().ljust(24, b"\0"))
            f.write(np.int64(len(b)).tobytes())
            f.write(b + b"\0" * ((16 - len(b) % 16) % 16))


@pytest.mark.parametrize("mode", ["ik", "trajectory"])
def test_fused_rollout_kernels_are_race_free_under_thread_sanitizer(mode, tmp_path):
    """The fused kernels are warp-synchronous code: lanes exchange data through shuffles and through shared memory followed by
    __syncwarp().  Here lanes are real threads, so an exchange through shared memory that lacks its barrier is a data race that
    ThreadSanitizer reports (and the emulated collectives are barriers themselves, as the hardware's are)."""
    from curobo_b200.robot_model import load_robot
    from curobo_b200.rollout import RolloutConfig, RolloutEngine, pack_robot_blob
    from curobo_b200.world import make_benchmark_cuboid_world
    from helpers import random_q, random_walk_q, small_voxel_world
    from oracle import rollout_oracle as O
    exe = _tsan_build("tsan_rollout_main.cpp", "tsan_rollout", ("tests/simt/simt_trajectory_abi.cpp",))
    rm = load_robot("franka")
    cub = make_benchmark_cuboid_world()
    S, L, D = rm.num_spheres, rm.num_tool_frames, rm.num_dof
    sec = {"blob": pack_robot_blob(rm), "cub_dims": cub.dims, "cub_inv_pose": cub.inv_pose, "cub_enable": cub.enable, "cub_count": cub.count}
    vox_dims = (0, 0, 0)
    if mode == "ik":
        B, H = 12, 1
        cfg = RolloutConfig.ik()
        q = random_q(rm, B, seed=3)[:, None, :]
    else:
        B, H = 2, 7
        cfg = RolloutConfig.trajopt()
        q = random_walk_q(rm, B, H, seed=63)
        rng = np.random.default_rng(2)
        vox = small_voxel_world()
        sec.update(vel=np.gradient(q, axis=1).astype(np.float32) / 0.05, acc=rng.normal(0, 5.0, size=q.shape).astype(np.float32),
                   jerk=rng.normal(0, 200.0, size=q.shape).astype(np.float32), dt=np.full(B, 0.05, np.float32),
                   vox_params=vox.params, vox_inv_pose=vox.inv_pose, vox_enable=vox.enable, vox_count=vox.count,
                   vox_features=vox.features.view(np.uint16), vox_max_dist=np.float32(vox.max_dist),
                   axes_nt=np.zeros((L, 6), np.float32))
        vox_dims = (vox.max_n, vox.num_envs, vox.features.shape[2])
    _, _, p, qt = O.fk_forward(rm, random_q(rm, B, seed=61))
    fake = type("E", (), {"cfg": cfg})()
    ccfg = RolloutEngine._make_ccfg(fake, 1)
    sec.update(cfg=bytes(ccfg), q=q.astype(np.float32), goal_pos=p[:, :, None, :].astype(np.float32),
               goal_quat=qt[:, :, None, :].astype(np.float32), idxs_goal=np.arange(B, dtype=np.int32),
               dims=np.array([B, H, D, S, L, cub.max_n, cub.num_envs, *vox_dims], np.int32))
    path = tmp_path / "case.bin"
    _write_sections(path, sec)
    r = _tsan_run([exe, str(path)])
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout[-200:], r.stderr[-2500:])
    assert float(r.stdout.split()[1]) > 0 and float(r.stdout.split()[2]) > 0


# ------------------------------------------------------------------------------------------------ drop-in cost kernels vs the
# reference-source fixture (tests/golden/warp_reference_golden.npz): kernel <-> reference source with nothing in between
@pytest.fixture
def host_cost(emu_main, monkeypatch):
    """curobo_b200.cost with the emulated library and host tensors (test-only swap, as host_engine)."""
    from curobo_b200 import cost as cb_cost
    from curobo_b200 import lib as cblib
    monkeypatch.setattr(cblib, "load", lambda: emu_main)
    monkeypatch.setattr(cb_cost, "check_tensors", lambda *a, **k: None)
    monkeypatch.setattr(cb_cost, "stream_ptr", lambda d: 0)
    monkeypatch.setattr(cb_cost.torch.cuda, "is_current_stream_capturing", lambda: False)   # no CUDA runtime on this box
    return cb_cost


def _golden(name):
    G = np.load(os.path.join(ROOT, "tests", "golden", "warp_reference_golden.npz"))
    pre = name + "/"
    return {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}


@pytest.mark.parametrize("retime", [0, 1])
def test_cspace_state_kernel_matches_the_reference_source_including_effort(host_cost, retime):
    """cspace_state_kernel itself against the output of the reference's forward_cspace_state_warp source: all five channels --
    the effort channel (bound hinge, squared-L2, energy term) is live in this fixture, which the GPU parity test of r21 did not
    exercise (it passes zero torques)."""
    import torch
    c = _golden(f"cspace_state_retime{retime}")
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x))  # noqa: E731
    B, H, D = c["q"].shape
    outs = [torch.zeros((B, H, D)) for _ in range(6)]
    host_cost.cspace_state_cost(t(c["q"]), t(c["v"]), t(c["a"]), t(c["j"]), t(c["tau"]), t(c["dt"]), t(c["target"]), t(c["idxs_target"]),
                                t(c["lim_p"]), t(c["lim_v"]), t(c["lim_a"]), t(c["lim_j"]), t(c["lim_tau"]), t(c["weight"]), t(c["act"]),
                                t(c["reg"]), t(c["target_weight"]), t(c["ntf"]), t(c["dof_weight"]), *outs, bool(retime), bool(retime))
    for got, key in zip(outs, ("cost", "grad_p", "grad_v", "grad_a", "grad_j", "grad_t")):
        w = c[key]
        assert np.allclose(got.numpy(), w, rtol=2e-5, atol=2e-6 * np.abs(w).max()), (key, float(np.abs(got.numpy() - w).max()))
    assert np.abs(c["grad_t"]).max() > 0


def test_cspace_position_kernel_matches_the_reference_source(host_cost):
    import torch
    c = _golden("cspace_position")
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x))  # noqa: E731
    B, H, D = c["q"].shape
    oc, gp, gt = (torch.zeros((B, H, D)) for _ in range(3))
    z = torch.zeros((B, H, D))
    host_cost.cspace_position_cost(t(c["q"]), z, t(c["target"]), t(c["idxs_target"]), t(c["lim_p"]), torch.ones((2, D)), t(c["weight"]),
                                   t(c["act"]), t(c["target_weight"]), t(c["dof_weight"]), torch.zeros(2), torch.zeros((1, D)),
                                   torch.zeros((1, D)), torch.zeros(B, dtype=torch.int32), torch.ones((2, D)), torch.zeros(B), oc, gp, gt)
    assert np.allclose(oc.numpy(), c["cost"], rtol=2e-5, atol=2e-6 * np.abs(c["cost"]).max())
    assert np.allclose(gp.numpy(), c["grad_p"], rtol=2e-5, atol=2e-6 * np.abs(c["grad_p"]).max())


@pytest.mark.parametrize("method", [0, 1])
def test_tool_pose_kernel_matches_the_reference_source(host_cost, method):
    import torch
    c = _golden(f"tool_pose_method{method}")
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x))  # noqa: E731
    B, H, L, _ = c["pos"].shape
    od, opd, ord_ = torch.zeros((B, H, 2 * L)), torch.zeros((B, H, L)), torch.zeros((B, H, L))
    opg, org, ogi = torch.zeros((B, H, L, 3)), torch.zeros((B, H, L, 4)), torch.zeros((B, H, L), dtype=torch.int32)
    host_cost.tool_pose_distance(t(c["pos"]), t(c["quat"]), t(c["goal_pos"]), t(c["goal_quat"]), t(c["idxs_goal"]).view(B, 1), t(c["weight"]),
                                 t(c["axes_t"]), t(c["axes_nt"]), t(c["tol_t"]), t(c["tol_nt"]), torch.zeros(L, dtype=torch.uint8), od,
                                 opd, ord_, opg, org, ogi, use_lie_group=bool(method))
    assert np.array_equal(ogi.numpy(), c["goalset_idx"])
    for got, key, rt in ((od, "distance", 1e-4), (opd, "pos_dist", 1e-4), (ord_, "rot_dist", 1e-4), (opg, "grad_pos", 5e-4),
                         (org, "grad_quat", 2e-3)):
        w = c[key]
        assert np.allclose(got.numpy(), w, rtol=rt, atol=rt * 0.1 * max(float(np.abs(w).max()), 1e-6)), key


@pytest.mark.parametrize("name", ["collision_discrete", "collision_multi_env", "collision_swept", "collision_swept_speed",
                                  "collision_edge_discrete", "collision_edge_swept", "collision_edge_swept_speed"])
def test_scene_collision_kernel_matches_the_reference_source(emu_main, monkeypatch, name):
    """scene_collision_kernel itself (one launch for all obstacle types, no atomics) against the output of the reference's Warp
    collision / swept / speed-metric kernel sources on the same spheres and worlds."""
    import torch
    from curobo_b200 import lib as cblib
    from curobo_b200 import scene as cb_scene
    from curobo_b200.world import CuboidWorld, VoxelWorld
    monkeypatch.setattr(cblib, "load", lambda: emu_main)
    monkeypatch.setattr(cb_scene, "check_tensors", lambda *a, **k: None)
    monkeypatch.setattr(cb_scene, "stream_ptr", lambda d: 0)
    c = _golden(name)
    cub = cb_scene.CuboidData.from_world(CuboidWorld(c["cub_dims"], c["cub_inv_pose"], c["cub_enable"], c["cub_count"]), "cpu") if "cub_dims" in c else None
    vox = cb_scene.VoxelData.from_world(VoxelWorld(c["vox_params"], c["vox_inv_pose"], c["vox_enable"], c["vox_count"], c["vox_features"],
                                                   float(c["vox_max_dist"])), "cpu") if "vox_params" in c else None
    sph = torch.as_tensor(np.ascontiguousarray(c["spheres"]))
    buf = cb_scene.CollisionBuffer.from_shape(sph.shape, "cpu")
    env = torch.as_tensor(c["env_query_idx"]) if "env_query_idx" in c else None
    speed = "speed_dt" in c
    cb_scene._launch("swept" in name, sph, buf, cb_scene.SceneData(cub, vox), torch.tensor([float(c["weight"])]),
                     torch.tensor([float(c["eta"])]), torch.tensor([float(c["speed_dt"])]) if speed else None, speed, env,
                     env is not None)
    w, g = c["cost"], c["grad"]
    assert np.allclose(buf.distance.numpy(), w, rtol=1e-4, atol=1e-5 * np.abs(w).max()), float(np.abs(buf.distance.numpy() - w).max())
    assert np.allclose(buf.gradient.numpy(), g, rtol=5e-4, atol=5e-5 * np.abs(g).max()), float(np.abs(buf.gradient.numpy() - g).max())
    assert np.array_equal(buf.distance.numpy() > 0, w > 0)
