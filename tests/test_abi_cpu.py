"""The C-ABI library loads and exports every symbol include/curobo_b200.h declares (no compute calls
without a GPU), and the host-side robot-blob packer produces the layout the kernels expect."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

from curobo_b200 import build, lib as cblib
from curobo_b200.robot_model import load_robot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build.build_product()
    return cblib.load()


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "curobo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(L):
    names = _declared_functions()
    assert len(names) >= 15
    raw = C.CDLL(cblib.lib_path())
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/curobo_b200.h but not exported"
    assert set(names) == set(cblib.EXPORTED_SYMBOLS), "ctypes binding and header disagree"
    assert L.cb200_abi_version() == 6 and L.cb200_sm_arch() == 100


def test_library_contains_sm100a_sass_and_tma_bulk_copy():
    """Native code check without a GPU: sm_100a cubin present, the fused kernel stages its constants
    with a bulk async copy (UBLKCP in SASS = cp.async.bulk)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", cblib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "rollout_fused_kernel" in sass
    assert "UBLKCP" in sass


@pytest.mark.parametrize("name", ["franka", "g1_29", "g1_43"])
def test_robot_blob_layout(L, name):
    from curobo_b200.rollout import pack_robot_blob
    rm = load_robot(name)
    blob = pack_robot_blob(rm)
    hdr = struct.unpack("<48i", blob[:192].tobytes())
    magic, total, smem, nl, D, S, Lt, P, n_levels = hdr[:9]
    offs = dict(zip(["fixed", "joff", "link_map", "joint_map", "joint_type", "tool_map", "spheres", "sph_link",
                     "padding", "link_sph_off", "link_sph_idx", "level_off", "level_links", "anc_mask", "jl_off",
                     "jl_idx", "limits", "pairs"], hdr[9:27]))
    n_cl, n_lp = hdr[27:29]
    offs.update(dict(zip(["cl_link", "cl_start", "cl_bound", "lp"], hdr[29:33])))
    assert magic == 0x30324243 and total == blob.shape[0] and smem % 16 == 0 and smem <= total
    assert (nl, D, S, Lt, P) == (rm.num_links, rm.num_dof, rm.num_spheres, rm.num_tool_frames, rm.collision_pairs.shape[0])
    assert all(o % 16 == 0 for o in offs.values())
    view = lambda k, dt, n: np.frombuffer(blob.tobytes(), dtype=dt, count=n, offset=offs[k])  # noqa: E731
    np.testing.assert_array_equal(view("fixed", np.float32, nl * 12), rm.fixed_transforms.reshape(-1))
    np.testing.assert_array_equal(view("link_map", np.int16, nl), rm.link_map)
    np.testing.assert_array_equal(view("spheres", np.float32, S * 4), rm.link_spheres.reshape(-1))
    np.testing.assert_array_equal(view("pairs", np.int16, 2 * P), rm.collision_pairs.reshape(-1))
    # depth levels are a valid schedule: every link appears once, parents in earlier levels
    lo = view("level_off", np.int16, n_levels + 1)
    ll = view("level_links", np.int16, nl)
    assert sorted(ll.tolist()) == list(range(nl)) and lo[0] == 0 and lo[-1] == nl
    level_of = {}
    for lev in range(n_levels):
        for l in ll[lo[lev]:lo[lev + 1]]:
            level_of[int(l)] = lev
    assert all(level_of[int(rm.link_map[l])] == level_of[l] - 1 for l in range(1, nl))
    # ancestor masks agree with the reference's link_chain CSR
    anc = view("anc_mask", np.uint64, nl)
    for l in range(nl):
        chain = rm.link_chain_data[rm.link_chain_offsets[l]:rm.link_chain_offsets[l + 1]]
        assert int(anc[l]) == sum(1 << int(k) for k in chain)
    # link -> spheres CSR covers every sphere exactly once
    lso = view("link_sph_off", np.int16, nl + 1)
    lsi = view("link_sph_idx", np.int16, S)
    assert sorted(lsi.tolist()) == list(range(S))
    for l in range(nl):
        assert all(rm.link_sphere_idx_map[s] == l for s in lsi[lso[l]:lso[l + 1]])
    # smem budget: at least 4 warps of per-eval state + the staged blob fit the B200 opt-in limit (227 KB)
    per_warp = (nl * 12 + S * 8 + nl * 8 + n_cl * 4 + nl + 2 * D + Lt * 8 + 3) // 4 * 4 * 4
    # broad-phase tables: collision links partition the spheres; link pairs reproduce the pair list exactly
    assert n_cl == len(rm.collision_link_names) and n_lp > 0
    cls = view("cl_start", np.int16, n_cl + 1)
    cll = view("cl_link", np.int16, n_cl)
    lp = view("lp", np.uint32, n_lp)
    assert cls[0] == 0 and cls[-1] == S and all(rm.link_sphere_idx_map[cls[a]] == cll[a] for a in range(n_cl))
    rebuilt = set()
    for w in lp:
        a, b = int(w) & 0xffff, int(w) >> 16
        rebuilt.update((i, j) for i in range(cls[a], cls[a + 1]) for j in range(cls[b], cls[b + 1]))
    assert rebuilt == set(map(tuple, rm.collision_pairs.tolist()))
    bnd = view("cl_bound", np.float32, 4 * n_cl).reshape(n_cl, 4)
    for a in range(n_cl):
        ids = np.arange(cls[a], cls[a + 1])
        r = rm.link_spheres[ids, 3] + rm.sphere_padding[ids]
        en = r >= 0
        if en.any():
            d = np.linalg.norm(rm.link_spheres[ids[en], :3] - bnd[a, :3], axis=1) + r[en]
            assert (d <= bnd[a, 3]).all()
        else:
            assert bnd[a, 3] < 0
    assert smem + 4 * per_warp <= 227 * 1024, (smem, per_warp)


def test_blob_packer_rejects_bad_input(L):
    rm = load_robot("franka")
    sz = cblib.RobotSizes(65, 7, 65, 1, 0)             # 65 links > 64
    assert L.cb200_robot_blob_bytes(C.byref(sz)) < 0


def test_cpu_tensors_are_refused():
    """There is no CPU fallback: handing CPU tensors to an op raises before any launch."""
    import torch
    from curobo_b200.backends import geometry
    t = torch.zeros(4)
    with pytest.raises(ValueError):
        geometry.self_collision_distance(t, t, t, t.to(torch.uint8), t, t, t, t.to(torch.int16), t, t.to(torch.int16),
                                         1, 64, 1, 1, 1, 1, False, True)


def test_entry_points_reject_bad_arguments_before_touching_a_device(L):
    """Error behaviour of the boundary (the reference launchers raise on bad arguments before launching,
    cuda_core_backend/optimization.py:173-176, launch_helper.py:13-19): null pointers / out-of-range sizes return
    cudaErrorInvalidValue (1) from the argument checks, which run before any CUDA call -- so this runs without a GPU."""
    raw = C.CDLL(cblib.lib_path())
    INVALID = 1
    nul = C.c_void_p(None)
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    raw.cb200_rnea_forward.restype = C.c_int
    raw.cb200_rnea_forward.argtypes = [C.c_void_p] * 15 + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    assert raw.cb200_rnea_forward(*([nul] * 15), 4, 3, 2, 1, nul, nul) == INVALID          # null tensors
    assert raw.cb200_rnea_forward(*([p] * 15), -1, 3, 2, 1, nul, nul) == INVALID           # negative batch
    assert raw.cb200_rnea_forward(*([p] * 15), 4, 0, 2, 1, nul, nul) == INVALID            # no links
    raw.cb200_rnea_backward.restype = C.c_int
    raw.cb200_rnea_backward.argtypes = [C.c_void_p] * 17 + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    assert raw.cb200_rnea_backward(*([nul] * 17), 4, 3, 2, 1, nul, nul) == INVALID
    raw.cb200_lbfgs_step.restype = C.c_int
    raw.cb200_lbfgs_step.argtypes = ([C.c_void_p] * 8 + [C.c_float] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int, C.c_void_p,
                                     C.c_int, C.c_int, C.c_void_p])
    ok8 = [p] * 8
    assert raw.cb200_lbfgs_step(*([nul] * 8), 0.01, 4, 7, 7, 1, nul, nul, nul, 0, nul, 0, 0, nul) == INVALID
    assert raw.cb200_lbfgs_step(*ok8, 0.01, 4, 32, 7, 1, nul, nul, nul, 0, nul, 0, 0, nul) == INVALID    # history > 31
    assert raw.cb200_lbfgs_step(*ok8, 0.01, 4, 7, 2000, 1, nul, nul, nul, 0, nul, 0, 0, nul) == INVALID  # v_dim > 1024
    assert raw.cb200_lbfgs_step(*ok8, 0.01, 4, 7, 7, 1, p, p, nul, 4, nul, 0, 0, nul) == INVALID         # x_set without magnitudes
    raw.cb200_rollout_cost_grad.restype = C.c_int
    raw.cb200_rollout_cost_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    assert raw.cb200_rollout_cost_grad(nul, nul, nul) == INVALID
    # the Python layer turns a non-zero status into an exception that names the call
    with pytest.raises(cblib.CudaCallError, match="probe"):
        cblib.check(INVALID, "probe")
