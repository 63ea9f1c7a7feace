"""Robot constants for the rollout hot path (host side, numpy only).

Mirrors the *outputs* of the reference's one-time CPU setup:
  - ``KinematicsParams``            curobo/_src/robot/types/kinematics_params.py:23-158
  - ``SelfCollisionKinematicsCfg``  curobo/_src/robot/types/self_collision_params.py:16-125
built by ``KinematicsLoader``       curobo/_src/robot/loader/kinematics_loader.py:215-264,367-486,848-915
from a URDF + robot YAML (parser:   curobo/_src/robot/parser/parser_urdf.py:133-300).

The reference's loader needs ``yourdfpy``/``trimesh`` (absent here); this is an independent,
dependency-free restatement (xml.etree + pyyaml) that produces the same tensors.  It is pinned by the
reference's FK golden vector (tests/_src/robot/kinematics/test_kinematics.py:57-82) and by the
self-collision pair counts (818 / 55,414 / 162,111).

Nothing here runs per optimizer iteration; the tensors are packed once into a device blob
(see ``pack_robot_blob``) that the fused kernel stages into shared memory with one bulk copy.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# joint type codes: curobo/_src/curobolib/kernels/kinematics/kinematics_constants.h:10-16
FIXED, X_PRISM, Y_PRISM, Z_PRISM, X_ROT, Y_ROT, Z_ROT = -1, 0, 1, 2, 3, 4, 5
_JT = {"FIXED": FIXED, "X_PRISM": X_PRISM, "Y_PRISM": Y_PRISM, "Z_PRISM": Z_PRISM,
       "X_ROT": X_ROT, "Y_ROT": Y_ROT, "Z_ROT": Z_ROT}

_ROBOT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "content", "robots")


@dataclass
class RobotModel:
    """Flat numpy mirror of KinematicsParams + SelfCollisionKinematicsCfg + joint limits."""

    name: str
    link_names: List[str]
    joint_names: List[str]
    tool_frames: List[str]
    fixed_transforms: np.ndarray       # [nl,3,4] f32
    link_map: np.ndarray               # [nl] i16 parent link index (parent < child)
    joint_map: np.ndarray              # [nl] i16 joint index or -1
    joint_map_type: np.ndarray         # [nl] i8
    joint_offset_map: np.ndarray       # [nl,2] f32 (scale, bias)
    tool_frame_map: np.ndarray         # [L] i16
    link_spheres: np.ndarray           # [S,4] f32 (x,y,z,r) in link frame; r<0 disabled
    link_sphere_idx_map: np.ndarray    # [S] i16 sphere -> link index
    link_chain_data: np.ndarray        # CSR: ancestors of each link (base..link), i16
    link_chain_offsets: np.ndarray     # [nl+1] i16
    joint_links_data: np.ndarray       # CSR: links driven by each joint, i16
    joint_links_offsets: np.ndarray    # [D+1] i16
    joint_affects_endeffector: np.ndarray  # [D*L] bool
    link_masses_com: np.ndarray        # [nl,4] f32
    collision_pairs: np.ndarray        # [P,2] i16, i<j
    sphere_padding: np.ndarray         # [S] f32
    position_limits: np.ndarray        # [2,D] f32
    velocity_limits: np.ndarray        # [2,D] f32
    acceleration_limits: np.ndarray    # [2,D] f32
    jerk_limits: np.ndarray            # [2,D] f32
    effort_limits: np.ndarray          # [2,D] f32
    default_joint_position: np.ndarray  # [D] f32
    collision_link_names: List[str] = field(default_factory=list)

    @property
    def num_links(self) -> int:
        return int(self.link_map.shape[0])

    @property
    def num_dof(self) -> int:
        return len(self.joint_names)

    @property
    def num_spheres(self) -> int:
        return int(self.link_spheres.shape[-2])

    @property
    def num_tool_frames(self) -> int:
        return int(self.tool_frame_map.shape[0])

    # self-collision launch sizing, curobo/_src/robot/types/self_collision_params.py:37-61
    @property
    def max_threads_per_block(self) -> int:
        return 512 if self.collision_pairs.shape[0] > 1000 else 64

    @property
    def num_checks_per_thread(self) -> int:
        return 256 if self.collision_pairs.shape[0] > 1000 else 32

    @property
    def num_blocks_per_batch(self) -> int:
        return int(math.ceil(self.collision_pairs.shape[0]
                             / (self.num_checks_per_thread * self.max_threads_per_block)))

    # ---- (de)serialisation ---------------------------------------------------------------
    _ARRAYS = ["fixed_transforms", "link_map", "joint_map", "joint_map_type", "joint_offset_map",
               "tool_frame_map", "link_spheres", "link_sphere_idx_map", "link_chain_data",
               "link_chain_offsets", "joint_links_data", "joint_links_offsets",
               "joint_affects_endeffector", "link_masses_com", "collision_pairs", "sphere_padding",
               "position_limits", "velocity_limits", "acceleration_limits", "jerk_limits",
               "effort_limits", "default_joint_position"]
    _LISTS = ["link_names", "joint_names", "tool_frames", "collision_link_names"]

    def save(self, path: str) -> None:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        for k in self._LISTS:
            d[k] = np.array(getattr(self, k), dtype=np.str_)
        d["name"] = np.array(self.name)
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path: str) -> "RobotModel":
        z = np.load(path, allow_pickle=False)
        kw = {k: z[k] for k in cls._ARRAYS}
        for k in cls._LISTS:
            kw[k] = [str(x) for x in z[k]]
        kw["name"] = str(z["name"])
        return cls(**kw)


def load_robot(name: str) -> RobotModel:
    """Load a packaged robot ("franka", "g1_29", "g1_43").

    The .npz files are generated from the reference's own URDF/YAML content by
    ``scripts/build_robot_fixtures.py`` (committed next to its outputs)."""
    path = os.path.join(_ROBOT_DIR, name + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"robot fixture {path} missing; run scripts/build_robot_fixtures.py")
    return RobotModel.load(path)


# ------------------------------------------------------------------------------------------
# URDF + YAML -> RobotModel
# ------------------------------------------------------------------------------------------

def _rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


def _quat_wxyz_to_mat(q):
    w, x, y, z = q
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def _floats(s, n, default):
    if s is None:
        return list(default)
    v = [float(t) for t in s.split()]
    assert len(v) == n
    return v


@dataclass
class _Body:
    link_name: str
    parent: Optional[str]
    joint_name: str
    joint_type: int
    fixed: np.ndarray                 # 4x4 float64
    offset: List[float]
    limits: Optional[List[float]] = None
    vel_limits: Optional[List[float]] = None
    effort: float = 10000.0
    mimic_of: Optional[str] = None
    mass_com: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 0.01]))


class _Urdf:
    """Minimal URDF reader: joints (origin/axis/limits/mimic), link inertial mass+com."""

    def __init__(self, path: str, extra_links: Dict[str, dict]):
        root = ET.parse(path).getroot()
        self.joints = {}
        self.parent_map: Dict[str, dict] = {}
        self.inertial = {}
        for ln in root.findall("link"):
            ine = ln.find("inertial")
            if ine is not None:
                mass = float(ine.find("mass").get("value")) if ine.find("mass") is not None else 0.01
                org = ine.find("origin")
                xyz = _floats(org.get("xyz") if org is not None else None, 3, [0, 0, 0])
                self.inertial[ln.get("name")] = (mass if mass > 0 else 0.01, xyz)
        for j in root.findall("joint"):
            name = j.get("name")
            org = j.find("origin")
            T = np.eye(4)
            if org is not None:
                T[:3, :3] = _rpy_to_mat(_floats(org.get("rpy"), 3, [0, 0, 0]))
                T[:3, 3] = _floats(org.get("xyz"), 3, [0, 0, 0])
            ax = j.find("axis")
            lim = j.find("limit")
            mim = j.find("mimic")
            rec = dict(name=name, type=j.get("type"), origin=T,
                       axis=_floats(ax.get("xyz") if ax is not None else None, 3, [1, 0, 0]),
                       parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                       limit=None if lim is None else dict(
                           lower=float(lim.get("lower", 0)), upper=float(lim.get("upper", 0)),
                           velocity=float(lim.get("velocity", 100.0)), effort=float(lim.get("effort", 100.0))),
                       mimic=None if mim is None else dict(
                           joint=mim.get("joint"), multiplier=float(mim.get("multiplier", 1.0)),
                           offset=float(mim.get("offset", 0.0))))
            self.joints[name] = rec
            self.parent_map[rec["child"]] = {"parent": rec["parent"], "joint_name": name}
        # extra links (parser_base.py:46-50): re-parent `child_link_name` under the extra link
        self.extra = extra_links or {}
        for k, e in self.extra.items():
            self.parent_map[k] = {"parent": e["parent_link_name"]}
            if e.get("child_link_name") is not None:
                self.parent_map[e["child_link_name"]]["parent"] = k

    def chain(self, base: str, ee: str) -> List[str]:
        out, link = [ee], ee
        while link != base:
            link = self.parent_map[link]["parent"]
            out.append(link)
        return out[::-1]

    def body(self, link_name: str, base: bool = False) -> _Body:
        if link_name in self.extra:
            e = self.extra[link_name]
            T = np.eye(4)
            ft = e["fixed_transform"]
            T[:3, :3] = _quat_wxyz_to_mat(ft[3:7])
            T[:3, 3] = ft[0:3]
            return _Body(link_name, e["parent_link_name"], e["joint_name"], _JT[e["joint_type"]], T,
                         list(e.get("joint_offset", [1.0, 0.0])), e.get("joint_limits"),
                         e.get("joint_velocity_limits", [-2.0, 2.0]))
        mass, com = self.inertial.get(link_name, (0.01, [0, 0, 0]))
        mc = np.array([com[0], com[1], com[2], mass])
        if base:
            return _Body(link_name, None, "base_joint", FIXED, np.eye(4), [1.0, 0.0], mass_com=mc)
        pm = self.parent_map[link_name]
        j = self.joints[pm["joint_name"]]
        b = _Body(link_name, pm["parent"], j["name"], FIXED, j["origin"], [1.0, 0.0], mass_com=mc)
        if j["type"] == "fixed":
            return b
        lim = dict(j["limit"]) if j["limit"] is not None else dict(lower=0, upper=0, velocity=100.0, effort=100.0)
        jt = j["type"]
        if jt == "continuous":
            jt, lim["lower"], lim["upper"] = "revolute", -6.28, 6.28
        offset = [1.0, 0.0]
        if j["mimic"] is not None:
            offset = [j["mimic"]["multiplier"], j["mimic"]["offset"]]
            b.mimic_of = j["name"]
            b.joint_name = j["mimic"]["joint"]
            act = self.joints[b.joint_name]["limit"]
            lim = dict(act)
        axis = j["axis"]
        k = int(np.argmax(np.abs(axis)))
        if abs(abs(axis[k]) - 1.0) > 1e-9:
            raise ValueError(f"joint {j['name']}: only axis-aligned joints supported, got {axis}")
        b.joint_type = (X_PRISM if jt == "prismatic" else X_ROT) + k
        if axis[k] < 0:
            offset[0] = -offset[0]
        b.offset, b.limits = offset, [lim["lower"], lim["upper"]]
        b.vel_limits, b.effort = [-lim["velocity"], lim["velocity"]], lim["effort"]
        return b


def _joint_motion(jtype: int, theta: float) -> np.ndarray:
    T = np.eye(4)
    if jtype == FIXED:
        return T
    if jtype <= Z_PRISM:
        T[jtype, 3] = theta
        return T
    c, s = math.cos(theta), math.sin(theta)
    a = jtype - X_ROT
    i, j = (a + 1) % 3, (a + 2) % 3
    T[i, i], T[i, j], T[j, i], T[j, j] = c, -s, s, c
    return T


def build_robot_model(name: str, urdf_path: str, cfg: dict) -> RobotModel:
    """cfg = the `kinematics` dict of a curobo robot YAML (content/configs/robot/*.yml)."""
    tool_frames = list(cfg["tool_frames"])
    coll_links = list(cfg.get("collision_link_names") or [])
    extra = cfg.get("extra_links") or {}
    urdf = _Urdf(urdf_path, extra)
    base = cfg["base_link"]

    # --- link ordering: kinematics_loader.py:100-110 (other_links) and :215-264 (_build_chain)
    other = list(tool_frames) + [c for c in coll_links if c not in tool_frames]
    for k, e in extra.items():
        p = e["parent_link_name"]
        if p not in tool_frames and p not in other:
            other.append(p)
    names = urdf.chain(base, tool_frames[0])
    for l in other:
        if l in names or l in extra:
            continue
        for k in urdf.chain(base, l):
            if k not in names:
                names.append(k)
    for k in extra:
        if k not in names:
            names.append(k)
    # parents must precede children (kinematics_loader.py:396-399); extra links appended last can
    # break this when they sit in the middle of the tree (G1 virtual base) -> stable topological sort.
    parent_of = {n: (None if n == base else urdf.parent_map[n]["parent"]) for n in names}
    ordered, placed = [], set()
    pending = list(names)
    while pending:
        rest = []
        for n in pending:
            if parent_of[n] is None or parent_of[n] in placed:
                ordered.append(n)
                placed.add(n)
            else:
                rest.append(n)
        if len(rest) == len(pending):
            raise ValueError("kinematic tree is not connected")
        pending = rest
    names = ordered
    idx = {n: i for i, n in enumerate(names)}
    bodies = [urdf.body(n, base=(n == base)) for n in names]

    # --- lock joints: joint becomes FIXED with fixed = origin * J(lock value) (loader :678-835)
    lock = cfg.get("lock_joints") or {}
    for b in bodies:
        if b.joint_type != FIXED and b.joint_name in lock:
            theta = b.offset[0] * float(lock[b.joint_name]) + b.offset[1]
            b.fixed = b.fixed @ _joint_motion(b.joint_type, theta)
            b.joint_type, b.offset = FIXED, [1.0, 0.0]

    # --- joint ordering: tree order of first appearance, then cspace.joint_names order if given
    joint_names: List[str] = []
    for b in bodies:
        if b.joint_type != FIXED and b.joint_name not in joint_names:
            joint_names.append(b.joint_name)
    cs = cfg.get("cspace") or {}
    cs_names = [j for j in (cs.get("joint_names") or []) if j in joint_names]
    if len(cs_names) == len(joint_names):
        joint_names = cs_names
    D = len(joint_names)
    nl = len(names)

    fixed = np.stack([b.fixed[:3, :4] for b in bodies]).astype(np.float32)
    link_map = np.array([0 if b.parent is None else idx[b.parent] for b in bodies], dtype=np.int16)
    assert all(link_map[i] < i for i in range(1, nl)), "parents must precede children"
    joint_map = np.array([-1 if b.joint_type == FIXED else joint_names.index(b.joint_name) for b in bodies], dtype=np.int16)
    joint_type = np.array([b.joint_type for b in bodies], dtype=np.int8)
    joint_off = np.array([b.offset for b in bodies], dtype=np.float32)
    tool_map = np.array([idx[t] for t in tool_frames], dtype=np.int16)
    masses = np.stack([b.mass_com for b in bodies]).astype(np.float32)

    # CSR of ancestors per link, base..link (kinematics_loader.py:421-441)
    chain_data, chain_off = [], [0]
    for n in names:
        chain_data += [idx[k] for k in urdf.chain(base, n)]
        chain_off.append(len(chain_data))
    # CSR links per joint (kinematics_loader.py:300-365)
    jl_data, jl_off = [], [0]
    for d in range(D):
        jl_data += [i for i in range(nl) if joint_map[i] == d]
        jl_off.append(len(jl_data))
    affects = np.zeros((D, len(tool_frames)), dtype=bool)
    for d in range(D):
        for e, t in enumerate(tool_frames):
            ch = set(idx[k] for k in urdf.chain(base, t))
            affects[d, e] = any(i in ch for i in jl_data[jl_off[d]:jl_off[d + 1]])

    # --- collision spheres (kinematics_loader.py:848-915; loader_cfg.py:177-183 extra spheres r=-100)
    spheres_cfg = dict(cfg.get("collision_spheres") or {})
    for k, n in (cfg.get("extra_collision_spheres") or {}).items():
        spheres_cfg[k] = [{"center": [0.0, 0.0, 0.0], "radius": -100.0} for _ in range(n)]
    buf = cfg.get("collision_sphere_buffer", 0.0)
    sph, sph_link = [], []
    for l in coll_links:
        off = buf if isinstance(buf, (int, float)) else buf.get(l, 0.0)
        for s in spheres_cfg[l]:
            sph.append(list(s["center"]) + [s["radius"] + off])
            sph_link.append(idx[l])
    link_spheres = np.array(sph, dtype=np.float32).reshape(-1, 4)
    sph_link = np.array(sph_link, dtype=np.int16)
    S = link_spheres.shape[0]

    # --- self-collision pair list (self_collision_params.py:63-125,127-205)
    ignore = cfg.get("self_collision_ignore") or {}
    pad_cfg = cfg.get("self_collision_buffer") or {}
    padding = np.zeros(S, dtype=np.float32)
    allowed = np.zeros((S, S), dtype=bool)
    for a in coll_links:
        ia = np.nonzero(sph_link == idx[a])[0]
        padding[ia] = pad_cfg.get(a, 0.0)
        for bname in coll_links:
            if bname == a or bname in ignore.get(a, []):
                continue
            ib = np.nonzero(sph_link == idx[bname])[0]
            allowed[np.ix_(ia, ib)] = True
    allowed = allowed & allowed.T        # torch.minimum(d, d^T): ignored in either direction
    ii, jj = np.nonzero(np.triu(allowed, k=1))
    pairs = np.stack([ii, jj], axis=1).astype(np.int16)

    # --- limits (cspace_params / joint limits)
    def per_joint(fn, default):
        out = np.zeros((2, D), dtype=np.float32)
        for d, jn in enumerate(joint_names):
            b = next(b for b in bodies if b.joint_type != FIXED and b.joint_name == jn and b.mimic_of is None)
            lo, hi = fn(b) if fn(b) is not None else default
            out[:, d] = [lo, hi]
        return out
    pos_lim = per_joint(lambda b: b.limits, (-6.28, 6.28))
    vel_lim = per_joint(lambda b: b.vel_limits, (-2.0, 2.0))
    eff = per_joint(lambda b: [-b.effort, b.effort], (-1e4, 1e4))

    def scalar_or_list(v, default):
        v = default if v is None else v
        a = np.full(D, v, dtype=np.float32) if np.isscalar(v) else np.array(
            [v[(cs.get("joint_names") or joint_names).index(j)] for j in joint_names], dtype=np.float32)
        return np.stack([-a, a])
    acc_lim = scalar_or_list(cs.get("max_acceleration"), 10.0)
    jerk_lim = scalar_or_list(cs.get("max_jerk"), 500.0)
    dq = cs.get("default_joint_position")
    if dq is not None:
        all_names = cs.get("joint_names") or joint_names
        dq = np.array([dq[all_names.index(j)] for j in joint_names], dtype=np.float32)
    else:
        dq = (0.5 * (pos_lim[0] + pos_lim[1])).astype(np.float32)

    return RobotModel(
        name=name, link_names=names, joint_names=joint_names, tool_frames=tool_frames,
        fixed_transforms=fixed, link_map=link_map, joint_map=joint_map, joint_map_type=joint_type,
        joint_offset_map=joint_off, tool_frame_map=tool_map, link_spheres=link_spheres,
        link_sphere_idx_map=sph_link, link_chain_data=np.array(chain_data, dtype=np.int16),
        link_chain_offsets=np.array(chain_off, dtype=np.int16),
        joint_links_data=np.array(jl_data, dtype=np.int16),
        joint_links_offsets=np.array(jl_off, dtype=np.int16),
        joint_affects_endeffector=affects.reshape(-1), link_masses_com=masses,
        collision_pairs=pairs, sphere_padding=padding, position_limits=pos_lim,
        velocity_limits=vel_lim, acceleration_limits=acc_lim, jerk_limits=jerk_lim,
        effort_limits=eff, default_joint_position=dq, collision_link_names=coll_links)
