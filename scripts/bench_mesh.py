"""Side measurement: sphere-vs-mesh collision (cb200_sphere_mesh_collision) on Franka-sized sphere batches against closed meshes
of growing triangle count, next to the cuboid / ESDF operator on the same spheres.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from curobo_b200.mesh import MeshData, MeshWorld, box_mesh, icosphere  # noqa: E402
from curobo_b200.scene import CollisionBuffer, CuboidData, SceneData, SphereObstacleCollision  # noqa: E402
from curobo_b200.world import make_benchmark_cuboid_world  # noqa: E402

DEV = "cuda:0"


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    rng = np.random.default_rng(0)
    B, S = 16384, 65
    sph = np.zeros((B, 1, S, 4), np.float32)
    sph[..., :3] = rng.uniform(-0.8, 0.8, (B, 1, S, 3))
    sph[..., 3] = rng.uniform(0.02, 0.06, (B, 1, S))
    sp = torch.as_tensor(sph).to(DEV)
    buf = CollisionBuffer.from_shape(sph.shape, DEV)
    w, eta = torch.tensor([5000.0], device=DEV), torch.tensor([0.02], device=DEV)
    out = {"spheres": B * S}
    cub = CuboidData.from_world(make_benchmark_cuboid_world(), DEV)
    out["cuboids_2_ms"] = timed(lambda: SphereObstacleCollision.apply(sp, buf, SceneData(cuboid=cub), w, eta, None, None, False, False))
    for sub in (1, 3, 5):
        v, f = icosphere(0.3, sub)
        vb, fb = box_mesh([2.2, 2.2, 0.2])
        mesh = MeshData.from_world(MeshWorld.create([{"vertices": v, "faces": f, "pose": [0.4, 0, 0.3, 1, 0, 0, 0]},
                                                     {"vertices": vb, "faces": fb, "pose": [0, 0, -0.1, 1, 0, 0, 0]}]), DEV)
        ms = timed(lambda: SphereObstacleCollision.apply(sp, buf, SceneData(mesh=mesh), w, eta, None, None, False, False))
        out[f"mesh_{f.shape[0]}+12_triangles_ms"] = ms
        out[f"mesh_{f.shape[0]}+12_queries_per_s"] = 2 * B * S / (ms * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
