#!/bin/bash
# round 2, session 3p: depth -> TSDF -> ESDF chain on the GPU (new kernels): tests + timing at 256^3
mkdir -p gpurun_out/r3p; O=gpurun_out/r3p
(timeout 900 python -m pytest tests/test_gpu_zz_edt.py -m gpu -q -x -p no:cacheprovider) 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r3p/timing.log
import sys, os, numpy as np, torch
sys.path.insert(0, "tests")
import test_gpu_zz_edt as t
from curobo_b200.esdf import DenseTSDF, DenseESDFBuilder
shape, voxel = (256, 256, 256), 0.01
trunc = 4 * voxel
K, pos, quat, depth, radius = t.depth_scene(shape, voxel, n_cam=2, hw=(480, 640), seed=1)
T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
tsdf = DenseTSDF(shape, voxel, trunc, "cuda:0", depth_min=0.05, depth_max=10.0, minimum_tsdf_weight=0.5)
b = DenseESDFBuilder(shape, voxel, trunc, "cuda:0")
d, k, p, q = T(depth), T(K), T(pos), T(quat)
def step():
    tsdf.integrate(d, k, p, q)
    return b.compute(tsdf.combined_sdf(), None)
for _ in range(3): step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
n = 20
ev[0].record()
for _ in range(n): tsdf.integrate(d, k, p, q)
ev[1].record()
for _ in range(n): c = tsdf.combined_sdf()
ev[2].record()
for _ in range(n): f = b.compute(c, None)
ev[3].record(); torch.cuda.synchronize()
print("256^3, 2 cameras 480x640: integrate %.3f ms, combined sdf %.3f ms, seed + transform + signed distance %.3f ms" % (ev[0].elapsed_time(ev[1]) / n, ev[1].elapsed_time(ev[2]) / n, ev[2].elapsed_time(ev[3]) / n))
ff = f.float()
print("observed voxels %.1f %%, sites %d, esdf min %.3f max %.3f" % (100 * float((c < 1e9).float().mean()), int((b.site_index.view(-1) >= 0).sum()), float(ff.min()), float(ff[ff < 1e3].max())))
PY
