"""Property cases restated from the reference's voxel-collision tests
(curobo/tests/_src/geom/sdf/test_voxel_collision.py:636-1156).  Each case: (world, spheres [B,H,S,4],
eta, sweep, check(cost, grad)).  Run against the oracle on CPU and the CUDA kernels on the GPU."""
import numpy as np

from curobo_b200.world import make_empty_esdf, make_single_box_esdf

_BOX = dict(grid_dims=(0.5, 0.5, 0.5), voxel_size=0.01, grid_center=(0, 0, 0), box_center=(0, 0, 0),
            box_half=(0.05, 0.05, 0.05))


def sph(*rows):
    return np.array(rows, np.float32).reshape(1, len(rows), 1, 4)


def cases():
    out = []
    empty = make_empty_esdf(dims=(1.0, 1.0, 1.0))
    out.append(("free_space_zero", empty, sph([0, 0, 0, 0.01]), 0.02, False,
                lambda c, g: abs(float(c.reshape(-1)[0])) < 1e-5))                                   # :640-653
    out.append(("multi_free_zero", empty, sph([0.1, 0, 0, 0.01], [-0.1, 0, 0, 0.01], [0, 0.1, 0, 0.01]), 0.02, False,
                lambda c, g: bool((c == 0).all())))                                                  # :656-672
    out.append(("outside_grid_zero", make_empty_esdf(dims=(0.2, 0.2, 0.2)), sph([5, 5, 5, 0.01]), 0.02, False,
                lambda c, g: abs(float(c.reshape(-1)[0])) < 1e-5))                                   # :675-685
    box = make_single_box_esdf(**_BOX)
    out.append(("inside_box_cost", box, sph([0, 0, 0, 0.01]), 0.02, False, lambda c, g: float(c.reshape(-1)[0]) > 0))
    out.append(("far_zero", box, sph([0.2, 0.2, 0.2, 0.01]), 0.02, False,
                lambda c, g: abs(float(c.reshape(-1)[0])) < 1e-5))
    # worked example :735-738: sphere at the box face, r=0.02, eta=0.02 -> pen = 0.04 > eta
    # -> cost = pen - eta/2 = 0.03 (weight 1); fp16 storage + trilinear smoothing of the face -> loose tol
    out.append(("surface_cost_pen_0.04", box, sph([0.05, 0, 0, 0.02]), 0.02, False,
                lambda c, g: abs(float(c.reshape(-1)[0]) - 0.03) < 2e-3))
    out.append(("grad_nonzero_inside", box, sph([0.02, 0, 0, 0.01]), 0.02, False,
                lambda c, g: float(np.abs(g.reshape(-1, 4)[0, :3]).sum()) > 0))
    out.append(("deeper_higher", box, sph([0, 0, 0, 0.01], [0.04, 0, 0, 0.01]), 0.02, False,
                lambda c, g: float(c.reshape(-1)[0]) > float(c.reshape(-1)[1])))
    # swept, stationary trajectory == static (:1014): same sphere at every h, sweep adds no samples
    stat = np.tile(np.array([0.045, 0.01, 0.0, 0.015], np.float32), (1, 4, 1, 1))
    out.append(("swept_stationary_equals_static", box, stat, 0.02, "both", None))
    return out
