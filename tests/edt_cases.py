"""Seeded occupancy grids for the nearest-site transform tests: random points, box shells (the analytic scenes of the
reference's ESDF tests are boxes), and the edge cases -- empty grid, a single site, every voxel a site, degenerate dimensions,
dimensions that are not multiples of the 32-column tile."""
import numpy as np


def occupancy(kind, shape, seed=0, p=0.02):
    rng = np.random.default_rng(seed)
    occ = np.zeros(shape, bool)
    if kind == "random":
        occ = rng.random(shape) < p
    elif kind == "empty":
        pass
    elif kind == "full":
        occ[:] = True
    elif kind == "single":
        occ[tuple(int(rng.integers(0, n)) for n in shape)] = True
    elif kind == "corner":
        occ[0, 0, 0] = True
        occ[-1, -1, -1] = True
    elif kind == "shells":                       # surfaces of a few random boxes
        for _ in range(3):
            lo = [int(rng.integers(0, max(1, n - 2))) for n in shape]
            hi = [int(min(n - 1, l + rng.integers(1, max(2, n // 2)))) for n, l in zip(shape, lo)]
            box = np.zeros(shape, bool)
            box[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] = True
            inner = np.zeros(shape, bool)
            inner[lo[0] + 1:hi[0], lo[1] + 1:hi[1], lo[2] + 1:hi[2]] = True
            occ |= box & ~inner
    elif kind == "plane":                        # many exact ties
        occ[shape[0] // 2, :, :] = True
    else:
        raise ValueError(kind)
    return occ


SMALL = [("random", (9, 7, 11), 0.03), ("random", (6, 6, 6), 0.3), ("random", (12, 5, 3), 0.06), ("random", (1, 1, 9), 0.3),
         ("random", (1, 7, 1), 0.3), ("random", (5, 1, 1), 0.5), ("empty", (8, 8, 8), 0), ("full", (4, 4, 4), 0),
         ("single", (7, 9, 5), 0), ("corner", (10, 6, 8), 0), ("shells", (14, 12, 10), 0), ("plane", (7, 6, 5), 0)]
MEDIUM = [("random", (40, 33, 50), 0.002), ("random", (33, 65, 31), 0.05), ("shells", (48, 40, 36), 0), ("single", (37, 5, 70), 0),
          ("plane", (20, 35, 33), 0), ("corner", (64, 64, 64), 0)]
