"""Exact nearest-site transform: our three in-place passes vs the reference's PBA+ kernels (oracle/_ref,
five launches + copy) on the same grids, CUDA events, grid re-seeded before every iteration (outside the timed region).
Prints one JSON line per grid.  NOT YET RUN ON A B200 (written after round 1's GPU budget was spent)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_kernels  # noqa: E402
from edt_cases import occupancy  # noqa: E402
from curobo_b200.esdf import ParallelBandingEDT, seed_sites_from_occupancy  # noqa: E402
from oracle import edt_oracle as E  # noqa: E402

DEV = "cuda:0"


def timed(fn, reseed, iters=10, warm=2):
    ts = []
    for i in range(iters + warm):
        reseed()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [128, 256]
    for n in sizes:
        shape = (n, n, n)
        occ = occupancy("shells", shape, seed=11) | occupancy("random", shape, seed=12, p=1e-4)
        fresh = seed_sites_from_occupancy(torch.as_tensor(occ).to(DEV))
        work = fresh.clone()
        edt = ParallelBandingEDT(shape, 0.01, torch.device(DEV))
        ours = timed(lambda: edt.propagate(work), lambda: work.copy_(fresh))
        out = dict(row="edt", grid=list(shape), sites=int(occ.sum()), ours_ms=ours, bytes_per_voxel=24,
                   ours_GBps=24 * n ** 3 / ours / 1e6)
        d2 = E.squared_distance(work.cpu().numpy())
        if ref_kernels.available() and "--no-ref" not in sys.argv:
            buf = torch.empty_like(work)
            rwork = fresh.clone()

            def ref():
                err = ref_kernels.lib().ref_pba3d(ref_kernels._p(rwork), ref_kernels._p(buf), n, n, n, 2, ref_kernels._stream(rwork.device))
                assert err == 0

            out["ref_ms"] = timed(ref, lambda: rwork.copy_(fresh))
            out["same_squared_distances"] = bool(np.array_equal(d2, E.squared_distance(rwork.cpu().numpy())))
        print(json.dumps(out))


if __name__ == "__main__":
    main()
