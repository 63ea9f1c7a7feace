"""RNEA row: our kernels vs the reference's kernels (oracle/_ref, serial path) on the same inputs,
CUDA events on the launching stream, L2 flushed between iterations. Prints one JSON line per (robot, batch)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_kernels  # noqa: E402
from dynamics_cases import make_case, model_args  # noqa: E402
from curobo_b200.backends import dynamics as dc  # noqa: E402

DEV = "cuda:0"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)  # noqa: E731


def timed(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    flush = torch.zeros(160 * 1024 * 1024 // 4, device=DEV)
    cases = (("franka", 16384), ("franka", 30720), ("g1_29", 30720))
    if "--only" in sys.argv:
        name = sys.argv[sys.argv.index("--only") + 1]
        cases = tuple(c for c in cases if c[0] == name)[:1]
    for robot, B in cases:
        c = make_case(robot, 64, 1)
        rep = B // 64
        m = model_args(c)
        model = tuple(T(x) for x in m) + (T(c["starts"]), T(c["order"]))
        nl, D, nlev = c["nl"], c["D"], c["n_levels"]
        q, qd, qdd, gt = (T(np.tile(c[k], (rep, 1))) for k in ("q", "qd", "qdd", "grad_tau"))
        tau = torch.zeros((B, D), device=DEV); cache = torch.zeros((B, nl * 20), device=DEV)
        g = [torch.zeros((B, D), device=DEV) for _ in range(3)]
        ours_f = timed(lambda: dc.launch_rnea_forward(tau, q, qd, qdd, *model, cache, B, nl, D, nlev), flush)
        ours_b = timed(lambda: dc.launch_rnea_backward(*g, gt, q, qd, *model, cache, B, nl, D, nlev), flush)
        out = dict(row="rnea", robot=robot, batch=B, n_links=nl, n_dof=D, ours_forward_ms=ours_f, ours_backward_ms=ours_b)
        # algorithmic bytes: forward reads q,qd,qdd and writes tau + the cache; the adjoint reads grad_tau,q,qd,cache, writes 3 grads
        fb = B * 4 * (4 * D + nl * 20); bb = B * 4 * (6 * D + nl * 20)
        out.update(forward_GBps=fb / ours_f / 1e6, backward_GBps=bb / ours_b / 1e6)
        if ref_kernels.available() and "--no-ref" not in sys.argv:
            rtau, rcache = torch.zeros_like(tau), torch.zeros_like(cache)
            rg = [torch.zeros_like(x) for x in g]

            def ref_bwd():  # the reference accumulates into its gradients: its host zeroes them every call
                for x in rg:
                    x.zero_()
                ref_kernels.rnea_backward(model, gt, q, qd, rcache, nl, D, nlev, out=rg)

            out["ref_forward_ms"] = timed(lambda: ref_kernels.rnea_forward(model, q, qd, qdd, nl, D, nlev, out=(rtau, rcache)), flush)
            out["ref_backward_ms"] = timed(ref_bwd, flush)
            out["note"] = "reference kernels, serial path, preallocated outputs; its backward includes the three gradient memsets it needs"
        print(json.dumps(out))


if __name__ == "__main__":
    main()
