#!/usr/bin/env python
"""Golden vectors for the L-BFGS step, produced by the REFERENCE's own torch implementation
(curobo/_src/optim/gradient/lbfgs_jit_helpers.py: jit_lbfgs_update_buffers + jit_lbfgs_compute_step_direction,
the code path `LBFGSOpt` takes when `use_cuda_kernel_step_direction=False`), imported from /root/reference in the
build container and run on the CPU.

    python tests/golden/make_optim_golden.py        # rewrites tests/golden/lbfgs_reference_torch.npz

Six consecutive optimizer iterations on strictly convex quadratics (so every curvature pair has y.s > 0 -- the
regime where the torch path and the CUDA kernel agree; they treat y.s <= 0 differently:
lbfgs_jit_helpers.py:64-66 vs lbfgs_step_helpers.cuh:134-137).  Stored per iteration: the kernel's inputs
(q, grad_q and the buffers BEFORE the call) and outputs (step, buffers AFTER).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("CUROBO_REFERENCE", "/root/reference"))

from curobo._src.optim.gradient.lbfgs_jit_helpers import (  # noqa: E402
    jit_lbfgs_compute_step_direction, jit_lbfgs_update_buffers)


def main():
    torch.manual_seed(0)
    out = {}
    for tag, (B, V, m) in {"ik": (12, 7, 7), "trajopt": (3, 112, 15), "small": (5, 4, 3)}.items():
        A = torch.randn(B, V, V)
        A = A @ A.transpose(1, 2) + 0.5 * torch.eye(V)
        bvec = torch.randn(B, V, 1)
        grad = lambda x: (A @ x.unsqueeze(-1) - bvec).squeeze(-1)  # noqa: E731
        s_buf, y_buf = torch.zeros(m, B, V, 1), torch.zeros(m, B, V, 1)
        rho, alpha = torch.zeros(m, B, 1, 1), torch.zeros(m, B, 1, 1)
        x = torch.randn(B, V)
        x_0, grad_0 = x.clone().unsqueeze(-1), grad(x).unsqueeze(-1)
        x = x - 0.05 * grad(x)
        rec = {k: [] for k in ("q", "grad_q", "rho_in", "y_in", "s_in", "x0_in", "g0_in", "step", "rho_out", "y_out", "s_out")}
        for it in range(6):
            g = grad(x)
            rec["q"].append(x.numpy().copy())
            rec["grad_q"].append(g.numpy().copy())
            rec["rho_in"].append(rho[:, :, 0, 0].numpy().copy())
            rec["y_in"].append(y_buf[..., 0].numpy().copy())
            rec["s_in"].append(s_buf[..., 0].numpy().copy())
            rec["x0_in"].append(x_0[..., 0].numpy().copy())
            rec["g0_in"].append(grad_0[..., 0].numpy().copy())
            s_buf, y_buf, rho, x_0, grad_0 = jit_lbfgs_update_buffers(x, g.unsqueeze(1), s_buf, y_buf, rho, x_0, grad_0, True)
            step = jit_lbfgs_compute_step_direction(alpha, rho, y_buf, s_buf, g.unsqueeze(1), m, 0.01, True)
            step = step.reshape(B, V)
            rec["step"].append(step.numpy().copy())
            rec["rho_out"].append(rho[:, :, 0, 0].numpy().copy())
            rec["y_out"].append(y_buf[..., 0].numpy().copy())
            rec["s_out"].append(s_buf[..., 0].numpy().copy())
            x = x + 0.3 * step
        for k, v in rec.items():
            out[f"{tag}_{k}"] = np.stack(v).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "lbfgs_reference_torch.npz"), **out)
    print("wrote", os.path.join(HERE, "lbfgs_reference_torch.npz"), {k: v.shape for k, v in out.items() if k.endswith("step")})


if __name__ == "__main__":
    main()
