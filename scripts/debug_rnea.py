"""Prints every pairwise difference between our RNEA kernels, the oracle and the reference's compiled kernels."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_kernels
from dynamics_cases import make_case, model_args, pack_cache
from curobo_b200.backends import dynamics as dc
from oracle import dynamics_oracle as do
DEV = "cuda:0"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
def d(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return f"max|d|={np.abs(a-b).max():.3e} (scale {np.abs(b).max():.3e}) worst col {np.abs(a-b).max(axis=0).argmax() if a.ndim==2 else -1}"
for robot, B, seed in [("franka", 9, 1), ("g1_29", 5, 2)]:
    c = make_case(robot, B, seed); m = model_args(c)
    model = tuple(T(x) for x in m) + (T(c["starts"]), T(c["order"]))
    nl, D, nlev = c["nl"], c["D"], c["n_levels"]
    q, qd, qdd, gt = T(c["q"]), T(c["qd"]), T(c["qdd"]), T(c["grad_tau"])
    tau = torch.zeros((B, D), device=DEV); cache = torch.zeros((B, nl*20), device=DEV)
    dc.launch_rnea_forward(tau, q, qd, qdd, *model, cache, B, nl, D, nlev)
    g = [torch.zeros((B, D), device=DEV) for _ in range(3)]
    dc.launch_rnea_backward(*g, gt, q, qd, *model, cache, B, nl, D, nlev)
    torch.cuda.synchronize()
    tw, cw = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    gw = do.rnea_backward(c["grad_tau"], c["q"], c["qd"], cw, *m)
    rt, rc = ref_kernels.rnea_forward(model, q, qd, qdd, nl, D, nlev)
    rg = ref_kernels.rnea_backward(model, gt, q, qd, rc, nl, D, nlev)
    torch.cuda.synchronize()
    print(robot, "nl", nl, "D", D, "levels", nlev)
    print(" tau ours-oracle", d(tau.cpu(), tw)); print(" tau ref-oracle ", d(rt.cpu(), tw))
    pc = pack_cache(cw, nl).reshape(B, nl, 20)[:, :, :18].reshape(B, -1)
    print(" cache ours-oracle", d(cache.cpu().view(B, nl, 20)[:, :, :18].reshape(B, -1), pc))
    print(" cache ref-oracle ", d(rc.cpu().view(B, nl, 20)[:, :, :18].reshape(B, -1), pc))
    for i, n in enumerate(("gq", "gqd", "gqdd")):
        print(f" {n} ours-oracle", d(g[i].cpu(), gw[i])); print(f" {n} ref-oracle ", d(rg[i].cpu(), gw[i])); print(f" {n} ours-ref   ", d(g[i].cpu(), rg[i].cpu()))
    np.set_printoptions(precision=4, linewidth=250, suppress=True)
    print(" row0 gq ours  ", g[0][0].cpu().numpy()); print(" row0 gq oracle", gw[0][0]); print(" row0 gq ref   ", rg[0][0].cpu().numpy())
    print(" row0 gqd ours  ", g[1][0].cpu().numpy()); print(" row0 gqd oracle", gw[1][0]); print(" row0 gqd ref   ", rg[1][0].cpu().numpy())
