import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ref_kernels
from optim_cases import lbfgs_case
from curobo_b200.backends import optimization as oc
from oracle import optim_oracle as oo
DEV = "cuda:0"
T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
for (B, V, m, fill) in [(3, 33, 3, None), (3, 64, 7, None), (3, 112, 7, None), (3, 112, 27, None), (3, 112, 27, 5), (3, 112, 15, None), (3, 40, 27, None), (2, 1000, 31, None), (3, 168, 15, None)]:
    c = lbfgs_case(seed=6, B=B, V=V, m=m, fill=fill)
    t = {k: T(v) for k, v in c.items()}
    step = torch.zeros((B, V), device=DEV)
    oc.launch_lbfgs_step(step, t["rho"], t["Y"], t["S"], t["q"], t["grad_q"], t["x_0"], t["grad_0"], 0.01, B, m, V, True, True)
    w = oo.lbfgs_step(c["rho"], c["Y"], c["S"], c["q"], c["grad_q"], c["x_0"], c["grad_0"], 0.01, True)[0]
    w64 = None
    res = []
    for shared in (True, False):
        r = {k: T(v) for k, v in c.items()}
        rs = torch.zeros_like(step)
        ref_kernels.lbfgs_step(rs, r["rho"], r["Y"], r["S"], r["q"], r["x_0"], r["grad_0"], r["grad_q"], 0.01, True, shared)
        torch.cuda.synchronize()
        res.append(rs)
    sc = float(np.abs(w).max())
    print(f"B{B} V{V} m{m} fill{fill}: ours-oracle {float(np.abs(step.cpu().numpy()-w).max())/sc:.2e}  ours-refS {float((step-res[0]).abs().max())/sc:.2e}  ours-refG {float((step-res[1]).abs().max())/sc:.2e}  refS-refG {float((res[0]-res[1]).abs().max())/sc:.2e}  rho_eq {torch.equal(t['rho'], r['rho'])}")
