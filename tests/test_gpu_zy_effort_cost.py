"""GPU parity of the dynamics-aware STATE cost (second half): RNEA -> effort channel of the c-space
STATE kernel -> RNEA adjoint, composed on the host by curobo_b200.dynamics.DynamicsStateCost, against the composition of the
two oracles (pinned on the CPU by finite differences, tests/test_dynamics_cpu.py).  Runs last in the suite on purpose: it
is the newest composition."""
import numpy as np
import pytest
import torch

from dynamics_cases import effort_cost_oracle, effort_cost_setup
from curobo_b200.dynamics import Dynamics, DynamicsStateCost

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("robot,B,H", [("franka", 6, 5), ("g1_29", 3, 4)])
def test_dynamics_state_cost_vs_oracle(robot, B, H):
    c, shape, jerk, dt, limits, weight, act, reg = effort_cost_setup(robot, B, H)
    want_c, want_g, want_tau = effort_cost_oracle(c, shape, jerk, dt, limits, weight, act, reg)
    dyn = Dynamics(c["rm"], c["mc"], c["inn"], gravity=(0.0, 0.0, -9.81), device=DEV)
    op = DynamicsStateCost(dyn, limits, weight, act, reg)
    r = lambda x: T(np.asarray(x, np.float32).reshape(shape))  # noqa: E731
    q, qd, qdd = r(c["q"]), r(c["qd"]), r(c["qdd"])
    cost, gp, gv, ga, gj, tau = op.evaluate(q, qd, qdd, T(jerk), T(dt))
    torch.cuda.synchronize()

    def close(got, want, rtol):
        got = got.cpu().numpy()
        assert np.allclose(got, want, rtol=rtol, atol=rtol * max(float(np.abs(want).max()), 1e-30)), float(np.abs(got - want).max())

    close(tau, want_tau, 1e-4)
    close(cost, want_c, 5e-4)
    for got, want in zip((gp, gv, ga, gj), want_g):
        close(got, want, 2e-3)
    # the effort channel is live in this case (hinge + L2 + energy), and a second call reuses the buffers bit for bit
    assert float(np.abs(want_g[2]).max()) > 0
    first = [x.clone() for x in (cost, gp, gv, ga, gj)]
    again = op.evaluate(q, qd, qdd, T(jerk), T(dt))
    torch.cuda.synchronize()
    for a_, b_ in zip(first, again[:5]):
        assert torch.equal(a_, b_)
