"""GPU parity of the B-spline knot -> state kernels and their adjoint, called through the C ABI
(curobo_b200.backends.trajectory), against
  * the numpy oracle (oracle/bspline_oracle.py), and
  * the REFERENCE's own kernels compiled from /root/reference into oracle/_ref (same nvcc flags): forward and adjoint
    are expected to agree to float rounding of identically ordered arithmetic -> tolerance 2 ulp-ish (rtol 1e-6),
    and bit-exact for the adjoint with power-of-two interpolation steps where the summation order is reproduced.
Tolerances vs the oracle (numpy divides exactly, the kernels use --prec-div=false): rel 2e-5 of the output scale.
"""
import numpy as np
import pytest
import torch

import ref_kernels
from bspline_cases import CASES, case_id, make_case
from curobo_b200.backends import trajectory as trajectory_cu
from curobo_b200.trajectory import (BSplineIdxKernel, ControlSpace, JointState, StateFromBSplineKnot,
                                    get_bspline_interpolation)
from oracle import bspline_oracle as bo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t.to(dt) if dt is not None else t


def dev_case(c):
    d = dict(c)
    d["knots_t"] = T(c["knots"])
    d["start_t"] = tuple(T(x) for x in c["start"])
    d["goal_t"] = tuple(T(x) for x in c["goal"])
    d["sidx_t"], d["gidx_t"] = T(c["start_idx"]), T(c["goal_idx"])
    d["dt_t"], d["imp_t"] = T(c["traj_dt"]), T(c["implicit"])
    d["grads_t"] = tuple(T(g) for g in c["grads"])
    return d


def ours_forward(c):
    B, Tn, D = c["B"], c["T"], c["D"]
    outs = [torch.full((B, Tn, D), float("nan"), device=DEV) for _ in range(4)]
    odt = torch.zeros((B,), device=DEV)
    trajectory_cu.launch_bspline_interpolation_forward_kernel(
        *outs, odt, c["knots_t"], *c["start_t"], *c["goal_t"], c["sidx_t"], c["gidx_t"], c["dt_t"], c["imp_t"], B, Tn, D,
        c["nk"], c["degree"])
    return outs + [odt]


def ours_backward(c):
    out = torch.full((c["B"], c["nk"], c["D"]), float("nan"), device=DEV)
    trajectory_cu.launch_bspline_interpolation_backward_kernel(
        out, *c["grads_t"], c["dt_t"], c["gidx_t"], c["imp_t"], c["B"], c["T"], c["D"], c["nk"], c["degree"])
    return out


@pytest.mark.parametrize("kw", CASES, ids=case_id)
def test_forward_vs_oracle_and_reference(kw):
    c = dev_case(make_case(**kw))
    got = ours_forward(c)
    want = bo.bspline_forward(c["knots"], c["start"], c["goal"], c["start_idx"], c["goal_idx"], c["traj_dt"], c["implicit"],
                              c["T"], c["degree"])
    for k in range(4):
        g, w = got[k].cpu().numpy(), want[k]
        assert np.isfinite(g).all()
        assert np.allclose(g, w, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(w).max())), f"derivative {k} vs oracle"
    assert np.array_equal(got[4].cpu().numpy(), want[4])
    if ref_kernels.available():
        ref = ref_kernels.bspline_forward(c["knots_t"], c["start_t"], c["goal_t"], c["sidx_t"], c["gidx_t"], c["dt_t"],
                                          c["imp_t"], c["T"], c["degree"])
        for k in range(4):
            g, r = got[k].cpu().numpy(), ref[k].cpu().numpy()
            assert np.allclose(g, r, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(r).max())), f"derivative {k} vs reference"
        assert torch.equal(got[4], ref[4])


@pytest.mark.parametrize("kw", CASES, ids=case_id)
def test_backward_vs_oracle_and_reference(kw):
    c = dev_case(make_case(**kw))
    got = ours_backward(c).cpu().numpy()
    want = bo.bspline_backward(*c["grads"], c["traj_dt"], c["goal_idx"], c["implicit"], c["nk"], c["degree"])
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5 * np.abs(want).max())
    steps = c["steps"]
    if ref_kernels.available() and (steps & (steps - 1)) == 0:
        # the reference's shuffle tree is only valid for power-of-two step counts (it mis-pairs lanes otherwise)
        ref = ref_kernels.bspline_backward(c["grads_t"], c["dt_t"], c["gidx_t"], c["imp_t"], c["nk"], c["degree"]).cpu().numpy()
        assert np.allclose(got, ref, rtol=1e-6, atol=1e-6 * np.abs(ref).max())


def test_single_dt_vs_oracle_and_reference():
    c = make_case(seed=21, B=6, nk=8, D=7, steps=4, degree=4, implicit=False)
    Tn = 70
    c = dev_case(dict(c, T=Tn))
    interp_h = np.array([52, 39, 69, 13, 200, 26], np.int32)
    out = JointState.zeros((6, Tn, 7), DEV)
    start, goal = JointState(*c["start_t"]), JointState(*c["goal_t"])
    idt = torch.tensor([0.025], device=DEV)
    get_bspline_interpolation(c["knots_t"], None, start, goal, c["sidx_t"], c["gidx_t"], idt, c["imp_t"], T(interp_h), out, 4)
    want = bo.bspline_forward(c["knots"], c["start"], c["goal"], c["start_idx"], c["goal_idx"], c["traj_dt"], c["implicit"],
                              Tn, 4, interpolation_horizon=interp_h, interpolation_dt=np.float32(0.025))
    got = [out.position, out.velocity, out.acceleration, out.jerk]
    for g, w in zip(got, want[:4]):
        assert np.allclose(g.cpu().numpy(), w, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(w).max()))
    if ref_kernels.available():
        ref = ref_kernels.bspline_single_dt(c["knots_t"], c["start_t"], c["goal_t"], c["sidx_t"], c["gidx_t"], idt, c["imp_t"],
                                            T(interp_h), Tn, 4)
        for g, r in zip(got, ref[:4]):
            assert np.allclose(g.cpu().numpy(), r.cpu().numpy(), rtol=1e-6, atol=1e-6 * max(1.0, float(r.abs().max())))


@pytest.mark.parametrize("implicit", [False, True])
def test_autograd_function_and_state_transition(implicit):
    """StateFromBSplineKnot.forward + loss.backward(): u_act.grad equals the oracle's adjoint of d loss / d state."""
    B, nk, D, steps = 8, 8, 7, 4
    c = dev_case(make_case(seed=31, B=B, nk=nk, D=D, steps=steps, degree=4, implicit=implicit))
    fn = StateFromBSplineKnot(DEV, D, batch_size=B, n_knots=nk, interpolation_steps=steps,
                              use_implicit_goal_state=implicit, control_space=ControlSpace.BSPLINE_4)
    assert fn.padded_horizon == c["T"]
    start = JointState(*c["start_t"])
    goal = JointState(*c["goal_t"], dt=c["dt_t"])
    out = JointState.zeros((B, c["T"], D), DEV)
    u = c["knots_t"].clone().requires_grad_(True)
    seq = fn.forward(start, u, out, start_state_idx=c["sidx_t"], goal_state=goal, goal_state_idx=c["gidx_t"],
                     use_implicit_goal_state=c["imp_t"])
    w = c["grads_t"]
    loss = (seq.position * w[0]).sum() + (seq.velocity * w[1]).sum() + (seq.acceleration * w[2]).sum() + (seq.jerk * w[3]).sum()
    loss.backward()
    want = bo.bspline_backward(*c["grads"], c["traj_dt"], c["goal_idx"], c["implicit"], nk, 4)
    assert np.allclose(u.grad.cpu().numpy(), want, rtol=1e-4, atol=1e-5 * np.abs(want).max())


def test_error_behaviour():
    c = dev_case(make_case(seed=41, B=2, nk=8, D=7, steps=4, degree=4, implicit=False))
    B, Tn, D = c["B"], c["T"], c["D"]
    outs = [torch.zeros((B, Tn, D), device=DEV) for _ in range(4)]
    odt = torch.zeros((B,), device=DEV)
    args = (*outs, odt, c["knots_t"], *c["start_t"], *c["goal_t"], c["sidx_t"], c["gidx_t"], c["dt_t"], c["imp_t"], B, Tn, D, c["nk"])
    with pytest.raises(RuntimeError, match="Unsupported B-spline degree"):
        trajectory_cu.launch_bspline_interpolation_forward_kernel(*args, 6)
    with pytest.raises(ValueError, match="dtype"):
        bad = list(args)
        bad[5] = c["knots_t"].double()
        trajectory_cu.launch_bspline_interpolation_forward_kernel(*bad, 4)
    with pytest.raises(ValueError, match="CUDA-only|device"):
        bad = list(args)
        bad[5] = c["knots_t"].cpu()
        trajectory_cu.launch_bspline_interpolation_forward_kernel(*bad, 4)
    g = [torch.zeros((B, 5, D), device=DEV) for _ in range(4)]
    out = torch.zeros((B, c["nk"], D), device=DEV)
    with pytest.raises(RuntimeError, match="horizon must be greater than 5"):
        trajectory_cu.launch_bspline_interpolation_backward_kernel(out, *g, c["dt_t"], c["gidx_t"], c["imp_t"], B, 5, D, c["nk"], 4)
    g = [torch.zeros((B, 10, D), device=DEV) for _ in range(4)]
    with pytest.raises(RuntimeError, match="interpolation_steps is 0"):
        trajectory_cu.launch_bspline_interpolation_backward_kernel(out, *g, c["dt_t"], c["gidx_t"], c["imp_t"], B, 10, D, c["nk"], 4)
    # after the rejected launches the library and torch are still healthy
    torch.cuda.synchronize()
    assert torch.isfinite(ours_forward(c)[0]).all()


def test_full_size_adjoint_property_and_graph_capture():
    """MPC / trajopt scale (1024 seeds x 16 knots x 7 dof, 4 steps, degree 4 -> 85 rows): <J du, g> == <du, J^T g> on the
    GPU alone (size-independent property), and both launches are CUDA-graph capturable."""
    B, nk, D, steps, deg = 1024, 16, 7, 4, 4
    c = dev_case(make_case(seed=51, B=B, nk=nk, D=D, steps=steps, degree=deg, implicit=False, mixed_implicit=True))
    base = ours_forward(c)[:4]
    du = torch.randn_like(c["knots_t"])
    c2 = dict(c, knots_t=(c["knots_t"] + du).contiguous())
    pert = ours_forward(c2)[:4]
    lhs = sum(((p.double() - b.double()) * g.double()).sum() for b, p, g in zip(base, pert, c["grads_t"]))
    scale = sum(((p.double() - b.double()) * g.double()).abs().sum() for b, p, g in zip(base, pert, c["grads_t"]))
    gk = ours_backward(c)
    rhs = (gk.double() * du.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-5 * float(scale)

    outs = [torch.zeros_like(base[0]) for _ in range(4)]
    odt = torch.zeros((B,), device=DEV)
    gout = torch.zeros_like(gk)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            trajectory_cu.launch_bspline_interpolation_forward_kernel(
                *outs, odt, c["knots_t"], *c["start_t"], *c["goal_t"], c["sidx_t"], c["gidx_t"], c["dt_t"], c["imp_t"], B,
                c["T"], D, nk, deg)
            trajectory_cu.launch_bspline_interpolation_backward_kernel(
                gout, *c["grads_t"], c["dt_t"], c["gidx_t"], c["imp_t"], B, c["T"], D, nk, deg)
        graph.replay()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], base[0]) and torch.equal(outs[3], base[3])
    assert torch.equal(gout, gk)
