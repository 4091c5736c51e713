"""GPU parity of ik::pik (ikb_pik_solve_batch, table-driven kernel) against the oracle's restatement of pik.cpp."""
import numpy as np
import pytest

import ik_b200 as ik
from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import make_workload, oracle_model, oracle_problem_like

pytestmark = pytest.mark.gpu
NT = 8


def _torch():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def _solve(pb, q0, tg, prm, dtype="f64"):
    torch = _torch()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    out = ik.pik_batch(pb, torch.tensor(q0.T.copy(), dtype=tdt, device="cuda:0"), torch.tensor(tg.T.copy(), dtype=tdt, device="cuda:0"), prm)
    torch.cuda.synchronize()
    return (out["q"].cpu().numpy().T.astype(np.float64), out["success"].cpu().numpy().astype(bool), out["iters"].cpu().numpy(),
            out["resid"].cpu().numpy().astype(np.float64))


@pytest.mark.parametrize("lambdas", [[1e-2, 1e-1], [1.0, 1.0], [1e-1, 1e-2]])
def test_pik_two_levels_cassie(lambdas):
    """The demo's declared task set: three priority-0 tasks (10 rows) and the posture task on level 1 (16 rows) -- the
    level-1 step lives in the null space of level 0 (pik.cpp:47-62).  Default damping 1.0 per level (pik.hpp:31) and two
    sharper settings.  The damped step is an LDL^T solve in the kernel and an SVD sum in the oracle (pik.cpp:5-21)."""
    pb = W.cassie_demo_posture_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 600
    q0, tg, _ = make_workload(pb, om, B, seed=71, standing=W.CASSIE_STANDING)
    q_ref, ok_ref, it_ref, res_ref = O.pik_batch(opb, q0, tg, O.pik_params(lambdas=lambdas), nthreads=NT)
    q, ok, it, res = _solve(pb, q0, tg, ik.pik_parameters(lambdas=lambdas))
    same = (ok == ok_ref) & (it == it_ref)
    print("pik cassie lambdas=%s: converged gpu/ref %d/%d, same flags+steps %.4f, mean steps %.1f, max|dq| %.2e"
          % (lambdas, ok.sum(), ok_ref.sum(), same.mean(), it_ref.mean(), np.abs(q[same & ok] - q_ref[same & ok]).max()))
    assert (ok == ok_ref).all() and same.mean() > 0.99
    assert np.abs(q[same & ok] - q_ref[same & ok]).max() < 1e-6
    assert np.abs(res[same & ok] - res_ref[same & ok]).max() < 1e-9


def test_pik_single_level_equals_dls_on_gpu():
    """One priority level: ik::pik's step is ik::dls's with damping = lambda; both GPU paths must agree with each other
    (specialised DLS kernel vs table-driven PIK kernel) and with the oracle."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, 1024, seed=72, standing=W.CASSIE_STANDING)
    q_ref, ok_ref, it_ref, _ = O.pik_batch(opb, q0, tg, O.pik_params(lambdas=[1e-2]), nthreads=NT)
    q, ok, it, res = _solve(pb, q0, tg, ik.pik_parameters(lambdas=[1e-2]))
    assert (ok == ok_ref).all() and (it == it_ref).mean() > 0.99
    torch = _torch()
    d = ik.dls_batch(pb, torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0"))
    okd = d["success"].cpu().numpy().astype(bool)
    both = ok & okd & (it == d["iters"].cpu().numpy()) & (it < 30)
    assert (ok == okd).mean() > 0.98 and both.mean() > 0.8
    assert np.abs(q[both] - d["q"].cpu().numpy().T[both]).max() < 1e-6


def test_pik_rank_deficient_level_and_host_path():
    """Two tasks that conflict on the lower level: the same foot position asked for on level 0 and, shifted, on level 1 --
    J_1 P loses rank (singular values 1.4 ... 0.04, 1.6e-15): the COD threshold drops the last direction (pik.cpp:59-61)
    and the level-1 damping (1.0, the reference's default) keeps the weak ones from blowing the step up.  Through the host
    entry point and the one-problem call."""
    m = W.cassie_model()
    pb = ik.InverseKinematicsProblem(m, 1)
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full), 0)
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position), 0)
    pb.add_frame_task("fl2", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position), 1)
    pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Position), 1)
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 200
    q0, tg, _ = make_workload(pb, om, B, seed=73, standing=W.CASSIE_STANDING)
    off = pb.target_offset(pb.get_frame_task("fl2"))
    tg[:, off + 9:off + 12] += 0.05            # level 1 disagrees with level 0 about the left foot
    prm = O.pik_params(max_iterations=30, lambdas=[1e-2, 1.0])
    q_ref, ok_ref, it_ref, res_ref = O.pik_batch(opb, q0, tg, prm, nthreads=NT)
    out = ik.pik_batch_host(pb, q0, tg, ik.pik_parameters(max_iterations=30, lambdas=[1e-2, 1.0]))
    ok = out["success"].astype(bool)
    same = (ok == ok_ref) & (out["iters"] == it_ref)
    assert (ok == ok_ref).mean() > 0.99 and same.mean() > 0.97
    assert np.abs(out["q"][same & ok] - q_ref[same & ok]).max() < 1e-6
    # one problem through ik.pik (pik.hpp:51-54)
    for name, col in (("pelvis", 0), ("fl", 12), ("fl2", 24), ("fr", 36)):
        pb.get_frame_task(name).target[:] = tg[0, col:col + 12]
    data = ik.pik_data(pb)
    q1 = ik.pik(pb, q0[0], data, None, ik.pik_parameters(max_iterations=30, lambdas=[1e-2, 1.0]))
    assert data.success == bool(ok_ref[0]) and data.iterations == it_ref[0] and np.abs(q1 - q_ref[0]).max() < 1e-6
