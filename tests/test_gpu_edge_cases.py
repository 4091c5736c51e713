"""GPU edge cases of the batched entry points: ragged / tiny / empty batches, optional outputs, batch sizes around the
one-wave threshold of the two-launch scheduler, concurrent solves on several streams, strided and broadcast views."""
import ctypes as C
import os

import numpy as np
import pytest

import ik_b200 as ik
from ik_b200 import _capi as capi
from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import make_workload, oracle_model, oracle_problem_like

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def cassie():
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    return pb, om, oracle_problem_like(pb, om)


def _gpu(pb, q0, tg, prm=None):
    torch = _torch()
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0"), prm)
    torch.cuda.synchronize()
    return out["q"].cpu().numpy().T, out["success"].cpu().numpy().astype(bool), out["iters"].cpu().numpy(), out["resid"].cpu().numpy()


def test_empty_batch_is_a_no_op(cassie):
    pb, _, _ = cassie
    out = ik.dls_batch_host(pb, np.zeros((0, 23)), np.zeros((0, 36)), None, "f64", "aos")
    assert out["q"].shape == (0, 23) and out["success"].shape == (0,)


@pytest.mark.parametrize("B", [1, 2, 31, 32, 33, 95, 9471, 9472, 9473, 18945])
def test_ragged_batch_sizes_around_the_scheduler_thresholds(cassie, B):
    """9472 = 2 x 32 x 148 is the largest batch the latency configuration takes in one wave on a 148-SM B200; above it
    the BULK + TAIL pair runs.  Every size must give the oracle's flags / iteration counts / q."""
    pb, om, opb = cassie
    q0, tg, _ = make_workload(pb, om, B, seed=1000 + B, standing=W.CASSIE_STANDING)
    q_ref, ok_ref, it_ref, res_ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    q, ok, it, res = _gpu(pb, q0, tg)
    assert np.array_equal(ok, ok_ref.astype(bool)) and np.array_equal(it, it_ref)
    # residuals: 1e-9 where converged; failed solves that oscillate for 100 steps amplify rounding differences (q 1e-8)
    assert np.abs(q - q_ref).max() < 1e-6 and np.abs(res - res_ref)[ok].max(initial=0) < 1e-9
    assert np.abs(res - res_ref).max() < 1e-6


def test_optional_outputs_may_be_null(cassie):
    """success / iters / resid are optional in ikb_batch_io -- also for a batch that takes the two-launch path, where
    the step counts of suspended problems then live in internal scratch."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    for B in (500, 20000):
        q0, tg, _ = make_workload(pb, om, B, seed=77, standing=W.CASSIE_STANDING)
        q_ref, ok_ref = O.dls_batch(opb, q0, tg, nthreads=NT)[:2]
        ok_ref = ok_ref.astype(bool)
        dq0 = torch.tensor(q0.T.copy(), device="cuda:0")
        dtg = torch.tensor(tg.T.copy(), device="cuda:0")
        dq = torch.empty_like(dq0)
        io = capi.BatchIO(dq0.data_ptr(), B, 1, dtg.data_ptr(), B, 1, dq.data_ptr(), B, 1, None, None, None)
        prm = ik.dls_parameters().c()
        capi.check(capi.lib.ikb_dls_solve_batch(pb._h, capi.F64, C.byref(prm), B, C.byref(io), None), "solve")
        torch.cuda.synchronize()
        err = np.abs(dq.cpu().numpy().T - q_ref).max(axis=1)
        assert err[ok_ref].max() < 1e-6
        # a problem that never converges ends after 100 steps of a non-contracting map: rounding differences between any two
        # FP64 implementations are amplified along the way (one of 508 such problems ends 1.1e-6 away, the rest < 1e-9)
        assert np.percentile(err[~ok_ref], 99) < 1e-6 and err[~ok_ref].max() < 1e-4


def test_broadcast_q0_and_aos_views(cassie):
    """batch_stride = 0 broadcasts one initial guess; AoS views need no transposition (include/ikb200.h)."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    B = 700
    q0, tg, _ = make_workload(pb, om, B, seed=5, standing=W.CASSIE_STANDING)
    q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, q0, tg, nthreads=NT)
    d_q0 = torch.tensor(q0[0].copy(), device="cuda:0")          # one vector of nq
    d_tg = torch.tensor(tg.copy(), device="cuda:0")             # AoS [B][36]
    d_q = torch.empty((B, 23), dtype=torch.float64, device="cuda:0")
    ok = torch.empty(B, dtype=torch.uint8, device="cuda:0")
    it = torch.empty(B, dtype=torch.int32, device="cuda:0")
    io = capi.BatchIO(d_q0.data_ptr(), 1, 0, d_tg.data_ptr(), 1, 36, d_q.data_ptr(), 1, 23, ok.data_ptr(), it.data_ptr(), None)
    prm = ik.dls_parameters().c()
    capi.check(capi.lib.ikb_dls_solve_batch(pb._h, capi.F64, C.byref(prm), B, C.byref(io), None), "solve")
    torch.cuda.synchronize()
    assert np.array_equal(ok.cpu().numpy().astype(bool), ok_ref.astype(bool)) and np.array_equal(it.cpu().numpy(), it_ref)
    assert np.abs(d_q.cpu().numpy() - q_ref).max() < 1e-6


def test_concurrent_solves_on_several_streams(cassie):
    """A finalized problem handle is immutable: solves on different streams may overlap (more streams than scratch
    slots, so slot reuse is exercised too)."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    B = 12000  # two-launch path
    n = 10
    sets = []
    for k in range(n):
        q0, tg, _ = make_workload(pb, om, B, seed=300 + k, standing=W.CASSIE_STANDING)
        sets.append((q0, tg, torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0")))
    streams = [torch.cuda.Stream() for _ in range(n)]
    torch.cuda.synchronize()
    outs = []
    for k in range(n):
        with torch.cuda.stream(streams[k]):
            outs.append(ik.dls_batch(pb, sets[k][2], sets[k][3]))  # enqueues on the current (k-th) stream
    torch.cuda.synchronize()
    for k in range(n):
        q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, sets[k][0], sets[k][1], nthreads=NT)
        assert np.array_equal(outs[k]["success"].cpu().numpy().astype(bool), ok_ref.astype(bool))
        assert np.array_equal(outs[k]["iters"].cpu().numpy(), it_ref)
        assert np.abs(outs[k]["q"].cpu().numpy().T - q_ref).max() < 1e-6


def test_tight_iteration_budgets(cassie):
    """max_iterations of 1, 2 and 17 (just above the BULK step cap) against the oracle."""
    pb, om, opb = cassie
    B = 15000
    q0, tg, _ = make_workload(pb, om, B, seed=9, standing=W.CASSIE_STANDING)
    for mi in (1, 2, 17):
        q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, q0, tg, O.params(mi), nthreads=NT)
        q, ok, it, _ = _gpu(pb, q0, tg, ik.dls_parameters(max_iterations=mi))
        assert np.array_equal(ok, ok_ref.astype(bool)) and np.array_equal(it, it_ref)
        assert np.abs(q - q_ref).max() < 1e-6


def test_table_driven_kernel_two_launch_schedule(monkeypatch):
    """The thread-per-problem table-driven kernel (dls_generic.cuh; since r2 only the fallback for problems beyond the
    team-per-problem kernel's table capacities, pinned here with IKB_GENERIC_LEGACY=1): above 2 048 problems it parks
    problems unfinished after 32 steps and continues them in a second launch.  Same answers as the single launch (bit for bit) and as the oracle -- with the caller's
    `iters` buffer and without it (internal scratch), for ik::dls and ik::pik."""
    torch = _torch()
    monkeypatch.setenv("IKB_GENERIC_LEGACY", "1")
    m = W.cassie_model()
    pb = ik.InverseKinematicsProblem(m, 1)
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
    pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Orientation), 1)
    assert pb.specialisation() is None
    pb.finalize(0)
    assert pb.kernel_name().startswith("generic<")
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 3000
    q0, tg, _ = make_workload(pb, om, B, seed=4242, standing=W.CASSIE_STANDING)
    dq0, dtg = torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0")
    monkeypatch.setenv("IKB_GENERIC_CAP", "0")
    one = ik.dls_batch(pb, dq0, dtg)
    one_p = ik.pik_batch(pb, dq0, dtg, ik.pik_parameters(lambdas=[1e-2, 1e-1]))
    torch.cuda.synchronize()
    launches = ik.kernel_launch_count()
    monkeypatch.delenv("IKB_GENERIC_CAP")
    two = ik.dls_batch(pb, dq0, dtg)
    torch.cuda.synchronize()
    assert ik.kernel_launch_count() - launches == 2
    two_p = ik.pik_batch(pb, dq0, dtg, ik.pik_parameters(lambdas=[1e-2, 1e-1]))
    torch.cuda.synchronize()
    for a, b in ((one, two), (one_p, two_p)):
        for k in ("q", "success", "iters", "resid"):
            assert torch.equal(a[k], b[k]), k
    assert (one["iters"] > 32).sum().item() > 0  # some problems did change launches
    # optional outputs absent: the step counts of parked problems live in internal scratch
    dq = torch.empty_like(dq0)
    io = capi.BatchIO(dq0.data_ptr(), B, 1, dtg.data_ptr(), B, 1, dq.data_ptr(), B, 1, None, None, None)
    prm = ik.dls_parameters().c()
    capi.check(capi.lib.ikb_dls_solve_batch(pb._h, capi.F64, C.byref(prm), B, C.byref(io), None), "solve")
    torch.cuda.synchronize()
    assert torch.equal(dq, one["q"])
    q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, q0, tg, nthreads=NT)
    assert np.array_equal(two["success"].cpu().numpy().astype(bool), ok_ref.astype(bool))
    assert np.array_equal(two["iters"].cpu().numpy(), it_ref)
    err = np.abs(two["q"].cpu().numpy().T - q_ref).max(axis=1)
    assert err[ok_ref.astype(bool)].max() < 1e-6 and np.percentile(err, 99) < 1e-6
