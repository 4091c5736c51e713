"""GPU tests of the round-2 boundary work: compact target wire format, sliced (gappy) host views, the multi-GPU handle,
dls_data::{dq, e, J}, an overridden stop test, the near-miss note, ticket-slot reuse across streams."""
import ctypes as C
import os

import numpy as np
import pytest

import ik_b200 as ik
from ik_b200 import _capi as capi
from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import make_workload, oracle_model, oracle_problem_like, urdf_text

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


@pytest.mark.parametrize("layout", ["aos", "soa"])
@pytest.mark.parametrize("B", [300, 20000])
def test_compact_targets_equal_se3_targets(cassie, layout, B):
    """IKB_TARGETS_COMPACT (quaternion + translation / translation only) through the host path = the SE3 records: 13 instead
    of 36 scalars per Cassie problem cross the link; the device expands them before the solve."""
    pb, om, opb = cassie
    q0, tg, _ = make_workload(pb, om, B, seed=77, standing=W.CASSIE_STANDING)
    assert pb.compact_target_size == 13 and pb.target_size == 36
    ctg = pb.compact_targets(tg)
    a = (lambda x: x) if layout == "aos" else (lambda x: np.ascontiguousarray(x.T))
    full = ik.dls_batch_host(pb, a(q0), a(tg), None, "f64", layout)
    os.environ["IKB_HOST_SLICES"] = "2"          # the sliced copy-in (B = 20 000) expands every slice on its compute stream
    try:
        comp = ik.dls_batch_host(pb, a(q0), a(ctg), None, "f64", layout, compact=True)
    finally:
        del os.environ["IKB_HOST_SLICES"]
    assert np.array_equal(full["success"], comp["success"]) and np.array_equal(full["iters"], comp["iters"])
    # R -> quaternion -> R costs an ulp or two of the target; the 100-step stragglers amplify that to ~4e-9
    err = np.abs(full["q"] - comp["q"]).max(axis=1 if layout == "aos" else 0)
    assert err.max() < 1e-6 and np.percentile(err, 99) < 1e-11
    # + one shared initial guess, only q and success read back: 289 B per solve instead of 669
    lean = ik.dls_batch_host(pb, q0[0].copy(), a(ctg), None, "f64", layout, compact=True, outputs=("q", "success"))
    assert lean["iters"] is None and lean["resid"] is None
    assert np.array_equal(lean["success"], full["success"]) and np.abs(lean["q"] - full["q"]).max() < 1e-6
    # the device-pointer entry points take SE3 records only
    torch = _torch()
    d = torch.zeros((36, 4), dtype=torch.float64, device="cuda:0")
    io = capi.BatchIO(d.data_ptr(), 4, 1, d.data_ptr(), 4, 1, d.data_ptr(), 4, 1, None, None, None, capi.TARGETS_COMPACT, 0)
    prm = ik.dls_parameters().c()
    assert capi.lib.ikb_dls_solve_batch(pb._h, capi.F64, C.byref(prm), 4, C.byref(io), None) == capi.ERR_INVALID_ARG


def test_sliced_host_views_copy_only_the_payload(cassie):
    """a host view with gaps (a column slice of a wider SoA array, records inside wider AoS rows) used to be
    copied as one flat extent -- the gaps of the OUTPUT array were overwritten.  Now 2-D copies move the payload only."""
    pb, om, opb = cassie
    pb.finalize(0)
    B, W0 = 700, 1000
    q0, tg, _ = make_workload(pb, om, W0, seed=5, standing=W.CASSIE_STANDING)
    ref = ik.dls_batch_host(pb, q0[100:100 + B], tg[100:100 + B], None, "f64", "aos")
    prm = ik.dls_parameters().c()
    # SoA: arrays [k][W0], the batch is columns 100 .. 100 + B
    q0s, tgs = np.ascontiguousarray(q0.T), np.ascontiguousarray(tg.T)
    qs = np.full((23, W0), -7.0)
    ok = np.zeros(B, dtype=np.uint8)
    io = capi.BatchIO(q0s[:, 100:].ctypes.data, W0, 1, tgs[:, 100:].ctypes.data, W0, 1, qs[:, 100:].ctypes.data, W0, 1, ok.ctypes.data, None, None, 0, 0)
    capi.check(capi.lib.ikb_dls_solve_batch_host(pb._h, capi.F64, C.byref(prm), B, C.byref(io)), "host solve, SoA slice")
    assert np.array_equal(qs[:, 100:100 + B].T, ref["q"]) and np.array_equal(ok, ref["success"])
    assert (qs[:, :100] == -7.0).all() and (qs[:, 100 + B:] == -7.0).all()       # the gaps are untouched
    # AoS: records of 23 / 36 scalars inside rows of 40 / 50
    q0a, tga, qa = np.zeros((B, 40)), np.zeros((B, 50)), np.full((B, 40), -7.0)
    q0a[:, 5:28], tga[:, 3:39] = q0[100:100 + B], tg[100:100 + B]
    io = capi.BatchIO(q0a[:, 5:].ctypes.data, 1, 40, tga[:, 3:].ctypes.data, 1, 50, qa[:, 5:].ctypes.data, 1, 40, ok.ctypes.data, None, None, 0, 0)
    capi.check(capi.lib.ikb_dls_solve_batch_host(pb._h, capi.F64, C.byref(prm), B, C.byref(io)), "host solve, AoS slice")
    assert np.array_equal(qa[:, 5:28], ref["q"]) and (qa[:, :5] == -7.0).all() and (qa[:, 28:] == -7.0).all()
    # anything else (negative strides, transposed-with-gaps) is refused, not mis-copied
    io = capi.BatchIO(q0a.ctypes.data, 2, 40, tga.ctypes.data, 1, 50, qa.ctypes.data, 1, 40, None, None, None, 0, 0)
    assert capi.lib.ikb_dls_solve_batch_host(pb._h, capi.F64, C.byref(prm), B, C.byref(io)) == capi.ERR_INVALID_ARG
    io = capi.BatchIO(q0a.ctypes.data, 1, 40, tga.ctypes.data, 1, 50, qa.ctypes.data, 1, -40, None, None, None, 0, 0)
    assert capi.lib.ikb_dls_solve_batch_host(pb._h, capi.F64, C.byref(prm), B, C.byref(io)) == capi.ERR_INVALID_ARG


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_multi_gpu_handle_shards_a_host_batch(cassie, layout):
    """ikb_multi_*: contiguous slices, one per listed device (here the visible GPUs, and device 0 listed three times so
    the slicing logic runs on a one-GPU box too): identical to the single-device host call."""
    torch = _torch()
    pb, om, opb = cassie
    B = 30001          # not a multiple of the slice count; each slice large enough for the two-launch path
    q0, tg, _ = make_workload(pb, om, B, seed=11, standing=W.CASSIE_STANDING)
    a = (lambda x: x) if layout == "aos" else (lambda x: np.ascontiguousarray(x.T))
    ref = ik.dls_batch_host(pb, a(q0), a(tg), None, "f64", layout)
    n = torch.cuda.device_count()
    for devices in ([0, 0, 0], list(range(n)) if n > 1 else [0]):
        multi = ik.MultiGPU(pb, devices=devices, depth=2, merge=1)
        out = multi.dls_batch_host(a(q0), a(tg), None, "f64", layout)
        for k in ("q", "success", "iters", "resid"):
            assert np.array_equal(out[k], ref[k]), (devices, k)
        # pipelined: three batches in flight, compact targets, shared q0
        ctg = a(pb.compact_targets(tg))
        tickets = [multi.submit_host(q0[0].copy(), ctg, None, "f64", layout, compact=True, outputs=("q", "success")) for _ in range(3)]
        for t, o in tickets:
            multi.wait(t)
            assert np.array_equal(o["success"], ref["success"]) and np.abs(o["q"] - ref["q"]).max() < 1e-6
        multi.drain()
        del multi


def test_same_problem_on_two_devices_in_one_process():
    """the dynamic shared-memory opt-in of the kernels is per device; a second handle finalized on another GPU
    of the same process must launch the BULK / TAIL / team kernels too."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    om = oracle_model("cassie")
    outs = []
    for dev in (0, 1):
        pb = W.cassie_feet_pelvis_problem()
        pb.finalize(dev)
        q0, tg, _ = make_workload(pb, om, 20000, seed=3, standing=W.CASSIE_STANDING)
        d = torch.device("cuda", dev)
        o = ik.dls_batch(pb, torch.tensor(q0.T.copy(), device=d), torch.tensor(tg.T.copy(), device=d))
        small = ik.dls_batch(pb, torch.tensor(q0[:500].T.copy(), device=d), torch.tensor(tg[:500].T.copy(), device=d))
        torch.cuda.synchronize(d)
        outs.append((o["q"].cpu(), o["success"].cpu(), small["q"].cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_dls_data_exports_dq_e_J(cassie):
    """data.hpp:15-28 / dls.hpp:34-65: after ik::dls the caller's dls_data holds dq, e and J of the last evaluation."""
    pb, om, opb = cassie
    q0 = W.standing_configuration(pb.model(), W.CASSIE_STANDING)
    lf = om.frame_placement(q0, om.frame_id("LeftFootFront"))[9:]
    rf = om.frame_placement(q0, om.frame_id("RightFootFront"))[9:]
    pb.get_frame_task("fl").target[9:] = lf + np.array([0.05, 0.0, 0.10])
    pb.get_frame_task("fr").target[9:] = rf
    data = ik.dls_data(pb)
    q = ik.dls(pb, q0, data)
    q_ref, ok, it, res, dq_ref = O.dls(opb, q0, pb.gather_targets())
    assert data.success and ok and data.iterations == it == 1 and np.abs(q - q_ref).max() < 1e-11
    e_ref, J_ref = opb.evaluate(q_ref, pb.gather_targets())[:2]     # success: the returned iterate is the evaluated one
    assert np.abs(data.e - e_ref).max() < 1e-12 and np.abs(data.J - np.asarray(J_ref).reshape(12, 22)).max() < 1e-11
    assert np.abs(data.dq - dq_ref).max() < 1e-9 and abs(data.residual - float(e_ref @ e_ref)) < 1e-15
    # a failed solve: e / J / dq belong to the last EVALUATED iterate, q is one step further (dls.cpp:67-77)
    data2 = ik.dls_data(pb)
    q2 = ik.dls(pb, q0, data2, p=ik.dls_parameters(max_iterations=1, tolerance=0.0))
    assert not data2.success and data2.iterations == 1
    e0, J0 = opb.evaluate(q0, pb.gather_targets())[:2]
    assert np.abs(data2.e - e0).max() < 1e-12 and np.abs(data2.J - np.asarray(J0).reshape(12, 22)).max() < 1e-11
    assert np.abs(q2 - om.clip(om.integrate(q0, data2.dq))).max() < 1e-12


def test_overridden_stop_test_is_honoured(cassie):
    """visitor.hpp:15-21: a visitor class may override should_stop -- the loop then runs on the host, one device iteration
    per step, and the override sees e / dq exactly where the reference calls it (dls.cpp:61)."""
    pb, om, opb = cassie
    q0 = W.standing_configuration(pb.model(), W.CASSIE_STANDING)
    lf = om.frame_placement(q0, om.frame_id("LeftFootFront"))[9:]
    pb.get_frame_task("fl").target[9:] = lf + np.array([0.05, 0.0, 0.10])
    pb.get_frame_task("fr").target[9:] = om.frame_placement(q0, om.frame_id("RightFootFront"))[9:]
    prm = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1)

    class SameTest(ik.inverse_kinematics_visitor):
        calls = 0

        def should_stop(self, problem, e, dq):
            SameTest.calls += 1
            return float(e[0] @ e[0]) < 1e-4

    class StepNorm(ik.inverse_kinematics_visitor):      # a different criterion: stop when the step is small
        def should_stop(self, problem, e, dq):
            return float(np.abs(dq).max()) < 5e-2

    stock, over, other = ik.dls_data(pb), ik.dls_data(pb), ik.dls_data(pb)
    q_stock = ik.dls(pb, q0, stock, p=prm)
    q_over = ik.dls(pb, q0, over, SameTest(), prm)
    assert stock.success and over.success and stock.iterations == over.iterations == 27 and SameTest.calls == 28
    assert np.abs(q_stock - q_over).max() < 1e-12
    q_other = ik.dls(pb, q0, other, StepNorm(), prm)
    assert other.success and other.iterations != 27 and float(np.abs(other.dq).max()) < 5e-2
    assert np.abs(q_other - q_stock).max() > 1e-6


def test_near_miss_of_a_specialisation_is_reported():
    """Cassie with ONE re-rounded URDF literal no longer matches the generated code bit for bit and runs
    on the table-driven kernel -- ikb_problem_status_string says so (and why); results still match the oracle."""
    xml = urdf_text("cassie").replace("1.57079632679", "1.5707963268", 1)
    assert xml != urdf_text("cassie")
    m = ik.Model.from_urdf(xml, free_flyer=True)
    pb = W.cassie_feet_pelvis_problem(m)
    assert pb.specialisation() is None
    pb.finalize(0)
    note = pb.status_string()
    assert "cassie_feet_pelvis" in note and "table-driven" in note and pb.kernel_name().startswith("coop<"), note
    exact = W.cassie_feet_pelvis_problem()
    exact.finalize(0)
    assert exact.status_string() == "" and exact.kernel_name() == "cassie_feet_pelvis"
    om = O.Model.from_urdf(xml, True)
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, 500, seed=8, standing=W.CASSIE_STANDING)
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    out = ik.dls_batch_host(pb, q0, tg, None, "f64", "aos")
    assert np.array_equal(out["success"].astype(bool), ref[1]) and np.array_equal(out["iters"], ref[2])
    assert np.abs(out["q"] - ref[0]).max() < 1e-6


def test_ticket_slots_survive_many_small_solves_beside_a_long_one(cassie):
    """the work counters come from a ring of 64 slots; a long kernel on stream X used to share its counter with
    the 64th small solve issued meanwhile on stream Y.  Each slot now carries an event of its last user."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    Bbig, Bsmall, n_small = 65536, 64, 200
    q0, tg, _ = make_workload(pb, om, 8192, seed=21, standing=W.CASSIE_STANDING)
    reps = Bbig // 8192
    q0b, tgb = np.tile(q0, (reps, 1)), np.tile(tg, (reps, 1))
    prm = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1)   # a long kernel pair
    oprm = O.params(200, 0.1, 0.1)
    ref = O.dls_batch(opb, q0, tg, oprm, nthreads=NT)
    d_q0, d_tg = torch.tensor(q0b.T.copy(), device="cuda:0"), torch.tensor(tgb.T.copy(), device="cuda:0")
    s_q0, s_tg = torch.tensor(q0[:Bsmall].T.copy(), device="cuda:0"), torch.tensor(tg[:Bsmall].T.copy(), device="cuda:0")
    sx, sy = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(sx):
        big = ik.dls_batch(pb, d_q0, d_tg, prm)
    smalls = []
    with torch.cuda.stream(sy):
        for _ in range(n_small):
            smalls.append(ik.dls_batch(pb, s_q0, s_tg, prm))
    torch.cuda.synchronize()
    q = big["q"].cpu().numpy().T
    for r in range(reps):
        sl = slice(r * 8192, (r + 1) * 8192)
        assert np.array_equal(big["success"].cpu().numpy()[sl].astype(bool), ref[1]) and np.array_equal(big["iters"].cpu().numpy()[sl], ref[2])
        assert np.abs(q[sl] - ref[0]).max() < 1e-6
    for o in smalls:
        assert np.array_equal(o["iters"].cpu().numpy(), ref[2][:Bsmall]) and np.abs(o["q"].cpu().numpy().T - ref[0][:Bsmall]).max() < 1e-6


def test_plugin_specialisation_for_another_robot():
    """ik_b200.specialise: the generator + nvcc build a specialised kernel for a robot / task list the library was not
    compiled for (UR5: RY joints, Position on level 0 + weighted Orientation on level 1); after ikb_load_specialisation the
    problem leaves the table-driven kernel and still matches the oracle."""
    from ik_b200 import specialise as SP

    torch = _torch()

    def make():
        m = ik.Model.builtin("ur5", free_flyer=False)
        m.set_limits(np.maximum(m.lowerPositionLimit, -3.0), np.minimum(m.upperPositionLimit, 3.0))
        pb = ik.InverseKinematicsProblem(m, 1)
        t_pos = ik.FrameTask(m, "tool0", ik.KinematicType.Position)
        t_pos.weighting()[:] = [1.0, 0.5, 2.0]
        pb.add_frame_task("pos", t_pos, 0)
        pb.add_frame_task("ori", ik.FrameTask(m, "ee_link", ik.KinematicType.Orientation), 1)
        return m, pb

    m, pb = make()
    pb.finalize(0)
    assert pb.kernel_name().startswith("coop<")
    SP.load_plugin(SP.build_plugin(pb, "ur5_tool_plugin"))
    m2, pb2 = make()
    assert pb2.specialisation() == "ur5_tool_plugin"
    pb2.finalize(0)
    assert pb2.kernel_name() == "ur5_tool_plugin"
    om = oracle_model("ur5", free_flyer=False)
    om.flat["lower"][:] = m.lowerPositionLimit
    om.flat["upper"][:] = m.upperPositionLimit
    om = O.Model(om.flat)
    opb = oracle_problem_like(pb2, om)
    B = 30000
    q0, tg, _ = make_workload(pb2, om, B, seed=21, start="near")
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    dq0, dtg = torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0")
    outs = {}
    for name, p in (("coop", pb), ("plugin", pb2)):
        o = ik.dls_batch(p, dq0, dtg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            o = ik.dls_batch(p, dq0, dtg)
        e1.record()
        torch.cuda.synchronize()
        ok, it = o["success"].cpu().numpy().astype(bool), o["iters"].cpu().numpy()
        agree = (ok == ref[1]) & (it == ref[2])
        err = np.abs(o["q"].cpu().numpy().T - ref[0])[agree & ok].max(axis=1)
        outs[name] = e0.elapsed_time(e1) / 3
        print("ur5 %s: %.3f ms for %d problems, agree %.5f, |dq| p99.9 %.2e max %.2e" % (name, outs[name], B, agree.mean(), np.percentile(err, 99.9), err.max()))
        # a 6R arm: a few problems per 10,000 pass near a singularity and amplify rounding (as in the full-size manipulator test)
        assert agree.mean() > 0.999 and np.percentile(err, 99.9) < 1e-6 and err.max() < 1e-4
    assert outs["plugin"] < outs["coop"]


def test_weights_masks_and_limits_edited_between_solves_take_effect():
    """The reference reads task->weighting(), PostureTask::mask and the position limits at every evaluation; the Python
    mirror rebuilds its finalized handle when any of them changed since the last solve."""
    pb = W.cassie_demo_posture_problem()
    om = oracle_model("cassie")
    q0, tg, _ = make_workload(pb, om, 64, seed=57, standing=W.CASSIE_STANDING)
    first = ik.dls_batch_host(pb, q0, tg, None, "f64", "aos")
    posture = pb.get_posture_task("posture")
    posture.weighting()[:] = 0.2
    posture.mask[3] = 0.0
    m = pb.model()
    hi = m.upperPositionLimit.copy()
    hi[10] = min(hi[10], -0.9)
    m.set_limits(m.lowerPositionLimit, hi)
    second = ik.dls_batch_host(pb, q0, tg, None, "f64", "aos")
    om.flat["upper"][:] = hi
    om2 = O.Model(om.flat)
    ref = O.dls_batch(oracle_problem_like(pb, om2), q0, tg, nthreads=NT)
    assert not np.array_equal(first["q"], second["q"])
    assert np.array_equal(second["success"].astype(bool), ref[1]) and np.array_equal(second["iters"], ref[2])
    assert np.abs(second["q"] - ref[0])[ref[1]].max() < 1e-6 and (second["q"][:, 10] <= -0.9 + 1e-15)[second["iters"] > 0].all()
