"""GPU tests of the pipelined queue (ikb_queue_*): merged launches must give the results of the per-batch calls."""
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


@pytest.fixture(scope="module")
def cassie():
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    return pb, om, oracle_problem_like(pb, om)


def _f32_routes_agree(torch, out, r):
    """Two FP32 routes through the kernels (merged / carried / per-batch) against each other.  The step at which a straggler
    changes kernels (BULK arithmetic -> team arithmetic of the TAIL launch) depends on scheduling (ikb200.h), and a carried
    launch hands over after ONE step where the per-batch call hands over after 16: the routes are two single-precision
    evaluations of the same iteration.  Their distance is bounded like the FP32 oracle's from the FP64 oracle
    (tests/test_oracle_f32.py: median 2e-6, p99 6e-5, p99.9 4e-4, max 4e-2 on the ill-conditioned 0.6 % of this workload),
    so it is asserted by quantiles, not by the maximum over a handful of problems that differs from run to run."""
    same = (out["success"] == r["success"]) & (out["iters"] == r["iters"])
    assert same.float().mean().item() > 0.97
    d = (out["q"] - r["q"]).abs().amax(dim=0)[r["success"].bool() & same].double()
    if d.numel() == 0:
        return
    q99, q999, dmax = torch.quantile(d, 0.99).item(), torch.quantile(d, 0.999).item(), d.max().item()
    print("f32 routes: same flags+steps %.5f |dq| p99 %.2e p99.9 %.2e max %.2e (n=%d)" % (same.float().mean().item(), q99, q999, dmax, d.numel()))
    assert q99 < 3e-4 and q999 < 3e-3 and dmax < 0.2, (q99, q999, dmax)   # measured over many runs: p99 4e-7 ... 2e-5, p99.9 8e-7 ... 2e-4, max 7e-6 ... 3e-3


def _dev(torch, a, dtype=None):
    return torch.tensor(np.ascontiguousarray(a.T), device="cuda:0", dtype=dtype)


def test_merged_large_batches_are_bit_identical_f64(cassie):
    """Four two-launch batches of different sizes in one kernel pair: FP64 BULK and TAIL share their arithmetic, so the
    merged launch reproduces ik.dls_batch bit for bit (and therefore the oracle to the parity bar)."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    sizes = [12000, 20000, 9600, 15000]
    data = [make_workload(pb, om, B, seed=100 + i, standing=W.CASSIE_STANDING) for i, B in enumerate(sizes)]
    dev = [(_dev(torch, q0), _dev(torch, tg)) for q0, tg, _ in data]
    ref = [ik.dls_batch(pb, a, b) for a, b in dev]
    torch.cuda.synchronize()
    queue = ik.SolveQueue(pb, depth=8, merge=4)
    got = [queue.submit(a, b) for a, b in dev]
    for t, _ in got:
        queue.wait(t)
    for (t, out), r in zip(got, ref):
        for k in ("q", "success", "iters", "resid"):
            assert torch.equal(out[k], r[k]), k
    # ... and the oracle, on the first batch
    q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, data[0][0], data[0][1], nthreads=NT)
    ok = got[0][1]["success"].cpu().numpy().astype(bool)
    assert np.array_equal(ok, ok_ref.astype(bool)) and np.array_equal(got[0][1]["iters"].cpu().numpy(), it_ref)
    assert np.abs(got[0][1]["q"].cpu().numpy().T - q_ref)[ok].max() < 1e-6


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_merged_small_batches_match_per_batch_calls(cassie, dtype):
    """Small batches: the merged group may take other kernels than a lone batch (team-per-problem vs thread-per-problem),
    so flags and step counts are equal and q agrees to rounding (FP64) / to the FP32 bar."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    sizes = [700, 33, 4096, 1]
    data = [make_workload(pb, om, B, seed=200 + i, standing=W.CASSIE_STANDING) for i, B in enumerate(sizes)]
    dev = [(_dev(torch, q0, tdt), _dev(torch, tg, tdt)) for q0, tg, _ in data]
    ref = [ik.dls_batch(pb, a, b) for a, b in dev]
    torch.cuda.synchronize()
    queue = ik.SolveQueue(pb, depth=4, merge=4)
    got = [queue.submit(a, b) for a, b in dev]
    queue.drain()
    for (t, out), r in zip(got, ref):
        ok = r["success"].bool()
        if dtype == "f64":
            assert torch.equal(out["success"], r["success"]) and torch.equal(out["iters"], r["iters"])
            assert (out["q"] - r["q"]).abs()[:, ok].max() < 1e-9
        else:
            same = (out["success"] == r["success"]) & (out["iters"] == r["iters"])
            assert same.float().mean() > 0.97
            assert (out["q"] - r["q"]).abs()[:, ok & same].max() < 3e-3


def test_host_mode_and_parameter_change(cassie):
    """Host buffers (SoA and AoS), a parameter change in the middle of the stream (starts a new group), waiting for
    tickets in any order, re-using a slot."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    B = 11000
    q0, tg, _ = make_workload(pb, om, B, seed=300, standing=W.CASSIE_STANDING)
    demo = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1)
    ref_default = ik.dls_batch_host(pb, q0, tg, None, "f64", "aos")
    ref_demo = ik.dls_batch_host(pb, q0, tg, demo, "f64", "aos")
    queue = ik.SolveQueue(pb, depth=3, merge=3)
    soa = (np.ascontiguousarray(q0.T), np.ascontiguousarray(tg.T))
    jobs = []
    for k in range(7):  # more batches than slots: submit blocks on the slot's previous occupant
        prm = demo if k in (2, 3) else None
        if k % 2:
            jobs.append((queue.submit_host(soa[0], soa[1], prm, "f64", "soa"), "soa", prm))
        else:
            jobs.append((queue.submit_host(q0, tg, prm, "f64", "aos"), "aos", prm))
    for (t, out), layout, prm in reversed(jobs):
        queue.wait(t)
        r = ref_demo if prm is not None else ref_default
        q = out["q"].T if layout == "soa" else out["q"]
        assert np.array_equal(q, r["q"]) and np.array_equal(out["success"], r["success"])
        assert np.array_equal(out["iters"], r["iters"]) and np.array_equal(out["resid"], r["resid"])


def test_host_mode_broadcast_initial_guess(cassie):
    """One initial guess for the whole batch (q0 of shape (nq,), batch_stride = 0): 23 values cross the host link
    instead of B x 23; results equal those of the tiled q0, SoA and AoS, merged and alone."""
    _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    B = 10000
    q0, tg, _ = make_workload(pb, om, B, seed=301, standing=W.CASSIE_STANDING)
    assert (q0 == q0[0]).all()
    ref = ik.dls_batch_host(pb, q0, tg, None, "f64", "aos")
    queue = ik.SolveQueue(pb, depth=4, merge=2)
    one = np.ascontiguousarray(q0[0])
    jobs = [queue.submit_host(one, np.ascontiguousarray(tg.T), None, "f64", "soa"),
            queue.submit_host(one, tg, None, "f64", "aos"),
            queue.submit_host(one, tg, None, "f64", "aos")]
    for k, (t, out) in enumerate(jobs):
        queue.wait(t)
        q = out["q"].T if k == 0 else out["q"]
        assert np.array_equal(q, ref["q"]) and np.array_equal(out["success"], ref["success"])
        assert np.array_equal(out["iters"], ref["iters"])


def test_queue_with_generic_kernel_and_empty_batch():
    """A problem without a specialised kernel goes through the queue batch by batch (no merged launch); B = 0 is legal."""
    torch = _torch()
    m = ik.Model.builtin("ur5", free_flyer=False)
    pb = ik.InverseKinematicsProblem(m)
    pb.add_frame_task("tool", ik.FrameTask(m, "tool0", ik.KinematicType.Full))
    pb.finalize(0)
    assert pb.specialisation() is None
    om = oracle_model("ur5", free_flyer=False)
    q0, tg, _ = make_workload(pb, om, 300, seed=3, start="near")
    a, b = _dev(torch, q0), _dev(torch, tg)
    ref = ik.dls_batch(pb, a, b)
    torch.cuda.synchronize()
    queue = ik.SolveQueue(pb, depth=4, merge=2)
    t0, o0 = queue.submit(a, b)
    t1, o1 = queue.submit(a[:, :0].contiguous(), b[:, :0].contiguous())
    t2, o2 = queue.submit(a, b)
    queue.drain()
    for o in (o0, o2):
        for k in ("q", "success", "iters", "resid"):
            assert torch.equal(o[k], ref[k])
    assert o1["q"].shape[1] == 0


def test_merged_large_batches_f32_and_stream_wait(cassie):
    """FP32 two-launch batches merged (BULK thread-per-problem + TAIL team-per-problem, both with the segment table):
    the step at which a straggler changes arithmetic depends on scheduling, so the comparison with the per-batch calls is
    to the FP32 bar, not bit for bit.  Also: depth = merge = 1 degenerates to the plain call; wait_on_stream chains a
    consumer stream instead of blocking the host."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    sizes = [11000, 16000, 12000]
    data = [make_workload(pb, om, B, seed=400 + i, standing=W.CASSIE_STANDING) for i, B in enumerate(sizes)]
    dev = [(_dev(torch, q0, torch.float32), _dev(torch, tg, torch.float32)) for q0, tg, _ in data]
    ref = [ik.dls_batch(pb, a, b) for a, b in dev]
    torch.cuda.synchronize()
    queue = ik.SolveQueue(pb, depth=4, merge=3)
    got = [queue.submit(a, b) for a, b in dev]
    consumer = torch.cuda.Stream()
    sums = []
    with torch.cuda.stream(consumer):
        for t, out in got:
            queue.wait_on_stream(t, consumer.cuda_stream)
            sums.append(out["success"].sum())          # enqueued on the consumer stream, after the batch
    consumer.synchronize()
    for (t, out), r, n in zip(got, ref, sums):
        assert int(n.item()) == int(out["success"].sum().item())
        _f32_routes_agree(torch, out, r)
    queue.drain()
    # degenerate queue: one slot, no merging -> exactly the plain FP64 call
    a64, b64 = _dev(torch, data[0][0]), _dev(torch, data[0][1])
    r64 = ik.dls_batch(pb, a64, b64)
    torch.cuda.synchronize()
    q1 = ik.SolveQueue(pb, depth=1, merge=1)
    for _ in range(3):  # the single slot is reused: submit blocks on the previous batch
        t, o = q1.submit(a64, b64)
    q1.wait(t)
    for k in ("q", "success", "iters", "resid"):
        assert torch.equal(o[k], r64[k])


def test_queue_argument_errors(cassie):
    from ik_b200 import _capi as capi

    torch = _torch()
    pb, _, _ = cassie
    pb.finalize(0)
    h = capi.C.c_void_p()
    assert capi.lib.ikb_queue_create(pb._h, 0, 1, capi.C.byref(h)) == 1      # IKB_ERR_INVALID_ARG: depth
    assert capi.lib.ikb_queue_create(pb._h, 4, 5, capi.C.byref(h)) == 1      # merge > depth
    assert capi.lib.ikb_queue_create(pb._h, 16, 9, capi.C.byref(h)) == 1     # merge > 8
    queue = ik.SolveQueue(pb, 2, 2)
    assert capi.lib.ikb_queue_wait(queue._h, 0) == 1                          # no such ticket yet
    assert capi.lib.ikb_queue_wait(queue._h, -1) == 1
    queue.drain()                                                             # draining an empty queue is fine


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_carried_stragglers_match_per_batch_calls(cassie, dtype, monkeypatch):
    """Device-buffer groups that take the two-launch path are launched WITHOUT their TAIL: the next group's BULK launch
    continues their stragglers (ikb_queue.cu, CarryState), and a wait / flush / drain / slot reuse that needs them earlier
    launches the TAIL.  Every route gives the results of the per-batch call -- bit for bit in FP64 -- and of the queue
    with IKB_QUEUE_CARRY=0."""
    torch = _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    sizes = [12000, 9800, 20000, 9600, 15000, 10000, 11000, 13000, 9900, 17000]   # each larger than one resident wave (9 472): the lone call takes BULK + TAIL too
    data = [make_workload(pb, om, B, seed=900 + i, standing=W.CASSIE_STANDING) for i, B in enumerate(sizes)]
    dev = [(_dev(torch, q0, tdt), _dev(torch, tg, tdt)) for q0, tg, _ in data]
    ref = [ik.dls_batch(pb, a, b) for a, b in dev]
    torch.cuda.synchronize()

    def check(got):
        for (t, out), r in zip(got, ref):
            if dtype == "f64":
                for k in ("q", "success", "iters", "resid"):
                    assert torch.equal(out[k], r[k]), k
            else:
                _f32_routes_agree(torch, out, r)

    # five groups of two: each group's stragglers ride in the next group's launch, the last group's get a TAIL at drain
    queue = ik.SolveQueue(pb, depth=6, merge=2)
    launches0 = ik.kernel_launch_count()
    got = [queue.submit(a, b) for a, b in dev]
    queue.drain()
    assert ik.kernel_launch_count() - launches0 == 5 + 1      # five BULK launches, ONE TAIL (the scratch is sized by the first group)
    check(got)
    # waiting for a ticket of the group that is being carried forces its TAIL; later groups go on carrying
    queue = ik.SolveQueue(pb, depth=6, merge=2)
    got = []
    for i, (a, b) in enumerate(dev):
        got.append(queue.submit(a, b))
        if i == 3:
            queue.wait(got[2][0])
    for t, _ in reversed(got):
        queue.wait(t)
    check(got)
    # a change of solver parameters cannot be carried across: results of both parameter sets are right
    demo = ik.dls_parameters(max_iterations=40, step_length=0.5, damping=1e-2)
    ref_demo = [ik.dls_batch(pb, a, b, demo) for a, b in dev[:4]]
    queue = ik.SolveQueue(pb, depth=8, merge=2)
    g1 = [queue.submit(a, b) for a, b in dev[:4]]
    g2 = [queue.submit(a, b, demo) for a, b in dev[:4]]
    queue.drain()
    check(g1)
    for (t, out), r in zip(g2, ref_demo):
        if dtype == "f64":
            for k in ("q", "success", "iters"):
                assert torch.equal(out[k], r[k]), k
    # ... and with the feature switched off
    monkeypatch.setenv("IKB_QUEUE_CARRY", "0")
    queue = ik.SolveQueue(pb, depth=6, merge=2)
    launches0 = ik.kernel_launch_count()
    got = [queue.submit(a, b) for a, b in dev]
    queue.drain()
    assert ik.kernel_launch_count() - launches0 == 10
    check(got)


def test_host_batches_carry_in_a_deep_queue(cassie):
    """A queue with three groups in flight (depth >= 3 x merge) carries the stragglers of HOST batches too: a group's results
    leave the device after the NEXT group's launch has finished them (or after the TAIL that wait / drain launch on demand).
    Whatever the route -- waiting in order, out of order, draining, compact or SE3 targets -- the host arrays hold the results
    of the blocking per-batch call, bit for bit in FP64."""
    _torch()
    pb, om, opb = cassie
    pb.finalize(0)
    sizes = [12000, 9800, 15000, 10000, 11000, 9900, 13000]   # each larger than one resident wave: two-launch batches
    data = [make_workload(pb, om, B, seed=700 + i, standing=W.CASSIE_STANDING) for i, B in enumerate(sizes)]
    ref = [ik.dls_batch_host(pb, q0, tg, None, "f64", "aos") for q0, tg, _ in data]

    def same(out, r, layout):
        q = out["q"].T if layout == "soa" else out["q"]
        assert np.array_equal(q, r["q"]) and np.array_equal(out["success"], r["success"])
        if "iters" in out and out["iters"] is not None:
            assert np.array_equal(out["iters"], r["iters"]) and np.array_equal(out["resid"], r["resid"])

    launches0 = ik.kernel_launch_count()
    queue = ik.SolveQueue(pb, depth=6, merge=2)
    jobs = []
    for k, (q0, tg, _) in enumerate(data):
        if k % 2:
            jobs.append((queue.submit_host(np.ascontiguousarray(q0.T), np.ascontiguousarray(tg.T), None, "f64", "soa"), "soa"))
        else:
            jobs.append((queue.submit_host(q0, tg, None, "f64", "aos"), "aos"))
    queue.drain()
    # 7 batches = 3 full groups + 1 open batch: 3 carried BULK launches; the odd batch takes the per-batch pair, in front of which
    # the carried stragglers get their TAIL (no TAIL per group: 3 + 1 + 2 launches, not 3 x 2 + 2)
    assert ik.kernel_launch_count() - launches0 <= 3 + 1 + 2
    for ((t, out), layout), r in zip(jobs, ref):
        same(out, r, layout)
    # compact targets, one shared initial guess, waits out of order (a wait on a carried group launches its TAIL)
    queue = ik.SolveQueue(pb, depth=6, merge=2)
    assert all((q0 == q0[0]).all() for q0, _, _ in data)
    one = np.ascontiguousarray(data[0][0][0])
    ctg = [pb.compact_targets(tg) for _, tg, _ in data[:6]]
    ref_c = [ik.dls_batch_host(pb, one, c, None, "f64", "aos", compact=True, outputs=("q", "success")) for c in ctg]
    jobs = [queue.submit_host(one, c, None, "f64", "aos", compact=True, outputs=("q", "success")) for c in ctg]
    for i in (1, 0, 5, 3, 2, 4):
        queue.wait(jobs[i][0])
        same(jobs[i][1], ref_c[i], "aos")
    # a shallow queue (depth < 3 x merge) launches a TAIL per host group, same results
    queue = ik.SolveQueue(pb, depth=4, merge=2)
    jobs = [queue.submit_host(q0, tg, None, "f64", "aos") for q0, tg, _ in data[:4]]
    queue.drain()
    for (t, out), r in zip(jobs, ref):
        same(out, r, "aos")
