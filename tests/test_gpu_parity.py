"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): converged/failed flags exactly equal; final joint vectors within 1e-6 rad (FP64)
/ 1e-4 rad (FP32) wherever the two solvers stopped at the same iteration; residuals within the stated tolerance.
"""
import os

import numpy as np
import pytest

import ik_b200 as ik
from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import make_workload, oracle_model, oracle_problem_like

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _solve_gpu(pb, q0, tg, params=None, dtype="f64"):
    torch = _torch()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    dev = torch.device("cuda:0")
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), dtype=tdt, device=dev),
                       torch.tensor(tg.T.copy(), dtype=tdt, device=dev), params)
    torch.cuda.synchronize()
    return (out["q"].cpu().numpy().T.astype(np.float64), out["success"].cpu().numpy().astype(bool),
            out["iters"].cpu().numpy(), out["resid"].cpu().numpy().astype(np.float64))


PARITY_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_r2.txt")


def _log(line):
    """Measured agreement of every parity case: printed (-s) and appended to gpurun_out/parity_r2.txt, which is copied to
    profiles/ after a GPU run (keep the measured same-iteration numbers)."""
    print(line)
    try:
        os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
        with open(PARITY_LOG, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def _compare(name, gpu, ref, qtol, expect_agree=1.0, flag_mismatch_allowed=0, converged_only=False, q_frac=1.0):
    """`agree` = problems whose converged flag AND iteration count equal the oracle's.  expect_agree = 1.0 asserts
    agree.all() (the FP64 bar: flags and iteration counts identical); where less than 100 % is real (chaotic failed
    trajectories of serial arms) the MEASURED fraction is pinned to +-0.5 %.
    converged_only: compare q on converged problems only -- a FAILED solve of a 6R arm bounces between joint limits
    for 100 iterations (chaotic: two FP64 implementations end at different limits), unlike Cassie's failures, which
    stagnate at a fixed point and are compared too."""
    q, ok, it, res = gpu
    q_ref, ok_ref, it_ref, res_ref = ref
    ok_ref = ok_ref.astype(bool)
    B = len(ok)
    flag_mismatch = int((ok != ok_ref).sum())
    agree = (it == it_ref) & (ok == ok_ref)
    same = agree & ok if converged_only else agree
    qerrs = np.abs(q[same] - q_ref[same]).max(axis=1) if same.any() else np.zeros(1)
    qerr = float(qerrs.max())
    within = float((qerrs < qtol).mean())
    conv = agree & ok
    rerr = float(np.abs(res[conv] - res_ref[conv]).max()) if conv.any() else 0.0
    _log("%s: B=%d converged gpu/ref=%d/%d flag mismatches=%d agree(flags+iterations)=%d (%.6f) max|q-q_ref|=%.3e "
         "(within %.0e: %.6f, p99.9 %.2e) max|resid diff|=%.3e mean iters=%.2f"
         % (name, B, ok.sum(), ok_ref.sum(), flag_mismatch, agree.sum(), agree.mean(), qerr, qtol, within,
            np.percentile(qerrs, 99.9), rerr, it_ref.mean()))
    assert flag_mismatch <= flag_mismatch_allowed
    if expect_agree is None:          # not pinned yet: floor only
        assert agree.mean() >= 0.9
    elif expect_agree >= 1.0:
        assert agree.all(), "%d of %d problems differ in flag or iteration count" % (B - agree.sum(), B)
    else:
        assert abs(agree.mean() - expect_agree) <= 0.005, "measured agreement %.4f, pinned %.4f" % (agree.mean(), expect_agree)
    if q_frac >= 1.0:
        assert qerr < qtol
    else:                              # chaotic outliers: the measured fraction within qtol is pinned (+-0.5 %)
        assert abs(within - q_frac) <= 0.005 and within >= q_frac - 1e-4, "within %.6f, pinned %.6f" % (within, q_frac)
    return qerr


def test_fk_matches_oracle():
    torch = _torch()
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    qs = W.sample_configurations(pb.model(), 512, seed=7)
    names = ["pelvis", "LeftFootFront", "RightFootFront", "LeftFootBack", "VectorNav"]
    out = ik.fk_batch(pb, torch.tensor(qs.T.copy(), device="cuda:0"), names)
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(len(names), 12, -1)
    for i, n in enumerate(names):
        ref = np.stack([om.frame_placement(q, om.frame_id(n)) for q in qs])
        assert np.abs(got[i].T - ref).max() < 1e-13, n


def test_single_solve_known_answer():
    """BASELINE config 1 / SURVEY 8c: one Cassie solve, library defaults -> success after 1 step."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    q0 = W.standing_configuration(pb.model(), W.CASSIE_STANDING)
    lf = om.frame_placement(q0, om.frame_id("LeftFootFront"))[9:]
    rf = om.frame_placement(q0, om.frame_id("RightFootFront"))[9:]
    pb.get_frame_task("fl").target[9:] = lf + np.array([0.05, 0.0, 0.10])
    pb.get_frame_task("fr").target[9:] = rf
    data = ik.dls_data(pb)
    q = ik.dls(pb, q0, data)
    assert data.success and data.iterations == 1
    assert abs(data.residual - 5.5754e-05) < 1e-8
    np.testing.assert_allclose(q[7:11], [1.034213001968e-02, -2.487453189466e-03, 4.961980803968e-01,
                                         -1.227703791318], atol=1e-9)
    opb = oracle_problem_like(pb, om)
    q_ref, ok, it, res, _ = O.dls(opb, q0, pb.gather_targets())
    assert ok and it == 1 and np.abs(q - q_ref).max() < 1e-12
    # demo parameters (cassie.cpp:107-109) -> success at iteration 27
    data2 = ik.dls_data(pb)
    ik.dls(pb, q0, data2, p=ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1))
    assert data2.success and data2.iterations == 27


def test_start_converged_returns_q0():
    """A solve that starts converged returns q0 with zero iterations (dls.cpp:52-63)."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    q0 = W.standing_configuration(pb.model(), W.CASSIE_STANDING)
    for name, frame in (("fl", "LeftFootFront"), ("fr", "RightFootFront")):
        pb.get_frame_task(name).target[9:] = om.frame_placement(q0, om.frame_id(frame))[9:]
    data = ik.dls_data(pb)
    q = ik.dls(pb, q0, data)
    assert data.success and data.iterations == 0 and np.array_equal(q, q0)


@pytest.fixture(params=["specialised", "generic"])
def kernel_path(request, monkeypatch):
    """Both solve kernels must meet the same bar: the generated topology-specialised one (the default for the
    benchmark problems) and the table-driven generic one (IKB_FORCE_GENERIC=1 is read by ikb_problem_finalize)."""
    monkeypatch.setenv("IKB_FORCE_GENERIC", "1" if request.param == "generic" else "0")
    return request.param


def _check_path(pb, kernel_path, spec_name):
    pb.finalize(0)
    name = pb.kernel_name()
    assert (name == spec_name) if kernel_path == "specialised" else name.startswith(("coop<", "generic<")), name


@pytest.mark.parametrize("B", [1, 33, 4096])
def test_cassie_f64_defaults(B, kernel_path):
    pb = W.cassie_feet_pelvis_problem()
    _check_path(pb, kernel_path, "cassie_feet_pelvis")
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, B, standing=W.CASSIE_STANDING)
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    _compare("cassie f64 defaults B=%d" % B, _solve_gpu(pb, q0, tg), ref, 1e-6)


def test_cassie_f64_demo_params(kernel_path):
    pb = W.cassie_feet_pelvis_problem()
    _check_path(pb, kernel_path, "cassie_feet_pelvis")
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, 1024, seed=99, standing=W.CASSIE_STANDING)
    prm = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1)
    ref = O.dls_batch(opb, q0, tg, O.params(200, 0.1, 0.1), nthreads=NT)
    _compare("cassie f64 demo params", _solve_gpu(pb, q0, tg, prm), ref, 1e-6)


@pytest.mark.parametrize("params", ["defaults", "demo"])
@pytest.mark.parametrize("B", [1, 700, 24000])
def test_cassie_demo_task_set(B, params, kernel_path):
    """the reference demo's own task set (cassie.cpp:43-81) -- foot Position relative to the MOVING
    pelvis frame (whose motion compute_jacobian does not differentiate, frame.hpp:169-181), pelvis Full, AlignAxisTask --
    on its specialised kernel (one launch and, for 24 000, BULK + TAIL) and on the generic one."""
    pb = W.cassie_demo_problem()
    _check_path(pb, kernel_path, "cassie_demo")
    if kernel_path == "generic" and B > 1000:
        pytest.skip("generic kernel: covered at the smaller sizes")
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, B, seed=31, standing=W.CASSIE_STANDING)
    if params == "defaults":
        prm, oprm = None, O.params()
    else:
        prm, oprm = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1), O.params(200, 0.1, 0.1)
    ref = O.dls_batch(opb, q0, tg, oprm, nthreads=NT)
    gpu = _solve_gpu(pb, q0, tg, prm)
    # Flags and step counts exactly.  q: 1e-6 rad on every problem that converges within half the iteration budget.  The
    # few that need 70-100 steps crawl along an ill-conditioned valley (the moving reference frame is not differentiated,
    # so the iteration is not a Gauss-Newton step there) and amplify rounding: the generic and the specialised kernel and
    # the oracle differ pairwise by up to 1e-5 rad on the same ~0.1 % of problems (tools/demo_diff.py) -- bar 1e-4.
    _compare("cassie demo tasks %s B=%d" % (params, B), gpu, ref, 1e-4,
             converged_only=True)
    q, ok, it, _ = gpu
    q_ref, ok_ref, it_ref, _ = ref
    fast = ok & (it_ref <= oprm.max_iterations // 2)
    assert np.abs(q[fast] - q_ref[fast]).max() < 1e-6
    assert (np.abs(q[ok] - q_ref[ok]).max(axis=1) < 1e-6).mean() > 0.998


@pytest.mark.parametrize("params", ["defaults", "demo"])
def test_cassie_demo_with_posture_task(params, kernel_path):
    """The demo's full declared task set (cassie.cpp:43-81 with the commented-out lines enabled): the three priority-0
    tasks plus a PostureTask (posture.hpp:17-86; masked, weighted) on priority level 1 -- 26 stacked rows, stop test on the
    priority-0 rows only (visitor.hpp:19).  Specialised kernel `cassie_demo_posture` and the table-driven one."""
    pb = W.cassie_demo_posture_problem()
    _check_path(pb, kernel_path, "cassie_demo_posture")
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 600 if kernel_path == "generic" else 12000    # 12 000: BULK + TAIL
    q0, tg, qstar = make_workload(pb, om, B, seed=57, standing=W.CASSIE_STANDING)
    off = pb.target_offset(pb.get_posture_task("posture"))
    assert np.array_equal(tg[:, off:off + 16], qstar[:, 7:])      # the posture the other targets were generated from
    if params == "defaults":
        prm, oprm = None, O.params()
    else:
        prm, oprm = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1), O.params(200, 0.1, 0.1)
    ref = O.dls_batch(opb, q0, tg, oprm, nthreads=NT)
    _compare("cassie demo + posture %s" % params, _solve_gpu(pb, q0, tg, prm), ref, 1e-6,
             converged_only=True)


def test_cassie_f32_defaults():
    """FP32 instantiation against the FP64 oracle.  The discrete stop decision may differ on a few problems (SURVEY 7
    'FP32 parity of the discrete stop decision'); where it agrees the bar is 1e-4 rad.  Measured: median 2e-6, 99th
    percentile 5e-5; ~0.2 % of problems exceed 1e-4 (max 7e-4 .. 1.5e-3) -- all in the foot-pitch joints, a direction
    the two foot-POSITION tasks barely observe, so FP32 rounding of FK (1e-7) is amplified by ~1/damping.  The test
    therefore states: >= 99 % within 1e-4 rad, all within 3e-3 rad, flags equal on >= 99 %, residuals within 1e-5."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 4096
    q0, tg, _ = make_workload(pb, om, B, standing=W.CASSIE_STANDING)
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    for path in ("0", "1"):
        os.environ["IKB_FORCE_GENERIC"] = path
        try:
            pb = W.cassie_feet_pelvis_problem()
            q, ok, it, res = _solve_gpu(pb, q0, tg, dtype="f32")
        finally:
            os.environ.pop("IKB_FORCE_GENERIC", None)
        q_ref, ok_ref, it_ref, res_ref = ref
        ok_ref = ok_ref.astype(bool)
        same = (it == it_ref) & ok & ok_ref
        err = np.abs(q[same] - q_ref[same]).max(axis=1)
        print("cassie f32 (%s): converged gpu/ref=%d/%d flag mismatches=%d same-iteration=%.4f |q-q_ref| median=%.2e "
              "p99=%.2e max=%.2e" % (pb.kernel_name("f32"), ok.sum(), ok_ref.sum(), (ok != ok_ref).sum(), same.mean(),
                                     np.median(err), np.percentile(err, 99), err.max()))
        assert same.mean() > 0.95
        assert np.percentile(err, 99) < 1e-4
        assert err.max() < 3e-3
        assert (ok != ok_ref).mean() < 0.01
        assert np.all(res[ok] < 1e-4)
        assert np.abs(res[same] - res_ref[same]).max() < 1e-5


def test_host_path_layouts_agree():
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    q0, tg, _ = make_workload(pb, om, 257, seed=3, standing=W.CASSIE_STANDING)
    a = ik.dls_batch_host(pb, q0, tg, layout="aos")
    s = ik.dls_batch_host(pb, q0.T, tg.T, layout="soa")
    d = _solve_gpu(pb, q0, tg)
    assert np.array_equal(a["q"], s["q"].T) and np.array_equal(a["q"], d[0])
    assert np.array_equal(a["success"], s["success"]) and np.array_equal(a["iters"], d[2])
    assert np.array_equal(a["resid"], d[3])


def test_humanoid_f64(kernel_path):
    """BASELINE config 4 (warm-started: W.near_start)."""
    pb = W.humanoid_problem()
    _check_path(pb, kernel_path, "humanoid_limbs")
    om = oracle_model("humanoid")
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, 512, seed=5, start="near")
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    _compare("humanoid f64", _solve_gpu(pb, q0, tg), ref, 1e-6)


def test_manipulator_f64(kernel_path):
    """BASELINE config 5 (warm-started: W.near_start -- the zero configuration of a serial arm is singular)."""
    pb = W.manipulator_problem()
    _check_path(pb, kernel_path, "manipulator_tool")
    om = oracle_model("manipulator", free_flyer=False)
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, 2048, seed=11, start="near")
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    _compare("manipulator f64", _solve_gpu(pb, q0, tg), ref, 1e-6)


def test_zero_iterations_returns_q0():
    """max_iterations = 0: the loop of dls.cpp:14 never runs -- q0 comes back, success = false, 0 iterations."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    q0, tg, _ = make_workload(pb, om, 64, standing=W.CASSIE_STANDING)
    q, ok, it, res = _solve_gpu(pb, q0, tg, ik.dls_parameters(max_iterations=0))
    assert np.array_equal(q, q0) and not ok.any() and not it.any()


def test_ur5_orientation_and_weights():
    """Second parser fixture (RY joints, ur5.urdf:93) with an Orientation task, a Position task and row weights."""
    m = ik.Model.builtin("ur5", free_flyer=False)
    pb = ik.InverseKinematicsProblem(m, 1)
    t_ori = ik.FrameTask(m, "ee_link", ik.KinematicType.Orientation)
    t_pos = ik.FrameTask(m, "tool0", ik.KinematicType.Position)
    t_pos.weighting()[:] = [1.0, 0.5, 2.0]
    pb.add_frame_task("ori", t_ori, 1)
    pb.add_frame_task("pos", t_pos, 0)
    om = oracle_model("ur5", free_flyer=False)
    opb = oracle_problem_like(pb, om)
    B = 1024
    lo = m.lowerPositionLimit
    hi = m.upperPositionLimit
    m.set_limits(np.maximum(lo, -3.0), np.minimum(hi, 3.0))
    om.flat["lower"][:] = np.maximum(lo, -3.0)
    om.flat["upper"][:] = np.minimum(hi, 3.0)
    om = O.Model(om.flat)
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, B, seed=21, start="near")  # warm start: far starts are chaotic for a 6R arm
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    _compare("ur5 mixed tasks", _solve_gpu(pb, q0, tg), ref, 1e-6, converged_only=True)


def test_full_size_properties():
    """BASELINE full size (65,536): size-independent properties instead of a CPU loop -- every converged problem
    satisfies ||e||^2 < 1e-4 when its residual is recomputed by GPU FK from the returned q, flags/iteration counts
    are consistent, and the first 2,048 problems equal a separate small solve (batch-size independence)."""
    torch = _torch()
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    B = 65536
    m = pb.model()
    qstar = W.sample_configurations(m, B)
    dev = torch.device("cuda:0")
    names = W.task_frames(pb)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses)
    q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1))
    q, ok, it, res = _solve_gpu(pb, q0, tg)
    assert ok.mean() > 0.97
    assert np.all(res[ok] < 1e-4) and np.all(it[~ok] == 100) and np.all(it[ok] < 100)
    # recompute the position residual of the feet at the returned q
    got = ik.fk_batch(pb, torch.tensor(q.T.copy(), device=dev), names).cpu().numpy()
    for i, n in enumerate(names):
        if n == "pelvis":
            continue
        perr = np.linalg.norm(got[12 * i + 9:12 * i + 12].T - poses[n][:, 9:12], axis=1)
        assert np.all(perr[ok] < 2e-2)
    # batch-size independence.  Same kernels (two-launch path, thread-per-problem arithmetic): bit-identical.
    q2, ok2, it2, res2 = _solve_gpu(pb, q0[:20000], tg[:20000])
    assert np.array_equal(q2, q[:20000]) and np.array_equal(ok2, ok[:20000]) and np.array_equal(it2, it[:20000])
    # A batch that fits the latency configuration runs the team-per-problem kernel (dls_team.cuh): other summation
    # orders, so equal flags / step counts and q to rounding on converged problems.
    q3, ok3, it3, res3 = _solve_gpu(pb, q0[:2048], tg[:2048])
    assert np.array_equal(ok3, ok[:2048]) and np.array_equal(it3, it[:2048])
    assert np.abs(q3 - q[:2048])[ok3].max() < 1e-9
    # lower/upper limits hold for every returned revolute joint (common.hpp:53-56)
    lo, hi = m.lowerPositionLimit, m.upperPositionLimit
    moved = it > 0
    assert np.all(q[moved][:, 7:] >= lo[7:] - 1e-15) and np.all(q[moved][:, 7:] <= hi[7:] + 1e-15)


def _full_size_case(make, B, start, dtype, chunk=65536):
    """Seeded workload of B problems built on the GPU (FK of sampled configurations), solved in one call."""
    torch = _torch()
    pb = make()
    pb.finalize(0)
    m = pb.model()
    dev = torch.device("cuda:0")
    tdt = torch.float64 if dtype == "f64" else torch.float32
    names = W.task_frames(pb)
    qstar = W.sample_configurations(m, B)
    poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + chunk].T.copy(), device=dev), names)
                         for i in range(0, B, chunk)], dim=1)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses)
    q0 = (np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)) if start == "standing" else W.near_start(m, qstar))
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), dtype=tdt, device=dev), torch.tensor(tg.T.copy(), dtype=tdt, device=dev))
    torch.cuda.synchronize()
    return pb, names, poses, q0, tg, out


@pytest.mark.parametrize("name,make,B,start,min_conv", [("humanoid", W.humanoid_problem, 262144, "near", 0.9),
                                                        ("manipulator", W.manipulator_problem, 1048576, "near", 0.99)])
def test_full_size_properties_other_configs(name, make, B, start, min_conv):
    """BASELINE configs 4 and 5 at their full batch sizes (262,144 / 1,048,576): size-independent properties -- every
    converged problem reaches its targets when the frames are recomputed by GPU FK from the returned q, failed problems
    used the whole iteration budget, joint limits hold, and a slice of the batch equals a separate smaller solve."""
    torch = _torch()
    pb, names, poses, q0, tg, out = _full_size_case(make, B, start, "f64")
    m = pb.model()
    ok = out["success"].bool()
    it = out["iters"]
    assert ok.float().mean().item() > min_conv
    assert (out["resid"][ok] < 1e-4).all() and (it[~ok] == 100).all() and (it[ok] < 100).all()
    got = torch.cat([ik.fk_batch(pb, out["q"][:, i:i + 65536].contiguous(), names) for i in range(0, B, 65536)], dim=1)
    want = torch.tensor(np.concatenate([poses[n] for n in names], axis=1).T.copy(), device=got.device)
    perr = (got - want).reshape(len(names), 12, B)[:, 9:12].norm(dim=1).max(dim=0).values  # worst frame position error
    assert (perr[ok] < 2e-2).all()                                                       # ||e||^2 < 1e-4 bounds it by 1e-2
    lo = torch.tensor(m.lowerPositionLimit, device=got.device)[:, None]
    hi = torch.tensor(m.upperPositionLimit, device=got.device)[:, None]
    moved = it > 0
    assert ((out["q"] >= lo - 1e-15) & (out["q"] <= hi + 1e-15))[:, moved].all()
    n2 = 20000
    dev = got.device
    out2 = ik.dls_batch(pb, torch.tensor(q0[:n2].T.copy(), device=dev), torch.tensor(tg[:n2].T.copy(), device=dev))
    torch.cuda.synchronize()
    assert torch.equal(out2["success"], out["success"][:n2]) and torch.equal(out2["iters"], out["iters"][:n2])
    assert (out2["q"] - out["q"][:, :n2]).abs().max().item() < 1e-9


def test_full_size_f32_cassie():
    """BASELINE config 3 in FP32 at 65,536: converged fraction, residuals and reached targets; against the FP64 solve of
    the same batch: flags equal on >= 99 %, q within 1e-4 rad on >= 99 % of the problems both converge on."""
    torch = _torch()
    pb, names, poses, q0, tg, o32 = _full_size_case(W.cassie_feet_pelvis_problem, 65536, "standing", "f32")
    dev = o32["q"].device
    o64 = ik.dls_batch(pb, torch.tensor(q0.T.copy(), device=dev), torch.tensor(tg.T.copy(), device=dev))
    torch.cuda.synchronize()
    ok32, ok64 = o32["success"].bool(), o64["success"].bool()
    assert ok32.float().mean().item() > 0.97 and (o32["resid"][ok32] < 1e-4).all()
    assert (ok32 == ok64).float().mean().item() > 0.99
    both = ok32 & ok64 & (o32["iters"] == o64["iters"])
    err = (o32["q"].double() - o64["q"]).abs().max(dim=0).values[both]
    q = torch.quantile(err, torch.tensor([0.5, 0.99, 0.999], dtype=err.dtype, device=err.device)).tolist()
    print("cassie f32 vs f64 at 65536: flags equal %.4f, same steps %.4f, |dq| median %.2e p99 %.2e p99.9 %.2e max %.2e"
          % ((ok32 == ok64).float().mean().item(), both.float().mean().item(), q[0], q[1], q[2], err.max().item()))
    assert both.float().mean().item() > 0.95
    # the foot-pitch direction is barely observed by the two foot POSITION tasks: FP32 rounding of FK is amplified by
    # ~1/damping there, so the tail of 65 536 samples reaches a few 1e-2 rad -- the bar is on quantiles
    assert (err < 1e-4).float().mean().item() > 0.99 and (err < 3e-3).float().mean().item() > 0.998 and err.max().item() < 0.2


@pytest.mark.parametrize("name,make,oname,ff,B,start,expect", [
    ("cassie", W.cassie_feet_pelvis_problem, "cassie", True, 65536, "standing", 1.0),
    ("humanoid", W.humanoid_problem, "humanoid", True, 262144, "near", None),
    ("manipulator", W.manipulator_problem, "manipulator", False, 1048576, "near", None)])
def test_full_size_oracle_parity(name, make, oname, ff, B, start, expect):
    """every BASELINE config against the ORACLE at its FULL batch size (65,536 / 262,144 /
    1,048,576), all problems: converged flags exact, iteration counts as pinned, |q - q_oracle| < 1e-6 rad wherever flag
    and iteration count agree, residuals within 1e-9.  The oracle loop is threaded over all host cores."""
    pb, names, poses, q0, tg, out = _full_size_case(make, B, start, "f64")
    om = oracle_model(oname, free_flyer=ff)
    opb = oracle_problem_like(pb, om)
    ref = O.dls_batch(opb, q0, tg, nthreads=NT)
    gpu = (out["q"].cpu().numpy().T.astype(np.float64), out["success"].cpu().numpy().astype(bool),
           out["iters"].cpu().numpy(), out["resid"].cpu().numpy().astype(np.float64))
    if name == "cassie":
        # failures stagnate at a fixed point and are compared too: every one of the 65,536 problems, exact flags and
        # iteration counts, q within 1e-6 rad (measured 6.0e-7 on the slowest stragglers, 6e-12 on 4,096)
        _compare("FULL SIZE %s f64 (%s)" % (name, pb.kernel_name()), gpu, ref, 1e-6)
        return
    # Serial chains started 0.3 rad away with full steps: a handful of problems per 100,000 follow chaotic trajectories
    # (failed solves bounce between joint limits; some converging ones pass near a singularity), on which ANY two FP64
    # evaluations drift apart.  The yardstick is the oracle itself compiled with fused multiply-adds (same algorithm,
    # other rounding): the kernel must agree with the oracle as well as the oracle's two builds agree with each other.
    ref_fma = O.dls_batch(opb, q0, tg, nthreads=NT, fma=True)

    def stats(x):
        agree = (x[2] == ref[2]) & (x[1] == ref[1].astype(bool))
        same = agree & x[1]
        err = np.abs(x[0][same] - ref[0][same]).max(axis=1)
        return dict(flags=int((x[1] != ref[1].astype(bool)).sum()), agree=float(agree.mean()), within=float((err < 1e-6).mean()),
                    p999=float(np.percentile(err, 99.9)), max=float(err.max()))

    g, o = stats(gpu), stats(ref_fma)
    for label, st in (("kernel %s" % pb.kernel_name(), g), ("oracle built with FMA", o)):
        _log("FULL SIZE %s f64, %s vs oracle: B=%d flag mismatches=%d agree(flags+iterations)=%.6f |dq|<1e-6 on %.6f of the "
             "converged agreeing problems, p99.9 %.2e, max %.2e" % (name, label, B, st["flags"], st["agree"], st["within"],
                                                                   st["p999"], st["max"]))
    assert g["flags"] <= max(3 * o["flags"], B // 20000) and g["agree"] >= min(o["agree"], 0.9999) - 1e-4
    assert g["within"] >= min(o["within"], 0.9999) - 1e-4 and g["p999"] < 1e-6


def test_f32_spread_is_inherent_to_single_precision():
    """VERDICT r1 item 1c / north_star 'within 1e-4 rad (FP32)'.  The oracle built with every scalar a float
    (libik_oracle_f32.so: number_t = float, common.hpp:13) is the yardstick: FP32-oracle vs FP64-oracle and FP32-KERNEL
    vs FP64-oracle must have the same |dq| distribution on the full 65,536 batch.  Measured on the CPU alone
    (tests/test_oracle_f32.py): median 2.1e-6, p99 6.0e-5, p99.9 3.9e-4, max 4.0e-2 rad -- 0.55 % of the problems
    exceed 1e-4 rad in ANY single-precision evaluation of this iteration (the foot-pitch direction is observed by the
    foot-position tasks only through a 4 cm lever, so FK rounding is amplified by ~1/damping^2 there)."""
    pb, names, poses, q0, tg, o32 = _full_size_case(W.cassie_feet_pelvis_problem, 65536, "standing", "f32")
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    q64, ok64, it64, _ = O.dls_batch(opb, q0, tg, nthreads=NT)
    stats = {}
    for label, (q, ok, it) in (("oracle f32", O.dls_batch_f32(opb, q0, tg, nthreads=NT)[:3]),
                               ("kernel f32", (o32["q"].cpu().numpy().T.astype(np.float64),
                                               o32["success"].cpu().numpy().astype(bool), o32["iters"].cpu().numpy()))):
        both = ok & ok64 & (it == it64)
        err = np.abs(q - q64).max(axis=1)[both]
        st = dict(flags=(ok == ok64).mean(), both=both.mean(), med=np.median(err), p99=np.percentile(err, 99),
                  p999=np.percentile(err, 99.9), max=err.max(), within=(err < 1e-4).mean())
        stats[label] = st
        _log("FP32 study, %s vs oracle f64 (B=65536): flags equal %.5f same steps %.5f |dq| median %.2e p99 %.2e p99.9 %.2e "
             "max %.2e within 1e-4: %.5f" % (label, st["flags"], st["both"], st["med"], st["p99"], st["p999"], st["max"], st["within"]))
    o, k = stats["oracle f32"], stats["kernel f32"]
    # the kernel is as close to the FP64 answer as the reference's own arithmetic in float is (within 1.5x on every
    # quantile), and both meet the 1e-4 bar on the same >= 99.4 % of the problems
    assert k["med"] < 1.5 * o["med"] and k["p99"] < 1.5 * o["p99"] and k["p999"] < 1.5 * o["p999"] and k["max"] < 3 * o["max"]
    assert k["within"] > 0.99 and abs(k["within"] - o["within"]) < 0.003
    assert k["flags"] > 0.9995 and k["both"] > 0.97


@pytest.mark.parametrize("params", ["defaults", "demo"])
def test_two_phase_scheduling_matches_oracle(params):
    """A batch larger than one resident wave runs BULK (with suspension of stragglers once the ticket queue is dry) +
    TAIL (continuation) -- results must be those of the plain loop: same flags, same iteration counts, same q."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 24000
    q0, tg, _ = make_workload(pb, om, B, seed=4242, standing=W.CASSIE_STANDING)
    if params == "defaults":
        prm, oprm = None, O.params()
    else:
        prm, oprm = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1), O.params(200, 0.1, 0.1)
    ref = O.dls_batch(opb, q0, tg, oprm, nthreads=NT)
    _compare("cassie two-phase %s" % params, _solve_gpu(pb, q0, tg, prm), ref, 1e-6)
    # the host path without an `iters` output uses the internal step-count scratch
    out = ik.dls_batch_host(pb, q0[:20000], tg[:20000], prm, "f64", "aos")
    assert np.array_equal(out["success"].astype(bool), ref[1][:20000].astype(bool))
    assert np.abs(out["q"] - ref[0][:20000]).max() < 1e-6
    # the host path cuts a two-launch batch into slices (H2D of slice c + 1 under the BULK launch of slice c): same
    # arithmetic per problem, so bit-identical to the unsliced call -- SoA (2-D slice copies) and AoS (dense slices)
    for layout, a, b in (("soa", q0.T.copy(), tg.T.copy()), ("aos", q0, tg)):
        os.environ["IKB_HOST_PIPELINE"] = "0"
        plain = ik.dls_batch_host(pb, a, b, prm, "f64", layout)
        del os.environ["IKB_HOST_PIPELINE"]
        os.environ["IKB_HOST_SLICES"] = "2"      # (a copy-in this short is not sliced by default: force the slices)
        try:
            piped = ik.dls_batch_host(pb, a, b, prm, "f64", layout)
        finally:
            del os.environ["IKB_HOST_SLICES"]
        for k in ("q", "success", "iters", "resid"):
            assert np.array_equal(plain[k], piped[k]), (layout, k)


@pytest.mark.parametrize("variant", ["dense", "arrow", "arrowb"])
def test_cassie_solve_variants_match_oracle(variant, monkeypatch):
    """Every compiled form of the Cassie step -- the dense 12 x 12 LDL^T on the solver role (r1) and the two layouts of
    the bordered-block-diagonal step (gen_solve_arrow: factor in shared memory / in registers) -- reproduces the oracle's flags, iteration counts and q on a BULK + TAIL batch, FP64."""
    monkeypatch.setenv("IKB_CASSIE_SOLVE", variant)
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    B = 24000
    q0, tg, _ = make_workload(pb, om, B, seed=777, standing=W.CASSIE_STANDING)
    ref = O.dls_batch(oracle_problem_like(pb, om), q0, tg, nthreads=NT)
    _compare("cassie solve variant %s" % variant, _solve_gpu(pb, q0, tg), ref, 1e-6)


@pytest.mark.parametrize("variant", ["uniform", "arrow"])
def test_humanoid_solve_variants_match_oracle(variant, monkeypatch):
    """The humanoid's two compiled steps (dense factorisation distributed over the roles / bordered block diagonal)."""
    monkeypatch.setenv("IKB_HUMANOID_SOLVE", variant)
    pb = W.humanoid_problem()
    om = oracle_model("humanoid")
    q0, tg, _ = make_workload(pb, om, 2048, seed=31, start="near")
    ref = O.dls_batch(oracle_problem_like(pb, om), q0, tg, nthreads=NT)
    _compare("humanoid solve variant %s" % variant, _solve_gpu(pb, q0, tg), ref, 1e-6, converged_only=True)


@pytest.mark.parametrize("ref", ["universe", "pelvis"])
@pytest.mark.parametrize("ktype", ["Full", "Position"])
def test_frame_constraint_null_space_projection(ktype, ref):
    """FrameConstraint (frame.hpp:333-465) in ik::dls (dls.cpp:26-34,44-52): the right foot is pinned (relative to the world
    or to the moving pelvis frame) while pelvis pose and left foot are tracked.  Table-driven kernel vs the oracle: flags,
    step counts, q; and the pinned frame really stays put."""
    m = W.cassie_model()
    pb = ik.InverseKinematicsProblem(m, 0)
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
    pb.add_frame_constraint("fr", ik.FrameConstraint(m, "RightFootFront", getattr(ik.KinematicType, ktype), ref))
    assert pb.c_size() == (6 if ktype == "Full" else 3) and pb.specialisation() is None
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 500
    q0, tg, _ = make_workload(pb, om, B, seed=91, standing=W.CASSIE_STANDING)
    # modest targets around the standing pose, so that a pinned foot is compatible with them
    s0 = W.standing_configuration(m, W.CASSIE_STANDING)
    lf = om.frame_placement(s0, om.frame_id("LeftFootFront"))[9:]
    rng = np.random.default_rng(3)
    tg[:, :9] = np.eye(3).reshape(-1)
    tg[:, 9:12] = rng.uniform(-0.03, 0.03, (B, 3))
    tg[:, 21:24] = lf + rng.uniform(-0.05, 0.05, (B, 3))
    prm, oprm = ik.dls_parameters(max_iterations=60, step_length=0.5), O.params(max_iterations=60, step_length=0.5)
    ref_out = O.dls_batch(opb, q0, tg, oprm, nthreads=NT)
    gpu = _solve_gpu(pb, q0, tg, prm)
    q, ok, it, res = gpu
    q_ref, ok_ref, it_ref, res_ref = ref_out
    assert (ok == ok_ref).all() and (it == it_ref).all()
    assert np.abs(q - q_ref).max() < 1e-6 and np.abs(res - res_ref).max() < 1e-9
    # the constrained frame has not moved relative to its reference frame (to first order along the path)
    f, r = om.frame_id("RightFootFront"), om.frame_id(ref)
    for b in range(0, B, 50):
        rel = lambda qq: O.se3_actinv(om.frame_placement(qq, r), om.frame_placement(qq, f))
        d = rel(q[b]) - rel(q0[b])
        assert np.abs(d[9:]).max() < 2e-3 and (ktype == "Position" or np.abs(d[:9]).max() < 2e-3)


@pytest.mark.parametrize("ref", ["universe", "LeftFootFront"])
def test_centre_of_mass_task(ref):
    """CentreOfMassTask (centre_of_mass.hpp:14-52; data.cpp:31-34 jacobianCenterOfMass): the centre of mass -- in the
    world, or relative to the moving stance foot -- is steered together with both foot positions and the pelvis
    orientation.  Table-driven kernel vs the oracle: flags, step counts, q, residuals; and converged solutions really put
    the centre of mass (computed independently from the URDF masses) on the target."""
    m = W.cassie_model()
    pb = ik.InverseKinematicsProblem(m, 0)
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
    com = pb.add_centre_of_mass_task(ik.CentreOfMassTask(m, ref))
    pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Position))
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Orientation))
    com.weighting()[:] = [2.0, 2.0, 0.5]
    assert pb.e_size(0) == 12 and pb.specialisation() is None and pb.get_centre_of_mass_task() is com
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 600
    q0, tg, qstar = make_workload(pb, om, B, seed=17, standing=W.CASSIE_STANDING)
    off, r = pb.target_offset(com), om.frame_id(ref)
    for b in range(B):
        oMr = om.frame_placement(qstar[b], r)
        tg[b, off:off + 3] = oMr[:9].reshape(3, 3).T @ (om.center_of_mass(qstar[b])[0] - oMr[9:])
    # Relative to a moving frame the task's Jacobian is inexact -- the reference does not differentiate the reference
    # frame's own motion -- so almost nothing converges and the iteration wanders off chaotically: there the first 6
    # steps of the trajectory are compared instead of its end.
    mi = 100 if ref == "universe" else 6
    q_ref, ok_ref, it_ref, res_ref = O.dls_batch(opb, q0, tg, O.params(mi), nthreads=NT)
    q, ok, it, res = _solve_gpu(pb, q0, tg, ik.dls_parameters(max_iterations=mi))
    assert (ok == ok_ref).all() and (it == it_ref).all() and (ok.mean() > 0.9 or ref != "universe")
    assert np.abs(q - q_ref)[ok].max(initial=0) < 1e-6 and np.abs(res - res_ref)[ok].max(initial=0) < 1e-9
    err = np.abs(q - q_ref).max(axis=1)
    print("com", ref, "converged", ok.mean(), "q err max", err.max(), "p99", np.percentile(err, 99))
    assert np.percentile(err, 99) < 1e-6 and err.max() < 1e-4
    for b in np.flatnonzero(ok)[:20]:
        oMr = om.frame_placement(q[b], r)
        c = oMr[:9].reshape(3, 3).T @ (om.center_of_mass(q[b])[0] - oMr[9:])
        assert np.abs(c - tg[b, off:off + 3]).max() < 2e-3  # stop test: |J^T e| < 1e-4, not |e|
    # FP32 build of the same kernel
    if ref == "universe":
        q32, ok32, _, _ = _solve_gpu(pb, q0, tg, dtype="f32")
        both = ok & ok32
        assert (ok32 == ok).mean() > 0.97 and np.percentile(np.abs(q32 - q)[both].max(axis=1), 99) < 1e-3
